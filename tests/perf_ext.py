"""Developer probe: steady-state throughput of the raw extension-job kernel on a large job set (the jobs mem_chain2aln
issued for a mid-size read set, dumped through the compiled reference's hook and replicated)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bwaflow_b200 import capi, synth  # noqa: E402
from oracle import api  # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    rl = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    ref = synth.make_reference(int(10e6), n_contigs=2, seed=64)
    td = tempfile.mkdtemp()
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    v = B.view
    api.write_index_files(os.path.join(td, "r"), v.primary, list(v.L2), v.seq_len, B.bwt_array(), B.sa_array(), v.sa_intv, B.pac_array(),
                          v.l_pac, ref.names, [int(x) for x in ref.lengths])
    R = api.Ref(os.path.join(td, "r"))
    reads = synth.simulate_pairs_indel(ref, pairs, rl, 0.05, seed=5) if rl >= 250 else synth.simulate_pairs(ref, pairs, rl, 0.01, seed=3)
    rb = R.batch(reads)
    rb.seq2chains()
    R.hook_begin()
    rb.chains2regions()
    jobs, qp, tp = R.hook_end()
    n0 = len(jobs)
    ej = np.zeros(n0 * rep, dtype=capi.EXT_JOB)
    for k in ("qlen", "tlen", "w", "h0", "end_bonus", "zdrop"):
        ej[k] = np.tile(jobs[k], rep)
    ej["q_off"] = (np.tile(jobs["q_off"], rep) + np.repeat(np.arange(rep, dtype=np.int64) * len(qp), n0))
    ej["t_off"] = (np.tile(jobs["t_off"], rep) + np.repeat(np.arange(rep, dtype=np.int64) * len(tp), n0))
    qpp, tpp = np.tile(qp, rep), np.tile(tp, rep)
    print("jobs", len(ej), "qbytes", len(qpp), "tbytes", len(tpp), "mean qlen %.1f tlen %.1f" % (jobs["qlen"].mean(), jobs["tlen"].mean()), flush=True)
    ctx = capi.Context(0)
    eb = capi.ExtBatch(ctx, len(ej), len(qpp) + 1, len(tpp) + 1)
    eb.submit(ej, qpp, tpp)
    res, t = eb.collect()
    eb.rerun(counted=True)
    res2, tc = eb.collect()
    assert all((res2[k] == res[k]).all() for k in capi.EXT_RES.names), "counted and plain kernel variants disagree"
    t.work_units = tc.work_units
    ok = all((res[k][:n0] == jobs[k]).all() for k in capi.EXT_RES.names) and all((res[k][-n0:] == jobs[k]).all() for k in capi.EXT_RES.names)
    cells = t.work_units
    print("parity %s; e2e %.2f ms (h2d %.2f kernel %.2f d2h %.2f); cells %.4g (%.0f per job)" % (ok, t.total_ms, t.h2d_ms, t.kernel_ms, t.d2h_ms, cells, cells / len(ej)))
    for _ in range(3):
        eb.rerun()
        _, t = eb.collect()
        print("   resident kernel+sort %.2f ms -> %.1f GCUPS, %.1f M jobs/s" % (t.kernel_ms, cells / t.kernel_ms / 1e6, len(ej) / t.kernel_ms / 1e3), flush=True)


if __name__ == "__main__":
    main()
