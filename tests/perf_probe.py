"""Developer probe (not a test, not the bench): times the three device stages on a mid-size synthetic workload whose
inputs are produced by the compiled reference (oracle/_ref).  Usage: python tests/perf_probe.py [ref_mbp] [pairs] [read_len]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bwaflow_b200 import capi, synth  # noqa: E402
from oracle import api  # noqa: E402


def conv(arr, dtype):
    out = np.zeros(len(arr), dtype=dtype)
    for k in dtype.names:
        out[k] = arr[k]
    return out


def main():
    ref_mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 20
    pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    rl = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    reps = int(os.environ.get("REPS", "3"))
    t0 = time.time()
    ref = synth.make_reference(int(ref_mbp * 1e6), n_contigs=4, seed=64)
    td = tempfile.mkdtemp()
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))
    print("index built %.1fs" % (time.time() - t0), flush=True)
    R = api.Ref(os.path.join(td, "r"))
    if rl >= 250:
        reads = synth.simulate_pairs_indel(ref, pairs, rl, 0.05, seed=5)
    else:
        reads = synth.simulate_pairs(ref, pairs, rl, 0.01, seed=3)
    t0 = time.time()
    rb = R.batch(reads)
    rb.seq2chains()
    t_seed = time.time() - t0
    chains, seeds, coff = rb.chains()
    R.hook_begin()
    t0 = time.time()
    rb.chains2regions()
    t_ext = time.time() - t0
    jobs, qp, tp = R.hook_end()
    regs, roff = rb.regions()
    print("cpu 1-thread: seq2chains %.2fs (%.1f us/read)  chains2regions %.2fs (%.1f us/read); %d chains %d seeds %d regs %d jobs"
          % (t_seed, t_seed / reads.n * 1e6, t_ext, t_ext / reads.n * 1e6, len(chains), len(seeds), len(regs), len(jobs)), flush=True)
    nt4 = rb.nt4()
    index, keep = capi.make_index(R.view, R.ctg_off, R.ctg_len)
    ctx = capi.Context(0, index, keepalive=keep)

    # ---- ext
    ej = conv(jobs, capi.EXT_JOB)
    eb = capi.ExtBatch(ctx, len(ej), len(qp) + 1, len(tp) + 1)
    eb.submit(ej, qp, tp)
    res, t = eb.collect()
    ok = all((res[k] == jobs[k]).all() for k in capi.EXT_RES.names)
    cells = t.work_units
    print("ext: parity %s; e2e %.3f ms (h2d %.3f kernel %.3f d2h %.3f) cells %.3g  -> %.1f GCUPS (kernel+sort)"
          % (ok, t.total_ms, t.h2d_ms, t.kernel_ms, t.d2h_ms, cells, cells / t.kernel_ms / 1e6), flush=True)
    for _ in range(reps):
        eb.rerun()
        _, t = eb.collect()
        print("   rerun kernel %.3f ms -> %.1f GCUPS" % (t.kernel_ms, cells / t.kernel_ms / 1e6), flush=True)

    # ---- c2a
    cb = capi.C2ABatch(ctx, reads.n, len(nt4), len(chains) + 1, len(seeds) + 1)
    cb.submit(nt4, reads.offsets, conv(chains, capi.CHAIN), coff, conv(seeds, capi.SEED))
    gregs, goff, t = cb.collect()
    ok = gregs.tobytes() == regs.tobytes() and (goff == roff).all()
    print("c2a: parity %s; e2e %.3f ms (h2d %.3f kernel %.3f d2h %.3f) cells %.3g -> %.1f GCUPS; %.2f Mreads/s"
          % (ok, t.total_ms, t.h2d_ms, t.kernel_ms, t.d2h_ms, t.work_units, t.work_units / t.kernel_ms / 1e6,
             reads.n / t.kernel_ms / 1e3), flush=True)
    for _ in range(reps):
        cb.rerun()
        _, _, t = cb.collect()
        print("   rerun kernel %.3f ms -> %.1f GCUPS" % (t.kernel_ms, t.work_units / t.kernel_ms / 1e6), flush=True)

    # ---- smem
    eintv, eoff = rb.intervals()
    sb = capi.SmemBatch(ctx, reads.n, len(nt4))
    sb.submit(nt4, reads.offsets)
    intv, ioff, t = sb.collect()
    ok = intv.tobytes() == eintv.tobytes() and (ioff == eoff).all()
    print("smem: parity %s; e2e %.3f ms (h2d %.3f kernel %.3f d2h %.3f) blocks %.3g -> %.1f GB/s; %.2f Mreads/s"
          % (ok, t.total_ms, t.h2d_ms, t.kernel_ms, t.d2h_ms, t.work_units, t.work_units * 64 / t.kernel_ms / 1e6,
             reads.n / t.kernel_ms / 1e3), flush=True)
    for _ in range(reps):
        sb.rerun()
        _, _, t = sb.collect()
        print("   rerun kernel %.3f ms -> %.1f GB/s" % (t.kernel_ms, t.work_units * 64 / t.kernel_ms / 1e6), flush=True)
    pos, poff, t = sb.seed_positions()
    print("seedpos: %d lookups kernel %.3f ms blocks %.3g -> %.1f GB/s" % (len(pos), t.kernel_ms, t.work_units, t.work_units * 64 / t.kernel_ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
