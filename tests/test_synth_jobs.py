"""CPU: bench.py's ksw_extend2 micro-benchmark input (synth.extension_jobs) is well formed and behaves like mem_chain2aln's jobs."""
import numpy as np


def test_synthetic_extension_jobs_are_well_formed():
    """Offsets tile the pools, flank/window sizes follow mem_chain2aln (bwa/bwamem.c:724-771), and the oracle extends most clean
    flanks to their end."""
    from bwaflow_b200 import capi, synth
    from oracle import api
    ref = synth.make_reference(300_000, n_contigs=2, seed=2)
    reads = synth.simulate_pairs(ref, 2000, 150, 0.01, seed=7)
    jobs, qp, tp = synth.extension_jobs(ref, reads)
    n = len(jobs["qlen"])
    assert n > 1000
    assert (jobs["qlen"] > 0).all() and (jobs["qlen"] < 150).all() and (jobs["h0"] >= 19).all()
    assert jobs["q_off"][0] == 0 and (np.diff(jobs["q_off"]) == jobs["qlen"][:-1]).all() and jobs["q_off"][-1] + jobs["qlen"][-1] == len(qp)
    assert jobs["t_off"][0] == 0 and (np.diff(jobs["t_off"]) == jobs["tlen"][:-1]).all() and jobs["t_off"][-1] + jobs["tlen"][-1] == len(tp)
    assert (jobs["tlen"] <= jobs["qlen"] * 2 + 200).all() and qp.max() <= 4 and tp.max() <= 3
    mat = np.array(list(_opt().mat), dtype=np.int8)
    opt = _opt()
    full = tried = 0
    for i in range(0, n, max(1, n // 60)):
        q = qp[jobs["q_off"][i]:jobs["q_off"][i] + jobs["qlen"][i]]
        t = tp[jobs["t_off"][i]:jobs["t_off"][i] + jobs["tlen"][i]]
        r = api.orc_ksw_extend2(q, t, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, int(jobs["w"][i]), int(jobs["end_bonus"][i]),
                                opt.zdrop, int(jobs["h0"][i]))
        assert r[0] >= jobs["h0"][i]
        tried += 1
        full += int(r[1] == jobs["qlen"][i])
    assert full * 2 >= tried          # most flanks of 1%-error reads extend to the end of the read


def _opt():
    from bwaflow_b200 import capi
    return capi.default_opt()
