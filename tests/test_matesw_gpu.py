"""GPU: mate rescue.  The ksw_align2 kernel against the reference's own ksw_align2 -- on the calls mem_matesw
really makes (observed by the link-time hook of oracle/_ref) and on random jobs covering the byte / word kernels, every xtra mode
and ragged shapes -- and the whole rescue block: regions after bwaflow_matesw_* byte-identical to the regions after the reference's
mem_matesw calls, on the golden cases and on a repeat-rich indel-heavy batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RES = ("score", "te", "qe", "score2", "te2", "tb", "qb")


def _sw_run(ctx, jobs, q, t):
    from bwaflow_b200 import capi
    ej = np.zeros(len(jobs), dtype=capi.SW_JOB)
    for k in ("qlen", "tlen", "xtra", "q_off", "t_off"):
        ej[k] = jobs[k]
    sb = capi.SwBatch(ctx, len(ej) + 1, len(q) + 16, len(t) + 16)
    sb.submit(ej, np.ascontiguousarray(q), np.ascontiguousarray(t))
    res, tm = sb.collect()
    res = np.array(res)
    sb.close()
    return res, tm


def _stage5_inputs(ix):
    """reference regions after mem_sort_dedup_patch + mem_pestat (inputs of the rescue block), and after the rescue block"""
    rb = ix.ref.batch(ix.reads, 0)
    rb.seq2chains()
    rb.chains2regions()
    pes = rb.dedup_pestat()
    regs1, off1 = rb.regions()
    regs1, off1 = np.array(regs1), np.array(off1)
    ix.ref.swhook_begin()
    n_sw = rb.matesw()
    jobs, q, t = ix.ref.swhook_end()
    regs2, off2 = rb.regions()
    regs2, off2 = np.array(regs2), np.array(off2)
    rb.free()
    return pes, regs1, off1, regs2, off2, n_sw, (jobs, q, t)


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_ksw_align2_on_the_calls_mem_matesw_makes(case, request):
    from bwaflow_b200 import capi
    ix = request.getfixturevalue("index_" + case)
    jobs, q, t = _stage5_inputs(ix)[6]
    assert len(jobs) > 0
    ctx = capi.Context(0)
    res, _ = _sw_run(ctx, jobs, q, t)
    for k in RES:
        assert (res[k] == jobs[k]).all(), (case, k, int((res[k] != jobs[k]).sum()), len(jobs))
    print("%s: %d ksw_align2 calls of mem_matesw, %d byte-mode" % (case, len(jobs), int((jobs["xtra"] & capi.KSW_XBYTE != 0).sum())))
    ctx.close()


def test_ksw_align2_random_jobs_all_modes():
    """random queries / targets with planted homologies: byte and word kernels, with and without KSW_XSUBO / KSW_XSTART, N bases in
    the query, queries from 1 to 260 bases (ragged last segment), empty and tiny targets"""
    from bwaflow_b200 import capi
    from oracle import api
    rng = np.random.default_rng(17)
    opt = capi.default_opt()
    mat = np.array(list(opt.mat), dtype=np.int8)
    recs, qs, ts = [], [], []
    qo = to = 0
    for it in range(400):
        ql = int(rng.choice([1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 64, 100, 149, 150, 151, 200, 249, 250, 260]))
        tl = int(rng.choice([0, 1, 5, 30, 150, 400, 700]))
        tgt = rng.integers(0, 4, tl).astype(np.uint8)
        qry = rng.integers(0, 4, ql).astype(np.uint8)
        if tl >= 30 and ql >= 8:                         # plant one or two diverged copies of pieces of the query
            for _ in range(int(rng.integers(1, 3))):
                ln = int(rng.integers(8, min(ql, tl) + 1))
                qs0, ts0 = int(rng.integers(0, ql - ln + 1)), int(rng.integers(0, tl - ln + 1))
                piece = qry[qs0:qs0 + ln].copy()
                mut = rng.random(ln) < 0.08
                piece[mut] = rng.integers(0, 4, int(mut.sum()))
                if rng.random() < 0.5 and ln > 20:       # an indel inside
                    cut = int(rng.integers(5, ln - 5))
                    piece = np.concatenate([piece[:cut], piece[cut + int(rng.integers(1, 4)):]])
                tgt[ts0:ts0 + len(piece)] = piece[:tl - ts0]
        if rng.random() < 0.2:
            qry[rng.integers(0, ql, max(1, ql // 20))] = 4
        xtra = 0
        if rng.random() < 0.8:
            xtra |= capi.KSW_XSUBO | int(rng.choice([5, 19, 30]))
        if rng.random() < 0.85:
            xtra |= capi.KSW_XSTART
        if ql * opt.a < 250 and rng.random() < 0.7:
            xtra |= capi.KSW_XBYTE
        exp = api.ref_ksw_align2(qry, tgt, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, xtra)
        recs.append((ql, tl, xtra, 0, qo, to) + tuple(int(x) for x in exp) + (0,))
        qs.append(qry); ts.append(tgt)
        qo += ql; to += tl
    jobs = np.array(recs, dtype=api.SWJOB)
    q, t = np.concatenate(qs), np.concatenate(ts + [np.zeros(1, np.uint8)])
    ctx = capi.Context(0)
    res, _ = _sw_run(ctx, jobs, q, t)
    for k in RES:
        bad = np.nonzero(res[k] != jobs[k])[0]
        assert len(bad) == 0, (k, bad[:5], [tuple(res[i]) for i in bad[:3]], [tuple(jobs[i])[6:13] for i in bad[:3]], [tuple(jobs[i])[:3] for i in bad[:3]])
    ctx.close()


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_matesw_regions_equal_reference(case, request):
    from bwaflow_b200 import capi
    g = request.getfixturevalue("golden_" + case)
    ix = request.getfixturevalue("index_" + case)
    pes, regs1, off1, regs2, off2, n_sw, hook = _stage5_inputs(ix)
    opt, po = capi.default_opt(), capi.default_pe_opt()
    mine = capi.pestat(opt, po, int(ix.ref.view.l_pac), regs1, off1)
    assert mine.tobytes() == pes.tobytes()                  # mem_pestat of the batch
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = ix.reads.nt4()
    mb = capi.MateswBatch(ctx, ix.reads.n, len(nt4), po)
    r1 = np.zeros(len(regs1), dtype=capi.ALNREG)
    for k in capi.ALNREG.names:
        r1[k] = regs1[k]
    mb.submit(nt4, ix.reads.offsets, r1, off1, mine)
    out = mb.collect()
    assert out["n_sw"] == n_sw
    assert (out["alnreg_off"] == off2).all()
    assert out["alnreg"].tobytes() == regs2.tobytes()
    print("%s: %d reads, %d windows run ahead, %d ksw_align2 calls consumed (reference: %d), regions %d -> %d"
          % (case, ix.reads.n, out["n_jobs"], out["n_sw"], len(hook[0]), len(regs1), len(regs2)))
    mb.close()
    ctx.close()
    B.close()


def test_matesw_repeat_rich_batch():
    """a batch where rescue fires often: 2x150 bp indel-heavy chimeric reads on a reference with a repeat family"""
    import os
    import tempfile
    from bwaflow_b200 import capi, synth
    from oracle import api
    ref = synth.make_reference(3_000_000, n_contigs=3, seed=21, repeat_copies=60, repeat_len=400)
    reads = synth.simulate_pairs_indel(ref, 6000, 150, 0.05, seed=23)
    td = tempfile.mkdtemp()
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))

    class IX:
        pass
    ix = IX()
    ix.ref = api.Ref(os.path.join(td, "r"))
    ix.reads = reads
    pes, regs1, off1, regs2, off2, n_sw, hook = _stage5_inputs(ix)
    opt, po = capi.default_opt(), capi.default_pe_opt()
    mine = capi.pestat(opt, po, int(ix.ref.view.l_pac), regs1, off1)
    assert mine.tobytes() == pes.tobytes()
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    mb = capi.MateswBatch(ctx, reads.n, len(nt4), po)
    r1 = np.zeros(len(regs1), dtype=capi.ALNREG)
    for k in capi.ALNREG.names:
        r1[k] = regs1[k]
    mb.submit(nt4, reads.offsets, r1, off1, mine)
    out = mb.collect()
    assert out["n_sw"] == n_sw and n_sw > 500
    assert (out["alnreg_off"] == off2).all()
    assert out["alnreg"].tobytes() == regs2.tobytes()
    print("repeat-rich: %d pairs, %d windows run ahead, %d consumed, regions %d -> %d, kernel %.2f ms"
          % (reads.n // 2, out["n_jobs"], out["n_sw"], len(regs1), len(regs2), out["timing"].kernel_ms))
    mb.close()
    ctx.close()
    B.close()
    ix.ref.close()
