"""GPU parity: stage 2 (reads -> sorted bwtintv_t lists) and SA lookups through the C ABI vs the golden vectors of the
compiled reference (mem_collect_intv_new src/bwa_wrapper.cpp:207-246, bwt_sa bwa/bwt.c:86-96) and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx(ix, **opt_over):
    from bwaflow_b200 import capi
    index, keep = capi.make_index(ix.ref.view, ix.ref.ctg_off, ix.ref.ctg_len)
    opt = capi.default_opt()
    for k, v in opt_over.items():
        setattr(opt, k, v)
    return capi.Context(0, index, opt, keepalive=keep)


@pytest.fixture(params=["13", "default", "5", "0"])
def kmer_k(request, monkeypatch):
    """Depth K of the level tables (intervals of all 1..K-mers, smem_fast.cu): deeper than the index warrants (most
    K-mers absent or unique), the default for the index size, shallow, and 0 = every step on the index."""
    if request.param != "default":
        monkeypatch.setenv("BWAFLOW_KMER_K", request.param)
    return request.param


def _oracle_blocks(ix, nt4, off, n):
    """occ-block touches of mem_collect_intv_new for reads [0, n) as the oracle restatement counts them."""
    from oracle import api
    O = api.Oracle.from_ref(ix.ref)
    O.reset()
    for i in range(n):
        O.collect_intv(nt4[off[i]:off[i + 1]])
    return O.blocks()


def _same_modulo_anchoring(ix, got, exp):
    """Anchored form: same info and size; a unique interval may carry bwt_sa(x0) and the ANCHORED marker instead of rows."""
    from bwaflow_b200 import capi
    assert len(got) == len(exp)
    assert (got["info"] == exp["info"]).all() and (got["x2"] == exp["x2"]).all()
    anch = got["x1"] == capi.ANCHORED
    assert (got["x2"][anch] == 1).all()
    plain = ~anch
    assert (got["x0"][plain] == exp["x0"][plain]).all() and (got["x1"][plain] == exp["x1"][plain]).all()
    idx = np.nonzero(anch)[0]
    for i in idx[:: max(1, len(idx) // 400)]:
        assert int(got["x0"][i]) == ix.ref.bwt_sa(int(exp["x0"][i])), i
    return int(anch.sum())


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_smem_matches_golden(case, request, kmer_k):
    from bwaflow_b200 import capi
    ix = request.getfixturevalue("index_" + case)
    g = request.getfixturevalue("golden_" + case)
    ctx = _ctx(ix)
    nt4 = ix.reads.nt4()
    b = capi.SmemBatch(ctx, ix.reads.n, len(nt4))
    b.submit(nt4, ix.reads.offsets)
    intv, off, t = b.collect()
    assert (off == g["intv_off"]).all(), np.nonzero(off != g["intv_off"])[0][:5]
    assert intv.tobytes() == g["intv"].tobytes()
    # seed positions = bwt_sa over each interval exactly as mem_chain_postprocess samples them
    pos, poff, _ = b.seed_positions()
    max_occ = ctx.opt.max_occ
    rng = np.random.default_rng(1)
    for i in rng.integers(0, len(intv), 300):
        x0, x2 = int(intv["x0"][i]), int(intv["x2"][i])
        step = x2 // max_occ if x2 > max_occ else 1
        ks = [x0 + k for k in range(0, x2, step)][:max_occ]
        exp = [ix.ref.bwt_sa(k) for k in ks]
        assert list(pos[poff[i]:poff[i + 1]]) == exp
    b.rerun()
    intv2, off2, _ = b.collect()
    assert intv2.tobytes() == g["intv"].tobytes() and (off2 == off).all()
    # the counting instances: same intervals, and the algorithmic work unit of the seeding roofline (64-byte occ-block
    # touches of the reference) equal to the oracle's count
    b.rerun(counted=True)
    intv3, off3, t3 = b.collect()
    assert intv3.tobytes() == g["intv"].tobytes() and (off3 == off).all()
    assert int(t3.work_units) == _oracle_blocks(ix, nt4, ix.reads.offsets, ix.reads.n) == int(g["blocks_total"])
    # the hot path's internal form: unique intervals anchored in the text (positions instead of rows)
    b.set_anchored(True)
    b.rerun()
    intv4, off4, _ = b.collect()
    assert (off4 == off).all()
    n_anch = _same_modulo_anchoring(ix, intv4, g["intv"])
    assert n_anch > 0
    pos4, poff4, _ = b.seed_positions()
    assert (poff4 == poff).all() and (pos4 == pos).all()
    b.close()
    ctx.close()


def test_sa_lookup_known_answers(index_c1, golden_c1):
    ctx = _ctx(index_c1)
    got = ctx.sa_lookup(golden_c1["sa_k"])
    assert (got == golden_c1["sa_v"]).all()
    ctx.close()


def test_smem_edge_reads(index_c1, kmer_k):
    """All-N reads, reads shorter than min_seed_len, N-riddled reads, exact reference substrings (both strands),
    a highly repetitive read, mixed lengths in one batch; checked against the oracle restatement."""
    from bwaflow_b200 import capi
    from oracle import api
    ix = index_c1
    O = api.Oracle.from_ref(ix.ref)
    rng = np.random.default_rng(11)
    pac = ix.ref.pac_array()
    L = ix.ref.view.l_pac

    def refsub(p, n):
        idx = np.arange(p, p + n)
        return ((pac[idx >> 2] >> ((~idx & 3) << 1)) & 3).astype(np.uint8)

    reads = [np.full(60, 4, np.uint8), np.zeros(10, np.uint8), np.zeros(18, np.uint8), np.zeros(19, np.uint8)]
    for _ in range(40):
        n = int(rng.integers(19, 250))
        p = int(rng.integers(0, L - n))
        s = refsub(p, n)
        if rng.random() < 0.5:
            s = (3 - s)[::-1].copy()
        k = rng.random()
        if k < 0.3:
            s[rng.random(n) < 0.1] = 4
        elif k < 0.6:
            m = rng.random(n) < 0.05
            s[m] = (s[m] + 1) & 3
        reads.append(s)
    reads.append(np.tile(np.array([0, 1], np.uint8), 70))
    reads.append(np.zeros(150, np.uint8))
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.concatenate(reads)
    ctx = _ctx(ix)
    b = capi.SmemBatch(ctx, len(reads), len(bases))
    b.submit(bases, off)
    intv, ioff, _ = b.collect()
    O.reset()
    exp_all = []
    for i, r in enumerate(reads):
        exp = O.collect_intv(r)
        exp_all.append(exp)
        got = intv[ioff[i]:ioff[i + 1]]
        assert got.tobytes() == exp.tobytes(), (i, len(r), got, exp)
    exp_blocks = O.blocks()
    b.rerun(counted=True)
    intv2, ioff2, t2 = b.collect()
    assert intv2.tobytes() == intv.tobytes() and (ioff2 == ioff).all()
    assert int(t2.work_units) == exp_blocks
    b.set_anchored(True)
    b.rerun()
    intv3, ioff3, _ = b.collect()
    assert (ioff3 == ioff).all()
    _same_modulo_anchoring(ix, intv3, np.concatenate(exp_all))
    b.close()
    ctx.close()


def test_smem_slot_overflow_regrows(index_c1, golden_c1):
    """Seeding knobs that make every read emit many intervals: the per-read slot overflows and the batch is redone
    on the GPU with larger slots (no CPU redo as in FPGAPipeline.cpp:706-714)."""
    from bwaflow_b200 import capi
    from oracle import api
    ix = index_c1
    ctx = _ctx(ix, min_seed_len=5, max_mem_intv=1000000)
    oo = ix.ref.orc_opt()
    oo.min_seed_len = 5
    oo.max_mem_intv = 1000000
    O = api.Oracle(ix.ref.view, oo, ix.ref.ctg_off, ix.ref.ctg_len, keepalive=ix.ref)
    nt4 = ix.reads.nt4()
    n = 64
    off = ix.reads.offsets[:n + 1]
    b = capi.SmemBatch(ctx, n, int(off[-1]))
    b.submit(nt4[:off[-1]], off)
    intv, ioff, _ = b.collect()
    assert (np.diff(ioff) > 32).any()
    for i in range(n):
        exp = O.collect_intv(nt4[off[i]:off[i + 1]])
        assert intv[ioff[i]:ioff[i + 1]].tobytes() == exp.tobytes(), i
    b.close()
    ctx.close()


@pytest.mark.parametrize("case,heavy_min", [("c1", 1), ("c5", 1), ("c5", 12)])
def test_smem_warp_per_read_kernel_matches_golden(case, heavy_min, request, monkeypatch):
    """smem_heavy_kernel (one warp per read: the reads the sweep kernel hands over because a call found many forward matches, i.e.
    reads inside repeats).  heavy_min = 1 hands EVERY read over, so the kernel is held to the golden intervals on all reads of both
    cases; 12 mixes the two kernels inside one batch (a read may be handed over in the middle of pass 1 or in pass 2).  Also in the
    hot path's anchored mode, where its rows and the sweep kernel's text-anchored entries meet in one interval list."""
    from bwaflow_b200 import capi
    monkeypatch.setenv("BWAFLOW_SMEM_HEAVY_MIN", str(heavy_min))
    ix = request.getfixturevalue("index_" + case)
    g = request.getfixturevalue("golden_" + case)
    ctx = _ctx(ix)
    nt4 = ix.reads.nt4()
    b = capi.SmemBatch(ctx, ix.reads.n, len(nt4))
    b.submit(nt4, ix.reads.offsets)
    intv, off, t = b.collect()
    assert (off == g["intv_off"]).all(), np.nonzero(off != g["intv_off"])[0][:5]
    assert intv.tobytes() == g["intv"].tobytes()
    b.set_anchored(True)
    b.rerun()
    intv4, off4, _ = b.collect()
    assert (off4 == off).all()
    _same_modulo_anchoring(ix, intv4, g["intv"])
    b.close()
    ctx.close()
