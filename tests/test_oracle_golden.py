"""CPU: the C restatement (oracle/oracle.c) against the golden vectors produced by the compiled reference
(tests/golden/make_golden.py) and, where oracle/_ref is available, against the compiled reference itself."""
import numpy as np
import pytest

from oracle import api


@pytest.mark.parametrize("case", ["golden_c1", "golden_c5"])
def test_ksw_extend2_golden_jobs(case, request):
    g = request.getfixturevalue(case)
    jobs, qp, tp = g["jobs"], g["qpool"], g["tpool"]
    a, b, od, ed, oi, ei = [int(x) for x in g["opt"][:6]]
    mat = np.array([[a if i == j else -b for j in range(4)] + [-1] for i in range(4)] + [[-1] * 5], dtype=np.int8).reshape(-1)
    for j in jobs:
        q = qp[j["q_off"]:j["q_off"] + j["qlen"]]
        t = tp[j["t_off"]:j["t_off"] + j["tlen"]]
        got = api.orc_ksw_extend2(q, t, mat, od, ed, oi, ei, int(j["w"]), int(j["end_bonus"]), int(j["zdrop"]), int(j["h0"]))
        exp = [j[k] for k in ("score", "qle", "tle", "gtle", "gscore", "max_off")]
        assert list(got) == exp


@pytest.mark.parametrize("case", ["index_c1", "index_c5"])
def test_restatement_against_golden(case, request):
    ix = request.getfixturevalue(case)
    g = request.getfixturevalue(case.replace("index", "golden"))
    O = api.Oracle.from_ref(ix.ref)
    nt4 = ix.reads.nt4()
    off = ix.reads.offsets
    for k, v in zip(g["sa_k"], g["sa_v"]):
        assert O.bwt_sa(int(k)) == int(v)
    for i in range(ix.reads.n):
        q = nt4[off[i]:off[i + 1]]
        mine = O.collect_intv(q)
        exp = g["intv"][g["intv_off"][i]:g["intv_off"][i + 1]]
        assert mine.tobytes() == exp.tobytes(), "read %d intervals" % i
        ch = g["chains"][g["chain_off"][i]:g["chain_off"][i + 1]]
        regs = O.chains2aln(q, ch, g["seeds"])
        exp = g["regs"][g["reg_off"][i]:g["reg_off"][i + 1]]
        assert regs.tobytes() == exp.tobytes(), "read %d regions" % i


def test_reference_reproduces_golden(index_c1, golden_c1):
    """The compiled reference run here gives the committed vectors (guards the fixture itself), incl. SAM bytes,
    and the accelerator return path (ChainsPipe, src/Pipeline.cpp:408) equals the CPU stage-2 path."""
    ix, g = index_c1, golden_c1
    b = ix.ref.batch(ix.reads, 0)
    intv, ioff = b.intervals()
    assert intv.tobytes() == g["intv"].tobytes() and (ioff == g["intv_off"]).all()
    b.chainspipe(intv, ioff)
    ch, sd, coff = b.chains()
    assert ch.tobytes() == g["chains"].tobytes() and sd.tobytes() == g["seeds"].tobytes()
    b.chains2regions()
    regs, roff = b.regions()
    assert regs.tobytes() == g["regs"].tobytes()
    sam, soff = b.regions2sam()
    assert sam == g["sam"].tobytes()
    b.free()


def test_extend_and_edge_cases(index_c1):
    ix = index_c1
    O = api.Oracle.from_ref(ix.ref)
    rng = np.random.default_rng(3)
    v = ix.ref.view
    # bwt_extend on random valid intervals, both directions
    for _ in range(300):
        c = int(rng.integers(0, 4))
        ik = [v.L2[c] + 1, v.L2[3 - c] + 1, v.L2[c + 1] - v.L2[c]]
        for _ in range(int(rng.integers(1, 12))):
            back = int(rng.integers(0, 2))
            a = ix.ref.bwt_extend(ik, back)
            m = O.bwt_extend(ik, back)
            assert (a == m).all()
            nxt = a[int(rng.integers(0, 4))]
            if nxt[2] == 0:
                break
            ik = [int(x) for x in nxt]
    # reads that are all N, shorter than min_seed_len, or empty give no intervals in both
    for q in (np.full(50, 4, np.uint8), np.zeros(10, np.uint8), np.zeros(0, np.uint8)):
        assert len(O.collect_intv(q)) == 0
        assert len(ix.ref.collect_intv(q)) == 0
    # ksw edge cases: qlen 1, tlen 1, tiny band, z-drop exit, N in query
    mat = ix.ref.mat()
    for _ in range(400):
        ql, tl = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        q = rng.integers(0, 5, ql).astype(np.uint8)
        t = rng.integers(0, 4, tl).astype(np.uint8)
        if rng.random() < 0.7:
            n = min(ql, tl)
            t[:n] = np.where(q[:n] < 4, q[:n], 0)
        args = (6, 1, 6, 1, int(rng.choice([1, 5, 100, 200])), int(rng.choice([0, 5])), int(rng.choice([0, 10, 100])), int(rng.integers(1, 120)))
        assert list(api.orc_ksw_extend2(q, t, mat, *args)) == list(api.ref_ksw_extend2(q, t, mat, *args))


@pytest.mark.parametrize("case", ["index_c1", "index_c5"])
def test_leftmost_begin_formulation(case, request):
    """The seeding kernel (smem_sweep.cu) does not replay the reference's back-enlarge / forward-enlarge routes: it computes,
    per forward match, the leftmost begin whose interval keeps min_intv rows.  The CPU twin of that formulation gives the
    golden intervals on the golden reads and agrees with the route-by-route restatement on adversarial reads (N runs,
    homopolymers, tandem repeats, exact reference substrings of both strands, mutated copies of the repeat family)."""
    ix = request.getfixturevalue(case)
    g = request.getfixturevalue(case.replace("index", "golden"))
    O = api.Oracle.from_ref(ix.ref)
    nt4 = ix.reads.nt4()
    off = ix.reads.offsets
    for i in range(ix.reads.n):
        exp = g["intv"][g["intv_off"][i]:g["intv_off"][i + 1]]
        assert O.collect_intv_leftmost(nt4[off[i]:off[i + 1]]).tobytes() == exp.tobytes(), i
    rng = np.random.default_rng(17)
    codes = g["ref_codes"]
    L = len(codes)
    for t in range(300):
        n = int(rng.integers(19, 250))
        p = int(rng.integers(0, L - n))
        s = codes[p:p + n].copy()
        s[s > 3] = 4
        if rng.random() < 0.5:
            s = np.where(s < 4, 3 - s, 4)[::-1].astype(np.uint8)
        k = t % 6
        if k == 1:
            s[rng.random(n) < 0.08] = 4
        elif k == 2:
            m = rng.random(n) < 0.04
            s[m] = (s[m] + 1) % 4
        elif k == 3:
            s[n // 3:n // 3 + 30] = int(rng.integers(0, 4))
        elif k == 4:
            unit = rng.integers(0, 4, int(rng.integers(2, 7)), dtype=np.uint8)
            s[n // 2:] = np.resize(unit, n - n // 2)
        elif k == 5:
            s[:n // 2] = s[n - n // 2:][::-1]
        s = np.ascontiguousarray(s, dtype=np.uint8)
        assert O.collect_intv_leftmost(s).tobytes() == O.collect_intv(s).tobytes(), (t, k)
