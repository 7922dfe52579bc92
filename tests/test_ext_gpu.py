"""GPU parity: batched ksw_extend2 through the C ABI vs golden vectors (compiled reference) and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _jobs_from_golden(g):
    from bwaflow_b200 import capi
    j = g["jobs"]
    jobs = np.zeros(len(j), dtype=capi.EXT_JOB)
    for k in ("qlen", "tlen", "w", "h0", "end_bonus", "zdrop", "q_off", "t_off"):
        jobs[k] = j[k]
    exp = np.zeros(len(j), dtype=capi.EXT_RES)
    for k in ("score", "qle", "tle", "gtle", "gscore", "max_off"):
        exp[k] = j[k]
    return jobs, np.ascontiguousarray(g["qpool"]), np.ascontiguousarray(g["tpool"]), exp


@pytest.fixture(scope="module")
def ctx():
    from bwaflow_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["packed", "int32"])
def engine(request, monkeypatch):
    """Both extension engines: the packed two-jobs-per-lane s16x2 one (default) and the one-job-per-lane int32 one."""
    monkeypatch.setenv("BWAFLOW_EXT_ENGINE", request.param)
    return request.param


@pytest.mark.parametrize("case", ["golden_c1", "golden_c5"])
def test_ext_matches_golden(case, request, ctx, engine):
    from bwaflow_b200 import capi
    g = request.getfixturevalue(case)
    jobs, qp, tp, exp = _jobs_from_golden(g)
    b = capi.ExtBatch(ctx, len(jobs), len(qp), len(tp))
    b.submit(jobs, qp, tp)
    got, t = b.collect()
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], exp[bad[:5]], jobs[bad[:5]])
    # re-running (plain and cell-counting kernel variants) and resubmitting are idempotent
    b.rerun()
    got1, _ = b.collect()
    assert (got1 == exp).all()
    b.rerun(counted=True)
    got1, t1 = b.collect()
    assert (got1 == exp).all()
    assert t1.work_units == float(g["job_cells"].sum())          # executed DP cells == the restatement's count per call
    b.submit(jobs, qp, tp)
    got2, _ = b.collect()
    assert (got2 == exp).all()
    b.close()


def _random_jobs(rng, n, qmax, mat, opt, zdrops=(0, 10, 100), ws=(1, 5, 100, 200)):
    from bwaflow_b200 import capi
    from oracle import api
    jobs = np.zeros(n, dtype=capi.EXT_JOB)
    qs, ts, exp = [], [], np.zeros(n, dtype=capi.EXT_RES)
    qo = to = 0
    for i in range(n):
        ql = int(rng.integers(1, qmax + 1))
        tl = int(rng.integers(0 if i % 50 == 0 else 1, 2 * ql + 30))
        q = rng.integers(0, 5 if i % 7 == 0 else 4, ql).astype(np.uint8)
        t = rng.integers(0, 4, tl).astype(np.uint8)
        mode = rng.random()
        if mode < 0.8:            # target = mutated copy of the query, with indels
            src = [int(x) if x < 4 else 0 for x in q]
            out = []
            for x in src:
                r = rng.random()
                if r < 0.03:
                    out.append((x + 1) & 3)
                elif r < 0.05:
                    out.extend([x, int(rng.integers(0, 4))])
                elif r < 0.07:
                    pass
                else:
                    out.append(x)
            out = (out + list(rng.integers(0, 4, 40)))[:tl]
            t = np.array(out + [0] * (tl - len(out)), dtype=np.uint8)[:tl]
        w, zd = int(rng.choice(ws)), int(opt.zdrop)      # ksw_extend2 always gets opt->zdrop; the batch rejects anything else
        h0, eb = int(rng.integers(1, 200)), int(rng.choice([0, 5]))
        jobs[i] = (ql, tl, w, h0, eb, zd, qo, to)
        qs.append(q)
        ts.append(t)
        qo += ql
        to += tl
        exp[i] = tuple(api.orc_ksw_extend2(q, t, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, eb, opt.zdrop, h0))
    return jobs, np.concatenate(qs), np.concatenate(ts + [np.zeros(1, np.uint8)]), exp


@pytest.mark.parametrize("qmax,zdrop", [(40, 100), (200, 100), (600, 30), (1000, 0)])
def test_ext_matches_oracle_random(qmax, zdrop, ctx, engine):
    """All qlen classes (shared-memory row-state sizes), ragged/empty targets, N bases, tiny bands, z-drop on/off."""
    from bwaflow_b200 import capi
    rng = np.random.default_rng(qmax)
    opt = capi.default_opt()
    opt.zdrop = zdrop
    ctx.set_opt(opt)
    mat = np.array(list(opt.mat), dtype=np.int8)
    n = 1500 if qmax <= 200 else 300
    from oracle import api
    api.orc_lib().orc_counters_reset()
    jobs, qp, tp, exp = _random_jobs(rng, n, qmax, mat, opt)
    cells = int(api.orc_lib().orc_cells())
    b = capi.ExtBatch(ctx, n, len(qp), len(tp))
    b.submit(jobs, qp, tp)
    got, _ = b.collect()
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], exp[bad[:5]], jobs[bad[:5]])
    b.rerun(counted=True)
    got, tc = b.collect()
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], exp[bad[:5]], jobs[bad[:5]])
    assert tc.work_units == float(cells), (tc.work_units, cells)
    b.close()
    ctx.set_opt(capi.default_opt())


@pytest.mark.parametrize("refill", [1, 64])
def test_ext_packed_refill_policies(refill, ctx, monkeypatch):
    """Slots refilled one at a time (maximal desynchronisation of the two jobs of a lane) and only when a whole warp is idle."""
    from bwaflow_b200 import capi
    monkeypatch.setenv("BWAFLOW_EXT_REFILL_PAIR", str(refill))
    rng = np.random.default_rng(refill)
    opt = capi.default_opt()
    mat = np.array(list(opt.mat), dtype=np.int8)
    jobs, qp, tp, exp = _random_jobs(rng, 2500, 150, mat, opt, ws=(3, 100))
    b = capi.ExtBatch(ctx, len(jobs), len(qp), len(tp))
    b.submit(jobs, qp, tp)
    got, _ = b.collect()
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], exp[bad[:5]], jobs[bad[:5]])
    b.close()


def test_ext_general_matrix_uses_int32_engine(ctx):
    """A scoring matrix outside the bwa_fill_scmat form (here: transitions cheaper than transversions) cannot be packed."""
    from bwaflow_b200 import capi
    rng = np.random.default_rng(11)
    opt = capi.default_opt()
    m = list(opt.mat)
    m[0 * 5 + 2] = m[2 * 5 + 0] = m[1 * 5 + 3] = m[3 * 5 + 1] = -2
    for i, v in enumerate(m):
        opt.mat[i] = v
    ctx.set_opt(opt)
    mat = np.array(list(opt.mat), dtype=np.int8)
    jobs, qp, tp, exp = _random_jobs(rng, 600, 120, mat, opt)
    b = capi.ExtBatch(ctx, len(jobs), len(qp), len(tp))
    b.submit(jobs, qp, tp)
    got, _ = b.collect()
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], exp[bad[:5]], jobs[bad[:5]])
    b.close()
    ctx.set_opt(capi.default_opt())


def test_ext_rejects_oversized_jobs(ctx):
    from bwaflow_b200 import capi
    jobs = np.zeros(1, dtype=capi.EXT_JOB)
    jobs[0] = (2000, 10, 100, 20, 5, 100, 0, 0)
    b = capi.ExtBatch(ctx, 1, 4096, 4096)
    b.submit(jobs, np.zeros(2000, np.uint8), np.zeros(10, np.uint8))
    with pytest.raises(capi.BwaflowError):
        b.collect()
    b.close()


def test_ext_rejects_foreign_zdrop(ctx):
    """A job's zdrop must be the context's (mem_chain2aln passes opt->zdrop to every call); a different value is an error, not a
    silently different result."""
    from bwaflow_b200 import capi
    jobs = np.zeros(2, dtype=capi.EXT_JOB)
    jobs[0] = (4, 4, 100, 20, 5, 100, 0, 0)
    jobs[1] = (4, 4, 100, 20, 5, 37, 4, 4)
    b = capi.ExtBatch(ctx, 2, 64, 64)
    b.submit(jobs, np.zeros(8, np.uint8), np.zeros(8, np.uint8))
    with pytest.raises(capi.BwaflowError):
        b.collect()
    b.close()


def test_ext_empty_batch(ctx):
    from bwaflow_b200 import capi
    b = capi.ExtBatch(ctx, 16, 64, 64)
    b.submit(np.zeros(0, dtype=capi.EXT_JOB), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    got, _ = b.collect()
    assert len(got) == 0
    b.close()
