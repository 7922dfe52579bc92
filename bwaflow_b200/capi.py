"""ctypes binding of libbwaflow_cuda.so (include/bwaflow_cuda.h).  Fails loudly when the CUDA library
is missing or no GPU is present -- there is no CPU fallback on this path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbwaflow_cuda.so")

INTV = np.dtype([("x0", "<u8"), ("x1", "<u8"), ("x2", "<u8"), ("info", "<u8")])
SEED = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4"), ("pad", "<i4")])
CHAIN = np.dtype([("pos", "<i8"), ("n_seeds", "<i4"), ("rid", "<i4"), ("frac_rep", "<f4"), ("flags", "<u4"),
                  ("seed_off", "<i8")])
ALNREG = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
                   ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"),
                   ("w", "<i4"), ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"),
                   ("seedlen0", "<i4"), ("n_comp_is_alt", "<u4"), ("frac_rep", "<f4"), ("hash", "<u8")])
EXT_JOB = np.dtype([("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"), ("h0", "<i4"), ("end_bonus", "<i4"),
                    ("zdrop", "<i4"), ("q_off", "<i8"), ("t_off", "<i8")])
EXT_RES = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"), ("gscore", "<i4"),
                    ("max_off", "<i4")])
SW_JOB = np.dtype([("qlen", "<i4"), ("tlen", "<i4"), ("xtra", "<i4"), ("pad", "<i4"), ("q_off", "<i8"), ("t_off", "<i8")])
SW_RES = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4")])
PESTAT = np.dtype([("low", "<i4"), ("high", "<i4"), ("failed", "<i4"), ("pad", "<i4"), ("avg", "<f8"), ("std", "<f8")])
ALN = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("flag", "<i4"), ("is_rev", "<i4"), ("is_alt", "<i4"), ("mapq", "<i4"), ("NM", "<i4"),
                ("n_cigar", "<i4"), ("score", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("global_score", "<i4"), ("pad", "<i4"),
                ("cigar_off", "<i8"), ("md_off", "<i8")])
assert ALN.itemsize == 72
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000
assert EXT_JOB.itemsize == 40 and EXT_RES.itemsize == 24 and ALNREG.itemsize == 88
assert SW_JOB.itemsize == 32 and SW_RES.itemsize == 28 and PESTAT.itemsize == 32


class Index(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.c_void_p), ("n_sa", C.c_uint64), ("sa_intv", C.c_int64), ("sa", C.c_void_p),
                ("l_pac", C.c_int64), ("pac", C.c_void_p), ("n_seqs", C.c_int32), ("_pad", C.c_int32),
                ("ctg_offset", C.c_void_p), ("ctg_len", C.c_void_p)]


class Opt(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_clip5", "pen_clip3", "w",
                                         "zdrop", "min_seed_len", "split_width", "max_occ", "max_mem_intv")] + \
               [("split_factor", C.c_float), ("mat", C.c_int8 * 25), ("_pad", C.c_int8 * 3)]


class ChainOpt(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("w", "max_chain_gap", "min_chain_weight", "max_chain_extend", "min_seed_len", "max_occ")] + \
               [("mask_level", C.c_float), ("drop_ratio", C.c_float)]


class PeOpt(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("pen_unpaired", "max_matesw", "max_ins", "max_chain_gap")] + \
               [("mask_level", C.c_float), ("mask_level_redun", C.c_float), ("no_rescue", C.c_int32), ("mapq_coef_len", C.c_float),
                ("mapq_coef_fac", C.c_int32), ("_pad", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("n_launches", C.c_int32), ("_pad", C.c_int32), ("work_units", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class HotTiming(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("h2d_ms", "smem_ms", "seedpos_ms", "chain_ms", "c2a_ms", "d2h_ms", "total_ms")] + \
               [("n_launches", C.c_int32), ("smem_blocks", C.c_double), ("seedpos_blocks", C.c_double), ("cells", C.c_double)] + \
               [(k, C.c_int64) for k in ("n_intv", "n_pos", "n_chains", "n_seeds")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BwaflowError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BwaflowError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(no CPU fallback exists for this path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.bwaflow_last_error.restype = C.c_char_p
        L.bwaflow_host_alloc.restype = C.c_void_p
        L.bwaflow_host_alloc.argtypes = [C.c_uint64]
        L.bwaflow_host_free.argtypes = [C.c_void_p]
        L.bwaflow_opt_default.argtypes = [C.POINTER(Opt)]
        L.bwaflow_ctx_create.argtypes = [C.c_int, C.POINTER(Index), C.POINTER(Opt), C.POINTER(C.c_void_p)]
        L.bwaflow_ctx_destroy.argtypes = [C.c_void_p]
        L.bwaflow_ctx_set_opt.argtypes = [C.c_void_p, C.POINTER(Opt)]
        L.bwaflow_ext_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        L.bwaflow_ext_batch_destroy.argtypes = [C.c_void_p]
        L.bwaflow_ext_submit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.bwaflow_ext_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(Timing)]
        L.bwaflow_ext_rerun.argtypes = [C.c_void_p]
        L.bwaflow_ext_rerun_counted.argtypes = [C.c_void_p]
        if hasattr(L, "bwaflow_c2a_batch_create"):
            L.bwaflow_c2a_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
            L.bwaflow_c2a_batch_destroy.argtypes = [C.c_void_p]
            L.bwaflow_c2a_submit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
            L.bwaflow_c2a_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Timing)]
            L.bwaflow_c2a_rerun.argtypes = [C.c_void_p]
            L.bwaflow_c2a_rerun_counted.argtypes = [C.c_void_p]
            L.bwaflow_c2a_exact_reads.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            L.bwaflow_c2a_second_round.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
            L.bwaflow_hot_second_round.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        if hasattr(L, "bwaflow_smem_batch_create"):
            L.bwaflow_smem_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
            L.bwaflow_smem_batch_destroy.argtypes = [C.c_void_p]
            L.bwaflow_smem_submit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
            L.bwaflow_smem_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Timing)]
            L.bwaflow_smem_rerun.argtypes = [C.c_void_p]
            L.bwaflow_smem_rerun_counted.argtypes = [C.c_void_p]
            L.bwaflow_smem_set_anchored.argtypes = [C.c_void_p, C.c_int]
            L.bwaflow_smem_seed_positions.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Timing)]
            L.bwaflow_sa_lookup.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.bwaflow_hot_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(ChainOpt), C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwaflow_hot_batch_destroy.argtypes = [C.c_void_p]
        L.bwaflow_hot_submit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.bwaflow_hot_rerun.argtypes = [C.c_void_p]
        L.bwaflow_hot_rerun_counted.argtypes = [C.c_void_p]
        L.bwaflow_hot_exact_reads.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.bwaflow_hot_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(HotTiming)]
        L.bwaflow_stopwatch_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwaflow_stopwatch_start.argtypes = [C.c_void_p]
        L.bwaflow_stopwatch_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.bwaflow_stopwatch_destroy.argtypes = [C.c_void_p]
        L.bwaflow_ctx_sync.argtypes = [C.c_void_p]
        L.bwaflow_dpx_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.bwaflow_index_build.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.bwaflow_index_build_gpu.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwaflow_built_index_view.restype = C.POINTER(Index)
        L.bwaflow_built_index_view.argtypes = [C.c_void_p]
        L.bwaflow_built_index_free.argtypes = [C.c_void_p]
        L.bwaflow_chain_opt_default.argtypes = [C.POINTER(ChainOpt)]
        L.bwaflow_chain_build.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.bwaflow_chain_result_view.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                                C.POINTER(C.c_int64), C.POINTER(C.c_void_p)]
        L.bwaflow_chain_result_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise BwaflowError("bwaflow error %d: %s" % (rc, lib().bwaflow_last_error().decode()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _view(ptr, count, dtype):
    """zero-copy numpy view of library-owned pinned memory"""
    if count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


def default_opt():
    o = Opt()
    lib().bwaflow_opt_default(C.byref(o))
    return o


def pinned(shape, dtype):
    """numpy array backed by pinned host memory (bwaflow_host_alloc); freed when garbage-collected."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    ptr = lib().bwaflow_host_alloc(max(1, n * dtype.itemsize))
    if not ptr:
        raise BwaflowError(lib().bwaflow_last_error().decode())
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PinnedOwner(arr, ptr)
    return arr


class _PinnedOwner:
    _live = {}

    def __init__(self, arr, ptr):
        import weakref
        self.ptr = ptr
        _PinnedOwner._live[ptr] = self
        weakref.finalize(arr, _PinnedOwner._release, ptr)

    @staticmethod
    def _release(ptr):
        if _PinnedOwner._live.pop(ptr, None) is not None and _lib is not None:
            _lib.bwaflow_host_free(ptr)


class Context:
    """One per GPU: index + options resident in HBM (BWAOCLEnv's role, src/fpga/BWAOCLEnv.h)."""

    def __init__(self, device=0, index: Index | None = None, opt: Opt | None = None, keepalive=None):
        self.L = lib()
        self.h = C.c_void_p()
        self._keep = keepalive
        self.opt = opt if opt is not None else default_opt()
        check(self.L.bwaflow_ctx_create(device, C.byref(index) if index is not None else None, C.byref(self.opt), C.byref(self.h)))

    def set_opt(self, opt: Opt):
        self.opt = opt
        check(self.L.bwaflow_ctx_set_opt(self.h, C.byref(opt)))

    def close(self):
        if self.h:
            self.L.bwaflow_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def dpx_peak(self):
        """(packed s16x2, scalar s32) DPX lane-instructions per second, measured."""
        a, b = C.c_double(), C.c_double()
        check(self.L.bwaflow_dpx_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sa_lookup(self, k):
        k = np.ascontiguousarray(k, dtype=np.uint64)
        out = np.zeros(len(k), dtype=np.uint64)
        check(self.L.bwaflow_sa_lookup(self.h, len(k), _p(k), _p(out)))
        return out


def make_index(view_like, ctg_off, ctg_len):
    """Build a bwaflow_index_t from any struct with the bwt_t/bntseq_t-shaped fields (e.g. the oracle's IndexView)
    plus contig tables.  Returns (Index, keepalive tuple)."""
    ix = Index()
    for k in ("primary", "seq_len", "bwt_size", "bwt", "n_sa", "sa_intv", "sa", "l_pac", "pac", "n_seqs"):
        setattr(ix, k, getattr(view_like, k))
    for i in range(5):
        ix.L2[i] = view_like.L2[i]
    co = np.ascontiguousarray(ctg_off, dtype=np.int64)
    cl = np.ascontiguousarray(ctg_len, dtype=np.int32)
    ix.ctg_offset = co.ctypes.data
    ix.ctg_len = cl.ctypes.data
    return ix, (co, cl, view_like)


class ExtBatch:
    """SWTask-shaped batch of raw ksw_extend2 jobs."""

    def __init__(self, ctx: Context, max_jobs, max_qbytes, max_tbytes):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        check(self.L.bwaflow_ext_batch_create(ctx.h, max_jobs, max_qbytes, max_tbytes, C.byref(self.h)))
        self.n = 0

    def submit(self, jobs, qpool, tpool):
        assert jobs.dtype == EXT_JOB
        self._keep = (jobs, qpool, tpool)
        self.n = len(jobs)
        check(self.L.bwaflow_ext_submit(self.h, len(jobs), _p(jobs), _p(qpool), len(qpool), _p(tpool), len(tpool)))

    def rerun(self, counted=False):
        """Device-resident re-run; counted=True uses the kernel variant that also counts the executed DP cells."""
        check(self.L.bwaflow_ext_rerun_counted(self.h) if counted else self.L.bwaflow_ext_rerun(self.h))

    def collect(self):
        out = C.c_void_p()
        t = Timing()
        check(self.L.bwaflow_ext_collect(self.h, C.byref(out), C.byref(t)))
        return _view(out.value, self.n, EXT_RES), t

    def close(self):
        if self.h:
            self.L.bwaflow_ext_batch_destroy(self.h)
            self.h = C.c_void_p()


class SwBatch:
    """Batch of raw ksw_align2 jobs (local SW with second-best score and start, bwa/ksw.c:343-365)."""

    def __init__(self, ctx: Context, max_jobs, max_qbytes, max_tbytes):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        self.L.bwaflow_sw_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        self.L.bwaflow_sw_submit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        self.L.bwaflow_sw_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(Timing)]
        self.L.bwaflow_sw_batch_destroy.argtypes = [C.c_void_p]
        check(self.L.bwaflow_sw_batch_create(ctx.h, max_jobs, max_qbytes, max_tbytes, C.byref(self.h)))
        self.n = 0

    def submit(self, jobs, qpool, tpool):
        assert jobs.dtype == SW_JOB
        self._keep = (jobs, qpool, tpool)
        self.n = len(jobs)
        check(self.L.bwaflow_sw_submit(self.h, len(jobs), _p(jobs), _p(qpool), len(qpool), _p(tpool), len(tpool)))

    def collect(self):
        out = C.c_void_p()
        t = Timing()
        check(self.L.bwaflow_sw_collect(self.h, C.byref(out), C.byref(t)))
        return _view(out.value, self.n, SW_RES), t

    def close(self):
        if self.h:
            self.L.bwaflow_sw_batch_destroy(self.h)
            self.h = C.c_void_p()


def default_pe_opt():
    po = PeOpt()
    lib().bwaflow_pe_opt_default(C.byref(po))
    return po


def pestat(opt, po, l_pac, regs, reg_off):
    """mem_pestat over one batch (host code of the library): returns PESTAT[4]."""
    L = lib()
    L.bwaflow_pestat.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    regs = np.ascontiguousarray(regs, dtype=ALNREG)
    reg_off = np.ascontiguousarray(reg_off, dtype=np.int64)
    pes = np.zeros(4, dtype=PESTAT)
    check(L.bwaflow_pestat(C.byref(opt), C.byref(po), l_pac, len(reg_off) - 1, _p(regs), _p(reg_off), _p(pes)))
    return pes


class MateswBatch:
    """The mate-rescue block of mem_sam_pe for a batch of pairs (bwaflow_matesw_*)."""

    def __init__(self, ctx: Context, max_reads, max_bases, po=None):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        L = self.L
        L.bwaflow_matesw_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwaflow_matesw_submit.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5
        L.bwaflow_matesw_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int64), C.POINTER(Timing)]
        L.bwaflow_matesw_batch_destroy.argtypes = [C.c_void_p]
        check(L.bwaflow_matesw_batch_create(ctx.h, max_reads, max_bases, C.byref(po) if po is not None else None, C.byref(self.h)))
        self.n = 0

    def submit(self, bases, read_off, regs, reg_off, pes):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        regs = np.ascontiguousarray(regs, dtype=ALNREG)
        reg_off = np.ascontiguousarray(reg_off, dtype=np.int64)
        pes = np.ascontiguousarray(pes, dtype=PESTAT)
        self._keep = (bases, read_off, regs, reg_off, pes)
        self.n = len(read_off) - 1
        check(self.L.bwaflow_matesw_submit(self.h, self.n, _p(bases), _p(read_off), _p(regs), _p(reg_off), _p(pes)))

    def collect(self):
        regs, off = C.c_void_p(), C.c_void_p()
        n, nsw, nj = C.c_int64(), C.c_int64(), C.c_int64()
        t = Timing()
        check(self.L.bwaflow_matesw_collect(self.h, C.byref(regs), C.byref(off), C.byref(n), C.byref(nsw), C.byref(nj), C.byref(t)))
        return {"alnreg": np.array(_view(regs.value, n.value, ALNREG)), "alnreg_off": np.array(_view(off.value, self.n + 1, np.dtype("<i8"))),
                "n_sw": nsw.value, "n_jobs": nj.value, "timing": t}

    def close(self):
        if self.h:
            self.L.bwaflow_matesw_batch_destroy(self.h)
            self.h = C.c_void_p()


class CigarBatch:
    """mem_reg2aln for a batch of (read, region) jobs (bwaflow_cigar_*)."""

    def __init__(self, ctx: Context, max_reads, max_bases, max_jobs, po=None):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        L = self.L
        L.bwaflow_cigar_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bwaflow_cigar_submit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.bwaflow_cigar_collect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(Timing)]
        L.bwaflow_cigar_batch_destroy.argtypes = [C.c_void_p]
        check(L.bwaflow_cigar_batch_create(ctx.h, max_reads, max_bases, max_jobs, C.byref(po) if po is not None else None, C.byref(self.h)))
        self.n = 0

    def submit(self, bases, read_off, job_read, job_reg):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        job_read = np.ascontiguousarray(job_read, dtype=np.int32)
        job_reg = np.ascontiguousarray(job_reg, dtype=ALNREG)
        self._keep = (bases, read_off, job_read, job_reg)
        self.n = len(job_read)
        check(self.L.bwaflow_cigar_submit(self.h, len(read_off) - 1, _p(bases), _p(read_off), self.n, _p(job_read), _p(job_reg)))

    def collect(self):
        """-> (ALN[n], list of CIGAR arrays, list of MD strings, timing)"""
        aln, cp, mp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        t = Timing()
        check(self.L.bwaflow_cigar_collect(self.h, C.byref(aln), C.byref(cp), C.byref(mp), C.byref(t)))
        a = np.array(_view(aln.value, self.n, ALN))
        n_cig = int(a["cigar_off"][-1] + a["n_cigar"][-1]) if self.n else 0
        cig = np.array(_view(cp.value, n_cig, np.dtype("<u4")))
        cigars = [cig[a["cigar_off"][j]:a["cigar_off"][j] + a["n_cigar"][j]] for j in range(self.n)]
        mds = [C.string_at(mp.value + int(a["md_off"][j])) for j in range(self.n)]
        return a, cigars, mds, t

    def dedup_patch(self, bases, read_off, regs, reg_off):
        """mem_sort_dedup_patch for every read (bwaflow_dedup_patch) -> (regs, reg_off, n_patch_jobs, rounds)"""
        L = self.L
        L.bwaflow_dedup_patch.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6 + [C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        regs = np.ascontiguousarray(regs, dtype=ALNREG)
        reg_off = np.ascontiguousarray(reg_off, dtype=np.int64)
        out = np.zeros(len(regs) + 1, dtype=ALNREG)
        out_off = np.zeros(len(reg_off), dtype=np.int64)
        nj, rounds = C.c_int64(), C.c_int32()
        check(L.bwaflow_dedup_patch(self.h, len(read_off) - 1, _p(bases), _p(read_off), _p(regs), _p(reg_off), _p(out), _p(out_off),
                                    C.byref(nj), C.byref(rounds)))
        return out[:out_off[-1]], out_off, nj.value, rounds.value

    def close(self):
        if self.h:
            self.L.bwaflow_cigar_batch_destroy(self.h)
            self.h = C.c_void_p()


class SamOpt(C.Structure):
    _fields_ = [("flag", C.c_int32), ("T", C.c_int32), ("drop_ratio", C.c_float), ("XA_drop_ratio", C.c_float),
                ("max_XA_hits", C.c_int32), ("max_XA_hits_alt", C.c_int32), ("n_threads", C.c_int32), ("_pad", C.c_int32)]


class SamTiming(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("dedup_ms", "matesw_ms", "pair_ms", "cigar_ms", "text_ms", "total_ms")] + \
               [(k, C.c_int64) for k in ("n_patch_jobs", "n_matesw_jobs", "n_matesw_used", "n_cigar_jobs", "sam_bytes")] + \
               [(k, C.c_float) for k in ("matesw_kernel_ms", "cigar_kernel_ms")]


def default_sam_opt():
    so = SamOpt()
    lib().bwaflow_sam_opt_default(C.byref(so))
    return so


def _blob(strings):
    """list of str/bytes -> (NUL-terminated blob as uint8 array, int64 offsets)"""
    bs = [x if isinstance(x, bytes) else x.encode() for x in strings]
    off = np.zeros(len(bs), dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(x) + 1 for x in bs[:-1]])
    blob = np.frombuffer(b"".join(x + b"\0" for x in bs) + b"\0", dtype=np.uint8).copy()
    return blob, off


class SamBatch:
    """Stage 5 for one batch: regions of stage 4 -> per-read SAM text (bwaflow_sam_*)."""

    def __init__(self, ctx: Context, max_reads, max_bases, ctg_names, po=None, so=None, ctg_is_alt=None, rg_id=None):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        L = self.L
        L.bwaflow_sam_batch_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 5 + [C.c_char_p, C.POINTER(C.c_void_p)]
        L.bwaflow_sam_process.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 9 + [C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                                                        C.c_void_p, C.POINTER(SamTiming)]
        L.bwaflow_sam_batch_destroy.argtypes = [C.c_void_p]
        L.bwaflow_sam_header.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.bwaflow_sam_header.restype = C.c_int64
        names = [x if isinstance(x, bytes) else x.encode() for x in ctg_names]
        arr = (C.c_char_p * len(names))(*names)
        alt = np.ascontiguousarray(ctg_is_alt, dtype=np.uint8) if ctg_is_alt is not None else None
        check(L.bwaflow_sam_batch_create(ctx.h, max_reads, max_bases, C.byref(po) if po is not None else None,
                                         C.byref(so) if so is not None else None, arr, _p(alt) if alt is not None else None, None,
                                         rg_id.encode() if rg_id else None, C.byref(self.h)))

    def process(self, bases, quals, read_off, names, regs, reg_off, start_idx=0, comments=None):
        """names: list of read names; quals: uint8 array parallel to bases or None -> (list of per-read SAM bytes, pestat, timing)"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        regs = np.ascontiguousarray(regs, dtype=ALNREG)
        reg_off = np.ascontiguousarray(reg_off, dtype=np.int64)
        quals = np.ascontiguousarray(quals, dtype=np.uint8) if quals is not None else None
        nb, no = _blob(names)
        cb = co = None
        if comments is not None:
            cb, co = _blob([c if c is not None else b"" for c in comments])
            co = np.where(np.array([c is None for c in comments]), -1, co).astype(np.int64)
        n = len(read_off) - 1
        sam, off = C.c_void_p(), C.c_void_p()
        pes = np.zeros(4, dtype=PESTAT)
        t = SamTiming()
        check(self.L.bwaflow_sam_process(self.h, n, _p(bases), _p(quals) if quals is not None else None, _p(read_off), _p(nb), _p(no),
                                         _p(cb) if cb is not None else None, _p(co) if co is not None else None, _p(regs), _p(reg_off),
                                         start_idx, C.byref(sam), C.byref(off), _p(pes), C.byref(t)))
        o = np.array(_view(off.value, n + 1, np.dtype("<i8")))
        text = C.string_at(sam.value, int(o[-1])) if o[-1] else b""
        return [text[o[i]:o[i + 1]] for i in range(n)], pes, t

    def header(self, hdr_line=None, pg_line=None):
        n = self.L.bwaflow_sam_header(self.h, hdr_line, pg_line, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        self.L.bwaflow_sam_header(self.h, hdr_line, pg_line, buf, n)
        return buf.raw[:n]

    def close(self):
        if self.h:
            self.L.bwaflow_sam_batch_destroy(self.h)
            self.h = C.c_void_p()


class C2ABatch:
    """ChainsToRegions-shaped batch: chains of many reads -> mem_alnreg_t lists."""

    def __init__(self, ctx: Context, max_reads, max_bases, max_chains, max_seeds):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        check(self.L.bwaflow_c2a_batch_create(ctx.h, max_reads, max_bases, max_chains, max_seeds, C.byref(self.h)))
        self.n = 0

    def submit(self, bases, read_off, chains, chain_off, seeds):
        assert chains.dtype == CHAIN and seeds.dtype == SEED
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        chain_off = np.ascontiguousarray(chain_off, dtype=np.int64)
        self._keep = (bases, read_off, chains, chain_off, seeds)
        self.n = len(read_off) - 1
        check(self.L.bwaflow_c2a_submit(self.h, self.n, _p(bases) if bases is not None else None, _p(read_off), _p(chains),
                                        _p(chain_off), _p(seeds), len(seeds)))

    def exact_reads(self):
        """Reads of the last collected launch that stage 4 ran on its exact lane-per-read kernel."""
        n = C.c_int64()
        check(self.L.bwaflow_c2a_exact_reads(self.h, C.byref(n)))
        return n.value

    def second_round(self):
        r, j = C.c_int64(), C.c_int64()
        check(self.L.bwaflow_c2a_second_round(self.h, C.byref(r), C.byref(j)))
        return r.value, j.value

    def rerun(self, counted=False):
        check(self.L.bwaflow_c2a_rerun_counted(self.h) if counted else self.L.bwaflow_c2a_rerun(self.h))

    def collect(self):
        regs, off = C.c_void_p(), C.c_void_p()
        n = C.c_int64()
        t = Timing()
        check(self.L.bwaflow_c2a_collect(self.h, C.byref(regs), C.byref(off), C.byref(n), C.byref(t)))
        return _view(regs.value, n.value, ALNREG), _view(off.value, self.n + 1, np.dtype("<i8")), t

    def close(self):
        if self.h:
            self.L.bwaflow_c2a_batch_destroy(self.h)
            self.h = C.c_void_p()


ANCHORED = (1 << 39) - 1


class SmemBatch:
    """SMemTask-shaped batch: reads -> sorted bwtintv_t lists (+ seed positions for chaining)."""

    def __init__(self, ctx: Context, max_reads, max_bases):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        check(self.L.bwaflow_smem_batch_create(ctx.h, max_reads, max_bases, C.byref(self.h)))
        self.n = 0

    def submit(self, bases, read_off):
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        self._keep = (bases, read_off)
        self.n = len(read_off) - 1
        check(self.L.bwaflow_smem_submit(self.h, self.n, _p(bases), _p(read_off)))

    def rerun(self, counted=False):
        """counted: the kernel instances that add up the reference's occ-block touches (timing.work_units)."""
        check(self.L.bwaflow_smem_rerun_counted(self.h) if counted else self.L.bwaflow_smem_rerun(self.h))

    def set_anchored(self, on):
        """Unique intervals may come back as {text position, ANCHORED, 1} (the hot path's internal form)."""
        check(self.L.bwaflow_smem_set_anchored(self.h, 1 if on else 0))

    def collect(self):
        a, off = C.c_void_p(), C.c_void_p()
        n = C.c_int64()
        t = Timing()
        check(self.L.bwaflow_smem_collect(self.h, C.byref(a), C.byref(off), C.byref(n), C.byref(t)))
        self.n_intv = n.value
        return _view(a.value, n.value, INTV), _view(off.value, self.n + 1, np.dtype("<i8")), t

    def seed_positions(self):
        a, off = C.c_void_p(), C.c_void_p()
        n = C.c_int64()
        t = Timing()
        check(self.L.bwaflow_smem_seed_positions(self.h, C.byref(a), C.byref(off), C.byref(n), C.byref(t)))
        return _view(a.value, n.value, np.dtype("<i8")), _view(off.value, self.n_intv + 1, np.dtype("<i8")), t

    def close(self):
        if self.h:
            self.L.bwaflow_smem_batch_destroy(self.h)
            self.h = C.c_void_p()


def default_chain_opt():
    o = ChainOpt()
    lib().bwaflow_chain_opt_default(C.byref(o))
    return o


def chain_build(l_pac, ctg_off, ctg_len, read_off, intv, intv_off, pos, pos_off, opt: ChainOpt | None = None, ctg_alt=None,
                n_threads=0, copy=True):
    """Host chaining (ChainsPipe's role): intervals + device-computed SA positions -> (chains, seeds, chain_off)."""
    L = lib()
    opt = opt if opt is not None else default_chain_opt()
    ctg_off = np.ascontiguousarray(ctg_off, dtype=np.int64)
    ctg_len = np.ascontiguousarray(ctg_len, dtype=np.int32)
    alt = np.ascontiguousarray(ctg_alt, dtype=np.int32) if ctg_alt is not None else None
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    intv_off = np.ascontiguousarray(intv_off, dtype=np.int64)
    pos_off = np.ascontiguousarray(pos_off, dtype=np.int64)
    intv = np.ascontiguousarray(intv)
    pos = np.ascontiguousarray(pos, dtype=np.int64)
    res = C.c_void_p()
    check(L.bwaflow_chain_build(C.byref(opt), int(l_pac), len(ctg_off), _p(ctg_off), _p(ctg_len), _p(alt) if alt is not None else None,
                                len(read_off) - 1, _p(read_off), _p(intv), _p(intv_off), _p(pos), _p(pos_off), int(n_threads), C.byref(res)))
    ch, sd, co = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nc, ns = C.c_int64(), C.c_int64()
    L.bwaflow_chain_result_view(res, C.byref(ch), C.byref(nc), C.byref(sd), C.byref(ns), C.byref(co))
    chains = _view(ch.value, nc.value, CHAIN).copy()
    seeds = _view(sd.value, ns.value, SEED).copy()
    off = _view(co.value, len(read_off), np.dtype("<i8")).copy()
    L.bwaflow_chain_result_free(res)
    return chains, seeds, off


class BuiltIndex:
    """FM-index of a synthetic reference built by the library (same arrays as `bwa index` + bwa_idx_load)."""

    def __init__(self, codes, contig_len, n_threads=0, device=None):
        """device=None: host builder (threads); device=i: the CUDA builder on GPU i."""
        self.L = lib()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        cl = np.ascontiguousarray(contig_len, dtype=np.int64)
        self.h = C.c_void_p()
        if device is None:
            check(self.L.bwaflow_index_build(_p(codes), len(codes), len(cl), _p(cl), int(n_threads), C.byref(self.h)))
        else:
            check(self.L.bwaflow_index_build_gpu(int(device), _p(codes), len(codes), len(cl), _p(cl), C.byref(self.h)))
        self.view = self.L.bwaflow_built_index_view(self.h).contents
        self.ctg_off = np.concatenate([[0], np.cumsum(cl)])[:-1].astype(np.int64)
        self.ctg_len = cl.astype(np.int32)

    def index(self):
        return self.view

    def bwt_array(self):
        return np.frombuffer((C.c_uint32 * self.view.bwt_size).from_address(self.view.bwt), dtype=np.uint32)

    def sa_array(self):
        return np.frombuffer((C.c_uint64 * self.view.n_sa).from_address(self.view.sa), dtype=np.uint64)

    def pac_array(self):
        return np.frombuffer((C.c_uint8 * (self.view.l_pac // 4 + 1)).from_address(self.view.pac), dtype=np.uint8)

    def close(self):
        if self.h:
            self.L.bwaflow_built_index_free(self.h)
            self.h = C.c_void_p()


class Stopwatch:
    """CUDA-event stopwatch on the context's device (bwaflow_stopwatch_*)."""

    def __init__(self, ctx: Context):
        self.L = ctx.L
        self.h = C.c_void_p()
        check(self.L.bwaflow_stopwatch_create(ctx.h, C.byref(self.h)))

    def start(self):
        check(self.L.bwaflow_stopwatch_start(self.h))

    def stop(self):
        ms = C.c_float()
        check(self.L.bwaflow_stopwatch_stop(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            self.L.bwaflow_stopwatch_destroy(self.h)
            self.h = C.c_void_p()


class HotBatch:
    """Fused reads -> regions batch (bwaflow_hot_*): every intermediate stays in HBM."""

    def __init__(self, ctx: Context, max_reads, max_bases, chain_opt: ChainOpt | None = None, ctg_alt=None):
        self.L = ctx.L
        self.ctx = ctx
        self.h = C.c_void_p()
        alt = np.ascontiguousarray(ctg_alt, dtype=np.int32) if ctg_alt is not None else None
        check(self.L.bwaflow_hot_batch_create(ctx.h, max_reads, max_bases, C.byref(chain_opt) if chain_opt is not None else None,
                                              _p(alt) if alt is not None else None, C.byref(self.h)))
        self.n = 0

    def submit(self, bases, read_off):
        read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        self._keep = (bases, read_off)
        self.n = len(read_off) - 1
        check(self.L.bwaflow_hot_submit(self.h, self.n, _p(bases), _p(read_off)))

    def exact_reads(self):
        n = C.c_int64()
        check(self.L.bwaflow_hot_exact_reads(self.h, C.byref(n)))
        return n.value

    def second_round(self):
        """(reads the first replay of stage 4 left, extension jobs the second round of waves ran ahead for them)"""
        r, j = C.c_int64(), C.c_int64()
        check(self.L.bwaflow_hot_second_round(self.h, C.byref(r), C.byref(j)))
        return r.value, j.value

    def rerun(self, counted=False):
        check(self.L.bwaflow_hot_rerun_counted(self.h) if counted else self.L.bwaflow_hot_rerun(self.h))

    def collect(self):
        regs, off = C.c_void_p(), C.c_void_p()
        n = C.c_int64()
        t = HotTiming()
        check(self.L.bwaflow_hot_collect(self.h, C.byref(regs), C.byref(off), C.byref(n), C.byref(t)))
        return _view(regs.value, n.value, ALNREG), _view(off.value, self.n + 1, np.dtype("<i8")), t

    def close(self):
        if self.h:
            self.L.bwaflow_hot_batch_destroy(self.h)
            self.h = C.c_void_p()
