"""TEST INFRASTRUCTURE ONLY: ctypes bindings of oracle/_ref/libbwaref.so (the reference's own
CPU path, oracle/ref_driver.cpp) and oracle/liboracle.so (our C restatement, oracle/oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under bwaflow_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libbwaref.so")
ORC_SO = os.path.join(HERE, "liboracle.so")

INTV = np.dtype([("x0", "<u8"), ("x1", "<u8"), ("x2", "<u8"), ("info", "<u8")])
SEED = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4"), ("pad", "<i4")])
CHAIN = np.dtype([("pos", "<i8"), ("n_seeds", "<i4"), ("rid", "<i4"), ("frac_rep", "<f4"), ("flags", "<u4"),
                  ("seed_off", "<i8")])
ALNREG = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
                   ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"),
                   ("w", "<i4"), ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"),
                   ("seedlen0", "<i4"), ("n_comp_is_alt", "<u4"), ("frac_rep", "<f4"), ("hash", "<u8")])
JOB = np.dtype([("qlen", "<i4"), ("tlen", "<i4"), ("w", "<i4"), ("h0", "<i4"), ("end_bonus", "<i4"), ("zdrop", "<i4"),
                ("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"), ("gscore", "<i4"),
                ("max_off", "<i4"), ("q_off", "<i8"), ("t_off", "<i8")])
assert INTV.itemsize == 32 and SEED.itemsize == 24 and CHAIN.itemsize == 32 and ALNREG.itemsize == 88 and JOB.itemsize == 64


class IndexView(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.c_void_p), ("n_sa", C.c_uint64), ("sa_intv", C.c_int64), ("sa", C.c_void_p),
                ("l_pac", C.c_int64), ("pac", C.c_void_p), ("n_seqs", C.c_int32), ("pad", C.c_int32)]


class OrcOpt(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_clip5", "pen_clip3", "w",
                                         "zdrop", "min_seed_len", "split_width", "max_occ", "max_mem_intv")] + \
               [("split_factor", C.c_float), ("mat", C.c_int8 * 25), ("pad", C.c_int8 * 3)]


def ensure_built():
    """Build liboracle.so and (when /root/reference is present) oracle/_ref; building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def have_ref():
    return os.path.exists(REF_SO)


_ref = None
_orc = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            ensure_built()
        L = C.CDLL(REF_SO)
        L.bwaref_open.restype = C.c_void_p
        L.bwaref_open.argtypes = [C.c_char_p, C.c_int]
        L.bwaref_close.argtypes = [C.c_void_p]
        L.bwaref_index_build.argtypes = [C.c_char_p, C.c_char_p]
        L.bwaref_opt.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.bwaref_opt_split_factor.argtypes = [C.c_void_p]
        L.bwaref_opt_split_factor.restype = C.c_float
        L.bwaref_opt_mat.argtypes = [C.c_void_p, C.c_void_p]
        L.bwaref_index_view.argtypes = [C.c_void_p, C.POINTER(IndexView)]
        L.bwaref_contigs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bwaref_contig_name.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.bwaref_sam_header.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.bwaref_bwt_sa.restype = C.c_uint64
        L.bwaref_bwt_sa.argtypes = [C.c_void_p, C.c_uint64]
        L.bwaref_bwt_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.bwaref_ksw_extend2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        L.bwaref_collect_intv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.bwaref_hook_count.restype = C.c_int64
        L.bwaref_hook_count.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.bwaref_hook_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bwaref_batch_new.restype = C.c_void_p
        L.bwaref_batch_new.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.bwaref_batch_free.argtypes = [C.c_void_p]
        L.bwaref_batch_nt4.argtypes = [C.c_void_p, C.c_void_p]
        for nm in ("bwaref_batch_intervals", "bwaref_batch_regions", "bwaref_batch_regions2sam"):
            f = getattr(L, nm)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.bwaref_batch_chains.restype = C.c_int64
        L.bwaref_batch_chains.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.bwaref_batch_seq2chains.argtypes = [C.c_void_p]
        L.bwaref_batch_chainspipe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bwaref_batch_chains2regions.argtypes = [C.c_void_p]
        L.bwaref_batch_dedup_pestat.argtypes = [C.c_void_p, C.c_void_p]
        L.bwaref_batch_matesw.argtypes = [C.c_void_p]
        L.bwaref_batch_matesw.restype = C.c_int64
        L.bwaref_swhook_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bwaref_swhook_count.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.bwaref_swhook_count.restype = C.c_int64
        L.bwaref_batch_set_regions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _ref = L
    return _ref


def orc_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORC_SO):
            ensure_built()
        L = C.CDLL(ORC_SO)
        L.orc_blocks.restype = C.c_uint64
        L.orc_cells.restype = C.c_uint64
        L.orc_bwt_sa.restype = C.c_uint64
        L.orc_bwt_sa.argtypes = [C.POINTER(IndexView), C.c_uint64]
        L.orc_bwt_extend.argtypes = [C.POINTER(IndexView), C.c_void_p, C.c_int, C.c_void_p]
        L.orc_collect_intv.argtypes = [C.POINTER(IndexView), C.POINTER(OrcOpt), C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_collect_intv_leftmost.argtypes = L.orc_collect_intv.argtypes
        L.orc_ksw_extend2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        L.orc_chains2aln.argtypes = [C.POINTER(IndexView), C.POINTER(OrcOpt), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _orc = L
    return _orc


def _as_array(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build_index(fasta, prefix):
    """`bwa index` through the reference library (bwa/bwtindex.c:256)."""
    r = ref_lib().bwaref_index_build(fasta.encode(), prefix.encode())
    if r != 0:
        raise RuntimeError("bwa_idx_build failed")


class Ref:
    """An opened reference index + default mem_opt_t (bwa/bwamem.c:48-84), -M on as in bwa-flow."""

    def __init__(self, prefix, paired=True):
        self.L = ref_lib()
        self.h = self.L.bwaref_open(prefix.encode(), 1 if paired else 0)
        if not self.h:
            raise RuntimeError("cannot load index " + prefix)
        self.view = IndexView()
        self.L.bwaref_index_view(self.h, C.byref(self.view))
        n = self.view.n_seqs
        self.ctg_off = np.zeros(n, dtype=np.int64)
        self.ctg_len = np.zeros(n, dtype=np.int32)
        self.ctg_alt = np.zeros(n, dtype=np.int32)
        self.L.bwaref_contigs(self.h, _p(self.ctg_off), _p(self.ctg_len), _p(self.ctg_alt))

    def close(self):
        if self.h:
            self.L.bwaref_close(self.h)
            self.h = None

    def opt(self, key, value=None):
        return self.L.bwaref_opt(self.h, key.encode(), 0 if value is None else 1, 0 if value is None else int(value))

    def orc_opt(self):
        o = OrcOpt()
        for k, _ in OrcOpt._fields_[:14]:
            setattr(o, k, self.opt(k))
        o.split_factor = self.L.bwaref_opt_split_factor(self.h)
        mat = np.zeros(25, dtype=np.int8)
        self.L.bwaref_opt_mat(self.h, _p(mat))
        for i in range(25):
            o.mat[i] = int(mat[i])
        return o

    def mat(self):
        mat = np.zeros(25, dtype=np.int8)
        self.L.bwaref_opt_mat(self.h, _p(mat))
        return mat

    # numpy views of the index arrays (zero-copy; owned by the library)
    def bwt_array(self):
        return np.frombuffer((C.c_uint32 * self.view.bwt_size).from_address(self.view.bwt), dtype=np.uint32)

    def sa_array(self):
        return np.frombuffer((C.c_uint64 * self.view.n_sa).from_address(self.view.sa), dtype=np.uint64)

    def pac_array(self):
        n = self.view.l_pac // 4 + 1
        return np.frombuffer((C.c_uint8 * n).from_address(self.view.pac), dtype=np.uint8)

    def contig_names(self):
        out = []
        buf = C.create_string_buffer(4096)
        for i in range(self.view.n_seqs):
            self.L.bwaref_contig_name(self.h, i, buf, 4096)
            out.append(buf.value.decode())
        return out

    def sam_header(self):
        n = self.L.bwaref_sam_header(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.L.bwaref_sam_header(self.h, buf, n + 1)
        return buf.value

    def bwt_sa(self, k):
        return int(self.L.bwaref_bwt_sa(self.h, int(k)))

    def bwt_extend(self, ik, is_back):
        a = np.array(list(ik) + [0], dtype=np.uint64)[:4].copy()
        out = np.zeros(16, dtype=np.uint64)
        self.L.bwaref_bwt_extend(self.h, _p(a), int(is_back), _p(out))
        return out.reshape(4, 4)[:, :3].copy()

    def collect_intv(self, nt4):
        nt4 = np.ascontiguousarray(nt4, dtype=np.uint8)
        out = np.zeros(4096, dtype=INTV)
        n = self.L.bwaref_collect_intv(self.h, _p(nt4), len(nt4), _p(out), len(out))
        assert n <= len(out)
        return out[:n].copy()

    def hook_begin(self):
        self.L.bwaref_hook_begin()

    def hook_end(self):
        """-> (jobs[JOB], query byte pool, target byte pool) observed since hook_begin on this thread."""
        self.L.bwaref_hook_end()
        qb, tb = C.c_int64(), C.c_int64()
        n = self.L.bwaref_hook_count(C.byref(qb), C.byref(tb))
        jobs = np.zeros(n, dtype=JOB)
        q = np.zeros(max(qb.value, 1), dtype=np.uint8)
        t = np.zeros(max(tb.value, 1), dtype=np.uint8)
        self.L.bwaref_hook_fetch(_p(jobs), _p(q), _p(t))
        return jobs, q[:qb.value], t[:tb.value]

    def swhook_begin(self):
        self.L.bwaref_swhook_begin()

    def swhook_end(self):
        """-> (jobs[SWJOB] with their kswr_t, query pool, target pool): the ksw_align2 calls of mem_matesw since swhook_begin."""
        self.L.bwaref_swhook_end()
        qb, tb = C.c_int64(), C.c_int64()
        self.L.bwaref_swhook_count.restype = C.c_int64
        n = self.L.bwaref_swhook_count(C.byref(qb), C.byref(tb))
        jobs = np.zeros(n, dtype=SWJOB)
        q = np.zeros(max(qb.value, 1), dtype=np.uint8)
        t = np.zeros(max(tb.value, 1), dtype=np.uint8)
        self.L.bwaref_swhook_fetch(_p(jobs), _p(q), _p(t))
        return jobs, q[:qb.value], t[:tb.value]

    def batch(self, reads, start_idx=0):
        return RefBatch(self, reads, start_idx)


SWJOB = np.dtype([("qlen", "<i4"), ("tlen", "<i4"), ("xtra", "<i4"), ("pad", "<i4"), ("q_off", "<i8"), ("t_off", "<i8"),
                  ("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4"), ("pad2", "<i4")])
PES = np.dtype([("low", "<i4"), ("high", "<i4"), ("failed", "<i4"), ("pad", "<i4"), ("avg", "<f8"), ("std", "<f8")])


def ref_fastq_dump(path1, path2=None, chunk_size=10_000_000):
    """the reference's own parser and batching (kseq_read + bseq_read, bwa/bwa.c:53-85) on FASTA/FASTQ files -> dump text"""
    L = ref_lib()
    L.bwaref_fastq_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.bwaref_fastq_dump.restype = C.c_int64
    L.bwaref_free.argtypes = [C.c_void_p]
    out = C.c_void_p()
    n = L.bwaref_fastq_dump(path1.encode(), path2.encode() if path2 else None, int(chunk_size), C.byref(out))
    if n < 0:
        raise RuntimeError("cannot open " + path1)
    s = C.string_at(out.value, n)
    L.bwaref_free(out)
    return s


def ref_ksw_align2(query, target, mat, o_del, e_del, o_ins, e_ins, xtra):
    """the reference's ksw_align2 (bwa/ksw.c:343-365) -> [score, te, qe, score2, te2, tb, qb]"""
    q = np.array(query, dtype=np.uint8)
    t = np.array(target, dtype=np.uint8)
    m = np.ascontiguousarray(mat, dtype=np.int8)
    out = np.zeros(7, dtype=np.int32)
    L = ref_lib()
    L.bwaref_ksw_align2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
    L.bwaref_ksw_align2(len(q), _p(q), len(t), _p(t), _p(m), o_del, e_del, o_ins, e_ins, xtra, _p(out))
    return out


FLAT_ALN = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("flag", "<i4"), ("is_rev", "<i4"), ("is_alt", "<i4"), ("mapq", "<i4"), ("NM", "<i4"),
                     ("n_cigar", "<i4"), ("score", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("pad", "<i4")])


def ref_reg2aln(ref, nt4_read, reg):
    """the reference's mem_reg2aln on one region of one read -> (FLAT_ALN record, cigar array, MD bytes)"""
    L = ref.L
    L.bwaref_reg2aln.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    q = np.ascontiguousarray(nt4_read, dtype=np.uint8)
    r = np.ascontiguousarray(reg).reshape(1)
    out = np.zeros(1, dtype=FLAT_ALN)
    cig = np.zeros(2048, dtype=np.uint32)
    md = np.zeros(8192, dtype=np.uint8)
    rc = L.bwaref_reg2aln(ref.h, len(q), _p(q), _p(r), _p(out), _p(cig), len(cig), _p(md), len(md))
    assert rc == 0
    return out[0], cig[:out[0]["n_cigar"]].copy(), bytes(md[:int(np.argmax(md == 0))])


def ref_ksw_extend2(query, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    m = np.ascontiguousarray(mat, dtype=np.int8)
    out = np.zeros(6, dtype=np.int32)
    ref_lib().bwaref_ksw_extend2(len(q), _p(q), len(t), _p(t), _p(m), o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, _p(out))
    return out


class RefBatch:
    """One kflow record's worth of reads moving through the reference stages (oracle/ref_driver.cpp)."""

    def __init__(self, ref: Ref, reads, start_idx=0):
        self.ref = ref
        self.L = ref.L
        self.n = reads.n
        blob, noff = reads.name_blob()
        bases = np.ascontiguousarray(reads.bases)
        quals = np.ascontiguousarray(reads.quals)
        off = np.ascontiguousarray(reads.offsets, dtype=np.int64)
        self.offsets = off
        self.h = self.L.bwaref_batch_new(ref.h, reads.n, blob, _p(noff), _p(bases), _p(quals), _p(off), int(start_idx))

    def free(self):
        if self.h:
            self.L.bwaref_batch_free(self.h)
            self.h = None

    def nt4(self):
        out = np.zeros(int(self.offsets[-1]), dtype=np.uint8)
        self.L.bwaref_batch_nt4(self.h, _p(out))
        return out

    def _two(self, fn, dtype):
        a, b = C.c_void_p(), C.c_void_p()
        n = fn(self.h, C.byref(a), C.byref(b))
        return _as_array(a.value, n, dtype), _as_array(b.value, self.n + 1, np.dtype("<i8"))

    def intervals(self):
        return self._two(self.L.bwaref_batch_intervals, INTV)

    def seq2chains(self):
        self.L.bwaref_batch_seq2chains(self.h)

    def chainspipe(self, intv, off):
        intv = np.ascontiguousarray(intv)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self.L.bwaref_batch_chainspipe(self.h, _p(intv), _p(off))

    def chains(self):
        a, s, o = C.c_void_p(), C.c_void_p(), C.c_void_p()
        ns = C.c_int64()
        n = self.L.bwaref_batch_chains(self.h, C.byref(a), C.byref(s), C.byref(o), C.byref(ns))
        return (_as_array(a.value, n, CHAIN), _as_array(s.value, ns.value, SEED), _as_array(o.value, self.n + 1, np.dtype("<i8")))

    def chains2regions(self):
        self.L.bwaref_batch_chains2regions(self.h)

    def set_regions(self, regs, off):
        regs = np.ascontiguousarray(regs)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self.L.bwaref_batch_set_regions(self.h, _p(regs), _p(off))

    def regions(self):
        return self._two(self.L.bwaref_batch_regions, ALNREG)

    def dedup_pestat(self):
        """stage 5, first part: mem_sort_dedup_patch + ALT marking of every read, then mem_pestat over the batch -> PES[4]"""
        out = np.zeros(20, dtype=np.float64)
        self.L.bwaref_batch_dedup_pestat(self.h, _p(out))
        pes = np.zeros(4, dtype=PES)
        for d in range(4):
            pes[d] = (int(out[d * 5]), int(out[d * 5 + 1]), int(out[d * 5 + 2]), 0, out[d * 5 + 3], out[d * 5 + 4])
        return pes

    def matesw(self):
        """stage 5, second part: the rescue block of mem_sam_pe for every pair -> sum of mem_matesw's return values"""
        self.L.bwaref_batch_matesw.restype = C.c_int64
        return self.L.bwaref_batch_matesw(self.h)

    def regions2sam(self):
        a, b = C.c_void_p(), C.c_void_p()
        n = self.L.bwaref_batch_regions2sam(self.h, C.byref(a), C.byref(b))
        sam = C.string_at(a.value, n) if n else b""
        return sam, _as_array(b.value, self.n + 1, np.dtype("<i8"))


# ---- our C restatement -------------------------------------------------------------------------

class Oracle:
    """oracle/oracle.c bound to index arrays that the CALLER keeps alive (numpy arrays or a Ref)."""

    def __init__(self, view: IndexView, opt: OrcOpt, ctg_off, ctg_len, keepalive=None):
        self.L = orc_lib()
        self.view = view
        self.opt = opt
        self.ctg_off = np.ascontiguousarray(ctg_off, dtype=np.int64)
        self.ctg_len = np.ascontiguousarray(ctg_len, dtype=np.int32)
        self._keep = keepalive

    @staticmethod
    def from_ref(ref: Ref):
        return Oracle(ref.view, ref.orc_opt(), ref.ctg_off, ref.ctg_len, keepalive=ref)

    def reset(self):
        self.L.orc_counters_reset()

    def blocks(self):
        return int(self.L.orc_blocks())

    def cells(self):
        return int(self.L.orc_cells())

    def bwt_sa(self, k):
        return int(self.L.orc_bwt_sa(C.byref(self.view), int(k)))

    def bwt_extend(self, ik, is_back):
        a = np.array(list(ik) + [0], dtype=np.uint64)[:4].copy()
        out = np.zeros(16, dtype=np.uint64)
        self.L.orc_bwt_extend(C.byref(self.view), _p(a), int(is_back), _p(out))
        return out.reshape(4, 4)[:, :3].copy()

    def collect_intv(self, nt4):
        nt4 = np.ascontiguousarray(nt4, dtype=np.uint8)
        out = np.zeros(4096, dtype=INTV)
        n = self.L.orc_collect_intv(C.byref(self.view), C.byref(self.opt), _p(nt4), len(nt4), _p(out), len(out))
        assert n <= len(out)
        return out[:n].copy()

    def collect_intv_leftmost(self, nt4):
        """passes 1-2 stated as "leftmost begin per forward match" (the formulation of smem_sweep.cu)"""
        nt4 = np.ascontiguousarray(nt4, dtype=np.uint8)
        out = np.zeros(4096, dtype=INTV)
        n = self.L.orc_collect_intv_leftmost(C.byref(self.view), C.byref(self.opt), _p(nt4), len(nt4), _p(out), len(out))
        assert n <= len(out)
        return out[:n].copy()

    def chains2aln(self, nt4, chains, seeds):
        """chains: CHAIN array of ONE read with seed_off relative to `seeds`."""
        nt4 = np.ascontiguousarray(nt4, dtype=np.uint8)
        chains = np.ascontiguousarray(chains)
        seeds = np.ascontiguousarray(seeds)
        cap = max(16, len(seeds) + 1)
        out = np.zeros(cap, dtype=ALNREG)
        n = self.L.orc_chains2aln(C.byref(self.view), C.byref(self.opt), _p(self.ctg_off), _p(self.ctg_len), len(nt4), _p(nt4),
                                  len(chains), _p(chains), _p(seeds), _p(out), cap)
        assert n >= 0
        return out[:n].copy()


def orc_ksw_extend2(query, target, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    m = np.ascontiguousarray(mat, dtype=np.int8)
    out = np.zeros(6, dtype=np.int32)
    orc_lib().orc_ksw_extend2(len(q), _p(q), len(t), _p(t), _p(m), o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, _p(out))
    return out


def write_index_files(prefix, primary, L2, seq_len, bwt, sa, sa_intv, pac, l_pac, names, ctg_len, codes=None):
    """Write bwa's on-disk index (.bwt .sa .pac .ann .amb; formats of bwt_dump_bwt/bwt_dump_sa bwa/bwt.c:385-407,
    bns_dump bwa/bntseq.c:66-97, .pac trailer bwa/bntseq.c:309-318) from in-memory arrays, so that the compiled reference
    can bwa_idx_load an index it did not have to rebuild.  `codes` (0..4 per base) adds the N holes to .amb/.ann."""
    L2 = [int(x) for x in L2]
    with open(prefix + ".bwt", "wb") as f:
        np.array([primary] + L2[1:5], dtype="<u8").tofile(f)
        np.ascontiguousarray(bwt, dtype="<u4").tofile(f)
    with open(prefix + ".sa", "wb") as f:
        np.array([primary] + L2[1:5] + [sa_intv, seq_len], dtype="<u8").tofile(f)
        np.ascontiguousarray(sa, dtype="<u8")[1:].tofile(f)
    with open(prefix + ".pac", "wb") as f:
        n = l_pac // 4 + (1 if l_pac % 4 else 0)
        np.ascontiguousarray(pac, dtype=np.uint8)[:n].tofile(f)
        if l_pac % 4 == 0:
            f.write(b"\0")
        f.write(bytes([l_pac % 4]))
    offs = np.concatenate([[0], np.cumsum(ctg_len)])
    holes = []
    if codes is not None:
        isn = np.concatenate([[0], (np.asarray(codes) > 3).astype(np.int8), [0]])
        d = np.diff(isn)
        starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
        for s, e in zip(starts, ends):          # a run never spans contigs in bwa (lasts resets per sequence)
            cuts = [s] + [int(o) for o in offs[1:-1] if s < o < e] + [e]
            for a, b in zip(cuts[:-1], cuts[1:]):
                holes.append((int(a), int(b - a)))
    with open(prefix + ".ann", "w") as f:
        f.write("%d %d %u\n" % (l_pac, len(names), 11))
        for i, nm in enumerate(names):
            f.write("0 %s\n" % nm)
            na = sum(1 for h in holes if offs[i] <= h[0] < offs[i + 1])
            f.write("%d %d %d\n" % (offs[i], ctg_len[i], na))
    with open(prefix + ".amb", "w") as f:
        f.write("%d %d %u\n" % (l_pac, len(names), len(holes)))
        for o, ln in holes:
            f.write("%d %d N\n" % (o, ln))
