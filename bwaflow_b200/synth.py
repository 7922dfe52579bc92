"""Seeded synthetic references and read pairs for the BASELINE.json configs.

Shapes follow the reference's inputs: a FASTA for `bwa index` (bwa/bwtindex.c:256) and
interleaved 2xL FASTQ-like records (name / bases / qualities) as `getKseqBatch`
(src/Pipeline.cpp:98-130) would hand them to stage 2.  Everything is numpy; nothing here
touches the GPU.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


class Reference:
    """A list of contigs, each a uint8 array of codes 0..3 (4 = N)."""

    def __init__(self, names, contigs, whole=None):
        self.names = list(names)
        self.contigs = list(contigs)
        self.lengths = np.array([len(c) for c in contigs], dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)])[:-1]
        self.total = int(self.lengths.sum())
        self._whole = whole            # the contigs are views of this array when it is given

    def concat(self):
        return self._whole if self._whole is not None else np.concatenate(self.contigs)

    def write_fasta(self, path, width=80):
        with open(path, "wb") as f:
            for name, c in zip(self.names, self.contigs):
                f.write(b">" + name.encode() + b"\n")
                asc = np.frombuffer(b"ACGTN", dtype=np.uint8)[c]
                n = len(asc)
                full = (n // width) * width
                if full:
                    body = np.empty((n // width, width + 1), dtype=np.uint8)
                    body[:, :width] = asc[:full].reshape(-1, width)
                    body[:, width] = 10
                    f.write(body.tobytes())
                if n > full:
                    f.write(asc[full:].tobytes() + b"\n")


def make_reference(n_bp, n_contigs=1, seed=5, n_run_frac=0.0, repeat_copies=0, repeat_len=400, repeat_div=0.02):
    """i.i.d. uniform ACGT reference; optional N runs and a diverged repeat family so that
    max_occ / frac_rep / pass-2 re-seeding paths fire."""
    rng = np.random.default_rng(seed)
    seq = rng.integers(0, 4, size=n_bp, dtype=np.uint8)
    if repeat_copies > 0:
        unit = rng.integers(0, 4, size=repeat_len, dtype=np.uint8)
        for _ in range(repeat_copies):
            pos = int(rng.integers(0, n_bp - repeat_len))
            cp = unit.copy()
            mut = rng.random(repeat_len) < repeat_div
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                cp = (3 - cp)[::-1]
            seq[pos:pos + repeat_len] = cp
    if n_run_frac > 0:
        n_runs = max(1, int(n_bp * n_run_frac / 50))
        for _ in range(n_runs):
            pos = int(rng.integers(0, n_bp - 100))
            seq[pos:pos + int(rng.integers(1, 100))] = 4
    bounds = np.linspace(0, n_bp, n_contigs + 1).astype(np.int64)
    contigs = [seq[bounds[i]:bounds[i + 1]] for i in range(n_contigs)]
    names = ["chr%d" % (i + 1) for i in range(n_contigs)]
    return Reference(names, contigs, whole=seq)


class ReadBatch:
    """Interleaved pairs: read 2i and 2i+1 are mates.  `bases`/`quals` are ASCII, concatenated;
    `offsets` has n+1 entries; `names` is a list of bytes (mates share a name, as bwa requires)."""

    def __init__(self, names, bases, quals, offsets):
        self.names = names
        self.bases = bases
        self.quals = quals
        self.offsets = offsets
        self.n = len(offsets) - 1

    def name_blob(self):
        blob = b"".join(self.names)
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum([len(x) for x in self.names], out=off[1:])
        return blob, off

    def slice(self, lo, hi):
        o = self.offsets[lo:hi + 1]
        return ReadBatch(self.names[lo:hi], self.bases[o[0]:o[-1]], self.quals[o[0]:o[-1]], o - o[0])

    def nt4(self):
        """0..3 for ACGT (either case), 4 otherwise: nst_nt4_table (bwa/bntseq.c:47-64)."""
        t = np.full(256, 4, dtype=np.uint8)
        for i, ch in enumerate(b"ACGT"):
            t[ch] = i
            t[ch + 32] = i
        return t[self.bases]


def simulate_pairs(ref: Reference, n_pairs, read_len=150, sub_rate=0.01, seed=1, ins_mean=400.0, ins_sd=50.0,
                   n_rate=0.0, first_id=0):
    """Vectorised FR pairs with substitutions only (configs 1/3/4)."""
    rng = np.random.default_rng(seed)
    L = read_len
    frag = np.clip(np.rint(rng.normal(ins_mean, ins_sd, n_pairs)), 2 * L if ins_mean >= 2 * L else L, 1000).astype(np.int64)
    ctg = rng.choice(len(ref.contigs), size=n_pairs, p=ref.lengths / ref.lengths.sum())
    clen = ref.lengths[ctg]
    frag = np.minimum(frag, clen)
    start = (rng.random(n_pairs) * (clen - frag + 1)).astype(np.int64) + ref.offsets[ctg]
    whole = ref.concat()
    idx = np.arange(L, dtype=np.int64)
    r1 = whole[start[:, None] + idx[None, :]]                       # forward strand, left end
    r2 = _COMP[whole[(start + frag - 1)[:, None] - idx[None, :]]]   # reverse strand, right end
    flip = rng.random(n_pairs) < 0.5                                # strand of the pair
    a = np.where(flip[:, None], r2, r1)
    b = np.where(flip[:, None], r1, r2)
    reads = np.empty((2 * n_pairs, L), dtype=np.uint8)
    reads[0::2] = a
    reads[1::2] = b
    mut = (rng.random(reads.shape) < sub_rate) & (reads < 4)
    reads[mut] = (reads[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    if n_rate > 0:
        reads[rng.random(reads.shape) < n_rate] = 4
    bases = np.frombuffer(b"ACGTN", dtype=np.uint8)[reads].reshape(-1)
    quals = np.full(bases.shape, ord("I"), dtype=np.uint8)
    offsets = np.arange(2 * n_pairs + 1, dtype=np.int64) * L
    names = []
    for i in range(n_pairs):
        nm = b"r%d" % (first_id + i)
        names.append(nm)
        names.append(nm)
    return ReadBatch(names, bases, quals, offsets)


def _mutate_indel(rng, seq, sub, ins, dele, max_indel=40, mean_indel=3.0):
    out = []
    i = 0
    n = len(seq)
    p_geo = 1.0 / mean_indel
    while i < n:
        r = rng.random()
        if r < sub:
            out.append((int(seq[i]) + int(rng.integers(1, 4))) & 3 if seq[i] < 4 else 4)
            i += 1
        elif r < sub + ins:
            k = min(int(rng.geometric(p_geo)), max_indel)
            out.extend(int(v) for v in rng.integers(0, 4, size=k))
        elif r < sub + ins + dele:
            i += min(int(rng.geometric(p_geo)), max_indel)
        else:
            out.append(int(seq[i]))
            i += 1
    return out


def simulate_pairs_indel(ref: Reference, n_pairs, read_len=250, err=0.05, seed=5, chimeric_frac=0.10,
                         long_indel_frac=0.05, first_id=0):
    """Config 5: indel-heavy errors (40% subst / 30% ins / 30% del, geometric lengths), plus chimeric
    reads and single long indels so that wide bands and the z-drop exit are provoked."""
    rng = np.random.default_rng(seed)
    L = read_len
    whole = ref.concat()
    N = len(whole)
    reads = []
    names = []
    for p in range(n_pairs):
        frag = int(np.clip(round(rng.normal(600, 80)), L, 1500))
        s = int(rng.integers(0, N - frag - 400))
        flip = rng.random() < 0.5
        mates = []
        for end in range(2):
            span = L + 200
            if end == 0:
                src = whole[s:s + span]
            else:
                src = _COMP[whole[s + frag - span:s + frag][::-1]]
            src = src.copy()
            if rng.random() < long_indel_frac:
                k = int(rng.integers(50, 151))
                at = int(rng.integers(40, L - 40))
                if rng.random() < 0.5:
                    src = np.concatenate([src[:at], src[at + k:], rng.integers(0, 4, size=k, dtype=np.uint8)])
                else:
                    src = np.concatenate([src[:at], rng.integers(0, 4, size=k, dtype=np.uint8), src[at:]])
            rd = _mutate_indel(rng, src, err * 0.4, err * 0.3, err * 0.3)[:L]
            if rng.random() < chimeric_frac:
                o = int(rng.integers(0, N - L))
                other = whole[o:o + L]
                if rng.random() < 0.5:
                    other = _COMP[other[::-1]]
                rd = rd[:L // 2] + [int(v) for v in other[:L - L // 2]]
            while len(rd) < L:
                rd.append(int(rng.integers(0, 4)))
            mates.append(np.array(rd, dtype=np.uint8))
        if flip:
            mates = mates[::-1]
        reads.extend(mates)
        nm = b"r%d" % (first_id + p)
        names.extend([nm, nm])
    arr = np.stack(reads)
    bases = np.frombuffer(b"ACGTN", dtype=np.uint8)[arr].reshape(-1)
    quals = np.full(bases.shape, ord("I"), dtype=np.uint8)
    offsets = np.arange(2 * n_pairs + 1, dtype=np.int64) * L
    return ReadBatch(names, bases, quals, offsets)


def batch_bounds(offsets, chunk_bp=10_000_000):
    """Read-index boundaries of the reference's batching rule (src/Pipeline.cpp:98-130,146):
    keep reading pairs until >= chunk_bp bases and an even read count."""
    n = len(offsets) - 1
    cuts = [0]
    size = 0
    for i in range(0, n, 2):
        size += int(offsets[min(i + 2, n)] - offsets[i])
        if size >= chunk_bp:
            cuts.append(min(i + 2, n))
            size = 0
    if cuts[-1] != n:
        cuts.append(n)
    return cuts


def extension_jobs(ref: Reference, reads: ReadBatch, a=1, w=100, end_bonus=5, zdrop=100, min_seed=19, seed=7):
    """Extension jobs of the shape mem_chain2aln issues (bwa/bwamem.c:724-771), generated without running the aligner:
    for every read a seed = a stretch of at least `min_seed` bases around a random anchor, then the LEFT job (reversed
    query prefix vs reversed reference window, h0 = seed length) and the RIGHT job (query suffix vs reference window,
    h0 = seed length + left flank length, i.e. roughly the left score), each window as long as the flank plus the largest
    gap cal_max_gap allows (bwa/bwamem.c:630-637, 651-666).  Reads are taken as placed by simulate_pairs (their true
    reference position is re-derived by exact search of the seed, so only reads whose seed is unique and error-free are
    used).  Returns (jobs[EXT_JOB-like dict of arrays], qpool, tpool) with byte-per-base pools (codes 0..4 / 0..3)."""
    rng = np.random.default_rng(seed)
    whole = ref.concat()
    nt4 = reads.nt4()
    L = int(reads.offsets[1] - reads.offsets[0])
    n = reads.n
    R = nt4.reshape(n, L)
    # anchor k-mer (32 bases) looked up in a hash of the reference's forward and reverse strands would be the general
    # way; the simulator's layout makes it simpler: reads are substrings of `whole` or of its reverse complement
    # with substitutions, so locate each read by its first exact 24-mer via a sorted k-mer table of the reference
    K = 24
    def kmers(arr):
        v = np.zeros(len(arr) - K + 1, dtype=np.uint64)
        for i in range(K):
            v = (v << np.uint64(2)) | arr[i:len(arr) - K + 1 + i].astype(np.uint64)
        return v
    fw = kmers(np.where(whole < 4, whole, 0))
    order = np.argsort(fw, kind="stable")
    fws = fw[order]
    jobs_q, jobs_t, meta = [], [], []
    sel = rng.permutation(n)
    comp = np.array([3, 2, 1, 0, 4], dtype=np.uint8)
    for strand in (0, 1):
        RR = R if strand == 0 else comp[R[:, ::-1]]
        for off in (0, L // 2, L - K):
            km = np.zeros(n, dtype=np.uint64)
            ok = np.ones(n, dtype=bool)
            for i in range(K):
                col = RR[:, off + i]
                ok &= col < 4
                km = (km << np.uint64(2)) | (col & 3).astype(np.uint64)
            pos = np.searchsorted(fws, km)
            pos = np.minimum(pos, len(fws) - 1)
            hit = ok & (fws[pos] == km)
            nxt = np.minimum(pos + 1, len(fws) - 1)
            uniq = hit & ((fws[nxt] != km) | (nxt == pos))
            meta.append((strand, off, uniq, order[pos] - off))
    placed = np.zeros(n, dtype=bool)
    rpos = np.zeros(n, dtype=np.int64)
    rstrand = np.zeros(n, dtype=np.int8)
    for strand, off, uniq, p in meta:
        take = uniq & ~placed & (p >= 0) & (p + L <= len(whole))
        rpos[take] = p[take]
        rstrand[take] = strand
        placed |= take
    idx = np.nonzero(placed)[0]
    # oriented read and its reference window start
    Q = np.where(rstrand[idx, None] == 0, R[idx], comp[R[idx][:, ::-1]])
    P = rpos[idx]
    T = whole[P[:, None] + np.arange(L)[None, :]]
    match = (Q == T) & (Q < 4)
    # seed: the longest exact run containing a random anchor column (>= min_seed), found by run-length labelling
    m = len(idx)
    anchor = rng.integers(0, L, size=m)
    run_id = np.cumsum(~match, axis=1)
    rid_a = run_id[np.arange(m), anchor]
    in_run = (run_id == rid_a[:, None]) & match
    qb = np.argmax(in_run, axis=1)
    ln = in_run.sum(axis=1)
    good = (ln >= min_seed) & match[np.arange(m), anchor]
    Q, P, qb, ln = Q[good], P[good], qb[good], ln[good]
    qe = qb + ln

    def max_gap(q):
        l = np.maximum((q * a - 6) // 1 + 1, 1)
        return np.minimum(l, w << 1)
    out = {k: [] for k in ("qlen", "tlen", "w", "h0", "end_bonus", "zdrop", "q_off", "t_off")}
    qparts, tparts = [], []
    qo = to = 0
    N = len(whole)
    for side in (0, 1):
        if side == 0:
            ql = qb
            keep = ql > 0
        else:
            ql = L - qe
            keep = ql > 0
        for i in np.nonzero(keep)[0]:
            qn = int(ql[i])
            g = int(max_gap(np.int64(qn)))
            if side == 0:
                t0 = max(0, int(P[i] + qb[i]) - (qn + g))
                t = whole[t0:int(P[i] + qb[i])][::-1]
                q = Q[i, :qn][::-1]
                h0 = int(ln[i]) * a
            else:
                t1 = min(N, int(P[i] + qe[i]) + qn + g)
                t = whole[int(P[i] + qe[i]):t1]
                q = Q[i, int(qe[i]):]
                h0 = (int(ln[i]) + int(qb[i])) * a
            t = np.where(t < 4, t, 0).astype(np.uint8)
            out["qlen"].append(qn); out["tlen"].append(len(t)); out["w"].append(w); out["h0"].append(max(h0, 1))
            out["end_bonus"].append(end_bonus); out["zdrop"].append(zdrop); out["q_off"].append(qo); out["t_off"].append(to)
            qparts.append(q.astype(np.uint8)); tparts.append(t)
            qo += qn; to += len(t)
    jobs = {k: np.array(v, dtype=np.int64 if k.endswith("_off") else np.int32) for k, v in out.items()}
    qpool = np.concatenate(qparts) if qparts else np.zeros(0, np.uint8)
    tpool = np.concatenate(tparts) if tparts else np.zeros(0, np.uint8)
    return jobs, qpool, tpool
