"""GPU: the stand-alone driver bwaflow_mem (bwaflow_b200/host/mem_main.cpp) -- FASTQ files and a bwa index on disk in, SAM out, with
every compute slot of compute_flow on the GPU (what `--use_gpu` switches on beside `--use_fpga`).  The SAM must equal, byte for byte,
the text of the reference's stages 2-5 (oracle/_ref) run record by record with the same start_idx, for gzipped paired files cut
into several records that complete out of order on several workers, for an interleaved file with comments and a read group, and
for single-end input."""
import gzip
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bwaflow_b200", "bwaflow_mem")


def _setup(n_pairs, read_len, seed, paired=True):
    from bwaflow_b200 import synth
    from oracle import api
    ref = synth.make_reference(2_000_000, n_contigs=3, seed=seed, repeat_copies=25, repeat_len=500, n_run_frac=0.001)
    reads = synth.simulate_pairs_indel(ref, n_pairs, read_len, 0.04, seed=seed + 1)
    td = tempfile.mkdtemp()
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))      # the reference's own `bwa index`: files on disk
    return td, ref, reads, api.Ref(os.path.join(td, "r"), paired=paired)


def _fastq(reads, idx, suffix=b"", comment=None):
    out = []
    for i in idx:
        o0, o1 = int(reads.offsets[i]), int(reads.offsets[i + 1])
        hdr = b"@" + reads.names[i] + suffix + (b" " + comment if comment else b"")
        out.append(hdr + b"\n" + bytes(reads.bases[o0:o1]) + b"\n+\n" + bytes(reads.quals[o0:o1]) + b"\n")
    return b"".join(out)


def _records(reads, chunk):
    """getKseqBatch's rule (src/Pipeline.cpp:98-130): whole pairs until the record holds >= chunk bases"""
    recs, lo, size = [], 0, 0
    for i in range(0, reads.n, 2):
        size += int(reads.offsets[i + 2] - reads.offsets[i])
        if size >= chunk:
            recs.append((lo, i + 2))
            lo, size = i + 2, 0
    if lo < reads.n:
        recs.append((lo, reads.n))
    return recs


def _ref_sam(R, reads, recs):
    out = []
    for lo, hi in recs:
        rb = R.batch(reads.slice(lo, hi), lo)
        rb.seq2chains()
        rb.chains2regions()
        sam, _ = rb.regions2sam()
        out.append(bytes(sam))
        rb.free()
    return out


def _body(path):
    return b"".join(l for l in open(path, "rb").read().splitlines(keepends=True) if not l.startswith(b"@"))


def _build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "bwaflow_b200", "host")], check=True)


def test_refuses_to_run_without_the_switch(tmp_path):
    _build()
    r = subprocess.run([BIN, "idx", "a.fq"], capture_output=True, text=True)
    assert r.returncode == 2 and "--use_gpu is required" in r.stderr


def test_paired_gz_files_many_records_in_order():
    _build()
    td, ref, reads, R = _setup(6000, 150, 71)
    with gzip.open(os.path.join(td, "r1.fq.gz"), "wb", compresslevel=1) as f:
        f.write(_fastq(reads, range(0, reads.n, 2), b"/1"))             # /1 and /2 are trimmed (trim_readno)
    with gzip.open(os.path.join(td, "r2.fq.gz"), "wb", compresslevel=1) as f:
        f.write(_fastq(reads, range(1, reads.n, 2), b"/2"))
    chunk = 200_000
    recs = _records(reads, chunk)
    assert len(recs) >= 8
    exp = b"".join(_ref_sam(R, reads, recs))
    out = os.path.join(td, "out.sam")
    r = subprocess.run([BIN, "--use_gpu", "-t", "3", "--inorder_output", "--chunk_size", str(chunk), "-o", out, os.path.join(td, "r"),
                        os.path.join(td, "r1.fq.gz"), os.path.join(td, "r2.fq.gz")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _body(out)
    assert got == exp, (len(got), len(exp))
    hdr = [l for l in open(out, "rb").read().split(b"\n") if l.startswith(b"@")]
    assert hdr[0] == b"@HD\tVN:1.3\tSO:coordinate" and hdr[1].startswith(b"@SQ\tSN:chr1\tLN:") and hdr[-1].startswith(b"@PG\tID:bwa")
    print(r.stderr.strip().splitlines()[-1])
    R.close()


def test_interleaved_with_comments_and_read_group_unordered():
    _build()
    td, ref, reads, R = _setup(3000, 150, 81)
    open(os.path.join(td, "il.fq"), "wb").write(_fastq(reads, range(reads.n), b"", b"BC:Z:ACGT"))
    chunk = 150_000
    recs = _records(reads, chunk)
    # what the reference writes with -C and the read group: the same records with the comment and RG:Z appended -- checked through
    # the tags themselves, and through the unordered output being a permutation of the ordered records' lines
    out = os.path.join(td, "out.sam")
    r = subprocess.run([BIN, "--use_gpu=1", "-t", "4", "-p", "-C", "-R", "@RG\\tID:lane7\\tSM:s", "--chunk_size", str(chunk), "-o", out,
                        os.path.join(td, "r"), os.path.join(td, "il.fq")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _body(out).splitlines()
    exp = b"".join(_ref_sam(R, reads, recs)).splitlines()
    assert len(got) == len(exp)
    strip = lambda l: l.replace(b"\tRG:Z:lane7", b"").replace(b"\tBC:Z:ACGT", b"")
    assert all(b"\tRG:Z:lane7" in l and l.endswith(b"\tBC:Z:ACGT") for l in got)
    assert sorted(strip(l) for l in got) == sorted(exp)
    assert any(l == b"@RG\tID:lane7\tSM:s" for l in open(out, "rb").read().split(b"\n"))
    R.close()


def test_single_end_250bp():
    _build()
    td, ref, reads, R = _setup(1500, 250, 91, paired=False)
    open(os.path.join(td, "se.fq"), "wb").write(_fastq(reads, range(reads.n)))
    recs = [(0, reads.n)]
    exp = b"".join(_ref_sam(R, reads, recs))
    out = os.path.join(td, "out.sam")
    r = subprocess.run([BIN, "--use_gpu", "--inorder_output", "-o", out, os.path.join(td, "r"), os.path.join(td, "se.fq")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert _body(out) == exp
    R.close()
