"""CPU: the ingest stage (bwaflow_b200/host/Ingest.cpp: FastqSource = KseqsRead + KseqsToBseqs, src/Pipeline.cpp:98-202) against the
reference's own parser and batching (kseq_read + bseq_read, bwa/bwa.c:53-85, compiled in oracle/_ref) on the same files: 4-line
and multi-line FASTQ, FASTA, comments, /1 /2 suffixes, CRLF line ends, blank lines, lower case and IUPAC codes, gzip, a missing
final newline, paired files of unequal length, records cut at every chunk size."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bwaflow_b200", "ingest_selftest")


def _ours(chunk, *paths):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "bwaflow_b200", "host"), "../ingest_selftest"], check=True)
    r = subprocess.run([BIN, str(chunk)] + list(paths), capture_output=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def _ref(chunk, *paths):
    from oracle import api
    if not api.have_ref():
        pytest.skip("oracle/_ref not built")
    return api.ref_fastq_dump(paths[0], paths[1] if len(paths) > 1 else None, chunk)


def _rand_reads(rng, n, lo=30, hi=260):
    out = []
    for i in range(n):
        l = int(rng.integers(lo, hi))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTACGTACGTNacgtnRYK", dtype=np.uint8), l))
        qual = bytes(rng.integers(33, 74, l).astype(np.uint8))
        out.append((b"read%d" % i, seq, qual))
    return out


def _fastq(reads, suffix=b"", comment=lambda i: b"", wrap=0, crlf=False, blank_every=0, final_newline=True):
    nl = b"\r\n" if crlf else b"\n"
    parts = []
    for i, (name, seq, qual) in enumerate(reads):
        c = comment(i)
        hdr = b"@" + name + suffix + ((b" " if i % 2 else b"\t") + c if c else b"")
        if wrap:
            sl = nl.join(seq[k:k + wrap] for k in range(0, len(seq), wrap))
            ql = nl.join(qual[k:k + wrap] for k in range(0, len(qual), wrap))
        else:
            sl, ql = seq, qual
        parts.append(hdr + nl + sl + nl + b"+" + (name if i % 3 == 0 else b"") + nl + ql + nl)
        if blank_every and i % blank_every == 0:
            parts.append(nl)
    s = b"".join(parts)
    return s if final_newline else s.rstrip(b"\r\n")


@pytest.mark.parametrize("chunk", [1, 500, 4000, 10_000_000])
def test_four_line_fastq_pairs_in_two_gz_files(tmp_path, chunk):
    rng = np.random.default_rng(3)
    r1, r2 = _rand_reads(rng, 180), _rand_reads(rng, 180)
    p1, p2 = str(tmp_path / "a_1.fq.gz"), str(tmp_path / "a_2.fq.gz")
    with gzip.open(p1, "wb") as f:
        f.write(_fastq(r1, b"/1", lambda i: b"BC:Z:ACGT-%d extra words" % i if i % 4 else b""))
    with gzip.open(p2, "wb") as f:
        f.write(_fastq(r2, b"/2"))
    a, b = _ours(chunk, p1, p2), _ref(chunk, p1, p2)
    assert a == b
    assert a.count(b"#record") == (180 if chunk == 1 else a.count(b"#record"))


@pytest.mark.parametrize("chunk", [300, 10_000_000])
def test_multiline_crlf_blank_lines_and_no_final_newline(tmp_path, chunk):
    rng = np.random.default_rng(5)
    reads = _rand_reads(rng, 120)
    p = str(tmp_path / "il.fq")
    open(p, "wb").write(_fastq(reads, wrap=37, crlf=True, blank_every=7, final_newline=False,
                               comment=lambda i: b"1:N:0:ATCACG" if i % 2 == 0 else b""))
    assert _ours(chunk, p) == _ref(chunk, p)
    q = str(tmp_path / "plain.fq")
    open(q, "wb").write(_fastq(reads, wrap=0, blank_every=5))
    assert _ours(chunk, q) == _ref(chunk, q)


def test_fasta_input_and_names_ending_in_slash_digit(tmp_path):
    rng = np.random.default_rng(7)
    reads = _rand_reads(rng, 60)
    p = str(tmp_path / "r.fa")
    with open(p, "wb") as f:
        for i, (name, seq, _) in enumerate(reads):
            f.write(b">" + name + (b"/7" if i % 3 == 0 else b"/x" if i % 3 == 1 else b"") + b" desc %d\n" % i)
            for k in range(0, len(seq), 60):
                f.write(seq[k:k + 60] + b"\n")
    assert _ours(700, p) == _ref(700, p)


def test_second_file_shorter(tmp_path):
    """bwa-flow's getKseqBatch (src/Pipeline.cpp:112-120) has already taken the read of file 1 when it finds file 2 exhausted, and
    KseqsRead keeps calling it until a batch comes back empty (:137-155): every left-over read of file 1 arrives as a record of its
    own (which the paired stage 5 then ignores: batch_num / 2 pairs).  Stock bseq_read tests file 2 first and stops at the last whole
    pair.  FastqSource follows bwa-flow."""
    rng = np.random.default_rng(9)
    r1, r2 = _rand_reads(rng, 40), _rand_reads(rng, 33)
    p1, p2 = str(tmp_path / "x_1.fq"), str(tmp_path / "x_2.fq")
    open(p1, "wb").write(_fastq(r1))
    open(p2, "wb").write(_fastq(r2))
    ours, ref = _ours(1000, p1, p2).split(b"\n"), _ref(1000, p1, p2).split(b"\n")
    reads_o = [l for l in ours if l and not l.startswith(b"#record")]
    reads_r = [l for l in ref if l and not l.startswith(b"#record")]
    assert len(reads_r) == 66 and reads_o[:66] == reads_r
    assert reads_o[66:] == [name + b"\t\t" + seq + b"\t" + qual for name, seq, qual in r1[33:]]
    recs = [l.split() for l in ours if l.startswith(b"#record")]
    assert [int(x[2]) for x in recs[-6:]] == [1] * 6 and int(recs[-7][2]) % 2 == 1 and int(recs[-1][1]) == 72
