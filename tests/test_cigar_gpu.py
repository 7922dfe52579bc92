"""GPU: CIGAR generation.  bwaflow_cigar_* against the reference's own mem_reg2aln (oracle/_ref) on every region
of the golden cases after mem_sort_dedup_patch and the mate rescue (so rescued hits, long indels, both strands, soft clips and
leading / trailing deletions are in), and on an indel-heavy 2x250 bp batch whose bands exceed the shared-memory ring."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("pos", "rid", "flag", "is_rev", "is_alt", "mapq", "NM", "n_cigar", "score", "sub", "alt_sc")


def _regions_for_reg2aln(ix):
    rb = ix.ref.batch(ix.reads, 0)
    rb.seq2chains()
    rb.chains2regions()
    rb.dedup_pestat()
    rb.matesw()
    regs, off = rb.regions()
    regs, off = np.array(regs), np.array(off)
    rb.free()
    # mem_mark_primary_se would set `secondary`; make it -1 for every other region so that both the mapq and the 0x100 paths run
    regs["secondary"] = np.where(np.arange(len(regs)) % 2 == 0, -1, 0)
    regs["sub"] = np.where(np.arange(len(regs)) % 3 == 0, 0, regs["score"] // 2)
    return regs, off


def _check(ix, ctx, regs, off, label):
    from bwaflow_b200 import capi
    from oracle import api
    nt4 = ix.reads.nt4()
    roff = ix.reads.offsets
    job_read = np.repeat(np.arange(len(off) - 1, dtype=np.int32), np.diff(off))
    r1 = np.zeros(len(regs), dtype=capi.ALNREG)
    for k in capi.ALNREG.names:
        r1[k] = regs[k]
    cb = capi.CigarBatch(ctx, ix.reads.n, len(nt4), len(regs) + 1)
    cb.submit(nt4, roff, job_read, r1)
    aln, cigars, mds, t = cb.collect()
    n_gap = n_clip = n_rev = 0
    for j in range(len(regs)):
        rd = int(job_read[j])
        exp, ecig, emd = api.ref_reg2aln(ix.ref, nt4[roff[rd]:roff[rd + 1]], regs[j])
        for k in FIELDS:
            assert int(aln[k][j]) == int(exp[k]), (label, j, k, int(aln[k][j]), int(exp[k]), tuple(regs[j]))
        assert (cigars[j] == ecig).all(), (label, j, cigars[j], ecig)
        assert mds[j] == emd, (label, j, mds[j], emd)
        n_gap += int(((ecig & 0xf) == 1).any() or ((ecig & 0xf) == 2).any())
        n_clip += int(((ecig & 0xf) == 3).any())
        n_rev += int(exp["is_rev"])
    print("%s: %d regions, %d with gaps, %d clipped, %d reverse strand; device %.2f ms" % (label, len(regs), n_gap, n_clip, n_rev, t.kernel_ms))
    cb.close()


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_reg2aln_matches_reference(case, request):
    from bwaflow_b200 import capi
    g = request.getfixturevalue("golden_" + case)
    ix = request.getfixturevalue("index_" + case)
    regs, off = _regions_for_reg2aln(ix)
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    _check(ix, ctx, regs, off, case)
    ctx.close()
    B.close()


def test_reg2aln_indel_heavy_wide_bands():
    import os
    import tempfile
    from bwaflow_b200 import capi, synth
    from oracle import api
    ref = synth.make_reference(2_000_000, n_contigs=2, seed=31, repeat_copies=30, repeat_len=500)
    reads = synth.simulate_pairs_indel(ref, 1500, 250, 0.05, seed=33)
    td = tempfile.mkdtemp()
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))

    class IX:
        pass
    ix = IX()
    ix.ref = api.Ref(os.path.join(td, "r"))
    ix.reads = reads
    regs, off = _regions_for_reg2aln(ix)
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    _check(ix, ctx, regs, off, "2x250 indel-heavy")
    ctx.close()
    B.close()
    ix.ref.close()
