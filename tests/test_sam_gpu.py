"""GPU: stage 5 as a whole (RegionsToSam::compute, src/Pipeline.cpp:546-648).  bwaflow_dedup_patch against the reference's
mem_sort_dedup_patch (with mem_patch_reg's merges), and bwaflow_sam_process -- regions of stage 4 in, per-read SAM text out -- against
the text the reference's own stage 5 writes (oracle/_ref: mem_sam_pe_hts / mem_reg2sam_hts / mem_aln2sam_hts): the committed golden
SAM of c1 and c5, a repeat-rich indel-heavy batch (rescues, merges, supplementary and XA records), single-end mode, and the whole
chain FASTQ-shaped reads -> SAM with nothing but this library between them."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _as_alnreg(regs):
    from bwaflow_b200 import capi
    r1 = np.zeros(len(regs), dtype=capi.ALNREG)
    for k in capi.ALNREG.names:
        r1[k] = regs[k]
    return r1


def _ref_stage4(ix, start_idx=0):
    """regions as stage 4 leaves them, after the reference's mem_sort_dedup_patch, and the reference's SAM text per read"""
    rb = ix.ref.batch(ix.reads, start_idx)
    rb.seq2chains()
    rb.chains2regions()
    regs0, off0 = rb.regions()
    regs0, off0 = np.array(regs0), np.array(off0)
    rb.dedup_pestat()
    regs1, off1 = rb.regions()
    regs1, off1 = np.array(regs1), np.array(off1)
    rb.free()
    rb = ix.ref.batch(ix.reads, start_idx)
    rb.seq2chains()
    rb.chains2regions()
    sam, soff = rb.regions2sam()
    sam = bytes(sam)
    soff = np.array(soff)
    rb.free()
    return regs0, off0, regs1, off1, [sam[soff[i]:soff[i + 1]] for i in range(len(soff) - 1)]


def _names(reads):
    return [n for n in reads.names]


def _first_diff(mine, exp):
    for i, (a, b) in enumerate(zip(mine, exp)):
        if a != b:
            return i, a[:600], b[:600]
    return None


def _make_ix(ref, reads, paired=True):
    from oracle import api
    td = tempfile.mkdtemp()
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))

    class IX:
        pass
    ix = IX()
    ix.ref = api.Ref(os.path.join(td, "r"), paired=paired)
    ix.reads = reads
    return ix


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_dedup_patch_equals_reference(case, request):
    from bwaflow_b200 import capi
    g = request.getfixturevalue("golden_" + case)
    ix = request.getfixturevalue("index_" + case)
    regs0, off0, regs1, off1, _ = _ref_stage4(ix)
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = ix.reads.nt4()
    cb = capi.CigarBatch(ctx, ix.reads.n, len(nt4), 4096)
    out, ooff, n_jobs, rounds = cb.dedup_patch(nt4, ix.reads.offsets, _as_alnreg(regs0), off0)
    assert (ooff == off1).all()
    assert out.tobytes() == _as_alnreg(regs1).tobytes()
    print("%s: %d -> %d regions, %d global alignments for mem_patch_reg in %d rounds" % (case, len(regs0), len(out), n_jobs, rounds))
    cb.close()
    ctx.close()
    B.close()


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_sam_equals_golden(case, request):
    """stage-4 regions of the golden file -> SAM text == the golden SAM (written by the reference's stage 5)"""
    from bwaflow_b200 import capi
    g = request.getfixturevalue("golden_" + case)
    ix = request.getfixturevalue("index_" + case)
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = ix.reads.nt4()
    sb = capi.SamBatch(ctx, ix.reads.n, len(nt4), ix.ref.contig_names())
    mine, pes, t = sb.process(nt4, g["quals"], ix.reads.offsets, _names(ix.reads), g["regs"], g["reg_off"], 0)
    gs = g["sam"].tobytes()
    exp = [gs[g["sam_off"][i]:g["sam_off"][i + 1]] for i in range(len(g["sam_off"]) - 1)]
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    assert b"".join(mine) == gs
    print("%s: %d reads, %d SAM bytes identical; %d patch, %d rescue windows (%d used), %d CIGAR jobs; %.1f ms (dedup %.1f rescue %.1f pair %.1f cigar %.1f text %.1f)"
          % (case, ix.reads.n, len(gs), t.n_patch_jobs, t.n_matesw_jobs, t.n_matesw_used, t.n_cigar_jobs, t.total_ms, t.dedup_ms, t.matesw_ms,
             t.pair_ms, t.cigar_ms, t.text_ms))
    hdr = sb.header(None, b"@PG\tID:bwa\tPN:bwa\tVN:0.7.17-r1194-dirty\tCL:bwaref")
    assert hdr == ix.ref.sam_header()
    sb.close()
    ctx.close()
    B.close()


def test_sam_repeat_rich_indel_batch_and_start_idx():
    """rescues, hit merges, supplementary (0x800 -> 0x100 under -M), XA and unmapped records; a non-zero start_idx (the pair id seeds
    the tie-breaking hashes); several worker threads"""
    from bwaflow_b200 import capi, synth
    ref = synth.make_reference(3_000_000, n_contigs=3, seed=41, repeat_copies=40, repeat_len=600)
    reads = synth.simulate_pairs_indel(ref, 5000, 150, 0.05, seed=43)
    ix = _make_ix(ref, reads)
    start = 133336
    regs0, off0, regs1, off1, exp = _ref_stage4(ix, start)
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    so = capi.default_sam_opt()
    so.n_threads = 5
    sb = capi.SamBatch(ctx, reads.n, len(nt4), ix.ref.contig_names(), so=so)
    mine, pes, t = sb.process(nt4, reads.quals, reads.offsets, _names(reads), _as_alnreg(regs0), off0, start)
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    text = b"".join(exp)
    n_xa, n_sa = text.count(b"\tXA:Z:"), text.count(b"\tSA:Z:")
    flags = [int(l.split(b"\t")[1]) for l in text.split(b"\n") if l]
    n_unm = sum(1 for f in flags if f & 4)
    n_sec = sum(1 for f in flags if f & 0x100)
    assert n_xa > 0 and n_sa > 0 and n_unm > 0 and n_sec > 0 and t.n_patch_jobs > 0 and t.n_matesw_used > 0
    print("repeat-rich: %d reads, %d records (%d unmapped, %d secondary, %d with XA, %d with SA), %d patch alignments, %d rescues; %.1f ms"
          % (reads.n, len(flags), n_unm, n_sec, n_xa, n_sa, t.n_patch_jobs, t.n_matesw_used, t.total_ms))
    sb.close()
    ctx.close()
    B.close()
    ix.ref.close()


def test_sam_single_end():
    from bwaflow_b200 import capi, synth
    ref = synth.make_reference(1_000_000, n_contigs=2, seed=51, repeat_copies=20, repeat_len=300)
    reads = synth.simulate_pairs_indel(ref, 1000, 150, 0.04, seed=53)
    ix = _make_ix(ref, reads, paired=False)
    regs0, off0, regs1, off1, exp = _ref_stage4(ix, 7)
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    so = capi.default_sam_opt()
    so.flag = 0x10                                       # -M, no MEM_F_PE
    sb = capi.SamBatch(ctx, reads.n, len(nt4), ix.ref.contig_names(), so=so)
    mine, pes, t = sb.process(nt4, reads.quals, reads.offsets, _names(reads), _as_alnreg(regs0), off0, 7)
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    sb.close()
    ctx.close()
    B.close()
    ix.ref.close()


def test_reads_to_sam_through_this_library_only():
    """reads -> bwaflow_hot_* (stages 2-4) -> bwaflow_sam_process (stage 5) == the reference's stages 2, 3, 4, 5"""
    from bwaflow_b200 import capi, stages, synth
    ref = synth.make_reference(2_000_000, n_contigs=4, seed=61, repeat_copies=25, repeat_len=450, n_run_frac=0.002)
    reads = synth.simulate_pairs(ref, 4000, 150, 0.01, seed=63)
    ix = _make_ix(ref, reads)
    exp = _ref_stage4(ix, 0)[4]
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    fp = stages.FusedHotPath(ctx, reads.n, len(nt4))
    out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
    sb = capi.SamBatch(ctx, reads.n, len(nt4), ix.ref.contig_names())
    mine, pes, t = sb.process(nt4, reads.quals, reads.offsets, _names(reads), out["alnreg"], out["alnreg_off"], 0)
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    print("reads -> SAM: %d reads, %d bytes identical to the reference's stages 2-5" % (reads.n, sum(len(x) for x in mine)))
    fp.close()
    sb.close()
    ctx.close()
    B.close()
    ix.ref.close()


def test_alt_contig_and_non_default_options():
    """An ALT contig (bns->anns[].is_alt from the .alt file) and -k 17 -w 80: the ALT flag travels through chaining (mem_chain_flt's
    ALT rules), the regions and stage 5 (mem_mark_primary_se's second round, alt_sc / pa:f, XA limits for ALT hits, the extra
    supplementary ALT record of mem_sam_pe); the options reach seeding, chaining, extension and CIGAR generation."""
    from bwaflow_b200 import capi, stages, synth
    from oracle import api
    base = synth.make_reference(1_200_000, n_contigs=2, seed=77)
    rng = np.random.default_rng(78)
    alt = base.contigs[0][200_000:230_000].copy()
    mut = rng.random(len(alt)) < 0.015
    alt[mut] = (alt[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
    ref = synth.Reference(list(base.names) + ["chr1_alt"], list(base.contigs) + [alt])
    region = synth.Reference(["sub"], [base.contigs[0][190_000:240_000]])
    ra = synth.simulate_pairs(region, 1500, 150, 0.01, seed=79)
    rb = synth.simulate_pairs(ref, 500, 150, 0.01, seed=80)
    reads = synth.ReadBatch(list(ra.names) + list(rb.names), np.concatenate([ra.bases, rb.bases]), np.concatenate([ra.quals, rb.quals]),
                            np.concatenate([ra.offsets, rb.offsets[1:] + ra.offsets[-1]]))
    td = tempfile.mkdtemp()
    ref.write_fasta(os.path.join(td, "r.fa"))
    api.build_index(os.path.join(td, "r.fa"), os.path.join(td, "r"))
    open(os.path.join(td, "r.alt"), "w").write("@SQ\tSN:ignored\nchr1_alt\t0\tchr1\t200001\t60\t30000M\n")
    R = api.Ref(os.path.join(td, "r"))
    R.opt("min_seed_len", 17)
    R.opt("w", 80)

    class IX:
        pass
    ix = IX()
    ix.ref, ix.reads = R, reads
    regs0, off0, regs1, off1, exp = _ref_stage4(ix, 0)
    is_alt = np.array([0, 0, 1], dtype=np.int32)
    assert int(((regs1["n_comp_is_alt"] >> 30) != 0).sum()) > 100          # the reference marked ALT regions
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    opt = capi.default_opt()
    opt.min_seed_len = 17
    opt.w = 80
    ctx = capi.Context(0, B.index(), opt, keepalive=B)
    co = capi.default_chain_opt()
    nt4 = reads.nt4()
    fp = stages.FusedHotPath(ctx, reads.n, len(nt4), co, is_alt)
    out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
    assert (out["alnreg_off"] == off0).all()
    assert out["alnreg"].tobytes() == _as_alnreg(regs0).tobytes()
    sb = capi.SamBatch(ctx, reads.n, len(nt4), R.contig_names(), ctg_is_alt=is_alt)
    mine, pes, t = sb.process(nt4, reads.quals, reads.offsets, _names(reads), out["alnreg"], out["alnreg_off"], 0)
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    text = b"".join(exp)
    print("ALT + options: %d reads, %d records, %d with pa:f, %d with XA, %d on chr1_alt" %
          (reads.n, text.count(b"\n"), text.count(b"\tpa:f:"), text.count(b"\tXA:Z:"), sum(1 for l in text.split(b"\n") if l and l.split(b"\t")[2] == b"chr1_alt")))
    assert text.count(b"\tpa:f:") > 0
    fp.close(); sb.close(); ctx.close(); B.close(); R.close()


@pytest.mark.parametrize("flag,T", [(0x2 | 0x8, 30), (0x2 | 0x200 | 0x10, 30), (0x2 | 0x4 | 0x10, 30), (0x2 | 0x20 | 0x10, 30), (0x2, 50), (0x8 | 0x200, 20)])
def test_sam_option_flags(flag, T):
    """mem_opt_t::flag and -T as stage 5 reads them: -a (all secondaries as records, no XA), -Y (soft clips on supplementary
    records), -P (no pairing), -S (no rescue), without -M (supplementary 0x800 instead of 0x100), single-end with -a -Y, and a higher
    output threshold (unmapped records for weak hits)"""
    from bwaflow_b200 import capi, synth
    ref = synth.make_reference(1_500_000, n_contigs=2, seed=91, repeat_copies=30, repeat_len=500)
    reads = synth.simulate_pairs_indel(ref, 1500, 150, 0.05, seed=93)
    ix = _make_ix(ref, reads, paired=bool(flag & 0x2))
    ix.ref.opt("flag", flag)
    ix.ref.opt("T", T)
    regs0, off0, regs1, off1, exp = _ref_stage4(ix, 10)
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    so = capi.default_sam_opt()
    so.flag = flag
    so.T = T
    sb = capi.SamBatch(ctx, reads.n, len(nt4), ix.ref.contig_names(), so=so)
    mine, pes, t = sb.process(nt4, reads.quals, reads.offsets, _names(reads), _as_alnreg(regs0), off0, 10)
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    sb.close()
    ctx.close()
    B.close()
    ix.ref.close()


def test_whole_path_with_scaled_scoring():
    """-A 2 -B 6 -O 8 -E 2 -d 150 -L 7 -U 30 -T 50 through every stage: the biased packed extension cells with a != 1, the word
    (16-bit) ksw_align2 kernel (l_ms * a >= 250), band inference and ksw_global2 with the scaled gaps, mapq with opt->a"""
    from bwaflow_b200 import capi, stages, synth
    ref = synth.make_reference(1_500_000, n_contigs=2, seed=95, repeat_copies=30, repeat_len=500)
    reads = synth.simulate_pairs_indel(ref, 1500, 150, 0.04, seed=97)
    ix = _make_ix(ref, reads)
    vals = dict(a=2, b=6, o_del=8, e_del=2, o_ins=8, e_ins=2, zdrop=150, pen_clip5=7, pen_clip3=7, pen_unpaired=30, T=50)
    for k, v in vals.items():
        ix.ref.opt(k, v)
    regs0, off0, regs1, off1, exp = _ref_stage4(ix, 0)
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    opt = capi.default_opt()
    for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "zdrop", "pen_clip5", "pen_clip3"):
        setattr(opt, k, vals[k])
    mat = [(vals["a"] if i == j else -vals["b"]) if i < 4 and j < 4 else -1 for i in range(5) for j in range(5)]   # bwa_fill_scmat
    for i, v in enumerate(mat):
        opt.mat[i] = v
    ctx = capi.Context(0, B.index(), opt, keepalive=B)
    nt4 = reads.nt4()
    fp = stages.FusedHotPath(ctx, reads.n, len(nt4))
    out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
    assert (out["alnreg_off"] == off0).all()
    assert out["alnreg"].tobytes() == _as_alnreg(regs0).tobytes()
    po = capi.default_pe_opt()
    po.pen_unpaired = vals["pen_unpaired"]
    so = capi.default_sam_opt()
    so.T = vals["T"]
    sb = capi.SamBatch(ctx, reads.n, len(nt4), ix.ref.contig_names(), po=po, so=so)
    mine, pes, t = sb.process(nt4, reads.quals, reads.offsets, _names(reads), out["alnreg"], out["alnreg_off"], 0)
    assert _first_diff(mine, exp) is None, _first_diff(mine, exp)
    print("scaled scoring: %d reads, %d rescue windows used, %d CIGAR jobs, %d reads on the exact stage-4 kernel"
          % (reads.n, t.n_matesw_used, t.n_cigar_jobs, fp.batch.exact_reads()))
    fp.close(); sb.close(); ctx.close(); B.close(); ix.ref.close()
