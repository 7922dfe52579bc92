"""GPU, end to end: reads -> (stage 2 on GPU) -> host chaining -> (stage 4 on GPU) -> regions, with the index built by the
product's own builder; then the reference's RegionsToSam (oracle/_ref) turns OUR regions into SAM and the bytes must
equal the golden SAM of the all-CPU reference path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_hot_path_regions_and_sam_bytes(case, request):
    from bwaflow_b200 import capi, stages
    g = request.getfixturevalue("golden_" + case)
    ix = request.getfixturevalue("index_" + case)          # only used for RegionsToSam at the end
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    reads = ix.reads
    nt4 = reads.nt4()
    hp = stages.HotPath(ctx, B.view, B.ctg_off, B.ctg_len, reads.n, len(nt4))
    out = hp.run(stages.seqs_record(0, nt4, reads.offsets))
    regs, off = out["alnreg"], out["alnreg_off"]
    assert (off == g["reg_off"]).all()
    assert regs.tobytes() == g["regs"].tobytes()
    # byte-identical SAM: our regions through the reference's own stage 5
    rb = ix.ref.batch(reads, 0)
    exp = np.zeros(len(regs), dtype=g["regs"].dtype)
    for k in exp.dtype.names:
        exp[k] = regs[k]
    rb.set_regions(exp, off)
    sam, soff = rb.regions2sam()
    assert sam == g["sam"].tobytes()
    rb.free()
    hp.close()
    ctx.close()
    B.close()


def test_hot_path_in_two_kflow_batches_equals_one(golden_c1, index_c1):
    """Splitting the reads into separate records (as getKseqBatch would, src/Pipeline.cpp:98-130) does not change
    per-read regions; start_idx bookkeeping restores order."""
    from bwaflow_b200 import capi, stages, dealer
    g, ix = golden_c1, index_c1
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    reads = ix.reads
    nt4 = reads.nt4()
    hp = stages.HotPath(ctx, B.view, B.ctg_off, B.ctg_len, reads.n, len(nt4))
    cuts = [0, 400, 1000, reads.n]
    ro = dealer.Reorder()
    got = []
    for i in (2, 0, 1):                                   # completion order differs from input order
        lo, hi = cuts[i], cuts[i + 1]
        off = reads.offsets[lo:hi + 1]
        rec = stages.seqs_record(lo, nt4[off[0]:off[-1]], off - off[0])
        out = hp.run(rec)["alnreg"].copy()              # views of the batch's pinned buffer die at its next submit
        for _, p in ro.push(lo, hi - lo, out):
            got.append(p)
    assert ro.drained() and len(got) == 3
    regs = np.concatenate([p for p in got])
    assert regs.tobytes() == g["regs"].tobytes()
    hp.close()
    ctx.close()
    B.close()


@pytest.mark.parametrize("case", ["c1", "c5"])
def test_fused_hot_path_equals_golden(case, request):
    """Device-only path (stage 2 -> device chaining -> stage 4): regions byte-identical to the golden vectors, and
    identical to the three-piece path with host chaining."""
    from bwaflow_b200 import capi, stages
    g = request.getfixturevalue("golden_" + case)
    ix = request.getfixturevalue("index_" + case)
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    ctx = capi.Context(0, B.index(), keepalive=B)
    reads = ix.reads
    nt4 = reads.nt4()
    fp = stages.FusedHotPath(ctx, reads.n, len(nt4))
    out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
    assert (out["alnreg_off"] == g["reg_off"]).all()
    assert out["alnreg"].tobytes() == g["regs"].tobytes()
    t = fp.last
    assert t.n_chains == len(g["chains"]) and t.n_seeds == len(g["seeds"]) and t.n_intv == len(g["intv"])
    regs2, off2, _ = fp.rerun()
    assert regs2.tobytes() == g["regs"].tobytes()
    # smaller second batch through the same object (buffers are reused)
    n = 100
    off = reads.offsets[:n + 1]
    out = fp.run(stages.seqs_record(0, nt4[:off[-1]], off))
    assert out["alnreg"].tobytes() == g["regs"][:g["reg_off"][n]].tobytes()
    fp.close()
    ctx.close()
    B.close()


def test_fused_hot_path_repeat_rich(tmp_path):
    """Many chains per read (multi-node B-tree, weight ties) on the device chaining path vs the compiled reference."""
    from bwaflow_b200 import capi, stages, synth
    from oracle import api
    ref = synth.make_reference(60000, n_contigs=2, seed=9, repeat_copies=120, repeat_len=300, repeat_div=0.03)
    fa = str(tmp_path / "r.fa")
    ref.write_fasta(fa)
    api.build_index(fa, str(tmp_path / "r"))
    R = api.Ref(str(tmp_path / "r"))
    reads = synth.simulate_pairs(ref, 300, 150, 0.02, seed=2)
    rb = R.batch(reads)
    rb.seq2chains()
    rb.chains2regions()
    regs, roff = rb.regions()
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    fp = stages.FusedHotPath(ctx, reads.n, len(nt4))
    out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
    assert (out["alnreg_off"] == roff).all()
    assert out["alnreg"].tobytes() == regs.tobytes()
    fp.close()
    ctx.close()
    B.close()
    rb.free()
    R.close()


def _against_compiled_reference(tmp_path, ref, reads, n_threads=8):
    """Fused hot path vs the compiled reference (oracle/_ref: mem_seq2chain_wrapper + mem_chain2aln) on a workload built
    here; returns (timing, reads redone on the exact stage-4 kernel, ksw_extend2 job table of the reference)."""
    from concurrent.futures import ThreadPoolExecutor
    from bwaflow_b200 import capi, stages
    from oracle import api
    fa = str(tmp_path / "r.fa")
    ref.write_fasta(fa)
    api.build_index(fa, str(tmp_path / "r"))
    R = api.Ref(str(tmp_path / "r"))
    cuts = np.linspace(0, reads.n, n_threads + 1).astype(np.int64) & ~1

    def part(i):
        R.hook_begin()
        rb = R.batch(reads.slice(int(cuts[i]), int(cuts[i + 1])), int(cuts[i]))
        rb.seq2chains()
        rb.chains2regions()
        jobs = R.hook_end()[0]
        out = rb.regions()
        rb.free()
        return out, jobs
    with ThreadPoolExecutor(n_threads) as ex:
        parts = list(ex.map(part, range(n_threads)))
    regs = np.concatenate([p[0][0] for p in parts])
    n_per = np.concatenate([np.diff(p[0][1]) for p in parts])
    jobs = np.concatenate([p[1] for p in parts])
    B = capi.BuiltIndex(ref.concat(), ref.lengths)
    ctx = capi.Context(0, B.index(), keepalive=B)
    nt4 = reads.nt4()
    fp = stages.FusedHotPath(ctx, reads.n, len(nt4))
    out = fp.run(stages.seqs_record(0, nt4, reads.offsets))
    assert (np.diff(out["alnreg_off"]) == n_per).all()
    assert out["alnreg"].tobytes() == regs.tobytes()
    t, n_exact = fp.last, fp.batch.exact_reads()
    fp.close()
    ctx.close()
    B.close()
    R.close()
    return t, n_exact, jobs


def test_fused_hot_path_c5_shape_6k_reads(tmp_path):
    """BASELINE config[4] shape at a size where the hard paths are populated: 3 000 indel-heavy 2x250 bp pairs (chimeric reads,
    long indels) plus 400 reads crafted so that a flank extension crosses an 80-95 bp deletion (max_off >= 75: mem_chain2aln's
    second band try with w = 200, bwa/bwamem.c:742,770) -- reads that leave the speculative stage-4 path."""
    from bwaflow_b200 import synth
    ref = synth.make_reference(3_000_000, n_contigs=3, seed=55, repeat_copies=60, repeat_len=600)
    reads = synth.simulate_pairs_indel(ref, 3000, 250, 0.05, seed=5)
    extra = synth.simulate_pairs(ref, 200, 250, 0.0, seed=6, first_id=3000)
    rng = np.random.default_rng(8)
    whole = ref.concat()
    asc = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(extra.n):
        k, a = int(rng.integers(80, 96)), int(rng.integers(105, 125))
        p = int(rng.integers(0, len(whole) // 3 - 600))
        seq = np.concatenate([whole[p:p + a], whole[p + a + k:p + a + k + 250 - a]])
        if i & 1:
            seq = (3 - seq)[::-1]
        extra.bases[extra.offsets[i]:extra.offsets[i + 1]] = asc[seq & 3]
    both = synth.ReadBatch(reads.names + extra.names, np.concatenate([reads.bases, extra.bases]),
                           np.concatenate([reads.quals, extra.quals]),
                           np.concatenate([reads.offsets, extra.offsets[1:] + reads.offsets[-1]]))
    t, n_exact, jobs = _against_compiled_reference(tmp_path, ref, both)
    second_try = int((jobs["w"] > 100).sum())
    print("c5 shape: %d reads, %d chains; %d ksw_extend2 calls, %d second band tries (max_off up to %d); %d reads on the exact kernel"
          % (both.n, t.n_chains, len(jobs), second_try, int(jobs["max_off"].max()), n_exact))
    assert second_try > 100 and n_exact > 100


def test_fused_hot_path_repeat_rich_c1_shape(tmp_path):
    """BASELINE config[0] shape on a reference with a diverged repeat family and N runs (the i.i.d. reference alone
    leaves max_occ sampling, pass-2 re-seeding and the containment skip between chains idle): 20 000 2x150 bp reads."""
    from bwaflow_b200 import synth
    ref = synth.make_reference(5_000_000, n_contigs=4, seed=5, n_run_frac=0.001, repeat_copies=800, repeat_len=500, repeat_div=0.03)
    reads = synth.simulate_pairs(ref, 10000, 150, 0.01, seed=1, n_rate=0.0005)
    t, n_exact, jobs = _against_compiled_reference(tmp_path, ref, reads)
    print("repeat-rich c1 shape: %d reads, %d chains, %d seeds; %d reads on the exact kernel" % (reads.n, t.n_chains, t.n_seeds, n_exact))
    assert t.n_chains > reads.n
