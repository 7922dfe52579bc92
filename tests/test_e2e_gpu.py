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
