"""CPU: the product's host chaining (bwaflow_chain_build, the role of ChainsPipe::compute src/Pipeline.cpp:408-438)
against the golden chains of the compiled reference, with SA positions supplied by the reference's bwt_sa."""
import numpy as np
import pytest


def _positions(ix, intv, max_occ=500):
    pos, off = [], [0]
    for p in intv:
        x0, x2 = int(p["x0"]), int(p["x2"])
        step = x2 // max_occ if x2 > max_occ else 1
        ks = [x0 + k for k in range(0, x2, step)][:max_occ]
        pos.extend(ix.ref.bwt_sa(k) for k in ks)
        off.append(len(pos))
    return np.array(pos, dtype=np.int64), np.array(off, dtype=np.int64)


def _cmp(chains, seeds, off, g):
    assert (off == g["chain_off"]).all(), np.nonzero(off != g["chain_off"])[0][:5]
    for k in ("pos", "n_seeds", "rid", "frac_rep", "flags", "seed_off"):
        bad = np.nonzero(chains[k] != g["chains"][k])[0]
        assert len(bad) == 0, (k, bad[:5], chains[bad[:5]], g["chains"][bad[:5]])
    assert seeds.tobytes() == g["seeds"].tobytes()


@pytest.mark.parametrize("case", ["c1", "c5"])
@pytest.mark.parametrize("threads", [1, 4])
def test_chain_build_matches_golden(case, threads, request):
    from bwaflow_b200 import capi
    ix = request.getfixturevalue("index_" + case)
    g = request.getfixturevalue("golden_" + case)
    intv = np.zeros(len(g["intv"]), dtype=capi.INTV)
    for k in capi.INTV.names:
        intv[k] = g["intv"][k]
    pos, poff = _positions(ix, intv)
    chains, seeds, off = capi.chain_build(ix.ref.view.l_pac, ix.ref.ctg_off, ix.ref.ctg_len, ix.reads.offsets, intv, g["intv_off"],
                                          pos, poff, n_threads=threads)
    _cmp(chains, seeds, off, g)


def test_chain_build_repeat_rich_many_chains(tmp_path):
    """A reference that is mostly one diverged repeat family: reads get tens of chains (multi-level B-tree, equal
    chain positions, weight ties in the introsort) -- the cases where klib's exact behaviour shows."""
    from bwaflow_b200 import capi, synth
    from oracle import api
    ref = synth.make_reference(60000, n_contigs=2, seed=9, repeat_copies=120, repeat_len=300, repeat_div=0.03)
    fa = str(tmp_path / "r.fa")
    ref.write_fasta(fa)
    api.build_index(fa, str(tmp_path / "r"))
    R = api.Ref(str(tmp_path / "r"))
    reads = synth.simulate_pairs(ref, 300, 150, 0.02, seed=2)
    # add reads made of two copies of the same 60-mer 120 bp apart (same reference position from two query offsets)
    b = R.batch(reads)
    intv, ioff = b.intervals()
    b.seq2chains()
    ch, sd, coff = b.chains()
    assert np.diff(coff).max() > 9, "want reads with more than one B-tree node"

    class IX:
        ref = R
    pos, poff = _positions(IX, intv)
    pi = np.zeros(len(intv), dtype=capi.INTV)
    for k in capi.INTV.names:
        pi[k] = intv[k]
    chains, seeds, off = capi.chain_build(R.view.l_pac, R.ctg_off, R.ctg_len, reads.offsets, pi, ioff, pos, poff, n_threads=2)
    _cmp(chains, seeds, off, dict(chains=ch, seeds=sd, chain_off=coff))
    b.free()
    R.close()
