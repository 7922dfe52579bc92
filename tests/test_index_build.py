"""CPU: the product's index builder gives the arrays `bwa index` + bwa_idx_load give (compiled reference)."""
import numpy as np
import pytest


@pytest.mark.parametrize("kw", [dict(n_bp=50_000, n_contigs=1, seed=1),
                                dict(n_bp=200_003, n_contigs=3, seed=2, n_run_frac=0.003, repeat_copies=30),
                                dict(n_bp=4_097, n_contigs=2, seed=3)])
def test_index_matches_bwa_index(kw, tmp_path):
    from bwaflow_b200 import capi, synth
    from oracle import api
    ref = synth.make_reference(**kw)
    fa = str(tmp_path / "r.fa")
    ref.write_fasta(fa)
    api.build_index(fa, str(tmp_path / "r"))
    R = api.Ref(str(tmp_path / "r"))
    B = capi.BuiltIndex(ref.concat(), ref.lengths, n_threads=4)
    v, e = B.view, R.view
    for k in ("primary", "seq_len", "bwt_size", "n_sa", "sa_intv", "l_pac", "n_seqs"):
        assert getattr(v, k) == getattr(e, k), k
    assert list(v.L2) == list(e.L2)
    assert (B.bwt_array() == R.bwt_array()).all()
    assert (B.sa_array() == R.sa_array()).all()
    n = v.l_pac // 4 + (1 if v.l_pac % 4 else 0)
    assert (B.pac_array()[:n] == R.pac_array()[:n]).all()
    assert (B.ctg_off == R.ctg_off).all() and (B.ctg_len == R.ctg_len).all()
    B.close()
    R.close()
