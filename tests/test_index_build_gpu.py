"""GPU: the CUDA index builder gives bit-identical arrays to the host builder (itself pinned to `bwa index`, test_index_build.py)
and to the compiled reference's own bwa_idx_build."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [dict(n_bp=50_000, n_contigs=1, seed=1),
         dict(n_bp=200_003, n_contigs=3, seed=2, n_run_frac=0.003, repeat_copies=30),
         dict(n_bp=4_097, n_contigs=2, seed=3),
         dict(n_bp=33, n_contigs=1, seed=4),
         dict(n_bp=3_000_001, n_contigs=5, seed=5, repeat_copies=200, repeat_len=1500, repeat_div=0.0)]


def _same(B, E):
    v, e = B.view, E.view
    for k in ("primary", "seq_len", "bwt_size", "n_sa", "sa_intv", "l_pac", "n_seqs"):
        assert getattr(v, k) == getattr(e, k), k
    assert list(v.L2) == list(e.L2)
    assert (B.bwt_array() == E.bwt_array()).all()
    assert (B.sa_array() == E.sa_array()).all()
    n = v.l_pac // 4 + (1 if v.l_pac % 4 else 0)
    assert (B.pac_array()[:n] == E.pac_array()[:n]).all()


@pytest.mark.parametrize("kw", CASES)
def test_gpu_index_matches_host_builder(kw):
    from bwaflow_b200 import capi, synth
    ref = synth.make_reference(**kw)
    G = capi.BuiltIndex(ref.concat(), ref.lengths, device=0)
    H = capi.BuiltIndex(ref.concat(), ref.lengths, n_threads=0)
    _same(G, H)
    G.close()
    H.close()


def test_gpu_index_matches_bwa_index(tmp_path):
    from bwaflow_b200 import capi, synth
    from oracle import api
    ref = synth.make_reference(n_bp=120_000, n_contigs=2, seed=9, n_run_frac=0.002, repeat_copies=20)
    fa = str(tmp_path / "r.fa")
    ref.write_fasta(fa)
    api.build_index(fa, str(tmp_path / "r"))
    R = api.Ref(str(tmp_path / "r"))
    G = capi.BuiltIndex(ref.concat(), ref.lengths, device=0)
    _same(G, R)
    G.close()
    R.close()


def test_gpu_index_palindromes_and_homopolymers():
    """Suffixes that stay equal to the end of the text (the reverse-complement half mirrors the forward half) and long
    runs of one base: exercises the tie resolution ('$' sorts first) rather than the radix keys."""
    from bwaflow_b200 import capi
    seq = np.concatenate([np.zeros(700, np.uint8), np.full(700, 3, np.uint8), np.tile(np.array([0, 3], np.uint8), 300),
                          np.tile(np.array([1, 2], np.uint8), 300), np.zeros(90, np.uint8)])
    G = capi.BuiltIndex(seq, [len(seq)], device=0)
    H = capi.BuiltIndex(seq, [len(seq)], n_threads=2)
    _same(G, H)
    G.close()
    H.close()
