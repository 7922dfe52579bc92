import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden_c1():
    return load_golden("c1_150bp")


@pytest.fixture(scope="session")
def golden_c5():
    return load_golden("c5_250bp_indel")


class GoldenIndex:
    """Rebuilds the FM-index of a golden case with the reference's own `bwa index` (oracle/_ref) in a temp dir."""

    def __init__(self, g, tmpdir):
        from bwaflow_b200 import synth
        from oracle import api
        codes = g["ref_codes"]
        lens = g["ctg_len"]
        bounds = np.concatenate([[0], np.cumsum(lens)])
        ref = synth.Reference(["chr%d" % (i + 1) for i in range(len(lens))],
                              [codes[bounds[i]:bounds[i + 1]] for i in range(len(lens))])
        fa = os.path.join(tmpdir, "ref.fa")
        ref.write_fasta(fa)
        api.build_index(fa, os.path.join(tmpdir, "ref"))
        self.ref = api.Ref(os.path.join(tmpdir, "ref"), paired=True)
        self.reads = synth.ReadBatch(
            [bytes(g["names"][g["name_off"][i]:g["name_off"][i + 1]]) for i in range(len(g["name_off"]) - 1)],
            g["bases"], g["quals"], g["offsets"])


@pytest.fixture(scope="session")
def index_c1(golden_c1, tmp_path_factory):
    return GoldenIndex(golden_c1, str(tmp_path_factory.mktemp("c1")))


@pytest.fixture(scope="session")
def index_c5(golden_c5, tmp_path_factory):
    return GoldenIndex(golden_c5, str(tmp_path_factory.mktemp("c5")))
