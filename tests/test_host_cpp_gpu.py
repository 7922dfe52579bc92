"""GPU: the C++ host side -- the src/cuda stage classes (bwaflow_b200/host/CUDAPipeline.{h,cpp}, the counterparts of
src/fpga/FPGAPipeline.{h,cpp}) wired as the compute_flow slice SeqsToChains -> ChainsPipe -> ChainsToRegions over a boost-free
kflow stand-in, several kflow records in flight on two workers per stage.  Regions must equal the golden vectors of the
compiled reference, read for read, whatever order the records complete in."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case,per_record,workers,mode,gpus", [("c1", 400, 2, "fast", 1), ("c1", 97, 3, "fast", 1), ("c5", 2000, 1, "fast", 1),
                                                               ("c1", 250, 2, "intervals", 1), ("c5", 300, 2, "intervals", 1),
                                                               ("c1", 97, 4, "fast", 0), ("c5", 150, 4, "intervals", 0)])
def test_cpp_stage_classes_match_golden(case, per_record, workers, mode, gpus, request, tmp_path):
    """fast = chains stay in HBM (tag 1); intervals = the FPGA flow: SMEM intervals to the host, ChainsPipe chains them, stage 4
    takes mem_chain_v lists (tag 0).  gpus = 0: CUDAEnv takes every visible GPU (index replicated), worker wid drives GPU
    wid % n_gpus and the kflow queue deals the records -- on a one-GPU box that is the single-GPU path again."""
    from bwaflow_b200 import capi
    g = request.getfixturevalue("golden_" + case)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "bwaflow_b200", "host")], check=True)
    B = capi.BuiltIndex(g["ref_codes"], g["ctg_len"])
    v = B.view
    off = np.ascontiguousarray(g["offsets"], dtype=np.int64)
    bases = np.ascontiguousarray(g["bases"], dtype=np.uint8)            # ASCII, as bseq1_t::seq arrives at stage 2
    n_reads = len(off) - 1
    path = str(tmp_path / "case.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<qqQ5QQQQqqqq", v.l_pac, v.n_seqs, v.primary, *list(v.L2), v.seq_len, v.bwt_size, v.n_sa, v.sa_intv,
                            n_reads, len(bases), per_record))
        B.bwt_array().tofile(f)
        B.sa_array().tofile(f)
        B.pac_array().tofile(f)
        np.ascontiguousarray(B.ctg_off, dtype=np.int64).tofile(f)
        np.ascontiguousarray(B.ctg_len, dtype=np.int32).tofile(f)
        off.tofile(f)
        bases.tofile(f)
    out = str(tmp_path / "regions.bin")
    r = subprocess.run([os.path.join(ROOT, "bwaflow_b200", "cuda_stage_demo"), path, out, str(workers), mode, str(gpus)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    n = struct.unpack_from("<q", raw, 0)[0]
    assert n == n_reads
    roff = np.frombuffer(raw, dtype="<i8", count=n + 1, offset=8)
    regs = raw[8 + 8 * (n + 1):]
    assert (roff == g["reg_off"]).all()
    assert regs == g["regs"].tobytes()
    B.close()
