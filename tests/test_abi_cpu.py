"""CPU: the C-ABI library loads and exports every symbol include/bwaflow_cuda.h declares (no compute calls)."""
import ctypes

import numpy as np
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bwaflow_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bwaflow_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bwaflow_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_gpu_is_an_error_not_a_fallback():
    from bwaflow_b200 import capi
    if capi.lib().bwaflow_device_count() > 0:
        return
    try:
        capi.Context(0)
    except capi.BwaflowError as e:
        assert "no CUDA device" in str(e) or "fallback" in str(e)
    else:
        raise AssertionError("context creation must fail without a GPU")


def test_struct_sizes_match_reference_layouts():
    from bwaflow_b200 import capi
    assert capi.INTV.itemsize == 32      # bwtintv_t
    assert capi.SEED.itemsize == 24      # mem_seed_t
    assert capi.ALNREG.itemsize == 88    # mem_alnreg_t
    assert ctypes.sizeof(capi.Opt) == 88


def test_max_read_len_follows_mem_flt_chained_seeds():
    """reads long enough for mem_flt_chained_seeds to re-score seeds (bwa/bwamem.c:607-624: MEM_MINSC_COEF * log(l) <= MEM_SEEDSW_COEF * l)
    are outside the path; bwaflow_max_read_len reports the last length below that (and below the 1 024-base row state)"""
    import math
    from bwaflow_b200 import capi
    L = capi.lib()
    L.bwaflow_max_read_len.argtypes = [ctypes.c_void_p]
    got = L.bwaflow_max_read_len(None)
    exp = max(l for l in range(1, 1024) if np.float32(5.5) * math.log(l) > np.float32(0.05) * l)
    assert got == exp == 724
    co = capi.default_chain_opt()
    co.min_chain_weight = 40            # then min_l = MEM_HSP_COEF * min_chain_weight = 44: re-scoring from 880 bases on
    assert L.bwaflow_max_read_len(ctypes.byref(co)) == 879
