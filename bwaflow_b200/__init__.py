"""bwaflow_b200: B200 (sm_100a) backend for bwa-flow's BWA-MEM per-read compute path.

The product is `libbwaflow_cuda.so` (csrc/, C ABI in include/bwaflow_cuda.h).  This package is the
thin Python host side used by tests and bench.py: ctypes bindings (`capi`), the stage mirrors
(`stages`) and synthetic workload generators (`synth`).  There is no CPU fallback: importing `capi`
without the built library, or creating a context without a GPU, raises.
"""
__version__ = "0.1.0"
