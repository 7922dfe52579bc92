"""Developer probe: time the CUDA index builder at a given reference size."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bwaflow_b200 import capi  # noqa: E402

n_bp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
n_ctg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
os.system("free -g | head -2; nproc; df -h /tmp /dev/shm | tail -2")
t0 = time.time()
rng = np.random.default_rng(3100)
codes = rng.integers(0, 4, size=n_bp, dtype=np.uint8)
t1 = time.time()
print("reference %.2f Gbp generated in %.1f s" % (n_bp / 1e9, t1 - t0), flush=True)
bounds = np.linspace(0, n_bp, n_ctg + 1).astype(np.int64)
lens = np.diff(bounds)
B = capi.BuiltIndex(codes, lens, device=0)
t2 = time.time()
v = B.view
print("index built in %.1f s: primary %d bwt_size %d n_sa %d L2 %s" % (t2 - t1, v.primary, v.bwt_size, v.n_sa, list(v.L2)), flush=True)
ctx = capi.Context(0, B.index())
t3 = time.time()
print("context (upload + occ32 re-layout) in %.1f s" % (t3 - t2), flush=True)
