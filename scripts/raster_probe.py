"""One launch of the min-distance kernel per raster setting (run under ncu to read DRAM bytes / clocks)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import _lib, ops, pipeline  # noqa: E402

lib = _lib.load()
n, P, D = 100, 784, 4096
torch.manual_seed(0)
Z = torch.randn(n, P, D, device="cuda") * 0.6
ps = pipeline.patchset_from_Z(Z, "f16")
del Z
out = torch.empty(n, n * P, dtype=torch.float32, device="cuda")
for gm in (4, 8, 16, 32):
    lib.ac_debug_set(1, gm)
    ops.min_dist_sym(ps.hi, None, ps.n2, 0, ps.hi, None, ps.n2, n, P, "f16")
    torch.cuda.synchronize()
for gm in (8, 16, 32):
    lib.ac_debug_set(1, gm)
    ops.min_dist(ps.hi, None, ps.n2, ps.hi, None, ps.n2, n, P, "f16", out=out)
    torch.cuda.synchronize()
