"""Epilogue-bound microbenchmark of the min-distance kernel: D = 64 (one K block per tile), so the
time per tile is the epilogue's.  python scripts/tune_epi.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import _lib, ops, pipeline  # noqa: E402

lib = _lib.load()
n, P = 100, 784
for D in (64, 256):
    torch.manual_seed(0)
    Z = torch.randn(n, P, D, device="cuda")
    ps = pipeline.patchset_from_Z(Z, "f16")
    out = torch.empty(n, n * P, dtype=torch.float32, device="cuda")

    def t(fn, reps=10):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    full = t(lambda: ops.min_dist(ps.hi, None, ps.n2, ps.hi, None, ps.n2, n, P, "f16", out=out))
    sym = t(lambda: ops.min_dist_sym(ps.hi, None, ps.n2, 0, ps.hi, None, ps.n2, n, P, "f16"))
    tiles_full = (n * P / 256) * n * 4 / 74
    print("D=%d  all-pairs %.3f ms (%.2f us per tile per CTA pair)   sym %.3f ms (%.2f us per tile per CTA pair)"
          % (D, full, 1e3 * full / tiles_full, sym, 1e3 * sym / (tiles_full / 2)), flush=True)
