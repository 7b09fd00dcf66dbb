"""Sweep tuning knobs of the tcgen05 min-distance kernel on config-2 geometry (sustained timing:
many back-to-back launches under the power cap).  python scripts/tune_tc.py [n_img]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import _lib, ops, pipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
P, D = 784, 4096
lib = _lib.load()
torch.manual_seed(0)
Z = torch.randn(n, P, D, device="cuda") * 0.6
ps = pipeline.patchset_from_Z(Z, "f16")
del Z
flops = 2.0 * (n * P) * ((n - 1) * P) * D
out = torch.empty(n, n * P, dtype=torch.float32, device="cuda")


def run(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.min_dist(ps.hi, None, ps.n2, ps.hi, None, ps.n2, n, P, "f16", out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for key, vals in ((1, [10, 12, 14, 16, 18, 20, 24]),):
    for v in vals:
        lib.ac_debug_set(key, v)
        run(2)
        ms = run(12)
        print("knob%d=%d  %.2f ms  %.0f TFLOP/s (algorithmic)" % (key, v, ms, flops / ms / 1e9), flush=True)
for g in (1, 2):
    lib.ac_debug_set(0, g)
    lib.ac_debug_set(1, 16)
    run(2)
    ms = run(12)
    print("cta_group=%d GM=16  %.2f ms  %.0f TFLOP/s" % (g, ms, flops / ms / 1e9), flush=True)


def run_sym(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.min_dist_sym(ps.hi, None, ps.n2, 0, ps.hi, None, ps.n2, n, P, "f16")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


lib.ac_debug_set(0, 2)
for v in (4, 8, 12, 16, 24, 32):
    lib.ac_debug_set(1, v)
    run_sym(2)
    ms = run_sym(12)
    print("SYM GM=%d  %.2f ms  %.0f TFLOP/s executed, %.0f algorithmic" % (v, ms, 0.5 * flops / ms / 1e9, flops / ms / 1e9), flush=True)
lib.ac_debug_set(1, 16)

for dyn in (0, 1):
    lib.ac_debug_set(4, dyn)
    run_sym(2)
    ms = run_sym(12)
    run(2)
    ms2 = run(6)
    print("dynamic=%d  SYM %.2f ms   all-pairs %.2f ms" % (dyn, ms, ms2), flush=True)
lib.ac_debug_set(4, 0)
