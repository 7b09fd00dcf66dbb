"""Live timing of stage 1 alone on config-2 geometry (clocks not depressed by the GEMM).
python scripts/tune_embed.py [n_img]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import _lib, ops, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lib = _lib.load()
feats, _ = synth.planted_features_device(range(n), [(768, 28, 28, True), (768, 28, 28, True)], device="cuda")
P, D = 784, 4096
Z = torch.empty(n * P, D, dtype=torch.float32, device="cuda")
hi = torch.empty(n * P, D, dtype=torch.float16, device="cuda")
in_bytes = sum(f[:, 1:].numel() * 4 for f in feats)


def run(reps, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.embed(feats, 3, 1, 2048, 4096, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for variant in (0, 1, 2):
    lib.ac_debug_set(2, variant)
    for name, kw, out_bytes in (
        ("Z+hi", dict(want_z=True, operand="f16", out_z=Z, out_hi=hi), n * P * D * 6),
        ("Z only", dict(want_z=True, out_z=Z), n * P * D * 4),
        ("hi only", dict(want_z=False, operand="f16", out_hi=hi), n * P * D * 2),
        ("Z+hi no-LN", dict(want_z=True, operand="f16", out_z=Z, out_hi=hi, layernorm=False), n * P * D * 6),
    ):
        run(3, **kw)
        ms = run(20, **kw)
        gb = (in_bytes + out_bytes) / 1e9
        print("variant %d %-11s %.3f ms  %.1f us/img  %.0f GB/s (algorithmic)" % (variant, name, ms, 1e3 * ms / n, gb / (ms * 1e-3)), flush=True)
lib.ac_debug_set(2, 0)
