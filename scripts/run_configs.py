"""Times the BASELINE configs that are not bench lines (3 and 5) on one GPU.  python scripts/run_configs.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import pipeline, synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


vitb = [(768, 28, 28, True), (768, 28, 28, True)]
q, _ = synth.planted_features_device(range(100), vitb, device="cuda")
bank, _ = synth.planted_features_device(range(1000, 1200), vitb, n_classes=1, device="cuda")
ms = timed(lambda: pipeline.run_path(q, 3, 1, 2048, 4096, "supervised", [1.0], bank_features=bank))
print("config3 supervised: 100 queries vs 200-image bank: %.2f ms  (%.0f images/s, %.0f TFLOP/s algorithmic)"
      % (ms, 100 / ms * 1e3, 2 * 78400 * 156800 * 4096 / ms / 1e9))
ms = timed(lambda: pipeline.run_path(q, 3, 1, 2048, 4096, "average"))
print("config3 average mode: %.3f ms" % ms)
del bank
vits = [(384, 56, 56, True), (384, 56, 56, True)]
q5, _ = synth.planted_features_device(range(64), vits, device="cuda")
taus = [0.1, 0.5, 1, 2, 5, 10]
ms = timed(lambda: pipeline.run_path(q5, 3, 1, 2048, 4096, "unsupervised", taus, precision="f16"), reps=3)
print("config5 (1 GPU, 64 images x 3136 patches, 6 taus from one pass, f16): %.1f ms  (%.0f images/s)" % (ms, 64 / ms * 1e3))
ms = timed(lambda: pipeline.run_path(q5, 3, 1, 2048, 4096, "unsupervised", taus, precision="f16x3"), reps=2)
print("config5 f16x3: %.1f ms" % ms)
