"""Exact refine kernel (precision f16r) at full config-2 size: time per pass by L2 group budget (ac_debug_set key 5), inside the
min-distance stage (so at the clocks it really runs at, right after the tensor-core pass) -- CUDA events at the stage marks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import _lib, pipeline, synth  # noqa: E402

lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
feats, _ = synth.planted_features_device(range(n), [(768, 28, 28, True), (768, 28, 28, True)], device="cuda")
q = pipeline.embed_images(feats, 3, 1, 2048, 4096, "f16r", want_z=True)
w_ref = None
for mb in (32, 48, 64, 96, 128, 192, 256):
    lib.ac_debug_set(5, mb)
    pipeline.PROFILE = []
    for _ in range(4):
        w = pipeline.min_distance_weights(q, q, "unsupervised", "f16r")
    torch.cuda.synchronize()
    ev = pipeline.PROFILE
    pipeline.PROFILE = None
    b = [e for t, e in ev if t == "refine_begin"]
    e_ = [e for t, e in ev if t == "refine_end"]
    same = True if w_ref is None else bool(torch.equal(w, w_ref))
    w_ref = w if w_ref is None else w_ref
    print("refine group budget %3d MB: %.2f ms per pass   (w bit-identical across budgets: %s)" % (
        mb, sum(x.elapsed_time(y) for x, y in zip(b[1:], e_[1:])) / (len(b) - 1), same), flush=True)
lib.ac_debug_set(5, 128)
