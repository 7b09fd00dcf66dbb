import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from anomaly_clustering_b200 import ops, pipeline, synth
layers = [(768, 28, 28, True), (768, 28, 28, True)]
feats, _ = synth.planted_features_device(range(100), layers, device="cuda")
a = torch.softmax(torch.randn(100, 784, device="cuda"), dim=1)
for name, fn in (("from_features", lambda: ops.weighted_embed_from_features(feats, a, 3, 1, 2048, 4096)),
                 ("embed operands only", lambda: ops.embed(feats, 3, 1, 2048, 4096, want_z=False, operand="f16")),
                 ("embed Z+hi", lambda: ops.embed(feats, 3, 1, 2048, 4096, want_z=True, operand="f16"))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-22s host enqueue %.3f ms/call   incl. GPU drain %.3f ms/call" % (name, (t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))
