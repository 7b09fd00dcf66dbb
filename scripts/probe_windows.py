"""One GPU, the launches one rank of an 8-GPU config-2 step issues (13 local query images against windows of the 100-image bank):
time and executed TFLOP/s per window, to see what short launches cost (CUDA events, 20 repetitions each, back to back)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import ops, pipeline, synth  # noqa: E402
from anomaly_clustering_b200.distributed import pair_owned  # noqa: E402

n, P, D = 100, 784, 4096
feats, _ = synth.planted_features_device(range(n), [(768, 28, 28, True), (768, 28, 28, True)], device="cuda")
q = pipeline.embed_images(feats, 3, 1, 2048, D, "f16", want_z=False)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 13
Qhi, Qn2 = q.hi[: nq * P], q.n2[: nq * P]


def run(windows, reps=20):
    out = None
    for k, w in enumerate(windows):
        out = ops.min_dist_sym(Qhi, None, Qn2, 0, q.hi, None, q.n2, n, P, "f16", bank_window=w, init=(k == 0), out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for k, w in enumerate(windows):
            out = ops.min_dist_sym(Qhi, None, Qn2, 0, q.hi, None, q.n2, n, P, "f16", bank_window=w, init=(k == 0), out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def pairs(w):
    return sum(pair_owned(i, (w[0] + k) % n, n) for i in range(nq) for k in range(w[1]))


for name, windows in (("local", [(0, nq)]), ("shards 1+2", [(nq, 26)]), ("shards 3+4", [(nq + 26, 24)]), ("all remote", [(nq, 50)]),
                      ("one launch", [(0, n)]), ("three launches", [(0, nq), (nq, 26), (nq + 26, 24)]),
                      ("two launches", [(0, nq), (nq, 50)])):
    ms = run(windows)
    pr = sum(pairs(w) for w in windows)
    print("%-15s %s: %.3f ms  %4d pairs  %.2f us/pair  %.0f TFLOP/s executed" % (name, windows, ms, pr, ms * 1e3 / max(pr, 1),
                                                                              2.0 * pr * P * P * D / (ms * 1e-3) / 1e12), flush=True)
