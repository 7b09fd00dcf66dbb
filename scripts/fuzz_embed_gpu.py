"""python scripts/fuzz_embed_gpu.py [n_cases] [seed] : ac_embed (every kernel variant the planner picks) against the
oracle on seeded random geometries -- the GPU counterpart of tests/test_oracle_golden.py::
test_oracle_fuzz_against_live_reference.  Prints one line per case; exit code 1 on any mismatch.
To be run first thing next round and promoted into tests/ once green (written after round 1's GPU budget was spent)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from anomaly_clustering_b200 import ops  # noqa: E402
from oracle import restated  # noqa: E402  (checker only)

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2023
rng = np.random.default_rng(seed)
gen = torch.Generator().manual_seed(seed)
bad = 0
for case in range(n_cases):
    k = int(rng.choice([1, 3, 3, 3, 5]))
    s = int(rng.choice([1, 1, 1, 2]))
    B = int(rng.integers(1, 4))
    kind = case % 4
    if kind == 0:       # ViT tokens, all layers on one grid; channel counts that hit and miss the periodic patterns
        g = int(rng.integers(4, 15))
        C = int(rng.choice([8, 24, 48, 96, 192, 384]))
        feats = [torch.randn(B, 1 + g * g, C, generator=gen) for _ in range(int(rng.integers(1, 4)))]
        Dp = int(rng.choice([C * 9 // 4, C * 9 // 8, C * 9 * 8 // 27, int(rng.integers(5, 300))])) if k == 3 else int(rng.integers(5, 300))
    elif kind == 1:     # CNN pyramid in NCHW, grid halves per layer (resampled layers)
        g = int(rng.choice([8, 12, 16, 28]))
        L = int(rng.integers(1, 4))
        feats = [torch.randn(B, int(rng.integers(6, 70)), max(g >> l, k), max(g >> l, k), generator=gen) for l in range(L)]
        Dp = int(rng.integers(5, 300))
    elif kind == 2:     # channels_last CNN maps
        g = int(rng.choice([7, 14, 28]))
        feats = [torch.randn(B, int(rng.choice([16, 32, 64, 128])), g, g, generator=gen).contiguous(memory_format=torch.channels_last)
                 for _ in range(int(rng.integers(1, 3)))]
        Dp = int(rng.choice([feats[0].shape[1] * 9 // 2, feats[0].shape[1] * 9, int(rng.integers(5, 300))])) if k == 3 else int(rng.integers(5, 300))
    else:               # non-square maps, odd sizes
        H, W = int(rng.integers(5, 20)), int(rng.integers(5, 20))
        feats = [torch.randn(B, int(rng.integers(3, 40)), max(H, k), max(W, k), generator=gen)]
        Dp = int(rng.integers(5, 300))
    Dp = max(Dp, 1)
    D = int(rng.choice([Dp, 2 * Dp, max(1, Dp // 2), int(rng.integers(5, 400))]))
    operand = [None, "f16", "bf16"][int(rng.integers(0, 3))]
    if operand is not None and D % 8:
        D += 8 - D % 8
    zo = restated.embed(feats, k, s, Dp, D)
    tag = "case %2d kind %d k=%d s=%d B=%d %s Dp=%d D=%d op=%s" % (case, kind, k, s, B, [tuple(f.shape) for f in feats], Dp, D, operand)
    try:
        Z, hi, lo, _ = ops.embed([f.cuda() for f in feats], k, s, Dp, D, operand=operand, want_lo=operand is not None)
        torch.cuda.synchronize()
    except Exception as e:   # noqa: BLE001
        print(tag, "RAISED", repr(e)[:200], flush=True)
        bad += 1
        continue
    err = (Z.cpu() - zo).abs().max().item()
    ok = err <= 2e-5
    msg = "max|dZ| %.1e" % err
    if hi is not None:
        rec = hi.float().cpu() + lo.float().cpu()
        e2 = (rec - zo).abs().max().item()
        ok &= e2 <= (2e-5 if operand == "f16" else 2e-4) * max(1.0, zo.abs().max().item())
        msg += "  max|hi+lo-Z| %.1e" % e2
    bad += not ok
    print(tag, msg, "OK" if ok else "MISMATCH", flush=True)
print("%d / %d cases failed" % (bad, n_cases))
sys.exit(1 if bad else 0)
