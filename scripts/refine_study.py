"""CPU study (no GPU needed) for DESIGN.md section 8 item 4b: does "one fp16 pass that tracks the arg-min column,
then an exact fp32 re-evaluation of the arg-min pairs" meet the alpha tolerance at small tau, where today
precision="auto" pays for the 3-pass split mode?

Simulation of the arithmetic on CPU at config-2 width (D = 4096, P = 784) with fewer images: operands rounded to
fp16 exactly as the embed kernel rounds them, |x|^2 + |y|^2 - 2 x.y with fp32 accumulation (torch CPU sgemm stands in
for the tensor core's fp32 accumulator), exact reference = fp64 cdist.  Usage: python scripts/refine_study.py [n_images]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from anomaly_clustering_b200 import synth  # noqa: E402
from oracle import restated  # noqa: E402  (study script = test infrastructure, not the product)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
layers = [(768, 28, 28, True), (768, 28, 28, True)]
feats, _ = synth.planted_features(n, layers, seed=2023)
Z = restated.embed(feats, 3, 1, 2048, 4096).reshape(n, 784, 4096)
P = Z.shape[1]
taus = [0.1, 0.2, 0.5, 1.0]


def alpha_of(w, tau):
    return restated.alpha_from_weights(w, tau, stable=True)


def weights(dmin_fn):
    """w[i] = mean over j != i of min_c dist(i-th image rows, image j)."""
    w = torch.zeros(n, P, dtype=torch.float64)
    for i in range(n):
        for j in range(n):
            if i != j:
                w[i] += dmin_fn(i, j)
    return (w / (n - 1)).float()


Zd = Z.double()
hi = Z.half()
hif = hi.float()
n2 = (hif * hif).sum(-1)                              # norms of the operand actually multiplied (ac_row_norms)
lo = (Z - hif).half().float()


def exact(i, j):
    return torch.cdist(Zd[i], Zd[j]).min(dim=1)[0]


def d2_f16(i, j):
    return n2[i][:, None] + n2[j][None, :] - 2.0 * (hif[i] @ hif[j].T)


def f16(i, j):
    return d2_f16(i, j).min(dim=1)[0].clamp_min(0).sqrt().double()


def f16x3(i, j):
    full_n2 = (Z * Z).sum(-1)
    dot = hif[i] @ hif[j].T + lo[i] @ hif[j].T + hif[i] @ lo[j].T
    return (full_n2[i][:, None] + full_n2[j][None, :] - 2.0 * dot).min(dim=1)[0].clamp_min(0).sqrt().double()


def f16_refine(i, j, k=1):
    """arg-min (top-k) from the fp16 pass, exact fp32 sum((x-y)^2) on those pairs, min of the k."""
    idx = d2_f16(i, j).topk(k, dim=1, largest=False)[1]                      # [P, k]
    cand = Z[j][idx]                                                          # [P, k, D] fp32 rows of the bank image
    d2 = ((Z[i][:, None, :] - cand) ** 2).sum(-1)                            # fp32, no expansion
    return d2.min(dim=1)[0].sqrt().double()


w_ex = weights(exact)
a_ex = [alpha_of(w_ex, t) for t in taus]
print("config-2 width, N=%d: patch norm mean %.1f, w mean %.2f, within-image w spread mean %.2f"
      % (n, Z.norm(dim=-1).mean().item(), w_ex.mean().item(), (w_ex.max(1)[0] - w_ex.min(1)[0]).mean().item()))
print("| mode | max abs dw | " + " | ".join("max abs dalpha tau=%g" % t for t in taus) + " |")
print("|---|---:|" + "---:|" * len(taus))
for name, fn in (("f16 (one pass)", f16), ("f16x3 (three passes)", f16x3), ("f16 + exact refine of arg-min", f16_refine),
                 ("f16 + exact refine of top-2", lambda i, j: f16_refine(i, j, 2))):
    w = weights(fn)
    da = [(alpha_of(w, t) - a).abs().max().item() for t, a in zip(taus, a_ex)]
    print("| %s | %.2e | " % (name, (w - w_ex).abs().max().item()) + " | ".join("%.1e" % x for x in da) + " |", flush=True)
