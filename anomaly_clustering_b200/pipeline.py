"""Feature maps -> (Z, w, alpha, X, Dmat) on one B200, and its query-sharded multi-GPU form.

This is the device-resident replacement of the body of make_category_data
(reference: Anomaly-Clustering/examples/main.py:266-296) from the hooked backbone features on:

    Z      = AnomalyClusteringCore.embed(...)                      main.py:266-267
    alpha  = Matrix_Alpha_Unsupervised / _Supervised / average     main.py:281-291
    X      = bmm(alpha, Z)                                         main.py:294-296
    Dmat   = pairwise Euclid on X (Ward's input)                   test.py:193-195

Nothing bounces through the host (the reference copies every patch row to numpy and back,
patchcore.py:358-361 + main.py:267) and one min-distance pass serves every tau.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

# Unsupervised self-bank runs multiply every unordered image pair once (ac_min_dist_sym); set to
# False to force the straightforward all-pairs kernel (bench.py --no-symmetry, A/B tests).
SYMMETRIC = True

# bench.py sets this to a list to collect (tag, cuda event) marks at stage boundaries
PROFILE = None


def _mark(tag: str) -> None:
    if PROFILE is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PROFILE.append((tag, ev))


# taus below this go to the refined mode under precision="auto" (one fp16 pass keeps alpha inside
# 1e-3 with a 2x margin from tau = 1 on, also at real-data patch norms; tau = 0 is the arg-max and needs exact w too)
AUTO_REFINE_BELOW_TAU = 1.0


def resolve_precision(precision: str, taus: Sequence[float]) -> str:
    """'auto' -> the cheapest mode that keeps alpha inside the 1e-3 max-abs tolerance with a 2x margin (measured table in
    DESIGN.md section 5): one fp16 tensor-core pass when every tau >= 1; otherwise 'f16r' = the same pass recording the
    arg-min + an exact fp32 re-evaluation of the selected pairs (ac_refine_min_dist, ~1.4x the one-pass cost)."""
    if precision != "auto":
        return precision
    small = [t for t in taus if abs(float(t)) < AUTO_REFINE_BELOW_TAU]
    return "f16r" if small else "f16"


@functools.lru_cache(maxsize=64)
def _fusable(n_layers: int, pretrain_dim: int, target_dim: int) -> bool:
    return ops.aggregator_fusable(n_layers, pretrain_dim, target_dim)


_OPERAND_OF = {"f16": ("f16", False), "bf16": ("bf16", False), "f16x3": ("f16", True), "bf16x3": ("bf16", True), "f32": (None, False),
               "f16r": ("f16", False), "bf16r": ("bf16", False)}
REFINED = ("f16r", "bf16r")


@dataclass
class PatchSet:
    """Embedded images: fp32 Z plus the tensor-core operands and their squared norms."""

    n_img: int
    P: int
    D: int
    grid: Tuple[int, int]
    Z: Optional[torch.Tensor] = None      # [n_img*P, D] fp32
    hi: Optional[torch.Tensor] = None     # [n_img*P, D] f16 | bf16
    lo: Optional[torch.Tensor] = None
    n2: Optional[torch.Tensor] = None     # [n_img*P] fp32 squared norms of the operand (hi [+ lo])


@dataclass
class PathResult:
    Z: Optional[torch.Tensor]             # [N, P, D] fp32 (None when run with keep_z=False)
    w: Optional[torch.Tensor]             # [N, P] fp32 (None in 'average' mode)
    alpha64: torch.Tensor                 # [T, N, P] float64  (reference dtype, utils.py:246)
    alpha32: torch.Tensor                 # [T, N, P] float32  (main.py:294 `.float()`)
    X: torch.Tensor                       # [T, N, D] fp32
    Dmat: torch.Tensor                    # [T, N, N] fp32
    taus: List[float] = field(default_factory=list)
    grid: Tuple[int, int] = (0, 0)


def embed_images(
    features: Sequence[torch.Tensor],
    patchsize: int,
    stride: int,
    pretrain_dim: int,
    target_dim: int,
    precision: str = "f16",
    want_z: bool = True,
    layernorm: bool = True,
    batch: Optional[int] = None,
    out=None,
) -> PatchSet:
    """Stage 1 for a set of images in one C-ABI call (ac_embed_ex: Z and / or the tensor-core operands, and the operands'
    squared norms); `batch` only caps the images per call.  `out` = (hi, lo | None, n2) preallocated operand buffers (the
    sharded path hands in its slice of the symmetric bank so that no staging copy is needed)."""
    operand, want_lo = _OPERAND_OF[precision]
    views = [ops.feature_view(f) for f in features]
    N = views[0].shape[0]
    _mark("embed_begin")
    h, w = ops.patch_grid(views[0].shape[2], views[0].shape[3], patchsize, stride)
    P = h * w
    dev = views[0].device
    # operand-only output needs Aggregator windows that stay inside one layer; otherwise Z is produced (and dropped below)
    need_z = want_z or operand is None or not _fusable(len(views), pretrain_dim, target_dim)
    Z = torch.empty(N * P, target_dim, dtype=torch.float32, device=dev) if need_z else None
    hi = lo = n2 = None
    if operand is not None:
        tdt = torch.float16 if operand == "f16" else torch.bfloat16
        if out is not None:
            hi, lo, n2 = out
            assert hi.shape == (N * P, target_dim) and hi.dtype == tdt and hi.is_contiguous() and n2.shape == (N * P,)
            assert (lo is not None) == want_lo
        else:
            hi = torch.empty(N * P, target_dim, dtype=tdt, device=dev)
            lo = torch.empty(N * P, target_dim, dtype=tdt, device=dev) if want_lo else None
            n2 = torch.empty(N * P, dtype=torch.float32, device=dev)
    step = N if not batch else batch
    for b0 in range(0, N, step):
        b1 = min(N, b0 + step)
        sl = slice(b0 * P, b1 * P)
        ops.embed(
            [v[b0:b1] for v in views], patchsize, stride, pretrain_dim, target_dim, layernorm=layernorm,
            want_z=need_z, operand=operand, want_lo=want_lo,
            out_z=None if Z is None else Z[sl], out_hi=None if hi is None else hi[sl], out_lo=None if lo is None else lo[sl],
            out_n2=None if n2 is None else n2[sl],
        )
    _mark("embed_end")
    if not (want_z or operand is None):
        Z = None
    return PatchSet(n_img=N, P=P, D=target_dim, grid=(h, w), Z=Z, hi=hi, lo=lo, n2=n2)


# element range inside which an arbitrary fp32 embedding may be rounded to fp16 operands: fp16 overflows at 65 504 and
# loses relative precision below its normal range (6e-5); LayerNorm'd embeddings of this path are O(1)
F16_OPERAND_RANGE = (2.0 ** -7, 2.0 ** 14)


def guard_operand_range(precision: str, *tensors: torch.Tensor) -> str:
    """Precision mode that is safe for embeddings that did NOT come out of this package's embed stage (the mirror
    Matrix_Alpha_* / Weight_Distance_* accept any fp32 Z, like the reference's torch.cdist, utils.py:226).  Non-finite
    values are refused; if the largest magnitude lies outside F16_OPERAND_RANGE the fp16 modes are replaced by the exact
    fp32 kernel (slow, correct) with a warning instead of silently overflowing / flushing the operands."""
    if precision == "f32":
        return precision
    fp16_mode = precision.startswith("f16")
    lo, hi = F16_OPERAND_RANGE
    for t in tensors:
        if t is None or not t.numel():
            continue
        amax = float(t.detach().abs().amax())
        if amax != amax or amax == float("inf"):
            raise ValueError("embedding contains non-finite values")
        if fp16_mode and amax > 0 and not (lo <= amax <= hi):
            import warnings

            warnings.warn("embedding magnitude %.3g is outside the fp16 operand range [%.3g, %.3g]: using the exact fp32 distance kernel "
                          "(precision='f32'); rescale Z to O(1) for the tensor-core path" % (amax, lo, hi), RuntimeWarning, stacklevel=3)
            return "f32"
    return precision


def patchset_from_Z(Z: torch.Tensor, precision: str = "f16") -> PatchSet:
    """Wrap an existing [N,P,D] fp32 embedding (e.g. the reference's `Z`) for the distance stage."""
    N, P, D = Z.shape
    Zf = Z.reshape(N * P, D).contiguous().float()
    operand, want_lo = _OPERAND_OF[precision]
    ps = PatchSet(n_img=N, P=P, D=D, grid=(0, 0), Z=Zf)
    if operand is not None:
        ps.hi, ps.lo = ops.split_operand(Zf, operand, want_lo)
        ps.n2 = ops.row_norms(ps.hi, ps.lo)
    return ps


def min_distance_weights(
    q: PatchSet,
    bank: PatchSet,
    mode: str,
    precision: str = "f16",
    q_self: Optional[torch.Tensor] = None,
    return_dmin: bool = False,
    groups: Optional[torch.Tensor] = None,
    bank_ready: Optional[torch.Tensor] = None,
    bank_first: int = 0,
    bank_landed=None,
):
    """Stage 2: w [Nq, P].  mode 'unsupervised' = mean over bank images != self (utils.py:222-227),
    'supervised' = min over bank images (utils.py:230-237).  q_self[i] = bank index of query image i.
    groups (ops.make_groups): the images are several categories back to back, each its own bank (symmetric form only).
    bank_ready / bank_first / bank_landed (sharded supervised runs): arrival flags of a bank whose remote shards are still being
    pulled, the first resident bank image, and a callable that orders the current stream after all pulls (called before the
    refine pass, which gathers from the whole bank)."""
    refined = precision in REFINED
    if (mode == "unsupervised" and SYMMETRIC and q is bank and precision != "f32" and q.P >= 32 and not return_dmin
            and q_self is None):
        _mark("mindist_begin")
        if refined:
            rowmin, rowarg, colkey = ops.min_dist_sym_arg(q.hi, q.lo, q.n2, 0, q.hi, q.lo, q.n2, q.n_img, q.P, precision, groups=groups)
        else:
            rowmin, colmin = ops.min_dist_sym(q.hi, q.lo, q.n2, 0, q.hi, q.lo, q.n2, q.n_img, q.P, precision, groups=groups)
        _mark("mindist_end")
        if not refined:
            return ops.reduce_weights_sym(rowmin, colmin, q.P, 0, groups=groups).reshape(q.n_img, q.P)
        _mark("refine_begin")
        dex = ops.refine_min_dist(q.Z, q.hi, q.lo, q.hi, q.lo, q.n_img, q.P, rowarg, colkey=colkey, q_img0=0, groups=groups, Bn2=q.n2)
        _mark("refine_end")
        own = torch.arange(q.n_img, dtype=torch.int32, device=dex.device)
        return ops.reduce_weights(dex, q.P, own, "mean", groups=groups).reshape(q.n_img, q.P)
    if groups is not None:
        raise ValueError("per-category groups need the symmetric unsupervised form (P >= 32, tensor-core precision)")
    _mark("mindist_begin")
    if precision == "f32":
        dmin = ops.min_dist(q.Z, None, None, bank.Z, None, None, bank.n_img, bank.P, "f32")
    elif refined:
        dmin, arg = ops.min_dist_arg(q.hi, q.lo, q.n2, bank.hi, bank.lo, bank.n2, bank.n_img, bank.P, precision, ready=bank_ready,
                                     first_image=bank_first)
    else:
        dmin = ops.min_dist(q.hi, q.lo, q.n2, bank.hi, bank.lo, bank.n2, bank.n_img, bank.P, precision, ready=bank_ready,
                            first_image=bank_first)
    _mark("mindist_end")
    if bank_landed is not None:
        bank_landed()
    if refined:
        if mode == "unsupervised" and q_self is None:
            q_self = torch.arange(q.n_img, dtype=torch.int32, device=dmin.device)
        _mark("refine_begin")
        dmin = ops.refine_min_dist(q.Z, q.hi, q.lo, bank.hi, bank.lo, bank.n_img, bank.P, arg,
                                   q_self=q_self if mode == "unsupervised" else None, Pq=q.P, Bn2=bank.n2)
        _mark("refine_end")
    if mode == "unsupervised":
        if q_self is None:
            q_self = torch.arange(q.n_img, dtype=torch.int32, device=dmin.device)
        w = ops.reduce_weights(dmin, q.P, q_self, "mean")
    elif mode == "supervised":
        w = ops.reduce_weights(dmin, q.P, None, "min")
    else:
        raise ValueError(mode)
    w = w.reshape(q.n_img, q.P)
    return (w, dmin) if return_dmin else w


def alpha_X_dist(ps: PatchSet, w: Optional[torch.Tensor], taus: Sequence[float], features=None, embed_args=None, want_dmat=True):
    """Stage 3 for every tau from one w.  w=None -> 'average' mode (main.py:290-291).
    When the PatchSet carries no fp32 Z, X is computed straight from `features` (ac_weighted_embed_from_features)."""
    N, P, D = ps.n_img, ps.P, ps.D
    dev = (ps.Z if ps.Z is not None else ps.hi).device
    if w is None:
        a32 = torch.full((1, N, P), 1.0 / P, dtype=torch.float32, device=dev)
        a64 = a32.double()
        taus = [float("nan")]
    else:
        a64, a32 = ops.alpha(w, taus)
    if ps.Z is not None:
        Z3 = ps.Z.reshape(N, P, D)
        X = ops.weighted_embed_multi(a32, Z3)                    # [T, N, D], Z read once for all taus
    else:
        patchsize, stride, Dp, layernorm = embed_args
        X = torch.stack([ops.weighted_embed_from_features(features, a32[t], patchsize, stride, Dp, D, layernorm=layernorm)
                         for t in range(a32.shape[0])])
    Dm = torch.stack([ops.pairwise_l2(X[t]) for t in range(a32.shape[0])]) if want_dmat else None
    return a64, a32, X, Dm


def z_free_supported(features, patchsize, stride, pretrain_dim, target_dim, precision) -> bool:
    """Shapes covered by ac_weighted_embed_from_features (see include/ac_b200.h)."""
    if precision == "f32" or patchsize != 3 or stride != 1:
        return False
    views = [ops.feature_view(f) for f in features]
    if any(v.shape[2:] != views[0].shape[2:] for v in views):
        return False
    L = len(views)
    if (L * pretrain_dim) % target_dim != 0 or pretrain_dim % ((L * pretrain_dim) // target_dim) != 0:
        return False
    return True


def run_path(
    features: Sequence[torch.Tensor],
    patchsize: int = 3,
    stride: int = 1,
    pretrain_dim: int = 1024,
    target_dim: int = 1024,
    mode: str = "unsupervised",
    taus: Sequence[float] = (1.0,),
    bank_features: Optional[Sequence[torch.Tensor]] = None,
    precision: str = "auto",
    layernorm: bool = True,
    keep_z: bool = True,
) -> PathResult:
    """Single-GPU hot path: hooked features (device tensors) -> PathResult (device tensors).
    keep_z=False never materialises the fp32 Z (1.3 GB at config 2): the embed kernel writes only the
    tensor-core operands and X comes straight from the feature maps; falls back to keep_z=True for shapes the
    Z-free form does not cover (resampled layers, patch size != 3, fp32 precision)."""
    precision = resolve_precision(precision, taus)
    if not keep_z and not z_free_supported(features, patchsize, stride, pretrain_dim, target_dim, precision):
        keep_z = True
    if precision in REFINED and mode != "average":
        keep_z = True      # the exact re-evaluation takes the QUERY rows from fp32 Z (their rounding would not average out)
    q = embed_images(features, patchsize, stride, pretrain_dim, target_dim, precision, want_z=keep_z, layernorm=layernorm)
    w = None
    if mode == "unsupervised":
        w = min_distance_weights(q, q, "unsupervised", precision)
    elif mode == "supervised":
        if bank_features is None:
            raise ValueError("supervised mode needs bank_features (the normal training images)")
        bank = embed_images(bank_features, patchsize, stride, pretrain_dim, target_dim, precision, want_z=False,
                            layernorm=layernorm)
        w = min_distance_weights(q, bank, "supervised", precision)
    elif mode != "average":
        raise ValueError("mode must be unsupervised | supervised | average")
    a64, a32, X, Dm = alpha_X_dist(q, w, list(taus), features, (patchsize, stride, pretrain_dim, layernorm))
    return PathResult(Z=None if q.Z is None else q.Z.reshape(q.n_img, q.P, q.D), w=w, alpha64=a64, alpha32=a32, X=X, Dmat=Dm,
                      taus=list(taus), grid=q.grid)


def run_categories(
    features: Sequence[torch.Tensor],
    sizes: Sequence[int],
    patchsize: int = 3,
    stride: int = 1,
    pretrain_dim: int = 1024,
    target_dim: int = 1024,
    taus: Sequence[float] = (1.0,),
    precision: str = "auto",
    keep_z: bool = True,
    layernorm: bool = True,
) -> List[PathResult]:
    """Several independent categories (the reference's own semantics: one make_category_data per category, each with
    its own bank -- examples/main.py:353) from ONE batch of hooked features and ONE launch sequence: `features` hold the
    images of all categories back to back, `sizes[c]` images each.  The embed, the tensor-core pass (image pairs only
    inside a category: ac_min_dist_sym_ex with a category table), the reductions, alpha and X run once for all
    categories; only the small per-category distance matrices are separate launches.  Returns one PathResult per category
    (views into the shared tensors).  Shapes the batched form does not cover run category by category."""
    precision = resolve_precision(precision, taus)
    sizes = [int(n) for n in sizes]
    views = [ops.feature_view(f) for f in features]
    N = views[0].shape[0]
    if sum(sizes) != N:
        raise ValueError("sizes sum to %d but the features hold %d images" % (sum(sizes), N))
    h, w_ = ops.patch_grid(views[0].shape[2], views[0].shape[3], patchsize, stride)
    batched = SYMMETRIC and precision != "f32" and h * w_ >= 32 and min(sizes) >= 2
    if not batched:
        out, start = [], 0
        for n in sizes:
            out.append(run_path([f[start:start + n] for f in features], patchsize, stride, pretrain_dim, target_dim, "unsupervised",
                                taus, precision=precision, keep_z=keep_z, layernorm=layernorm))
            start += n
        return out
    if not keep_z and not z_free_supported(features, patchsize, stride, pretrain_dim, target_dim, precision):
        keep_z = True
    if precision in REFINED:
        keep_z = True
    q = embed_images(features, patchsize, stride, pretrain_dim, target_dim, precision, want_z=keep_z, layernorm=layernorm)
    groups = ops.make_groups(sizes, q.hi.device)
    w = min_distance_weights(q, q, "unsupervised", precision, groups=groups)
    a64, a32, X, _ = alpha_X_dist(q, w, list(taus), features, (patchsize, stride, pretrain_dim, layernorm), want_dmat=False)
    out, start = [], 0
    Z3 = None if q.Z is None else q.Z.reshape(q.n_img, q.P, q.D)
    for n in sizes:
        sl = slice(start, start + n)
        Dm = torch.stack([ops.pairwise_l2(X[t][sl]) for t in range(len(taus))])
        out.append(PathResult(Z=None if Z3 is None else Z3[sl], w=w[sl], alpha64=a64[:, sl], alpha32=a32[:, sl], X=X[:, sl], Dmat=Dm,
                              taus=list(taus), grid=q.grid))
        start += n
    return out
