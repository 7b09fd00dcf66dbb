"""Mirror of Weight_Distance_* / Matrix_Alpha_* (models/patchcore/utils.py:222-277).

Same names, argument order and return dtypes (w: fp32 [P]; alpha: float64 [N,P]).  The N*(N-1)
torch.cdist launches of the reference become one tcgen05 GEMM with a fused per-bank-image row-min;
extra keyword `precision` selects the tensor-core operand mode (default: module-level PRECISION)."""
from __future__ import annotations

import math
import weakref
from typing import Dict, Optional, Tuple

import torch

from .. import ops, pipeline

PRECISION = "auto"  # auto | f16 | f16r | bf16 | f16x3 | bf16x3 | f32; auto = f16 for tau >= 1, else f16r

_cache: Dict[str, tuple] = {}


def clear_cache() -> None:
    """Drops the cached tensor-core operands of the last Z (they hold a view of the caller's Z plus an fp16 copy: about
    2 GB at config 2).  The cache only exists because the reference API is invoked once per image."""
    _cache.clear()


def _patchset(Z: torch.Tensor, precision: str) -> pipeline.PatchSet:
    """The reference API is stateless per call (Weight_Distance_*(Z, i, ...) is invoked once per image); the
    operands of an UNCHANGED Z are reused.  The entry is tied to the tensor object itself (weak reference) and its
    in-place version counter, so a different tensor that happens to reuse the same memory never hits."""
    ent = _cache.get("z")
    if ent is not None:
        ref, version, prec, ps = ent
        if ref() is Z and version == Z._version and prec == precision:
            return ps
    ps = pipeline.patchset_from_Z(Z, precision)
    _cache["z"] = (weakref.ref(Z), Z._version, precision, ps)
    return ps


def _to_device(Z, device):
    Z = torch.as_tensor(Z)
    return Z.to(device) if device is not None else Z


def _rows(ps: pipeline.PatchSet, i: int) -> pipeline.PatchSet:
    sl = slice(i * ps.P, (i + 1) * ps.P)
    pick = lambda t: None if t is None else t[sl]  # noqa: E731
    return pipeline.PatchSet(1, ps.P, ps.D, ps.grid, pick(ps.Z), pick(ps.hi), pick(ps.lo), pick(ps.n2))


def _prec(precision, taus, *Zs):
    """auto -> mode for these taus, then the fp16 range guard on the caller's embeddings (pipeline.guard_operand_range)."""
    return pipeline.guard_operand_range(pipeline.resolve_precision(precision or PRECISION, taus), *[torch.as_tensor(z) for z in Zs])


def Weight_Distance_Unsupervised(Z, i, device, precision: Optional[str] = None):
    """utils.py:222-227 -> w_i [P]: mean over j != i of min_q ||Z[i,p] - Z[j,q]||."""
    precision = _prec(precision, [0.1], Z)     # tau unknown here: assume the most demanding one
    ps = _patchset(_to_device(Z, device), precision)
    q_self = torch.tensor([i], dtype=torch.int32, device=ps.Z.device)
    return pipeline.min_distance_weights(_rows(ps, i), ps, "unsupervised", precision, q_self=q_self)[0]


def Weight_Distance_Supervised(Z, Z_train, i, device, precision: Optional[str] = None):
    """utils.py:230-237 -> w_i [P]: min over bank images and bank patches."""
    precision = _prec(precision, [0.1], Z, Z_train)
    ps = _patchset(_to_device(Z, device), precision)
    bank = pipeline.patchset_from_Z(_to_device(Z_train, device), precision)
    return pipeline.min_distance_weights(_rows(ps, i), bank, "supervised", precision)[0]


def _alpha(w: torch.Tensor, tau: float) -> torch.Tensor:
    a64, _ = ops.alpha(w, [0.0 if math.isclose(tau, 0) else float(tau)], want32=False)
    return a64[0]


def Matrix_Alpha_Unsupervised(tau, k, Z, device, precision: Optional[str] = None):
    """utils.py:240-257 -> [N,P] float64 (k cancels in the normalisation, as in the reference)."""
    print("{:-^80}".format("Calculating Unsupervised Alpha Matrix"))
    precision = _prec(precision, [tau], Z)
    ps = _patchset(_to_device(Z, device), precision)
    return _alpha(pipeline.min_distance_weights(ps, ps, "unsupervised", precision), tau)


def Matrix_Alpha_Supervised(tau, k, Z, Z_train, device, precision: Optional[str] = None):
    """utils.py:260-277 -> [N,P] float64."""
    print("{:-^80}".format("Calculating Supervised Alpha Matrix"))
    precision = _prec(precision, [tau], Z, Z_train)
    ps = _patchset(_to_device(Z, device), precision)
    bank = pipeline.patchset_from_Z(_to_device(Z_train, device), precision)
    return _alpha(pipeline.min_distance_weights(ps, bank, "supervised", precision), tau)
