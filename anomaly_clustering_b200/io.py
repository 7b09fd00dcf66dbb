"""On-disk format of the reference: the (matrix_alpha, X) tuple that
examples/main.py:302-309 torch.save()s and examples/test.py:150-156 / utils/draw_alpha.py:64-71 load."""
from __future__ import annotations

import os
from typing import Sequence, Tuple

import numpy as np
import torch


def run_dir(save_path: str, layers: Sequence[str], pretrain_dim: int, target_dim: int, tau: float, train_ratio: float) -> str:
    """main.py:302-305 directory naming: <layers>_<Dp>_<D>_<float tau>_<float train_ratio>."""
    return os.path.join(save_path, "_".join(layers) + "_" + str(pretrain_dim) + "_" + str(target_dim) + "_" + str(float(tau)) + "_"
                        + str(float(train_ratio)))


def save_matrix_alpha_X(save_path, layers, pretrain_dim, target_dim, tau, train_ratio, category, supervised, matrix_alpha, X) -> str:
    """Writes exactly what the reference writes: (alpha [N,1,P] float32 tensor, X [N,D] float32 ndarray)."""
    d = run_dir(save_path, layers, pretrain_dim, target_dim, tau, train_ratio)
    os.makedirs(d, exist_ok=True)
    a = torch.as_tensor(matrix_alpha).detach()
    if a.dim() == 2:
        a = a.unsqueeze(1)
    a = a.float()
    Xn = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
    path = os.path.join(d, "matrix_alpha_X_" + category + "_" + supervised + ".pickle")
    torch.save((a, Xn.astype(np.float32)), path)
    return path


def load_matrix_alpha_X(path: str) -> Tuple[torch.Tensor, np.ndarray]:
    """test.py:154 -- torch.load(..., map_location='cpu'); also reads the reference's shipped files."""
    import numpy

    with torch.serialization.safe_globals([numpy.ndarray, numpy.dtype, numpy._core.multiarray._reconstruct,
                                           type(numpy.dtype("float32"))]):
        try:
            alpha, X = torch.load(path, map_location="cpu", weights_only=True)
        except Exception:
            alpha, X = torch.load(path, map_location="cpu", weights_only=False)
    return alpha, np.asarray(X)
