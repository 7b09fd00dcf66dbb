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
    # own, compact CPU storage: torch.save serialises the WHOLE underlying storage of a view (a slice of the [T,N,P]
    # alpha tensor would drag every other tau's alpha into each file), and the reference's pickles load without CUDA
    a = a.float().cpu().contiguous().clone()
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


# ---- info_<category>.pickle (main.py:253-262 collects it; test.py:156 / draw_alpha.py read it) ---------------------------

def info_path(outputs_root: str, dataset: str, category: str) -> str:
    """test.py:156 -- <outputs>/<dataset>/info/info_<category>.pickle."""
    return os.path.join(outputs_root, dataset, "info", "info_" + category + ".pickle")


def make_info(classname: str, anomalies: Sequence[str], image_names: Sequence[str] = None, image_paths: Sequence[str] = None):
    """The list the reference builds from its batch_size=1 loader (main.py:253-262): one dict per image, every
    string wrapped in a 1-element list and `is_anomaly` a 1-element tensor (default collate of a batch of one)."""
    out = []
    for i, a in enumerate(anomalies):
        name = image_names[i] if image_names is not None else "%s/test/%s/%03d.png" % (classname, a, i)
        out.append({"classname": [classname], "anomaly": [str(a)], "is_anomaly": torch.tensor([int(a != "good")]),
                    "image_name": [name], "image_path": [image_paths[i] if image_paths is not None else name]})
    return out


def save_info(outputs_root: str, dataset: str, category: str, info) -> str:
    p = info_path(outputs_root, dataset, category)
    os.makedirs(os.path.dirname(p), exist_ok=True)
    torch.save(info, p)
    return p


def load_info(path: str):
    return torch.load(path, map_location="cpu", weights_only=False)


def anomaly_names(info) -> list:
    """test.py:183-186 -- info[i]['anomaly'][0]."""
    return [str(rec["anomaly"][0]) if isinstance(rec["anomaly"], (list, tuple)) else str(rec["anomaly"]) for rec in info]


# ---- <layers>_<Dp>_<D>_tau_result.csv (test.py:246-325) -------------------------------------------------------------------

OBJECT = ["bottle", "cable", "capsule", "hazelnut", "metal_nut", "pill", "screw", "toothbrush", "transistor", "zipper"]
TEXTURE = ["carpet", "grid", "leather", "tile", "wood"]


def result_csv_path(mode_dir: str, layers: Sequence[str], pretrain_dim: int, target_dim: int, variable: str = "tau") -> str:
    """test.py:247-249."""
    return os.path.join(mode_dir, "_".join(layers) + "_" + str(pretrain_dim) + "_" + str(target_dim) + "_" + variable + "_result.csv")


def write_result_csv(path: str, supervised: str, blocks, variable: str = "TAU") -> str:
    """`blocks` = [(tau, [(row_name, NMI, ARI, F1), ...]), ...] in the order to be written; the aggregate rows
    ('MVTec(object)', 'MVTec(texture)') are ordinary rows supplied by the caller (cluster.evaluate_runs).
    Same writer, quoting and float formatting (repr) as test.py:252-325, so a file parsed by read_result_csv
    and written back is byte-identical."""
    import csv

    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w", newline="", encoding="gbk") as f:
        wr = csv.writer(f)
        wr.writerow([supervised])
        wr.writerow(["Category", "NMI", "ARI", "F1"])
        for tau, rows in blocks:
            wr.writerow(["---"] * 4)
            wr.writerow([variable + "=" + str(tau)])
            for name, nmi, ari, f1 in rows:
                wr.writerow([name, nmi, ari, f1])
    return path


def read_result_csv(path: str):
    """-> (supervised, [(tau_string, [(row_name, NMI, ARI, F1), ...]), ...]); tau is kept as written ('0', '0.2', '1')."""
    import csv

    with open(path, newline="", encoding="gbk") as f:
        rows = [r for r in csv.reader(f)]
    supervised = rows[0][0]
    assert rows[1] == ["Category", "NMI", "ARI", "F1"], rows[1]
    blocks = []
    for r in rows[2:]:
        if not r or r[0] == "---":
            continue
        if len(r) == 1 and "=" in r[0]:
            blocks.append((r[0].split("=", 1)[1], []))
        else:
            blocks[-1][1].append((r[0], float(r[1]), float(r[2]), float(r[3])))
    return supervised, blocks


# ---- optional: the tau-independent weights, so a tau sweep re-runs only alpha -> X -> Dmat ------------

def save_weights(run_root: str, category: str, supervised: str, w) -> str:
    os.makedirs(run_root, exist_ok=True)
    p = os.path.join(run_root, "weights_" + category + "_" + supervised + ".pickle")
    torch.save(torch.as_tensor(w).detach().float().cpu(), p)
    return p


def load_weights(path: str) -> torch.Tensor:
    return torch.load(path, map_location="cpu", weights_only=True)
