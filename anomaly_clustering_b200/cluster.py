"""Consumer of the hot path: Ward clustering on the GPU-produced distance matrix + the reference's
metrics.  CPU code by design -- BASELINE config 1 keeps Ward on the CPU;
N <= a few hundred images per category.

Reference: Anomaly-Clustering/examples/test.py:109-131 (best_map), :177-226 (label filtering,
AgglomerativeClustering(n_clusters=k) = Ward/Euclidean, NMI / ARI / F1-micro)."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def ward_labels_from_dmat(Dmat, k: int) -> np.ndarray:
    """Ward linkage cut at k clusters from a precomputed Euclidean distance matrix.  Equivalent to
    sklearn AgglomerativeClustering(n_clusters=k).fit_predict(X) (test.py:193-195), which runs
    scipy's ward on pdist(X) internally; labels may be permuted (best_map removes that)."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform

    D = np.asarray(Dmat, dtype=np.float64)
    D = 0.5 * (D + D.T)
    np.fill_diagonal(D, 0.0)
    Zl = linkage(squareform(D, checks=False), method="ward")
    return fcluster(Zl, t=k, criterion="maxclust") - 1


def best_map(L1, L2) -> np.ndarray:
    """test.py:109-131 with scipy's Hungarian solver in place of munkres."""
    from scipy.optimize import linear_sum_assignment

    L1, L2 = np.asarray(L1), np.asarray(L2)
    u1, u2 = np.unique(L1), np.unique(L2)
    n = max(len(u1), len(u2))
    G = np.zeros((n, n))
    for i, a in enumerate(u1):
        for j, b in enumerate(u2):
            G[i, j] = np.sum((L1 == a) & (L2 == b))
    rows, cols = linear_sum_assignment(-G.T)
    c = np.zeros(n, dtype=int)
    c[rows] = cols
    out = np.zeros(L2.shape)
    for j, b in enumerate(u2):
        out[L2 == b] = u1[c[j]] if c[j] < len(u1) else -1
    return out


def calculate_metrics(Dmat, anomaly_names: Sequence[str]):
    """test.py:177-226 from the distance matrix: drop 'combined' images, encode labels, Ward(k),
    best_map, NMI / ARI / F1-micro.  Returns (NMI, ARI, F1, label, predict)."""
    from sklearn import metrics
    from sklearn.preprocessing import LabelEncoder

    keep = np.array([i for i, a in enumerate(anomaly_names) if a != "combined"], dtype=int)
    D = np.asarray(Dmat, dtype=np.float64)[np.ix_(keep, keep)]
    label = LabelEncoder().fit_transform([anomaly_names[i] for i in keep]).astype(int)
    predict = best_map(label, ward_labels_from_dmat(D, len(set(label)))).astype(int)
    return (metrics.normalized_mutual_info_score(label, predict), metrics.adjusted_rand_score(label, predict),
            metrics.f1_score(label, predict, average="micro"), label, predict)


def size_weighted_mean(values: Sequence[float], sizes: Sequence[int]) -> float:
    """test.py:286-325 -- object / texture aggregate rows of the tau-result CSV."""
    v, s = np.asarray(values, dtype=np.float64), np.asarray(sizes, dtype=np.float64)
    return float((v * s).sum() / s.sum())


def evaluate_runs(outputs_root: str, dataset: str, backbone: str, supervised: str, layers: Sequence[str], pretrain_dim: int,
                  target_dim: int, taus: Sequence[float], train_ratio: float = 1, objects: Sequence[str] = None,
                  textures: Sequence[str] = None, dmat_fn=None, write_csv: bool = True):
    """The `__main__` loop of test.py:228-325: for every tau and category load the (alpha, X) pickle and
    info_<category>.pickle, cluster, and write <layers>_<Dp>_<D>_tau_result.csv with the size-weighted
    'MVTec(object)' / 'MVTec(texture)' rows.  `dmat_fn(X ndarray) -> [N,N]` defaults to the library's
    ac_pairwise_l2 on the GPU (stage a13); categories whose pickle is missing are skipped (the reference would
    crash).  Returns the blocks written."""
    import os

    from . import io

    if dmat_fn is None:
        import torch

        from . import ops

        def dmat_fn(X):
            return ops.pairwise_l2(torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()).cpu().numpy()

    objects = io.OBJECT if objects is None else list(objects)
    textures = io.TEXTURE if textures is None else list(textures)
    mode_dir = os.path.join(outputs_root, dataset, backbone, supervised)
    blocks = []
    for tau in taus:
        rows, aggregates = [], []
        for group, title in ((objects, "MVTec(object)"), (textures, "MVTec(texture)")):
            vals, sizes = [], []
            for category in group:
                p = os.path.join(io.run_dir(mode_dir, layers, pretrain_dim, target_dim, tau, train_ratio),
                                 "matrix_alpha_X_" + category + "_" + supervised + ".pickle")
                ip = io.info_path(outputs_root, dataset, category)
                if not (os.path.exists(p) and os.path.exists(ip)):
                    continue
                _, X = io.load_matrix_alpha_X(p)
                nmi, ari, f1, label, _ = calculate_metrics(dmat_fn(X), io.anomaly_names(io.load_info(ip)))
                rows.append((category, nmi, ari, f1))
                vals.append((nmi, ari, f1))
                sizes.append(len(label))
            if vals:    # test.py:303-325: both aggregate rows follow all category rows
                aggregates.append((title,) + tuple(size_weighted_mean([v[k] for v in vals], sizes) for k in range(3)))
        blocks.append((tau, rows + aggregates))
    if write_csv:
        io.write_result_csv(io.result_csv_path(mode_dir, layers, pretrain_dim, target_dim), supervised, blocks)
    return blocks
