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
