"""Restatement of the reference's clustering + metrics consumer.  TEST INFRASTRUCTURE ONLY.

Follows examples/test.py:109-131 (best_map) and :177-226 (calculate_metrics tail).
`munkres` is not installed; scipy.optimize.linear_sum_assignment solves the same
assignment problem (minimise -G^T)."""
from __future__ import annotations

import numpy as np


def best_map(L1, L2):
    """examples/test.py:109-131 -- relabel clustering L2 to best match labels L1."""
    from scipy.optimize import linear_sum_assignment

    L1 = np.asarray(L1)
    L2 = np.asarray(L2)
    Label1 = np.unique(L1)
    Label2 = np.unique(L2)
    nClass1, nClass2 = len(Label1), len(Label2)
    nClass = max(nClass1, nClass2)
    G = np.zeros((nClass, nClass))
    for i in range(nClass1):
        for j in range(nClass2):
            G[i, j] = np.sum((L2 == Label2[j]) & (L1 == Label1[i]))
    rows, cols = linear_sum_assignment(-G.T)
    c = np.zeros(nClass, dtype=int)
    c[rows] = cols
    newL2 = np.zeros(L2.shape)
    for i in range(nClass2):
        newL2[L2 == Label2[i]] = Label1[c[i]]
    return newL2


def ward_labels(X, k):
    """examples/test.py:193-195 -- AgglomerativeClustering(n_clusters=k) defaults = Ward/Euclidean."""
    from sklearn import cluster

    return cluster.AgglomerativeClustering(n_clusters=k).fit_predict(np.asarray(X))


def metrics_from_X(X, anomaly_names):
    """examples/test.py:177-226 -- drop 'combined', LabelEncoder, Ward(k), best_map, NMI/ARI/F1-micro."""
    from sklearn import metrics
    from sklearn.preprocessing import LabelEncoder

    keep = [i for i, a in enumerate(anomaly_names) if a != "combined"]
    Xk = np.asarray(X)[keep].astype(np.float64)
    label = LabelEncoder().fit_transform([anomaly_names[i] for i in keep]).astype(int)
    predict = ward_labels(Xk, len(set(label)))
    predict = best_map(label, predict).astype(int)
    NMI = metrics.normalized_mutual_info_score(label, predict)
    ARI = metrics.adjusted_rand_score(label, predict)
    F1 = metrics.f1_score(label, predict, average="micro")
    return NMI, ARI, F1, label, predict
