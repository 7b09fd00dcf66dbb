"""CPU: the consumer side (Ward on a distance matrix, best_map, metrics) and the on-disk format,
checked against the reference's shipped results (tests/golden/shipped_cluster_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from anomaly_clustering_b200 import cluster, io


def test_metrics_from_distance_matrix_match_published_csv(golden_dir):
    g = np.load(os.path.join(golden_dir, "shipped_cluster_golden.npz"), allow_pickle=False)
    for key in g["cases"]:
        key = str(key)
        X = g[key + "_Xlow"]
        D = np.sqrt(np.maximum(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1), 0))
        nmi, ari, f1, _, _ = cluster.calculate_metrics(D.astype(np.float32), [str(a) for a in g[key + "_anomaly"]])
        assert np.allclose([nmi, ari, f1], g[key + "_csv"], atol=1e-9), key


def test_best_map_permutation_invariance():
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 4, size=60)
    perm = np.array([2, 0, 3, 1])
    assert np.array_equal(cluster.best_map(lab, perm[lab]), lab)


def test_size_weighted_mean():
    assert cluster.size_weighted_mean([1.0, 0.0], [3, 1]) == 0.75


def test_pickle_roundtrip_matches_reference_layout(tmp_path):
    alpha = torch.rand(5, 784, dtype=torch.float64)
    alpha /= alpha.sum(1, keepdim=True)
    X = np.random.default_rng(1).normal(size=(5, 64)).astype(np.float32)
    p = io.save_matrix_alpha_X(str(tmp_path), ["blocks.10", "blocks.11"], 2048, 4096, 2, 1, "bottle", "unsupervised", alpha, X)
    assert p.endswith(os.path.join("blocks.10_blocks.11_2048_4096_2.0_1.0", "matrix_alpha_X_bottle_unsupervised.pickle"))
    a, Xl = io.load_matrix_alpha_X(p)
    assert a.shape == (5, 1, 784) and a.dtype == torch.float32
    assert np.array_equal(Xl, X)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Anomaly-Clustering/outputs"), reason="reference artefacts absent")
def test_reads_the_reference_shipped_pickle():
    p = ("/root/reference/Anomaly-Clustering/outputs/mvtec_ad/dino_vitbase8/unsupervised/"
         "blocks.10_blocks.11_2048_4096_2.0_1.0/matrix_alpha_X_bottle_unsupervised.pickle")
    a, X = io.load_matrix_alpha_X(p)
    assert a.shape == (83, 1, 784) and X.shape == (83, 4096)
