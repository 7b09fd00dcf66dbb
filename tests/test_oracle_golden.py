"""CPU: the oracle restatement against the committed reference outputs (tests/golden, generated
by oracle/make_golden.py from the UNMODIFIED reference) and against the reference's shipped results."""
import os

import numpy as np
import pytest
import torch

from oracle import cluster as ocluster
from oracle import restated

EMBED_CASES = ["embed_vit_small", "embed_wrn_small", "embed_ragged", "embed_k5s2", "embed_single"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", EMBED_CASES)
def test_embed_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    feats = [torch.from_numpy(g["feat%d" % i]) for i in range(int(g["L"]))]
    z = restated.embed(feats, int(g["patchsize"]), int(g["stride"]), int(g["Dp"]), int(g["D"]))
    assert z.shape == g["Z"].shape
    assert np.abs(z.numpy() - g["Z"]).max() <= 2e-6


def test_alpha_w_X_match_reference(golden_dir):
    g = load(golden_dir, "alpha_small")
    Z, Zb = torch.from_numpy(g["Z"]), torch.from_numpy(g["Z_train"])
    wu = restated.weight_distance_unsupervised(Z)
    ws = restated.weight_distance_supervised(Z, Zb)
    assert np.abs(wu.numpy() - g["w_unsup"]).max() <= 1e-5
    assert np.abs(ws.numpy() - g["w_sup"]).max() <= 1e-5
    for t in g["taus"]:
        au = restated.alpha_from_weights(wu, float(t))
        asup = restated.alpha_from_weights(ws, float(t))
        assert np.abs(au.numpy() - g["alpha_unsup_%g" % t]).max() <= 1e-5
        assert np.abs(asup.numpy() - g["alpha_sup_%g" % t]).max() <= 1e-5
        assert np.abs(restated.weighted_embedding(au, Z) - g["X_unsup_%g" % t]).max() <= 1e-4
        # stabilised softmax == reference softmax wherever the reference is finite
        assert np.abs(restated.alpha_from_weights(wu, float(t), stable=True).numpy() - au.numpy()).max() <= 1e-12


def test_tau_zero_is_onehot_with_ties():
    w = torch.tensor([[1.0, 3.0, 3.0, 2.0]])
    a = restated.alpha_from_weights(w, 0.0)
    assert torch.allclose(a, torch.tensor([[0.0, 0.5, 0.5, 0.0]], dtype=torch.float64))


def test_reference_softmax_overflows_where_stable_does_not():
    w = torch.tensor([[800.0, 799.0]])
    assert torch.isnan(restated.alpha_from_weights(w, 1.0)).any()  # reference behaviour (utils.py:253)
    a = restated.alpha_from_weights(w, 1.0, stable=True)
    assert torch.isfinite(a).all() and abs(a.sum().item() - 1) < 1e-12


def test_patchify_pool_match_reference(golden_dir):
    g = load(golden_dir, "patchify_small")
    x = torch.from_numpy(g["x"])
    for k, s in [(3, 1), (5, 2), (1, 1)]:
        u, grid = restated.patchify(x, k, s)
        assert grid == list(g["grid_k%d_s%d" % (k, s)])
        assert np.array_equal(u.contiguous().numpy(), g["patch_k%d_s%d" % (k, s)])
    pre = restated.preprocessing_forward([torch.from_numpy(g["pre_in0"]), torch.from_numpy(g["pre_in1"])], 20)
    assert np.array_equal(pre.numpy(), g["pre_out"])
    assert np.array_equal(restated.aggregator_forward(pre, 13).numpy(), g["agg_out"])


def test_shipped_results_reproduce_published_metrics(golden_dir):
    """X -> Ward -> best_map -> NMI/ARI/F1 equals the reference's own tau_result.csv (TAU=2)."""
    g = load(golden_dir, "shipped_cluster_golden")
    for key in g["cases"]:
        key = str(key)
        nmi, ari, f1, _, _ = ocluster.metrics_from_X(g[key + "_Xlow"], [str(a) for a in g[key + "_anomaly"]])
        assert np.allclose([nmi, ari, f1], g[key + "_csv"], atol=1e-9), key
        assert np.allclose(g[key + "_alpha_rowsum"], 1.0, atol=1e-4)
        assert int(g[key + "_alpha_shape"][2]) == 784


def test_pairwise_matches_scipy():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(7, 33)).astype(np.float32)
    D = restated.pairwise_euclidean(X)
    ref = np.sqrt(((X[:, None, :].astype(np.float64) - X[None, :, :]) ** 2).sum(-1))
    assert np.allclose(D, ref, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Anomaly-Clustering"), reason="reference tree absent")
def test_oracle_against_live_reference():
    """Build container only: run the imported reference on fresh seeded inputs."""
    from oracle import ref_import

    gen = torch.Generator().manual_seed(11)
    feats = [torch.randn(1, 1 + 64, 48, generator=gen), torch.randn(1, 1 + 64, 48, generator=gen)]
    zr = ref_import.reference_embed(feats, 3, 1, 128, 256)
    zo = restated.embed(feats, 3, 1, 128, 256)
    assert (zr - zo).abs().max().item() <= 2e-6
    Z = torch.randn(4, 30, 32, generator=gen)
    ar = ref_import.reference_alpha_unsupervised(1.0, Z)
    ao = restated.matrix_alpha_unsupervised(1.0, Z)
    assert (ar - ao).abs().max().item() <= 1e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/Anomaly-Clustering"), reason="reference tree absent")
def test_oracle_fuzz_against_live_reference():
    """Build container only: 24 seeded random geometries (CNN pyramids with 1-3 layers of different grids, ViT
    token layers, patch size 1/3/5, stride 1/2, pooling up and down, aggregator windows that straddle layers)
    through the imported reference's _embed, and random-tau alpha in both modes, against the restatement."""
    from oracle import ref_import

    rng = np.random.default_rng(2023)
    gen = torch.Generator().manual_seed(2023)
    for case in range(24):
        k = int(rng.choice([1, 3, 3, 3, 5]))
        s = int(rng.choice([1, 1, 1, 2]))
        B = int(rng.integers(1, 3))
        if case % 3 == 0:      # ViT tokens: all layers share the grid
            g = int(rng.integers(4, 9))
            feats = [torch.randn(B, 1 + g * g, int(rng.integers(8, 40)), generator=gen) for _ in range(int(rng.integers(1, 4)))]
        else:                  # CNN pyramid: grid halves per layer
            g = int(rng.choice([8, 12, 16]))
            L = int(rng.integers(1, 4))
            feats = [torch.randn(B, int(rng.integers(6, 30)), max(g >> l, k), max(g >> l, k), generator=gen) for l in range(L)]
        Dp, D = int(rng.integers(5, 70)), int(rng.integers(5, 90))
        zr = ref_import.reference_embed(feats, k, s, Dp, D)
        zo = restated.embed(feats, k, s, Dp, D)
        assert zr.shape == zo.shape, (case, zr.shape, zo.shape)
        assert (zr - zo).abs().max().item() <= 5e-6, (case, k, s, [tuple(f.shape) for f in feats], Dp, D)
    for case in range(8):
        N, P, Dd = int(rng.integers(2, 6)), int(rng.integers(4, 30)), int(rng.integers(8, 40))
        Z = torch.randn(N, P, Dd, generator=gen) * float(rng.uniform(0.5, 3.0))
        Zt = torch.randn(int(rng.integers(1, 5)), P, Dd, generator=gen)
        tau = float(rng.choice([0.0, 0.3, 1.0, 2.5, 20.0]))
        assert (ref_import.reference_alpha_unsupervised(tau, Z) - restated.matrix_alpha_unsupervised(tau, Z)).abs().max().item() <= 1e-6
        assert (ref_import.reference_alpha_supervised(tau, Z, Zt) - restated.matrix_alpha_supervised(tau, Z, Zt)).abs().max().item() <= 1e-6
