"""GPU parity AT THE SIZES BASELINE.json STATES: the CUDA path (through the C ABI) against the CPU oracle
(oracle/restated.py, the pinned restatement of the reference's torch path) on the same synthetic inputs.

    config 1  WRN50 layer2+layer3 shape, 20 images, 1024 -> 1024, unsupervised tau = 1          (full)
    config 2  ViT-B/8 shape, 100 images x 784 patches x 4096-d, unsupervised                     (full, ~40 s of oracle)
    config 3  config-2 queries against a 200-image normal bank, supervised + average             (full GPU run; the
              oracle evaluates a subsample of the query images against the FULL bank -- rows are independent)
    config 4  1,210 images in 10 MVTec-object-sized categories: per-category banks in one batched launch (one category
              in full + sampled rows of another against the oracle) and the joint 1,210-image bank (sampled image pairs
              against the oracle's cdist + a float64 reduction of every row)
    config 5  ViT-S/8 at 448x448 shape (3136 patches), tau sweep incl. tau = 0.1                  (5 images)
    stress    config-2 shape with real-data patch norms (45-50) and precision="auto"

Tolerances are the north_star's: alpha max-abs <= 1e-3, X and Dmat <= 1e-3 relative L2, identical Ward labels and
NMI/ARI at the ground-truth k.  The oracle's cost is a few minutes of host CPU in total."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

from anomaly_clustering_b200 import pipeline, synth  # noqa: E402
from oracle import cluster as ocluster  # noqa: E402
from oracle import restated  # noqa: E402

VITB = [(768, 28, 28, True), (768, 28, 28, True)]
VITS448 = [(384, 56, 56, True), (384, 56, 56, True)]
WRN = [(512, 28, 28, False), (1024, 14, 14, False)]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def oracle_embed(feats_cpu, Dp, D, chunk=10):
    """restated.embed in image chunks (the unfolded tensor of 100 ViT-B images would be 2 GB per layer)."""
    n = feats_cpu[0].shape[0]
    out = [restated.embed([f[a:a + chunk] for f in feats_cpu], 3, 1, Dp, D) for a in range(0, n, chunk)]
    Z = torch.cat(out, dim=0)
    return Z.reshape(n, -1, D)


def oracle_stage3(w, Z, taus):
    """alpha (fp64), X, Dmat per tau from the oracle's w and Z (utils.py:246-255, main.py:294-296, test.py:193-195).
    The reference's softmax has no max-subtraction: exp(w / tau) overflows to inf -> NaN rows once w / tau > 709
    (defect patches reach w ~ 97 here, so tau = 0.1 overflows).  The comparison uses the mathematically identical
    max-subtracted form and checks that it EQUALS the literal form on every row where the literal form is finite."""
    res = []
    for t in taus:
        a = restated.alpha_from_weights(w, t, stable=True)
        lit = restated.alpha_from_weights(w, t)
        fin = torch.isfinite(lit).all(dim=1)
        assert fin.any() or t < 0.2
        assert (a[fin] - lit[fin]).abs().max().item() <= 1e-12 if fin.any() else True
        X = restated.weighted_embedding(a, Z)
        res.append((a, X, restated.pairwise_euclidean(X)))
    return res


def check_tau(res, ti, want, labels=None, k=None, alpha_tol=1e-3):
    a, X, Dm = want
    assert torch.isfinite(a).all()
    da = (res.alpha64[ti].cpu() - a).abs().max().item()
    assert da <= alpha_tol, (res.taus[ti], da)                             # north_star: alpha max-abs <= 1e-3
    assert rel_l2(res.X[ti].cpu().numpy(), X) <= 1e-3                      # X within 1e-3 relative L2
    assert rel_l2(res.Dmat[ti].cpu().numpy(), Dm) <= 1e-3                  # Dmat within 1e-3 relative L2
    if labels is not None:
        from sklearn import metrics

        lab_ref = ocluster.ward_labels(X, k)
        lab_gpu = ocluster.ward_labels(res.X[ti].cpu().numpy().astype(np.float64), k)
        assert metrics.adjusted_rand_score(lab_ref, lab_gpu) == 1.0        # identical Ward partition
        m_ref = [metrics.normalized_mutual_info_score(labels, lab_ref), metrics.adjusted_rand_score(labels, lab_ref)]
        m_gpu = [metrics.normalized_mutual_info_score(labels, lab_gpu), metrics.adjusted_rand_score(labels, lab_gpu)]
        assert m_ref == m_gpu
    return da


# ------------------------------------------------------------------------------------------------ config 2
@pytest.fixture(scope="module")
def config2():
    """100 config-2 images on the device + the oracle's Z and w for them (the 9,900-pair cdist loop, ~40 s on 16 cores)."""
    feats, labels = synth.planted_features_device(range(100), VITB, device="cuda")
    Z = oracle_embed([f.cpu() for f in feats], 2048, 4096)
    w = restated.weight_distance_unsupervised(Z)
    return feats, labels.numpy(), Z, w


def test_config2_full_size_vs_oracle(config2):
    """BASELINE config 2 at full size, precision='auto', tau in {0.5, 1, 2}: every output against the oracle."""
    feats, labels, Z, w = config2
    taus = [0.5, 1.0, 2.0]
    res = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", taus, precision="auto")
    assert (res.Z.cpu() - Z).abs().max().item() <= 2e-5
    assert ((res.w.cpu() - w).abs() / w).max().item() <= 2e-4
    want = oracle_stage3(w, Z, taus)
    for ti in range(len(taus)):
        check_tau(res, ti, want[ti], labels, 4)


def test_config2_full_size_small_tau_and_argmax_vs_oracle(config2):
    """Same inputs, the taus 'auto' sends to the high-precision mode (0.1, 0.25) and tau = 0 (one-hot arg-max)."""
    feats, labels, Z, w = config2
    taus = [0.1, 0.25, 0.0]
    res = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", taus, precision="auto", keep_z=True)
    want = oracle_stage3(w, Z, taus)
    check_tau(res, 0, want[0], labels, 4)
    check_tau(res, 1, want[1], labels, 4)
    # tau = 0: alpha is the indicator of the arg-max patch (utils.py:248-250) -- identical positions, not a tolerance
    assert torch.equal(res.alpha64[2].cpu().argmax(dim=1), want[2][0].argmax(dim=1))
    check_tau(res, 2, want[2], labels, 4, alpha_tol=0.0)


def test_config2_all_pairs_kernel_and_z_free_vs_oracle(config2):
    """The straightforward all-pairs kernel and the Z-free form (operands only, X from the feature maps) give the
    same answers as the oracle; symmetric and all-pairs runs are each bit-reproducible."""
    feats, labels, Z, w = config2
    want = oracle_stage3(w, Z, [1.0])
    try:
        pipeline.SYMMETRIC = False
        r_full = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", [1.0], precision="auto")
        r_full2 = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", [1.0], precision="auto")
    finally:
        pipeline.SYMMETRIC = True
    assert torch.equal(r_full.w, r_full2.w) and torch.equal(r_full.Dmat, r_full2.Dmat)
    check_tau(r_full, 0, want[0], labels, 4)
    r_zf = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", [1.0], precision="auto", keep_z=False)
    assert r_zf.Z is None
    check_tau(r_zf, 0, want[0], labels, 4)
    r_zf2 = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", [1.0], precision="auto", keep_z=False)
    assert torch.equal(r_zf.w, r_zf2.w) and torch.equal(r_zf.X, r_zf2.X)


def test_config2_full_size_properties(config2):
    """Size-independent properties at full size: alpha rows sum to 1, Dmat is a symmetric zero-diagonal metric,
    a permutation of the images permutes the outputs, the planted classes are recovered."""
    feats, labels, _, _ = config2
    r1 = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", [1.0, 2.0])
    assert (r1.alpha64.sum(dim=2) - 1).abs().max().item() <= 1e-12
    D = r1.Dmat[0]
    assert torch.equal(D, D.T) and (D.diagonal() == 0).all() and (D >= 0).all()
    assert (D[:, None, :] <= D[:, :, None] + D[None, :, :] + 1e-3).all()        # triangle inequality
    perm = torch.randperm(100, generator=torch.Generator().manual_seed(0)).cuda()
    r3 = pipeline.run_path([f[perm] for f in feats], 3, 1, 2048, 4096, "unsupervised", [1.0])
    assert ((r3.w - r1.w[perm]).abs() / r1.w[perm]).max().item() <= 2e-4
    assert (r3.X[0] - r1.X[0][perm]).norm().item() / r1.X[0].norm().item() <= 1e-3
    from anomaly_clustering_b200 import cluster

    nmi, ari, f1, _, _ = cluster.calculate_metrics(D.cpu().numpy(), [str(int(c)) for c in labels])
    assert nmi > 0.9


# ------------------------------------------------------------------------------------------------ stress
def test_auto_precision_with_real_data_norms_vs_oracle():
    """real DINO patch norms are 35-46 and the expansion's cancellation error grows with the norm.
    Config-2 shape with a per-channel offset (patch norm ~46, nearest-neighbour distance ~21), 40 images:
    precision='auto' must keep alpha inside the tolerance with a 2x margin at every tau it serves with one
    fp16 pass, and inside the tolerance everywhere."""
    n = 40
    feats, labels = synth.planted_features_device(range(n), VITB, device="cuda", channel_bias=1.45)
    Z = oracle_embed([f.cpu() for f in feats], 2048, 4096)
    norms = Z.norm(dim=2).mean().item()
    assert 44.0 <= norms <= 52.0, norms
    w = restated.weight_distance_unsupervised(Z)
    taus = [0.1, 0.5, 1.0, 2.0]
    res = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", taus, precision="auto")
    want = oracle_stage3(w, Z, taus)
    for ti in range(len(taus)):
        check_tau(res, ti, want[ti], labels.numpy(), 4)
    # per-tau resolution: the cheap mode only where it has a 2x margin
    for t in (1.0, 2.0, 5.0):
        mode = pipeline.resolve_precision("auto", [t])
        r = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", [t], precision=mode)
        a = restated.alpha_from_weights(w, t, stable=True)
        da = (r.alpha64[0].cpu() - a).abs().max().item()
        assert da <= 5e-4, (t, mode, da)


# ------------------------------------------------------------------------------------------------ config 1
def test_config1_full_size_vs_oracle():
    """BASELINE config 1 in full: WRN50 layer2+layer3 shape, 20 images, 1024 -> 1024, unsupervised tau = 1."""
    feats, labels = synth.planted_features(20, WRN, n_classes=4, seed=2023)
    want = restated.full_path(feats, 3, 1, 1024, 1024, 1.0, "unsupervised")
    res = pipeline.run_path([f.cuda() for f in feats], 3, 1, 1024, 1024, "unsupervised", [1.0], precision="auto")
    assert (res.Z.cpu() - want[0]).abs().max().item() <= 2e-5
    assert ((res.w.cpu() - want[1]).abs() / want[1]).max().item() <= 2e-4
    check_tau(res, 0, (want[2], want[3], want[4]), labels.numpy(), 4)


# ------------------------------------------------------------------------------------------------ config 3
def test_config3_full_size_query_subsample_vs_oracle():
    """BASELINE config 3 at the stated 100 queries x 200-image normal bank (supervised, utils.py:230-237, 260-277)
    + the average mode.  The GPU runs all 100 queries; the oracle embeds the FULL bank and evaluates 5 query images
    (w_i, alpha_i, X_i depend on query image i and the bank only)."""
    nq, nb = 100, 200
    feats, labels = synth.planted_features_device(range(nq), VITB, device="cuda")
    bank, _ = synth.planted_features_device(range(1000, 1000 + nb), VITB, n_classes=1, device="cuda")
    taus = [1.0, 2.0, 0.5]
    res = pipeline.run_path(feats, 3, 1, 2048, 4096, "supervised", taus, bank_features=bank, precision="auto")
    sample = [0, 17, 42, 63, 99]
    Zq = oracle_embed([f[sample].cpu() for f in feats], 2048, 4096)
    Zb = oracle_embed([f.cpu() for f in bank], 2048, 4096)
    w = restated.weight_distance_supervised(Zq, Zb)
    got_w = res.w[sample].cpu()
    assert ((got_w - w).abs() / w).max().item() <= 5e-4
    for ti, t in enumerate(taus):
        a = restated.alpha_from_weights(w, t, stable=True)
        assert (res.alpha64[ti][sample].cpu() - a).abs().max().item() <= 1e-3
        X = restated.weighted_embedding(a, Zq)
        assert rel_l2(res.X[ti][sample].cpu().numpy(), X) <= 1e-3
        Dm = restated.pairwise_euclidean(X)
        got_D = res.Dmat[ti][sample][:, sample].cpu().numpy()
        assert rel_l2(got_D, Dm) <= 1e-3
    # average mode (main.py:290-291) on the same queries
    res_avg = pipeline.run_path(feats, 3, 1, 2048, 4096, "average")
    a = restated.matrix_alpha_average(len(sample), 784)
    assert rel_l2(res_avg.X[0][sample].cpu().numpy(), restated.weighted_embedding(a, Zq)) <= 1e-4
    # the planted classes separate in the supervised distance matrix of all 100 queries
    from anomaly_clustering_b200 import cluster

    nmi, _, _, _, _ = cluster.calculate_metrics(res.Dmat[0].cpu().numpy(), [str(int(c)) for c in labels])
    assert nmi > 0.9


# ------------------------------------------------------------------------------------------------ config 5
def test_config5_geometry_tau_sweep_vs_oracle():
    """BASELINE config 5 geometry at reduced image count: ViT-S/8 tokens at 448x448 (3136 patches; 13 column tiles
    per bank image), the six taus of the reference's sweep INCLUDING 0.1 from one distance pass, against the oracle."""
    n = 5
    feats, labels = synth.planted_features_device(range(n), VITS448, device="cuda")
    taus = [0.1, 0.5, 1.0, 2.0, 5.0, 10.0]
    res = pipeline.run_path(feats, 3, 1, 2048, 4096, "unsupervised", taus, precision="auto")
    assert res.w.shape == (n, 3136)
    Z = oracle_embed([f.cpu() for f in feats], 2048, 4096, chunk=1)
    assert (res.Z.cpu() - Z).abs().max().item() <= 2e-5
    w = restated.weight_distance_unsupervised(Z)
    assert ((res.w.cpu() - w).abs() / w).max().item() <= 2e-4
    want = oracle_stage3(w, Z, taus)
    for ti in range(len(taus)):
        check_tau(res, ti, want[ti])
    assert (res.alpha64.sum(dim=2) - 1).abs().max().item() <= 1e-12


# ------------------------------------------------------------------------------------------------ config 4
MVTEC_OBJECT_SIZES = [83, 150, 132, 110, 115, 167, 160, 42, 100, 151]


@pytest.fixture(scope="module")
def config4():
    """The 1,210 config-4 images (10 MVTec-object-sized categories back to back, bench.py's ids) on the device."""
    ids = [1000 * c + i for c, n in enumerate(MVTEC_OBJECT_SIZES) for i in range(n)]
    feats, _ = synth.planted_features_device(ids, VITB, device="cuda")
    return feats


def test_config4_per_category_full_size_vs_oracle(config4):
    """BASELINE config 4 with the reference's semantics (one make_category_data per category, examples/main.py:353): all 10
    categories, 1,210 images, in ONE batched launch sequence (category table).  The oracle runs the smallest category
    (42 images) on its own, in full: every w / alpha / X row and the category's distance matrix; a second category is
    checked through three of its images against that category's full bank."""
    feats = config4
    sizes = MVTEC_OBJECT_SIZES
    res = pipeline.run_categories(feats, sizes, 3, 1, 2048, 4096, [1.0], precision="auto", keep_z=False)
    assert len(res) == len(sizes) and all(r.w.shape == (n, 784) for r, n in zip(res, sizes))
    starts = np.cumsum([0] + sizes)
    c = 7
    sl = slice(int(starts[c]), int(starts[c + 1]))
    Z = oracle_embed([f[sl].cpu() for f in feats], 2048, 4096)
    w = restated.weight_distance_unsupervised(Z)
    assert ((res[c].w.cpu() - w).abs() / w).max().item() <= 2e-4
    check_tau(res[c], 0, oracle_stage3(w, Z, [1.0])[0])
    c = 0
    sl = slice(int(starts[c]), int(starts[c + 1]))
    Zc = oracle_embed([f[sl].cpu() for f in feats], 2048, 4096)
    sample = [0, 40, 82]
    dm = restated.per_image_min_dist(Zc[sample], Zc)                      # [3, 784, 83]
    for k, i in enumerate(sample):
        keep = torch.ones(sizes[c], dtype=torch.bool)
        keep[i] = False
        wi = dm[k][:, keep].mean(dim=1)
        assert ((res[c].w[i].cpu() - wi).abs() / wi).max().item() <= 2e-4
        a = restated.alpha_from_weights(wi[None], 1.0, stable=True)
        assert (res[c].alpha64[0][i].cpu() - a[0]).abs().max().item() <= 1e-3
        assert rel_l2(res[c].X[0][i].cpu().numpy(), restated.weighted_embedding(a, Zc[i:i + 1])[0]) <= 1e-3


def test_config4_joint_bank_full_size_sampled_pairs_vs_oracle(config4):
    """BASELINE config 4 as ONE joint bank (every image against the other 1,209; 948,640 patch rows, 7.4 PFLOP of
    algorithmic distance work) through the C ABI at full size.  The oracle cannot run 1.46 M image pairs, but the path
    factorises: (i) the per-pair minima d(r, j) = min_c |z_r - z_(j,c)| (utils.py:226) are checked against the oracle's
    cdist for sampled image pairs spread over the whole raster -- both orientations of a pair, i.e. the row-min and the
    column-min side of the symmetric kernel -- and (ii) w = mean over the 1,209 other images is checked for EVERY row
    against a float64 reduction of those minima."""
    from anomaly_clustering_b200 import ops

    feats = config4
    n, P, D = sum(MVTEC_OBJECT_SIZES), 784, 4096
    q = pipeline.embed_images(feats, 3, 1, 2048, D, "f16", want_z=False)
    rowmin, colmin = ops.min_dist_sym(q.hi, None, q.n2, 0, q.hi, None, q.n2, n, P, "f16")
    w = ops.reduce_weights_sym(rowmin, colmin, P, 0).reshape(n, P)
    torch.cuda.synchronize()
    assert torch.isfinite(w).all()
    gen = torch.Generator().manual_seed(4)
    pairs = [(0, 1), (0, 605), (1209, 0), (604, 1209), (83, 82), (700, 95)] + [tuple(int(v) for v in torch.randint(0, n, (2,), generator=gen)) for _ in range(10)]
    pairs = [(i, j) for i, j in pairs if i != j]
    imgs = sorted({i for pr in pairs for i in pr})
    Zs = oracle_embed([f[imgs].cpu() for f in feats], 2048, D, chunk=8)
    pos = {g: k for k, g in enumerate(imgs)}
    from anomaly_clustering_b200.distributed import pair_owned

    def d_gpu(i, j):          # query image i, bank image j
        # both minima are indexed [other image, own patch row]: the owner j's tile wrote colmin[j, i*P + c] for the patches of i
        d2 = (rowmin if pair_owned(i, j, n) else colmin)[j, i * P:(i + 1) * P]
        return d2.sqrt().cpu()

    for i, j in pairs:
        for a, b in ((i, j), (j, i)):
            want = torch.cdist(Zs[pos[a]], Zs[pos[b]]).min(dim=1)[0]
            got = d_gpu(a, b)
            assert ((got - want).abs() / want).max().item() <= 3e-4, (a, b)
    # (ii) the mean over the other images, every row, in float64 on the device
    ii, jj = torch.arange(n, device="cuda")[:, None], torch.arange(n, device="cuda")[None, :]
    dd = (jj - ii) % n
    own = (dd != 0) & ((2 * dd < n) | ((2 * dd == n) & (ii < jj)))                                     # [i, j], the kernel's rule
    assert all(bool(own[i, j]) == pair_owned(i, j, n) for i, j in pairs)
    rm = rowmin.reshape(n, n, P).permute(1, 0, 2)            # [i, j, P]: rowmin[j, i*P + p]
    cm = colmin.reshape(n, n, P).permute(1, 0, 2)            # [i, j, P]: colmin[j, i*P + p], written by the owner j
    acc = torch.zeros(n, P, dtype=torch.float64, device="cuda")
    for j0 in range(0, n, 64):                               # chunks of bank images keep the temporaries small
        j1 = min(n, j0 + 64)
        d2 = torch.where(own[:, j0:j1, None], rm[:, j0:j1], cm[:, j0:j1]).double().sqrt()
        eye = (torch.arange(n, device="cuda")[:, None] == torch.arange(j0, j1, device="cuda")[None, :])
        acc += torch.where(eye[:, :, None], torch.zeros((), dtype=torch.float64, device="cuda"), d2).sum(dim=1)
    w64 = acc / (n - 1)
    assert ((w.double() - w64).abs() / w64).max().item() <= 2e-6
