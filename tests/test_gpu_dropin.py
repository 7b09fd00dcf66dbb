"""GPU: the reference's call sequence through the mirror classes (drop-in surface), end to end with a
random-init backbone whose forward stays in torch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from anomaly_clustering_b200 import backbones, driver, io  # noqa: E402
from anomaly_clustering_b200.patchcore import common, patchcore, utils  # noqa: E402
from oracle import restated  # noqa: E402


def _images(n, seed=0, size=224):
    gen = torch.Generator().manual_seed(seed)
    base = torch.randn(1, 3, size, size, generator=gen)
    x = base + 0.3 * torch.randn(n, 3, size, size, generator=gen)
    for i in range(n):
        if i % 2:
            x[i, :, 60:100, 80:140] += 2.0
    return x


def test_anomaly_clustering_core_wideresnet50_matches_oracle():
    """BASELINE config 1 pipeline: WRN50 layer2+layer3 hooks -> _embed (1024 -> 1024) -> Matrix_Alpha."""
    dev = torch.device("cuda")
    net = backbones.load("wideresnet50", allow_random_init=True)
    core = patchcore.AnomalyClusteringCore(dev).load(
        backbone=net, layers_to_extract_from=["layer2", "layer3"], device=dev, input_shape=(3, 224, 224),
        pretrain_embed_dimension=1024, target_embed_dimension=1024, patchsize=3)
    torch.backends.cudnn.allow_tf32 = False      # the backbone is torch: keep its features bit-stable across calls
    torch.backends.cuda.matmul.allow_tf32 = False
    imgs = _images(5)
    rows, shapes = core._embed(imgs[:2], "unsupervised", provide_patch_shapes=True)     # reference return convention
    assert isinstance(rows, list) and len(rows) == 2 * 784 and rows[0].shape == (1024,)
    assert shapes == [[28, 28], [14, 14]]
    feats2 = [f.float().cpu() for f in core._features(imgs[:2].to(dev))]
    assert feats2[0].shape == (2, 512, 28, 28) and feats2[1].shape == (2, 1024, 14, 14)
    assert np.abs(np.stack(rows) - restated.embed(feats2, 3, 1, 1024, 1024).numpy()).max() <= 5e-5
    # main.py:266-267: batch_size = 1 loop, list of per-patch numpy rows -> torch.tensor(Z)
    Zs, Zws = [], []
    for i in range(5):
        Zs.append(torch.from_numpy(np.stack(core._embed(imgs[i:i + 1], "unsupervised"))))
        Zws.append(restated.embed([f.float().cpu() for f in core._features(imgs[i:i + 1].to(dev))], 3, 1, 1024, 1024))
    Z = torch.stack(Zs)
    Zw = torch.cat(Zws)
    assert (Z.reshape(-1, 1024) - Zw).abs().max().item() <= 5e-5
    alpha = utils.Matrix_Alpha_Unsupervised(1.0, 1, Z, dev)
    assert alpha.dtype == torch.float64 and alpha.shape == (5, 784)
    want = restated.matrix_alpha_unsupervised(1.0, Zw.reshape(5, 784, 1024))
    assert (alpha.cpu() - want).abs().max().item() <= 1e-3
    w0 = utils.Weight_Distance_Unsupervised(Z, 0, dev)
    assert (w0.cpu() - restated.weight_distance_unsupervised(Zw.reshape(5, 784, 1024))[0]).abs().max().item() <= 5e-2
    a_sup = utils.Matrix_Alpha_Supervised(2.0, 1, Z[:3], Z[3:], dev)
    want_sup = restated.matrix_alpha_supervised(2.0, Zw.reshape(5, 784, 1024)[:3], Zw.reshape(5, 784, 1024)[3:])
    assert (a_sup.cpu() - want_sup).abs().max().item() <= 1e-3


def test_mirror_alpha_arbitrary_scale_embeddings_vs_oracle():
    """The mirror Matrix_Alpha_* accepts ANY fp32 Z like the reference's torch.cdist (utils.py:226, :233): embeddings far
    outside the fp16 range (x 1e6: fp16 operands would overflow; x 1e-6: they would flush to zero) take the exact fp32
    kernel with a warning and still match the oracle to the alpha tolerance; non-finite input is refused."""
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(11)
    Z1 = torch.randn(4, 64, 128, generator=gen)
    for scale in (1e6, 1e-6):
        Z = Z1 * scale
        with pytest.warns(RuntimeWarning, match="fp16 operand range"):
            a = utils.Matrix_Alpha_Unsupervised(scale, 1, Z, dev)
        want = restated.matrix_alpha_unsupervised(scale, Z)
        assert torch.isfinite(a).all() and (a.cpu() - want).abs().max().item() <= 1e-3
        with pytest.warns(RuntimeWarning, match="fp16 operand range"):
            a = utils.Matrix_Alpha_Supervised(2 * scale, 1, Z[:2], Z[2:], dev)
        assert (a.cpu() - restated.matrix_alpha_supervised(2 * scale, Z[:2], Z[2:])).abs().max().item() <= 1e-3
    bad = Z1.clone()
    bad[1, 2, 3] = float("nan")
    with pytest.raises(ValueError, match="non-finite"):
        utils.Matrix_Alpha_Unsupervised(1.0, 1, bad, dev)


def test_mirror_modules_standalone():
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 8, 10, 10, generator=gen)
    pm = patchcore.PatchMaker(3, stride=1)
    u, grid = pm.patchify(x.to(dev), return_spatial_info=True)
    uw, gw = restated.patchify(x, 3, 1)
    assert grid == gw and torch.equal(u.cpu(), uw.contiguous())
    feats = [torch.randn(50, 8, 3, 3, generator=gen), torch.randn(50, 16, 3, 3, generator=gen)]
    pre = common.Preprocessing([8, 16], 24)([f.to(dev) for f in feats])
    agg = common.Aggregator(target_dim=10)(pre)
    assert (pre.cpu() - restated.preprocessing_forward(feats, 24)).abs().max().item() <= 1e-6
    assert (agg.cpu() - restated.aggregator_forward(restated.preprocessing_forward(feats, 24), 10)).abs().max().item() <= 1e-6


def test_make_category_data_vit_writes_reference_pickle(tmp_path):
    """DINO ViT-S/8 shape (random init): driver == oracle, tau list from one pass, pickle layout."""
    dev = torch.device("cuda")
    net = backbones.load("dino_vitsmall8", allow_random_init=True)
    imgs = _images(4, seed=3)
    loader = [{"image": imgs[i:i + 1], "is_anomaly": torch.tensor([i % 2])} for i in range(4)]   # batch_size=1 like main.py:211
    layers = ["blocks.10", "blocks.11"]
    res = driver.make_category_data(None, "bottle", 512, 1024, ["dino_vitsmall8"], layers, 3, str(tmp_path), tau=[1.0, 2.0],
                                    supervised="unsupervised", test_dataloader=loader, backbone=net, device=dev, allow_random_init=True)
    assert len(res) == 2
    with pytest.raises(ValueError):      # a random-init stand-in is never accepted silently (nor is a missing backbone)
        driver.make_category_data(None, "bottle", 512, 1024, ["dino_vitsmall8"], layers, 3, None, test_dataloader=loader, backbone=net)
    with pytest.raises(ValueError):
        driver.make_category_data(None, "bottle", 512, 1024, ["dino_vitsmall8"], layers, 3, None, test_dataloader=loader)
    agg = common.NetworkFeatureAggregator(net, layers, dev)
    feats = [agg(imgs.to(dev))[l].float().cpu() for l in layers]
    assert feats[0].shape == (4, 785, 384)
    for (alpha, X), tau in zip(res, [1.0, 2.0]):
        _, _, a_w, X_w, _ = restated.full_path(feats, 3, 1, 512, 1024, tau, "unsupervised")
        assert alpha.shape == (4, 1, 784) and alpha.dtype == torch.float32
        assert (alpha.squeeze(1).cpu().double() - a_w).abs().max().item() <= 1e-3
        assert np.linalg.norm(X - X_w) / np.linalg.norm(X_w) <= 1e-3
        a_l, X_l = io.load_matrix_alpha_X(str(tmp_path / ("blocks.10_blocks.11_512_1024_%s_1.0" % float(tau))
                                              / "matrix_alpha_X_bottle_unsupervised.pickle"))
        assert np.array_equal(X_l, X) and torch.equal(a_l, alpha.cpu())


def test_cli_synthetic_run(tmp_path):
    """python -m anomaly_clustering_b200.main with the reference's argument names (WRN50 config-1 shape)."""
    import os

    from anomaly_clustering_b200 import main as cli

    rows = cli.main(["--dataset", "synthetic", "--backbone_names", "wideresnet50", "--layers_to_extract_from", "layer2", "layer3",
                     "--pretrain_embed_dimension", "1024", "--target_embed_dimension", "1024", "--tau", "1", "2",
                     "--synthetic_images", "8", "--synthetic_classes", "2", "--output_dir", str(tmp_path)])
    assert len(rows) == 2 and all(0.0 <= r[2] <= 1.0 for r in rows)
    p = (tmp_path / "synthetic" / "wideresnet50-randinit" / "unsupervised" / "layer2_layer3_1024_1024_2.0_1.0"
         / "matrix_alpha_X_bottle_unsupervised.pickle")
    assert os.path.exists(p)
    # each per-tau pickle holds only its own alpha (a view of the [T,N,P] tensor would serialise every tau's storage)
    a, X = io.load_matrix_alpha_X(str(p))
    assert a.untyped_storage().nbytes() == a.numel() * 4 and not a.is_cuda


def test_embed_over_dataloader_reference_convention():
    """AnomalyClusteringCore.embed(DataLoader, supervised) -> (list of per-batch row lists, list of is_anomaly)
    exactly as examples/main.py:266-267 consumes it: torch.tensor(Z) -> [N, P, D]."""
    dev = torch.device("cuda")
    net = backbones.load("dino_vitsmall8", allow_random_init=True)
    core = patchcore.AnomalyClusteringCore(dev).load(net, ["blocks.10", "blocks.11"], dev, (3, 224, 224), 256, 512, patchsize=3)
    imgs = _images(3, seed=5)

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 3

        def __getitem__(self, i):
            return {"image": imgs[i], "is_anomaly": i % 2}

    loader = torch.utils.data.DataLoader(DS(), batch_size=1, shuffle=False)
    Z, labels = core.embed(loader, "unsupervised")
    assert len(Z) == 3 and len(Z[0]) == 784 and [int(l) for l in labels] == [0, 1, 0]
    Zt = torch.tensor(np.array(Z))
    assert Zt.shape == (3, 784, 512)
    direct = core.embed(imgs[1:2], "unsupervised")            # non-loader input goes straight to _embed
    assert np.abs(np.stack(direct) - Zt[1].numpy()).max() <= 5e-5
    assert patchcore.PatchMaker(3, 1).score(torch.tensor([[1.0, 5.0], [3.0, 2.0]])).tolist() == [5.0, 3.0]


def test_plain_c_consumer_of_the_abi(tmp_path):
    """tests/c/abi_smoke.c: a C program (no Python, no torch) links libac_b200.so through include/ac_b200.h."""
    import os
    import shutil
    import subprocess

    from anomaly_clustering_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no C toolchain / CUDA headers on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([gcc, os.path.join(root, "tests", "c", "abi_smoke.c"), "-I" + os.path.join(root, "include"),
                    "-I/usr/local/cuda/include", "-L" + pkg, "-lac_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-lm",
                    "-Wl,-rpath," + pkg, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max abs error" in r.stdout
