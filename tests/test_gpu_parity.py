"""GPU parity: the CUDA path (through the C ABI) against the oracle and the committed reference
outputs.  Tolerances are the north_star's: alpha max-abs <= 1e-3, X and Dmat <= 1e-3 relative L2,
identical Ward labels; kernel-level checks are much tighter and say so."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from anomaly_clustering_b200 import ops, pipeline, synth  # noqa: E402
from oracle import cluster as ocluster  # noqa: E402
from oracle import restated  # noqa: E402

EMBED_CASES = ["embed_vit_small", "embed_wrn_small", "embed_ragged", "embed_k5s2", "embed_single"]


def gload(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_device_is_b200():
    assert ops.device_ok()


# ------------------------------------------------------------------------------- stage 1
@pytest.mark.parametrize("name", EMBED_CASES)
def test_embed_golden(golden_dir, name):
    g = gload(golden_dir, name)
    feats = [torch.from_numpy(g["feat%d" % i]).cuda() for i in range(int(g["L"]))]
    Z, hi, lo, grid = ops.embed(feats, int(g["patchsize"]), int(g["stride"]), int(g["Dp"]), int(g["D"]), operand="f16",
                                want_lo=True)
    err = np.abs(Z.cpu().numpy() - g["Z"]).max()
    assert err <= 1e-5, err   # fp32 kernel vs the reference's fp32 torch ops: summation-order noise only
    # operands: hi = fp16(Z), lo = fp16(Z - hi)
    assert torch.equal(hi, Z.half())
    assert torch.equal(lo, (Z - hi.float()).half())


@pytest.mark.parametrize("layernorm", [True, False])
def test_embed_config2_shape_vs_oracle(layernorm):
    """DINO ViT-B/8 shape: two [B,785,768] token tensors, 2048 -> 4096 (BASELINE config 2)."""
    feats, _ = synth.planted_features(3, [(768, 28, 28, True), (768, 28, 28, True)], seed=5)
    want = restated.embed(feats, 3, 1, 2048, 4096, layernorm=layernorm)
    Z, hi, _, grid = ops.embed([f.cuda() for f in feats], 3, 1, 2048, 4096, layernorm=layernorm, operand="bf16")
    assert grid == (28, 28) and Z.shape == (3 * 784, 4096)
    err = (Z.cpu() - want).abs().max().item()
    assert err <= 2e-5, err
    assert torch.equal(hi, Z.bfloat16())


def test_embed_config1_shape_vs_oracle():
    """WideResNet50 layer2+layer3 shape, 1024 -> 1024, incl. the 14->28 bilinear step (config 1)."""
    feats, _ = synth.planted_features(2, [(512, 28, 28, False), (1024, 14, 14, False)], seed=6)
    want = restated.embed(feats, 3, 1, 1024, 1024)
    Z, _, _, grid = ops.embed([f.cuda() for f in feats], 3, 1, 1024, 1024)
    assert grid == (28, 28)
    err = (Z.cpu() - want).abs().max().item()
    assert err <= 2e-5, err


@pytest.mark.parametrize("C,Dp,D,L", [(64, 128, 128, 2), (32, 288, 288, 1), (64, 144, 288, 2), (96, 256, 256, 2), (48, 256, 512, 2),
                                      (128, 128, 128, 1), (64, 64, 64, 2), (48, 64, 64, 1), (64, 128, 128, 1), (64, 256, 256, 1),
                                      (64, 32, 32, 1), (64, 256, 256, 2), (48, 256, 256, 2)])
def test_embed_periodic_fast_path_variants(C, Dp, D, L):
    """channels_last maps take the register sliding-window kernel: every instantiated (A,B,R) period
    pattern (9C/Dp in lowest terms, Aggregator ratio R) against the oracle."""
    gen = torch.Generator().manual_seed(C + Dp)
    feats = [torch.randn(2, C, 9, 13, generator=gen) for _ in range(L)]
    want = restated.embed(feats, 3, 1, Dp, D)
    cl = [f.cuda().contiguous(memory_format=torch.channels_last) for f in feats]
    Z, hi, lo, _ = ops.embed(cl, 3, 1, Dp, D, operand="bf16", want_lo=True)
    assert (Z.cpu() - want).abs().max().item() <= 1e-5
    assert torch.equal(hi, Z.bfloat16()) and torch.equal(lo, (Z - hi.float()).bfloat16())


def test_embed_kernel_variants_agree():
    """TMA-staged, LDG-periodic and generic tap kernels compute the same Z (config-2 geometry)."""
    from anomaly_clustering_b200 import _lib

    lib = _lib.load()
    feats, _ = synth.planted_features(2, [(768, 28, 28, True), (768, 28, 28, True)], seed=9)
    f = [x.cuda() for x in feats]
    outs = []
    try:
        for variant in (0, 1, 2, 3):     # 0 = single-pass fused form, 1 = LDG periodic, 2 = generic taps, 3 = per-layer TMA launches
            assert lib.ac_debug_set(2, variant) == 0
            Z, hi, _, _ = ops.embed(f, 3, 1, 2048, 4096, operand="f16")
            outs.append((Z.clone(), hi.clone()))
    finally:
        lib.ac_debug_set(2, 0)
    for k in (1, 2, 3):                  # summation order / where the LayerNorm affine is applied differ: fp32 round-off only
        assert (outs[0][0] - outs[k][0]).abs().max().item() <= 3e-6, k
    assert (outs[0][1].float() - outs[1][1].float()).abs().max().item() <= 4e-3


@pytest.mark.parametrize("layers,Dp,D,want_lo", [([(768, 28, 28, True), (768, 28, 28, True)], 2048, 4096, False),
                                                 ([(384, 20, 20, True), (384, 20, 20, True)], 2048, 4096, True),
                                                 ([(96, 12, 12, True), (96, 12, 12, True)], 256, 512, True),
                                                 ([(1024, 9, 13, True)], 1024, 1024, False)])
def test_embed_fused_single_pass_norms_and_batches(layers, Dp, D, want_lo):
    """The single-pass fused embed (statistics slices run ahead of the embedding inside one persistent launch) against the
    oracle for batch sizes around its look-ahead, operands = round(Z), and the operand norms it emits (ac_embed_ex)
    against ac_row_norms of the operands it wrote; repeated launches are bit-identical."""
    if layers[0][1] != layers[0][2]:
        layers = [(c, 12, 12, t) for c, _, _, t in layers]
    for n in (1, 2, 5):
        feats, _ = synth.planted_features(n, layers, seed=40 + n)
        f = [x.cuda() for x in feats]
        want = restated.embed(feats, 3, 1, Dp, D)
        P = layers[0][1] * layers[0][2]
        n2 = torch.empty(n * P, dtype=torch.float32, device="cuda")
        Z, hi, lo, _ = ops.embed(f, 3, 1, Dp, D, operand="f16", want_lo=want_lo, out_n2=n2)
        assert (Z.cpu() - want).abs().max().item() <= 1e-5
        assert torch.equal(hi, Z.half())
        if want_lo:
            assert torch.equal(lo, (Z - hi.float()).half())
        ref = ops.row_norms(hi, lo)
        assert ((n2 - ref).abs() / ref).max().item() <= 2e-6
        n2b = torch.empty_like(n2)
        Zb, hib, _, _ = ops.embed(f, 3, 1, Dp, D, operand="f16", want_lo=want_lo, out_n2=n2b)
        assert torch.equal(Z, Zb) and torch.equal(hi, hib) and torch.equal(n2, n2b)
        # operands only (Z-free) and no LayerNorm (PatchCore._embed)
        _, hi2, _, _ = ops.embed(f, 3, 1, Dp, D, want_z=False, operand="f16")
        # (without norms the general fused kernel runs; with them possibly the lean one, whose edge positions round differently)
        assert torch.equal(hi2, hi) or ((hi2.float() - hi.float()).abs().max().item() <= 4e-3 and (hi2 != hi).float().mean().item() < 1e-3)
        Zn, _, _, _ = ops.embed(f, 3, 1, Dp, D, layernorm=False)
        assert (Zn.cpu() - restated.embed(feats, 3, 1, Dp, D, layernorm=False)).abs().max().item() <= 1e-5


@pytest.mark.parametrize("layers,Dp,D", [([(768, 28, 28, True), (768, 28, 28, True)], 2048, 4096),     # ViT-B, config 2
                                         ([(768, 17, 17, True), (768, 17, 17, True)], 2048, 4096),     # ragged x segments (9 + 8)
                                         ([(384, 20, 20, True), (384, 20, 20, True)], 2048, 4096),     # ViT-S, config 5 (64 consumers)
                                         ([(768, 12, 12, True)], 1024, 1024)])                          # one layer, 27 : 4
@pytest.mark.parametrize("want_z", [False, True])
def test_embed_lean_fused_kernel_vs_general_and_oracle(layers, Dp, D, want_z):
    """embed_fused_fast_kernel (fp16 operands + norms, taps outside the map staged as zeros and corrected on edge positions)
    against the general fused kernel and the oracle: Z, operands = round(Z), norms of the rounded operands; sliced token
    tensors (strided images) included."""
    from anomaly_clustering_b200 import _lib

    lib = _lib.load()
    n = 3
    feats, _ = synth.planted_features(n + 1, layers, seed=77)
    feats = [x + 2.5 for x in feats]                       # a non-zero map mean: the edge correction is mu * rstd * (missing taps)
    f = [x.cuda()[1:] for x in feats]                      # batch slice: base pointer not at the allocation start
    want = restated.embed([x[1:] for x in feats], 3, 1, Dp, D)
    P = layers[0][1] * layers[0][2]
    res = []
    try:
        for lean in (1, 0):
            assert lib.ac_debug_set(9, lean) == 0
            n2 = torch.empty(n * P, dtype=torch.float32, device="cuda")
            Z, hi, _, _ = ops.embed(f, 3, 1, Dp, D, want_z=want_z, operand="f16", out_n2=n2)
            res.append((None if Z is None else Z.clone(), hi.clone(), n2.clone()))
    finally:
        lib.ac_debug_set(9, 1)
    (Zl, hil, n2l), (Zg, hig, n2g) = res
    ref_n2 = ops.row_norms(hil, None)
    assert ((n2l - ref_n2).abs() / ref_n2).max().item() <= 2e-6
    assert bool(((hil.float().cpu() - want).abs() <= want.abs() * 2.0 ** -11 * 1.02 + 2e-5).all())   # one fp16 rounding of Z
    # edge positions sum in a different order: a rounding boundary of the fp16 operand is crossed now and then
    # (windows that lie entirely outside the map are exactly 0 here, like the reference's zero padding, and fma(mu, rstd, -mu*rstd)
    # = +-1 fp16 subnormal in the general kernel: not counted)
    dh = (hil.float() - hig.float()).abs()
    assert (dh > 1e-6).float().mean().item() < 1e-2 and dh.max().item() <= 4e-3
    assert ((n2l - n2g).abs() / n2g).max().item() <= 1e-4
    if want_z:
        assert (Zl.cpu() - want).abs().max().item() <= 1e-5
        assert (Zl - Zg).abs().max().item() <= 3e-6
        assert torch.equal(hil, Zl.half())


def test_embed_reads_strided_views_in_place():
    """Non-contiguous inputs (channels-last maps, sliced batches) are read through their strides."""
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(4, 20, 9, 11, generator=gen)
    want = restated.embed([x[1:3]], 3, 1, 64, 64)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    Z, _, _, _ = ops.embed([xc[1:3]], 3, 1, 64, 64)
    assert (Z.cpu() - want).abs().max().item() <= 1e-5


def test_embed_constant_image_is_finite():
    """zero-variance map: LayerNorm divides by sqrt(eps), no NaN (same as torch)."""
    x = torch.ones(1, 8, 6, 6)
    want = restated.embed([x], 3, 1, 16, 16)
    Z, _, _, _ = ops.embed([x.cuda()], 3, 1, 16, 16)
    assert torch.isfinite(Z).all()
    assert (Z.cpu() - want).abs().max().item() <= 1e-5


def test_patchify_and_pools_golden(golden_dir):
    g = gload(golden_dir, "patchify_small")
    x = torch.from_numpy(g["x"]).cuda()
    for k, s in [(3, 1), (5, 2), (1, 1)]:
        u, grid = ops.patchify(x, k, s)
        assert grid == list(g["grid_k%d_s%d" % (k, s)])
        assert np.array_equal(u.cpu().numpy(), g["patch_k%d_s%d" % (k, s)])  # pure data movement: bit exact
    for key, dim in (("pre_in0", 20), ("pre_in1", 20)):
        got = ops.adaptive_pool1d(torch.from_numpy(g[key]).cuda(), dim).cpu()
        want = restated.preprocessing_forward([torch.from_numpy(g[key])], dim)[:, 0]
        assert (got - want).abs().max().item() <= 1e-6
    agg = ops.adaptive_pool1d(torch.from_numpy(g["pre_out"]).cuda(), 13).cpu().numpy()
    assert np.abs(agg - g["agg_out"]).max() <= 1e-6


# ------------------------------------------------------------------------------- stage 2
def _exact_dmin(opq, opb, n_b, P):
    """fp64 min distance per bank image from given operands (CPU, checker only)."""
    q = opq.double().cpu()
    b = opb.double().cpu().reshape(n_b, P, -1)
    out = torch.empty(n_b, q.shape[0], dtype=torch.float64)
    for j in range(n_b):
        out[j] = torch.cdist(q, b[j]).min(dim=1)[0]
    return out


def test_mindist_f32_matches_reference_weights(golden_dir):
    g = gload(golden_dir, "alpha_small")
    Z = torch.from_numpy(g["Z"]).cuda()
    Zb = torch.from_numpy(g["Z_train"]).cuda()
    q = pipeline.patchset_from_Z(Z, "f32")
    b = pipeline.patchset_from_Z(Zb, "f32")
    wu = pipeline.min_distance_weights(q, q, "unsupervised", "f32")
    ws = pipeline.min_distance_weights(q, b, "supervised", "f32")
    assert np.abs(wu.cpu().numpy() - g["w_unsup"]).max() <= 2e-5
    assert np.abs(ws.cpu().numpy() - g["w_sup"]).max() <= 2e-5


@pytest.mark.parametrize("precision", ["f16", "bf16", "f16x3", "bf16x3"])
@pytest.mark.parametrize("shape", [(3, 784, 256), (4, 100, 320), (2, 36, 64), (5, 260, 128)])
def test_mindist_tensor_core_kernel_arithmetic(precision, shape):
    """The tcgen05 kernel against an fp64 evaluation of the SAME rounded operands: isolates kernel
    correctness (tiling, swizzle, segmented min, masking) from operand rounding."""
    n, P, D = shape
    gen = torch.Generator().manual_seed(n * 1000 + P)
    base = torch.randn(1, P, D, generator=gen)
    Z = (base + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, precision)
    dmin = ops.min_dist(ps.hi, ps.lo, ps.n2, ps.hi, ps.lo, ps.n2, n, P, precision)
    op = ps.hi.double() + (ps.lo.double() if ps.lo is not None else 0.0)
    want = _exact_dmin(op, op, n, P)
    # off-diagonal (different image) entries: fp32 accumulation error only
    got = dmin.double().cpu()
    scale = want.max().item()
    # self-distances are sqrt of a cancelled difference of two ~|x|^2 sums: ~sqrt(eps_fp32 * |x|^2 * few) ~ 5e-2
    assert (got - want).abs().max().item() <= 2e-3 * scale + 8e-2
    mask = torch.ones_like(want, dtype=torch.bool)
    for j in range(n):
        mask[j, j * P:(j + 1) * P] = False
    rel = ((got - want).abs() / want.clamp_min(1e-6))[mask].max().item()
    assert rel <= (5e-4 if "x3" not in precision else 5e-5), rel


# f16x3 is bounded by the tensor core's fp32 accumulation over K = 3*4096 (products are exact, the
# running sum is not re-rounded to nearest), not by operand rounding: ~0.1-0.2 absolute in d^2 ~ 1e3.
@pytest.mark.parametrize("precision,tol", [("f16", 3e-3), ("f16x3", 1e-3), ("f32", 1e-4)])
def test_mindist_vs_oracle_config2_width(precision, tol):
    """P = 784, D = 4096 (config-2 geometry), 4 query images vs 5 bank images (supervised form)."""
    feats, _ = synth.planted_features(9, [(768, 28, 28, True), (768, 28, 28, True)], seed=8)
    Zall = restated.embed(feats, 3, 1, 2048, 4096).reshape(9, 784, 4096)
    Zq, Zb = Zall[:4], Zall[4:]
    want = restated.per_image_min_dist(Zq, Zb)  # [4, 784, 5]
    q = pipeline.patchset_from_Z(Zq.cuda(), precision)
    b = pipeline.patchset_from_Z(Zb.cuda(), precision)
    _, dmin = pipeline.min_distance_weights(q, b, "supervised", precision, return_dmin=True)
    got = dmin.reshape(5, 4, 784).permute(1, 2, 0).cpu()
    rel = ((got - want).abs() / want.abs().clamp_min(1e-3)).max().item()
    assert rel <= tol, rel


@pytest.mark.parametrize("n,P,D", [(7, 100, 256), (6, 784, 512), (2, 64, 64), (9, 36, 128), (5, 260, 320)])
@pytest.mark.parametrize("precision", ["f16", "bf16x3"])
def test_mindist_symmetric_equals_all_pairs(n, P, D, precision):
    """ac_min_dist_sym multiplies every unordered image pair once (row-min + column-min of the same
    tile); the resulting w must equal the straightforward all-pairs kernel's."""
    gen = torch.Generator().manual_seed(n * 31 + P)
    base = torch.randn(1, P, D, generator=gen)
    Z = (base + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, precision)
    try:
        pipeline.SYMMETRIC = False
        w_full = pipeline.min_distance_weights(ps, ps, "unsupervised", precision)
        pipeline.SYMMETRIC = True
        w_sym = pipeline.min_distance_weights(ps, ps, "unsupervised", precision)
    finally:
        pipeline.SYMMETRIC = True
    assert torch.isfinite(w_sym).all()
    assert ((w_sym - w_full).abs() / w_full.abs().clamp_min(1e-3)).max().item() <= 2e-4
    # and against the exact fp64 weights of the same operands
    op = (ps.hi.double() + (ps.lo.double() if ps.lo is not None else 0.0)).cpu().reshape(n, P, D)
    want = torch.empty(n, P, dtype=torch.float64)
    for i in range(n):
        cols = [torch.cdist(op[i], op[j]).min(dim=1)[0] for j in range(n) if j != i]
        want[i] = torch.stack(cols, 1).mean(1)
    assert ((w_sym.double().cpu() - want).abs() / want).max().item() <= 5e-4


@pytest.mark.parametrize("arg", [False, True])
def test_mindist_symmetric_waits_for_arriving_bank(arg):
    """ac_min_dist_sym_ready (one launch over a bank whose remote shards are still travelling): the query slice and its own
    bank images are resident, the other images' rows and norms are copied in by ANOTHER stream long after the kernel has
    started, which then sets their arrival flags.  The result must be the one of the resident bank -- bit for bit."""
    n, P, D, nq = 9, 96, 256, 3
    gen = torch.Generator().manual_seed(17)
    Z = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, "f16")
    launch = ops.min_dist_sym_arg if arg else ops.min_dist_sym
    sl = slice(0, nq * P)
    want = launch(ps.hi[sl], None, ps.n2[sl], 0, ps.hi, None, ps.n2, n, P, "f16")
    torch.cuda.synchronize()
    hi = torch.full_like(ps.hi, float("nan"))                 # remote rows: poison until they "arrive"
    n2 = torch.full_like(ps.n2, float("nan"))
    hi[sl], n2[sl] = ps.hi[sl], ps.n2[sl]
    ready = torch.zeros(n, dtype=torch.int32, device="cuda")
    ready[:nq] = 1
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for a in range(nq, n, 2):                             # "shards" of two images, ~10 ms apart
            b = min(n, a + 2)
            torch.cuda._sleep(int(2.0e7))
            hi[a * P : b * P].copy_(ps.hi[a * P : b * P], non_blocking=True)
            n2[a * P : b * P].copy_(ps.n2[a * P : b * P], non_blocking=True)
            ready[a:b].fill_(1)
    got = launch(hi[sl], None, n2[sl], 0, hi, None, n2, n, P, "f16", ready=ready)     # starts at once on the current stream
    torch.cuda.synchronize()
    # rowmin [n, nq*P] is written only for the pairs the query image owns (the rest is uninitialised memory by contract)
    from anomaly_clustering_b200.distributed import pair_owned

    own = torch.tensor([[pair_owned(i, j, n) for i in range(nq)] for j in range(n)], device="cuda")      # [j, i]
    mask = own[:, :, None].expand(n, nq, P).reshape(n, nq * P)
    assert mask.any() and torch.equal(got[0][mask], want[0][mask]) and torch.isfinite(got[0][mask]).all()
    for g, w_ in zip(got[1:], want[1:]):                      # column minima / keys (initialised everywhere), row arg-mins
        assert torch.equal(g, w_)


@pytest.mark.parametrize("arg", [False, True])
def test_mindist_all_pairs_waits_for_arriving_bank(arg):
    """ac_min_dist_ready (sharded supervised runs): the walk over the bank starts at the first resident image, the other
    shards' rows and norms are copied in by another stream after the kernel has started; same bits as with a resident bank."""
    nb, P, D, nq = 8, 96, 256, 3
    gen = torch.Generator().manual_seed(23)
    Zb = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(nb, P, D, generator=gen)).cuda()
    Zq = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(nq, P, D, generator=gen)).cuda()
    bank, q = pipeline.patchset_from_Z(Zb, "f16"), pipeline.patchset_from_Z(Zq, "f16")
    launch = ops.min_dist_arg if arg else ops.min_dist
    want = launch(q.hi, None, q.n2, bank.hi, None, bank.n2, nb, P, "f16")
    torch.cuda.synchronize()
    first, n_local = 5, 2                                      # resident: images 5, 6; arrival order 7, 0, 1, 2, 3, 4
    hi = torch.full_like(bank.hi, float("nan"))
    n2 = torch.full_like(bank.n2, float("nan"))
    ready = torch.zeros(nb, dtype=torch.int32, device="cuda")
    loc = slice(first * P, (first + n_local) * P)
    hi[loc], n2[loc] = bank.hi[loc], bank.n2[loc]
    ready[first:first + n_local] = 1
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for k in range(n_local, nb, 2):
            torch.cuda._sleep(int(2.0e7))
            for j in ((first + k) % nb, (first + k + 1) % nb):
                sl = slice(j * P, (j + 1) * P)
                hi[sl].copy_(bank.hi[sl], non_blocking=True)
                n2[sl].copy_(bank.n2[sl], non_blocking=True)
                ready[j:j + 1].fill_(1)
    got = launch(q.hi, None, q.n2, hi, None, n2, nb, P, "f16", ready=ready, first_image=first)
    torch.cuda.synchronize()
    if arg:
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    else:
        assert torch.equal(got, want)
    # the rotated walk alone (resident bank) changes nothing either
    rot = launch(q.hi, None, q.n2, bank.hi, None, bank.n2, nb, P, "f16", first_image=3)
    assert torch.equal(rot[0] if arg else rot, want[0] if arg else want)


def test_mindist_symmetric_sharded_slices():
    """Two query slices of one bank (what two ranks compute) + the column-block exchange reproduce the
    single-slice result."""
    from anomaly_clustering_b200 import distributed

    n, P, D = 7, 96, 128
    gen = torch.Generator().manual_seed(5)
    Z = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, "f16")
    w_one = pipeline.min_distance_weights(ps, ps, "unsupervised", "f16")
    bounds = distributed.shard_bounds(n, 2)
    parts = []
    for a, b in bounds:
        sl = slice(a * P, b * P)
        parts.append(ops.min_dist_sym(ps.hi[sl], None, ps.n2[sl], a, ps.hi, None, ps.n2, n, P, "f16"))
    for r, (a, b) in enumerate(bounds):
        colfull = torch.cat([parts[s][1][:, a * P:b * P] for s in range(2)], dim=0).contiguous()
        w = ops.reduce_weights_sym(parts[r][0], colfull, P, a).reshape(b - a, P)
        assert ((w - w_one[a:b]).abs() / w_one[a:b]).max().item() <= 2e-4


def test_mindist_symmetric_bank_windows_accumulate():
    """Two launches over complementary circular windows of bank images (what the sharded path does to overlap
    the shard transfers) give bit-identical minima to one launch."""
    n, P, D = 9, 64, 128
    gen = torch.Generator().manual_seed(11)
    Z = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, "f16")
    a, b = 3, 6                                   # this "rank" owns query images [3, 6)
    sl = slice(a * P, b * P)
    one = ops.min_dist_sym(ps.hi[sl], None, ps.n2[sl], a, ps.hi, None, ps.n2, n, P, "f16")
    two = ops.min_dist_sym(ps.hi[sl], None, ps.n2[sl], a, ps.hi, None, ps.n2, n, P, "f16", bank_window=(a, b - a), init=True)
    two = ops.min_dist_sym(ps.hi[sl], None, ps.n2[sl], a, ps.hi, None, ps.n2, n, P, "f16", bank_window=(b % n, n - (b - a)),
                           init=False, out=two)
    from anomaly_clustering_b200 import distributed

    assert torch.equal(one[1], two[1])                       # column minima: every entry
    for j in range(n):                                       # row minima: only owned pairs are defined
        for i in range(a, b):
            if distributed.pair_owned(i, j, n):
                r = slice((i - a) * P, (i - a + 1) * P)
                assert torch.equal(one[0][j, r], two[0][j, r])


# ------------------------------------------------------------------------------- refined mode (arg-min + exact re-evaluation)
@pytest.mark.parametrize("shape", [(3, 784, 256), (4, 100, 320), (5, 260, 128)])
def test_mindist_arg_and_refine_kernel(shape):
    """ac_min_dist_arg names the bank row that won (up to near-ties) and ac_refine_min_dist returns the fp32-exact
    distance of exactly that pair: compared with fp64 on the same operands."""
    n, P, D = shape
    gen = torch.Generator().manual_seed(n * 977 + P)
    Z = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, "f16r")
    assert ps.lo is None
    dmin, arg = ops.min_dist_arg(ps.hi, None, ps.n2, ps.hi, None, ps.n2, n, P, "f16r")
    assert arg.dtype == torch.int32 and int(arg.min()) >= 0 and int(arg.max()) < P
    op = ps.hi.double().cpu().reshape(n, P, D)
    q = op.reshape(n * P, D)
    a = arg.cpu().long()
    for j in range(n):
        d_all = torch.cdist(q, op[j])                                  # [n*P, P] fp64
        d_sel = d_all.gather(1, a[j][:, None])[:, 0]
        d_min = d_all.min(dim=1)[0]
        other = torch.ones(n * P, dtype=torch.bool)
        other[j * P:(j + 1) * P] = False                               # own image: distance 0 to itself, any tie is fine
        assert ((d_sel - d_min)[other] <= 2e-3 * d_min[other] + 1e-6).all()   # the selected row IS the nearest (near-ties aside)
        # refined value from the operand rows (query side = the operand as well here): exact distance of the selected pair
    dex = ops.refine_min_dist(None, ps.hi, None, ps.hi, None, n, P, arg).double().cpu()
    dz = ops.refine_min_dist(Z.reshape(n * P, D).contiguous(), None, None, ps.hi, None, n, P, arg).double().cpu()
    dzn = ops.refine_min_dist(Z.reshape(n * P, D).contiguous(), None, None, ps.hi, None, n, P, arg, Bn2=ps.n2).double().cpu()
    zq = Z.double().cpu().reshape(n * P, D)
    for j in range(n):
        sel = op[j][a[j]]                                              # [n*P, D] selected bank rows
        want = (q - sel).norm(dim=1)
        assert ((dex[j] - want).abs() <= 2e-6 * want + 1e-6).all()     # fp32 sum of squares, no cancellation
        wantz = (zq - sel).norm(dim=1)
        assert ((dz[j] - wantz).abs() <= 2e-6 * wantz + 1e-6).all()    # query row taken from fp32 Z
        off = torch.ones(n * P, dtype=torch.bool)
        off[j * P:(j + 1) * P] = False       # |q|^2+|b|^2-2q.b form (bank norms given): cancellation only matters at d ~ 0 (own image)
        assert ((dzn[j] - wantz).abs()[off] <= 2e-5 * wantz[off] + 1e-5).all()


@pytest.mark.parametrize("n,P,D", [(7, 100, 256), (6, 784, 512), (5, 260, 320)])
def test_refined_symmetric_equals_refined_all_pairs_and_fp64(n, P, D):
    """'f16r' through the symmetric kernel (row arg-mins + (distance, row) column keys) and through the all-pairs kernel
    give the same w; both agree with the fp64 weights of the operands to fp32 round-off (not 5e-4 like one pass)."""
    gen = torch.Generator().manual_seed(n * 131 + P)
    Z = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, "f16r")
    ps.Z = None                                                        # refine from the operand on both sides
    try:
        pipeline.SYMMETRIC = False
        w_full = pipeline.min_distance_weights(ps, ps, "unsupervised", "f16r")
        pipeline.SYMMETRIC = True
        w_sym = pipeline.min_distance_weights(ps, ps, "unsupervised", "f16r")
    finally:
        pipeline.SYMMETRIC = True
    op = ps.hi.double().cpu().reshape(n, P, D)
    want = torch.empty(n, P, dtype=torch.float64)
    for i in range(n):
        want[i] = torch.stack([torch.cdist(op[i], op[j]).min(dim=1)[0] for j in range(n) if j != i], 1).mean(1)
    assert ((w_sym.double().cpu() - want).abs() / want).max().item() <= 2e-5
    assert ((w_full.double().cpu() - want).abs() / want).max().item() <= 2e-5
    assert ((w_sym - w_full).abs() / w_full).max().item() <= 2e-5


def test_refined_sharded_slices_and_windows():
    """Two query slices of one bank in 'f16r' (what two ranks compute: windows accumulate, key column blocks are
    exchanged, every slice refines its own rows against the whole bank) reproduce the single-slice result bit for bit."""
    from anomaly_clustering_b200 import distributed

    n, P, D = 7, 96, 128
    gen = torch.Generator().manual_seed(5)
    Z = (torch.randn(1, P, D, generator=gen) + 0.5 * torch.randn(n, P, D, generator=gen)).cuda()
    ps = pipeline.patchset_from_Z(Z, "f16r")
    w_one = pipeline.min_distance_weights(ps, ps, "unsupervised", "f16r")
    bounds = distributed.shard_bounds(n, 2)
    parts = []
    for a, b in bounds:
        sl = slice(a * P, b * P)
        out = ops.min_dist_sym_arg(ps.hi[sl], None, ps.n2[sl], a, ps.hi, None, ps.n2, n, P, "f16r", bank_window=(a, b - a), init=True)
        out = ops.min_dist_sym_arg(ps.hi[sl], None, ps.n2[sl], a, ps.hi, None, ps.n2, n, P, "f16r", bank_window=(b % n, n - (b - a)),
                                   init=False, out=out)
        parts.append(out)
    for r, (a, b) in enumerate(bounds):
        sl = slice(a * P, b * P)
        colfull = torch.cat([parts[s][2][:, a * P:b * P] for s in range(2)], dim=0).contiguous()
        dex = ops.refine_min_dist(ps.Z[sl], ps.hi[sl], None, ps.hi, None, n, P, parts[r][1], colkey=colfull, q_img0=a, Bn2=ps.n2)
        own = torch.arange(a, b, dtype=torch.int32, device="cuda")
        w = ops.reduce_weights(dex, P, own, "mean").reshape(b - a, P)
        assert torch.equal(w, w_one[a:b])


# ------------------------------------------------------------------------------- stage 3
def test_alpha_golden(golden_dir):
    g = gload(golden_dir, "alpha_small")
    taus = [float(t) for t in g["taus"]]
    for key in ("unsup", "sup"):
        w = torch.from_numpy(g["w_" + key]).cuda()
        a64, a32 = ops.alpha(w, taus)
        for i, t in enumerate(taus):
            want = g["alpha_%s_%g" % (key, t)]
            assert np.abs(a64[i].cpu().numpy() - want).max() <= 1e-12
            assert np.abs(a32[i].cpu().numpy() - want).max() <= 1e-7


def test_alpha_edge_cases():
    w = torch.tensor([[1.0, 3.0, 3.0, 2.0], [800.0, 799.0, 0.0, 0.0]]).cuda()
    a64, _ = ops.alpha(w, [0.0, 1.0])
    assert torch.allclose(a64[0, 0].cpu(), torch.tensor([0.0, 0.5, 0.5, 0.0], dtype=torch.float64))
    assert torch.isfinite(a64).all()  # the reference overflows to NaN here (utils.py:253); documented divergence
    assert abs(a64[1, 1].sum().item() - 1.0) < 1e-12
    want = restated.alpha_from_weights(w.cpu(), 1.0, stable=True)
    assert (a64[1].cpu() - want).abs().max().item() <= 1e-12


@pytest.mark.parametrize("T,N,P,D", [(1, 3, 50, 128), (3, 4, 96, 512), (6, 2, 200, 1024), (17, 3, 64, 256), (5, 2, 40, 130)])
def test_weighted_embed_all_taus_in_one_pass(T, N, P, D):
    """ac_weighted_embed_multi (X for every tau of a sweep from one pass over Z) is bit-identical to one ac_weighted_embed per
    tau (examples/main.py:294-296 inside the tau loop), including tau groups of 8 / 4 / 2 / 1 and a D the vector kernel refuses."""
    gen = torch.Generator(device="cuda").manual_seed(T * 100 + D)
    Z = torch.randn(N, P, D, generator=gen, device="cuda")
    a = torch.softmax(torch.randn(T, N, P, generator=gen, device="cuda") * 3, dim=2)
    got = ops.weighted_embed_multi(a, Z)
    want = torch.stack([ops.weighted_embed(a[t], Z) for t in range(T)])
    torch.cuda.synchronize()
    assert got.shape == (T, N, D) and torch.equal(got, want)
    ref = torch.bmm(a.double().reshape(T * N, 1, P), Z.double().repeat(T, 1, 1)).reshape(T, N, D)
    assert (got.double() - ref).abs().max().item() <= 1e-5


def test_copy_blocks_strided_batched():
    """ac_copy_blocks (the sharded path's collection of column-minimum blocks): several strided 2-D blocks in one launch,
    16-byte and 4-byte aligned shapes, float and 64-bit elements -- bit-exact against torch's copy."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    for dtype, cols in ((torch.float32, 96), (torch.float32, 7), (torch.int64, 33)):
        big = [torch.randint(-1000, 1000, (9, 400), generator=gen, device="cuda").to(dtype) for _ in range(5)]
        srcs = [b[: 3 + k, 11 * k : 11 * k + cols] for k, b in enumerate(big)]          # different rows, offsets, strides
        out = torch.zeros(sum(s_.shape[0] for s_ in srcs), cols, dtype=dtype, device="cuda")
        dsts, r = [], 0
        for s_ in srcs:
            dsts.append(out[r : r + s_.shape[0]])
            r += s_.shape[0]
        ops.copy_blocks(dsts, srcs)
        torch.cuda.synchronize()
        assert torch.equal(out, torch.cat(srcs, dim=0))
    ops.copy_blocks([], [])


def test_weighted_embed_and_pairwise_vs_oracle(golden_dir):
    g = gload(golden_dir, "alpha_small")
    Z = torch.from_numpy(g["Z"]).cuda()
    for t in (0.5, 2.0):
        a = torch.from_numpy(g["alpha_unsup_%g" % t]).float().cuda()
        X = ops.weighted_embed(a, Z)
        assert np.abs(X.cpu().numpy() - g["X_unsup_%g" % t]).max() <= 1e-5
    gen = torch.Generator().manual_seed(1)
    X = torch.randn(77, 300, generator=gen)
    D = ops.pairwise_l2(X.cuda()).cpu().numpy()
    want = restated.pairwise_euclidean(X.numpy())
    assert rel_l2(D, want) <= 1e-6
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    # large N takes the tiled kernel
    X = torch.randn(500, 129, generator=gen)
    D = ops.pairwise_l2(X.cuda()).cpu().numpy()
    assert rel_l2(D, restated.pairwise_euclidean(X.numpy())) <= 1e-6
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    # odd sizes
    X = torch.randn(5, 7, generator=gen)
    assert rel_l2(ops.pairwise_l2(X.cuda()).cpu().numpy(), restated.pairwise_euclidean(X.numpy())) <= 1e-6
    Z2 = torch.randn(3, 10, 6, generator=gen)
    a2 = torch.softmax(torch.randn(3, 10, generator=gen), dim=1)
    assert np.abs(ops.weighted_embed(a2.cuda(), Z2.cuda()).cpu().numpy() - restated.weighted_embedding(a2, Z2)).max() <= 1e-6


# ------------------------------------------------------------------------------- whole path
def _check_path(res, want, labels, k, tau_idx=0):
    Zw, ww, aw, Xw, Dw = want
    a = res.alpha64[tau_idx].cpu()
    assert (a - aw).abs().max().item() <= 1e-3                    # north_star: alpha max-abs <= 1e-3
    assert rel_l2(res.X[tau_idx].cpu().numpy(), Xw) <= 1e-3       # X within 1e-3 relative L2
    assert rel_l2(res.Dmat[tau_idx].cpu().numpy(), Dw) <= 1e-3    # distance matrix within 1e-3 relative L2
    lab_ref = ocluster.ward_labels(Xw, k)
    lab_gpu = ocluster.ward_labels(res.X[tau_idx].cpu().numpy().astype(np.float64), k)
    from sklearn import metrics

    assert metrics.adjusted_rand_score(lab_ref, lab_gpu) == 1.0   # identical Ward partition
    m_ref = [metrics.normalized_mutual_info_score(labels, lab_ref), metrics.adjusted_rand_score(labels, lab_ref)]
    m_gpu = [metrics.normalized_mutual_info_score(labels, lab_gpu), metrics.adjusted_rand_score(labels, lab_gpu)]
    assert m_ref == m_gpu


@pytest.mark.parametrize("precision", ["f16", "f16x3", "f32"])
def test_full_path_unsupervised_wrn_shape(precision):
    """BASELINE config 1 geometry at reduced image count: WRN50 layer2+layer3, 1024 -> 1024, tau = 1."""
    n, k = 8, 4
    feats, labels = synth.planted_features(n, [(512, 28, 28, False), (1024, 14, 14, False)], n_classes=k, seed=2023)
    want = restated.full_path(feats, 3, 1, 1024, 1024, 1.0, "unsupervised")
    res = pipeline.run_path([f.cuda() for f in feats], 3, 1, 1024, 1024, "unsupervised", [1.0], precision=precision)
    assert (res.Z.cpu() - want[0]).abs().max().item() <= 2e-5
    _check_path(res, want, labels.numpy(), k)


def test_full_path_supervised_and_average_vit_shape():
    """Config 3 geometry at reduced counts: ViT-B/8 tokens, 2048 -> 4096, supervised bank + average."""
    n, nb, k = 6, 4, 3
    layers = [(768, 28, 28, True), (768, 28, 28, True)]
    feats, labels = synth.planted_features(n, layers, n_classes=k, seed=2023)
    bank, _ = synth.planted_features(nb, layers, n_classes=1, seed=77)
    want = restated.full_path(feats, 3, 1, 2048, 4096, 2.0, "supervised", bank_features=bank)
    res = pipeline.run_path([f.cuda() for f in feats], 3, 1, 2048, 4096, "supervised", [2.0],
                            bank_features=[f.cuda() for f in bank], precision="f16")
    _check_path(res, want, labels.numpy(), k)
    want_avg = restated.full_path(feats, 3, 1, 2048, 4096, 1.0, "average")
    res_avg = pipeline.run_path([f.cuda() for f in feats], 3, 1, 2048, 4096, "average")
    _check_path(res_avg, want_avg, labels.numpy(), k)


@pytest.mark.parametrize("layers,Dp,D", [([(768, 28, 28, True), (768, 28, 28, True)], 2048, 4096),
                                         ([(64, 12, 12, False), (64, 12, 12, False)], 128, 128),
                                         ([(96, 10, 14, True)], 256, 256)])
def test_z_free_path_matches_z_path(layers, Dp, D):
    """keep_z=False (operands only from the embed kernel, X as a 3x3 correlation of alpha with the feature maps)
    against the Z-based path and the oracle."""
    tokens = layers[0][3]
    if tokens and layers[0][1] != layers[0][2]:
        layers = [(c, 12, 12, t) for c, _, _, t in layers]
    feats, _ = synth.planted_features(4, layers, seed=21)
    f = [x.cuda() for x in feats]
    a = pipeline.run_path(f, 3, 1, Dp, D, "unsupervised", [1.0, 2.0], precision="f16", keep_z=True)
    b = pipeline.run_path(f, 3, 1, Dp, D, "unsupervised", [1.0, 2.0], precision="f16", keep_z=False)
    assert b.Z is None and a.Z is not None
    assert torch.equal(a.w, b.w) and torch.equal(a.alpha64, b.alpha64)
    assert ((a.X - b.X).norm() / a.X.norm()).item() <= 1e-5
    want = restated.full_path(feats, 3, 1, Dp, D, 2.0, "unsupervised")
    assert rel_l2(b.X[1].cpu().numpy(), want[3]) <= 1e-3
    assert rel_l2(b.Dmat[1].cpu().numpy(), want[4]) <= 1e-3
    # average mode without Z
    c = pipeline.run_path(f, 3, 1, Dp, D, "average", keep_z=False)
    want_avg = restated.full_path(feats, 3, 1, Dp, D, 1.0, "average")
    assert rel_l2(c.X[0].cpu().numpy(), want_avg[3]) <= 1e-4


def test_tau_sweep_reuses_one_distance_pass():
    """Config 5 behaviour: every tau from one min-distance pass equals per-tau runs."""
    feats, _ = synth.planted_features(5, [(96, 12, 12, True), (96, 12, 12, True)], seed=4)
    f = [x.cuda() for x in feats]
    taus = [0.5, 1.0, 2.0, 5.0, 10.0]
    res = pipeline.run_path(f, 3, 1, 256, 512, "unsupervised", taus, precision="f16x3")
    Z = restated.embed(feats, 3, 1, 256, 512).reshape(5, 144, 512)
    w = restated.weight_distance_unsupervised(Z)
    for i, t in enumerate(taus):
        a = restated.alpha_from_weights(w, t)
        assert (res.alpha64[i].cpu() - a).abs().max().item() <= 1e-3
        assert rel_l2(res.X[i].cpu().numpy(), restated.weighted_embedding(a, Z)) <= 1e-3


# The full BASELINE sizes (configs 1, 2, 3, 5 against the oracle) live in tests/test_gpu_baseline_sizes.py.


# ------------------------------------------------------------------------------- per-category banks in one launch
@pytest.mark.parametrize("precision", ["f16", "f16r"])
def test_batched_categories_equal_per_category_runs(precision):
    """run_categories (one launch sequence, category table in the unit list -- reference semantics: one
    make_category_data per category, main.py:353) gives exactly what separate run_path calls per category give."""
    sizes = [5, 9, 4, 7, 2]
    layers = [(96, 12, 12, True), (96, 12, 12, True)]
    feats, _ = synth.planted_features(sum(sizes), layers, seed=31)
    f = [x.cuda() for x in feats]
    taus = [1.0, 0.25]
    got = pipeline.run_categories(f, sizes, 3, 1, 256, 512, taus, precision=precision)
    start = 0
    for c, n in enumerate(sizes):
        ref = pipeline.run_path([x[start:start + n] for x in f], 3, 1, 256, 512, "unsupervised", taus, precision=precision)
        assert torch.equal(got[c].w, ref.w), c
        assert torch.equal(got[c].alpha64, ref.alpha64) and torch.equal(got[c].X, ref.X) and torch.equal(got[c].Dmat, ref.Dmat)
        assert torch.equal(got[c].Z, ref.Z)
        start += n
    # and against the oracle for one category
    want = restated.full_path([x[5:14] for x in feats], 3, 1, 256, 512, 1.0, "unsupervised")
    assert (got[1].alpha64[0].cpu() - want[2]).abs().max().item() <= 1e-3
    assert rel_l2(got[1].X[0].cpu().numpy(), want[3]) <= 1e-3


def test_batched_categories_config2_width_straddling_blocks():
    """Categories whose boundaries fall inside 256-row query blocks and 32-row epilogue warps (P = 784)."""
    sizes = [3, 4, 2]
    layers = [(768, 28, 28, True), (768, 28, 28, True)]
    f, _ = synth.planted_features_device(range(sum(sizes)), layers, device="cuda")
    got = pipeline.run_categories(f, sizes, 3, 1, 2048, 4096, [1.0], precision="f16", keep_z=False)
    start = 0
    for c, n in enumerate(sizes):
        ref = pipeline.run_path([x[start:start + n] for x in f], 3, 1, 2048, 4096, "unsupervised", [1.0], precision="f16", keep_z=False)
        assert torch.equal(got[c].w, ref.w) and torch.equal(got[c].Dmat, ref.Dmat), c
        start += n


def test_supervised_with_straddling_aggregator_windows():
    """L = 3 layers, Dp = 32, D = 64: Aggregator windows straddle layers, so the operands cannot be written without Z.
    The supervised bank (embedded without Z) must still work (it used to fail with AC_ERR_UNSUPPORTED)."""
    gen = torch.Generator().manual_seed(9)
    feats = [torch.randn(4, 8, 6, 6, generator=gen) for _ in range(3)]
    bank = [torch.randn(3, 8, 6, 6, generator=gen) for _ in range(3)]
    assert not ops.aggregator_fusable(3, 32, 64)
    want = restated.full_path(feats, 3, 1, 32, 64, 2.0, "supervised", bank_features=bank)
    res = pipeline.run_path([f.cuda() for f in feats], 3, 1, 32, 64, "supervised", [2.0], bank_features=[f.cuda() for f in bank],
                            precision="f16")
    assert (res.alpha64[0].cpu() - want[2]).abs().max().item() <= 1e-3
    assert rel_l2(res.X[0].cpu().numpy(), want[3]) <= 1e-3
    with pytest.raises(Exception):       # the C entry point refuses before enqueuing anything
        ops.embed([f.cuda() for f in bank], 3, 1, 32, 64, want_z=False, operand="f16")
