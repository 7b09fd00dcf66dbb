"""CPU: the C-ABI library builds, loads, and exports exactly what include/ac_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from anomaly_clustering_b200 import build

    return build.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ac_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ac_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "ac_embed" in syms and "ac_min_dist" in syms and "ac_pairwise_l2" in syms
    assert len(syms) >= 15


def test_library_exports_every_header_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in header_symbols():
        assert hasattr(lib, s), "missing export: " + s


def test_binding_covers_header(lib_path):
    from anomaly_clustering_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.ac_version() >= 100
    assert lib.ac_strerror(-3).decode().startswith("device is not sm_100")
    # workspace sizing is host-only arithmetic
    tok = (_lib.AcLayer * 2)(_lib.AcLayer(0, 768, 28, 28, 785 * 768, 1, 28 * 768, 768), _lib.AcLayer(0, 768, 28, 28, 785 * 768, 1, 28 * 768, 768))
    assert lib.ac_embed_workspace_bytes(tok, 2, 4, 3, 1, 2048, 4096) < (1 << 20)    # token layout, fused aggregator: no scratch
    cnn = (_lib.AcLayer * 3)(*[_lib.AcLayer(0, 40, 10, 10, 4000, 100, 10, 1)] * 3)
    assert lib.ac_embed_workspace_bytes(cnn, 3, 3, 3, 1, 100, 77) > 3 * 100 * 300 * 4   # straddling Aggregator windows need the concat
    assert lib.ac_min_dist_workspace_bytes(1000, 10, 784, 4096, 0) >= 256


def test_no_cpu_fallback_without_gpu(lib_path):
    import torch

    from anomaly_clustering_b200 import ops

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ValueError):
        ops.pairwise_l2(torch.zeros(4, 8))


def test_sass_is_blackwell_native(lib_path):
    """tcgen05 / TMEM / TMA mnemonics must be present in the shipped SASS (B200_PROFILING.md)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic


def test_argument_validation_needs_no_gpu(lib_path):
    """Null pointers / bad sizes are rejected before any CUDA call; valid-looking calls fail loudly
    (AC_ERR_CUDA or AC_ERR_DEVICE) on a machine without a B200 -- never a silent CPU path."""
    import torch

    from anomaly_clustering_b200 import _lib

    lib = _lib.load()
    assert lib.ac_pairwise_l2(None, 4, 8, None, None) == _lib.AC_ERR_INVALID
    assert lib.ac_weighted_embed(None, None, 1, 1, 1, None, None) == _lib.AC_ERR_INVALID
    assert lib.ac_reduce_weights(None, 1, 1, 1, None, 0, None, None) == _lib.AC_ERR_INVALID
    assert lib.ac_min_dist(None, None, None, 1, None, None, None, 1, 1, 8, 0, None, None, 0, None) == _lib.AC_ERR_INVALID
    assert lib.ac_embed(None, 1, 1, 3, 1, 8, 8, 1, 1e-5, None, None, None, 0, None, 0, None) == _lib.AC_ERR_INVALID
    assert lib.ac_debug_set(0, 7) == _lib.AC_ERR_INVALID
    assert lib.ac_copy_blocks(1, None, None, None, None, None, None, None) == _lib.AC_ERR_INVALID
    assert lib.ac_copy_blocks(17, None, None, None, None, None, None, None) == _lib.AC_ERR_UNSUPPORTED     # at most 16 blocks per launch
    assert lib.ac_copy_blocks(0, None, None, None, None, None, None, None) == _lib.AC_OK
    if not torch.cuda.is_available():
        buf = (ctypes.c_float * 64)()
        rc = lib.ac_pairwise_l2(ctypes.cast(buf, ctypes.c_void_p), 4, 8, ctypes.cast(buf, ctypes.c_void_p), None)
        assert rc in (_lib.AC_ERR_CUDA, _lib.AC_ERR_DEVICE)
        with pytest.raises(_lib.AcError):
            _lib.check(rc, "ac_pairwise_l2")


def test_plain_c_program_links_the_abi(lib_path, tmp_path):
    """tests/c/abi_smoke.c compiles against include/ac_b200.h and links libac_b200.so with gcc (no Python,
    no torch); without a B200 it must exit with the 'no device' code instead of computing anything."""
    import shutil
    import subprocess

    import torch

    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no C toolchain / CUDA headers")
    pkg = os.path.dirname(lib_path)
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([gcc, os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-I" + os.path.join(ROOT, "include"),
                    "-I/usr/local/cuda/include", "-L" + pkg, "-lac_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-lm",
                    "-Wl,-rpath," + pkg, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout
    else:
        assert r.returncode == 2 and "no sm_100 device" in r.stdout


def test_header_is_valid_c99_and_cxx(tmp_path):
    """include/ac_b200.h on its own, as strict C99 and as C++17, with warnings as errors: the boundary a non-Python
    consumer compiles against must not depend on anything but <stddef.h>/<stdint.h> and the CUDA runtime header."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no C toolchain / CUDA headers")
    inc = ["-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]
    c = tmp_path / "t.c"
    c.write_text('#include "ac_b200.h"\nint main(void) { int (*f)(void) = ac_version; (void)f; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only", str(c)] + inc, check=True)
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "ac_b200.h"\nint main() { int (*f)() = ac_version; (void)f; return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only", str(cpp)] + inc, check=True)


def test_mirror_surface_never_computes_on_cpu(lib_path):
    """Every compute entry point of the reference-facing mirror refuses CPU tensors instead of quietly running
    torch on the host: the product has no CPU path (the oracle is test infrastructure only)."""
    import torch

    from anomaly_clustering_b200 import ops, pipeline
    from anomaly_clustering_b200.patchcore import common, patchcore, utils

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = torch.zeros(2, 6, 5, 5)
    Z = torch.zeros(3, 25, 16)
    calls = [
        lambda: common.MeanMapper(8)(x),
        lambda: common.Preprocessing([54, 54], 8)([x, x]),
        lambda: common.Aggregator(8)(torch.zeros(4, 2, 8)),
        lambda: patchcore.PatchMaker(3, stride=1).patchify(x),
        lambda: utils.Matrix_Alpha_Unsupervised(1.0, 1, Z, torch.device("cpu")),
        lambda: utils.Matrix_Alpha_Supervised(1.0, 1, Z, Z, torch.device("cpu")),
        lambda: utils.Weight_Distance_Unsupervised(Z, 0, torch.device("cpu")),
        lambda: pipeline.run_path([x, x], 3, 1, 8, 16),
        lambda: ops.embed([x], 3, 1, 8, 16),
        lambda: ops.alpha(torch.zeros(2, 4), [1.0]),
        lambda: ops.weighted_embed(torch.zeros(3, 25), Z),
    ]
    for call in calls:
        with pytest.raises(ValueError, match="CUDA tensors only"):
            call()


def test_fp16_operand_range_guard_is_host_logic():
    """pipeline.guard_operand_range: embeddings handed to the mirror from outside keep the requested tensor-core mode only
    inside the fp16 operand range; outside they go to the exact fp32 kernel (with a warning), non-finite input is refused."""
    import torch

    from anomaly_clustering_b200 import pipeline

    Z = torch.randn(2, 8, 16)
    assert pipeline.guard_operand_range("f16", Z) == "f16"
    assert pipeline.guard_operand_range("f16r", Z, None, Z * 100) == "f16r"
    assert pipeline.guard_operand_range("f32", Z * 1e9) == "f32"
    assert pipeline.guard_operand_range("bf16", Z * 1e9) == "bf16"          # bf16 has the fp32 exponent range
    for scale in (1e6, 1e-6):
        with pytest.warns(RuntimeWarning, match="fp16 operand range"):
            assert pipeline.guard_operand_range("f16r", Z, Z * scale) == "f32"
    assert pipeline.guard_operand_range("f16", torch.zeros(2, 3, 4)) == "f16"
    with pytest.raises(ValueError, match="non-finite"):
        pipeline.guard_operand_range("f16", torch.full((1, 2, 3), float("inf")))
