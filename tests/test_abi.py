"""The drop-in boundary: libkernels.so loads without a GPU and exports every
symbol include/*.h declares and every symbol the reference's Go loader resolves
(internal/cuda/kernels/purego.go:141-273; the gate of symbol_parity_test.go:57)."""
import ctypes
import os
import subprocess

import pytest

from zerfoo_b200 import lib


@pytest.fixture(scope="module")
def so_path():
    p = lib.lib_path()
    if not os.path.exists(p):
        lib.build()
    return p


def exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_headers_declare_the_expected_counts():
    assert len(lib.declared_symbols("zerfoo_kernels.h")) == 79   # 79 extern "C" launchers
    assert len(lib.declared_symbols("zb200.h")) >= 17


def test_every_declared_symbol_is_exported(so_path):
    have = exported(so_path)
    for header in ("zerfoo_kernels.h", "zb200.h"):
        missing = [n for _, n, _ in lib.declared_symbols(header) if n not in have]
        assert not missing, f"{header}: not exported: {missing}"


def test_reference_loader_symbols_present(so_path, golden_dir):
    """TestForkParitySymbols in dynamic mode: every name purego.go dlsym's."""
    want = open(os.path.join(golden_dir, "ref_purego_symbols.txt")).read().split()
    assert len(want) >= 60
    have = exported(so_path)
    assert not [s for s in want if s not in have]


def test_library_loads_and_binds_without_gpu(so_path):
    L = lib.load()
    assert isinstance(L, ctypes.CDLL)
    assert L.gemv_q4k_check_sm121() == 0
    assert L.gemm_q4_f32.argtypes[-1] is ctypes.c_void_p     # trailing cudaStream_t
    assert L.fused_add_rmsnorm_f32.argtypes[5] is ctypes.c_uint  # eps as IEEE-754 bits


def test_no_oracle_in_product():
    """The product path must not reach into oracle/ (tier rule 3)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "zerfoo_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "zoracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_engine_fails_loudly_without_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zerfoo_b200 import engine
    with pytest.raises(engine.EngineError, match="no CPU fallback|CUDA"):
        engine.load_file("/nonexistent.gguf")


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setenv("ZERFOO_KERNEL_LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(lib, "_lib", None)
    with pytest.raises(lib.KernelLibraryError):
        lib.load()
