"""CPU tests: the C-ABI library builds/loads here (no GPU) and exports every symbol include/ggml_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ggml_b200.h")).read()
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ["b200_graph_compute", "b200_supports_op", "b200_malloc", "b200_ctx_create", "b200_event_record", "b200_quantize_act"]:
        assert must in syms


def test_library_exports_every_declared_symbol(b200):
    if not os.path.exists(b200.KERNELS_SO):
        pytest.fail("libggml_b200_kernels.so not built: run __graft_entry__.build()")
    L = C.CDLL(b200.KERNELS_SO)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, "header declares symbols the library does not export: %s" % missing
    assert L.b200_abi_version() == 1


def test_binding_matches_header(b200):
    L = b200.lib()          # sets argtypes for every symbol it knows; AttributeError if one is missing
    assert L.b200_device_count() >= 0
    assert C.sizeof(b200.Tensor) == 8 + 4 + 4 + 32 + 32
    assert C.sizeof(b200.Op) == 4 + 4 + 64 + 5 * C.sizeof(b200.Tensor)


def test_no_gpu_means_loud_failure(b200):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.B200Error):
        b200.Context(0)


def test_ggml_backend_exports_entry_points(b200):
    """the ggml plugin must export the two symbols ggml_backend_load() looks up (ggml-backend-reg.cpp:227-271)"""
    if not os.path.exists(b200.BACKEND_SO):
        pytest.skip("libggml-b200.so not built (needs the reference headers at build time)")
    import subprocess
    out = subprocess.run(["nm", "-D", b200.BACKEND_SO], capture_output=True, text=True).stdout
    assert " T ggml_backend_init" in out and " T ggml_backend_score" in out
