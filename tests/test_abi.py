"""CPU: the C-ABI library builds/loads without a GPU, exports every symbol include/medmoe_b200.h
declares, and the ctypes signatures in medmoe_b200/_lib.py agree with the header (count and kind)."""
import ctypes
import os
import re

from medmoe_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    h = open(os.path.join(ROOT, "include", "medmoe_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return re.findall(r"(?:const char\*|int|long long|void)\s+(mm_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    decls = _header_decls()
    assert len(decls) >= 25
    for name, _ in decls:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.mm_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_ctypes_signatures_match_header():
    kinds = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float}
    names = set()
    for name, args in _header_decls():
        names.add(name)
        assert name in _lib.SIGNATURES, f"{name} missing from _lib.SIGNATURES"
        sig = _lib.SIGNATURES[name][1]
        args = args.strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        assert len(params) == len(sig), f"{name}: header has {len(params)} params, ctypes {len(sig)}"
        for p, t in zip(params, sig):
            if "*" in p:
                assert t in (ctypes.c_void_p, ctypes.c_char_p), f"{name}: {p}"
            else:
                base = re.sub(r"\s+\w+$", "", p).replace("const ", "").strip()
                assert t is kinds[base], f"{name}: {p} vs {t}"
    assert names == set(_lib.SIGNATURES), set(_lib.SIGNATURES) ^ names


def test_compute_call_fails_loudly_without_a_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import medmoe_b200
    moe = medmoe_b200.MoE(num_experts=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        moe([torch.randn(1, 64, 96), torch.randn(1, 16, 192), torch.randn(1, 4, 384), torch.randn(1, 1, 768)], torch.randn(1, 768))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        medmoe_b200.GLORIAGlobalContrastiveLoss()(torch.randn(4, 768), torch.randn(4, 768))


def test_epilogue_flag_values_match_the_header():
    """The Python host layer passes the epilogue flags as plain ints: they must be the header's enum values."""
    from medmoe_b200 import ops
    h = open(os.path.join(ROOT, "include", "medmoe_b200.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"(MM_EPI_\w+)\s*=\s*(\d+)", h))
    assert enum["MM_EPI_RELU"] == ops.EPI_RELU and enum["MM_EPI_ZERO_PAD"] == ops.EPI_ZERO_PAD
    assert enum["MM_EPI_PAIR_OK"] == ops.EPI_PAIR_OK
    # and the kernels' own enum (csrc/gemm.cuh)
    g = open(os.path.join(ROOT, "medmoe_b200", "csrc", "gemm.cuh")).read()
    kern = dict((k, int(v)) for k, v in re.findall(r"\b(EPI_\w+)\s*=\s*(\d+)", g))
    for name, val in enum.items():
        assert kern[name[3:]] == val, name
