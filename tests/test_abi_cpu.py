"""The C-ABI library builds, loads without a GPU, and exports exactly what include/vcla.h declares.  CPU only
(no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from visualcla import _native
    return _native.load()


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "vcla.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vcla_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(lib):
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vcla.h but not exported by libvcla.so"


def test_binding_covers_header():
    from visualcla import _native
    assert sorted(_native.EXPORTED_SYMBOLS) == _header_symbols()


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.vcla_version()
    assert isinstance(lib.vcla_last_error(), bytes)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from visualcla import _native
    cfg = _native.VclaConfig(v_hidden=128, v_layers=1, v_heads=2, v_ffn=256, v_patch=14, v_image=56, v_eps=1e-5,
                             r_hidden=128, r_layers=1, r_heads=2, r_ffn=256, r_queries=8, r_eps=1e-12,
                             t_hidden=256, t_layers=1, t_heads=2, t_ffn=448, t_vocab=100, t_eps=1e-6, rope_theta=1e4,
                             max_batch=1, max_seq=32, max_prefill_tokens=32, page_tokens=16)
    ctx = C.c_void_p()
    rc = lib.vcla_create(C.byref(cfg), C.byref(ctx))
    assert rc != 0 and b"no CUDA device" in lib.vcla_last_error()
    import visualcla
    with pytest.raises(_native.NativeError):
        visualcla.VisualCLAModel.from_synthetic("7b")


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may execute oracle/ (comments may cite it)."""
    pkg = os.path.join(ROOT, "visual-chinese-llama-alpaca_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(d, f)
            if f.endswith(".py"):
                for line in open(path):
                    assert not re.match(r"\s*(import|from)\s+\S*oracle", line), (path, line)
                    assert "sys.path" not in line or "oracle" not in line, (path, line)
            elif f.endswith((".cu", ".cuh", ".h")):
                for line in open(path):
                    assert not ("#include" in line and "oracle" in line), (path, line)
