"""The per-thread code of the image pre-processing kernels (csrc/preprocess_core.h: the staging / filtering / store
phases and the host-side plan) replayed on the host by tests/preprocess_harness.cpp and compared bit for bit with the
Pillow/HF golden vectors and with oracle/clip_preprocess_oracle.py.  No GPU: this checks the arithmetic and the index
maths the device executes; the launch itself is covered by tests/test_preprocess_gpu.py."""
import ctypes as C
import hashlib
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clip_preprocess_oracle as P  # noqa: E402

CSRC = os.path.join(ROOT, "visual-chinese-llama-alpaca_b200", "csrc")
MEAN = np.asarray(P.CLIP_MEAN, np.float32)
STD = np.asarray(P.CLIP_STD, np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("pp") / "libpp_harness.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, os.path.join(ROOT, "tests", "preprocess_harness.cpp"),
                    "-o", so], check=True)
    lib = C.CDLL(so)
    lib.harness_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.harness_taps.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.harness_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


def run(lib, img, S=224, nthr=256):
    img = np.ascontiguousarray(img)
    px = np.full((3, S, S), np.nan, np.float32)
    u8 = np.zeros((S, S, 3), np.uint8)
    rc = lib.harness_preprocess(img.ctypes.data, img.shape[0], img.shape[1], S, MEAN.ctypes.data, STD.ctypes.data, nthr,
                                px.ctypes.data, u8.ctypes.data)
    assert rc == 0
    return px, u8


def test_taps_match_oracle_and_golden(harness):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))
    for n_in, n_out in [(640, 298), (480, 224), (53, 317), (224, 224), (5, 224), (4000, 224), (225, 226), (1, 7), (7, 1)]:
        ks = harness.harness_ksize(n_in, n_out)
        first, count = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32)
        taps = np.full((n_out, ks), -1, np.int32)
        harness.harness_taps(n_in, n_out, first.ctypes.data, count.ctypes.data, taps.ctypes.data)
        xmin, cnt, kk = P.resample_coeffs(n_in, n_out)
        assert kk.shape[1] == ks
        assert np.array_equal(first, xmin) and np.array_equal(count, cnt) and np.array_equal(taps, kk), (n_in, n_out)
        tag = f"{n_in}_{n_out}"
        if f"coef_{tag}_k" in gold:
            assert np.array_equal(taps, gold[f"coef_{tag}_k"])


def test_geometry_matches_hf_rules(harness):
    for h, w in [(480, 640), (640, 480), (225, 223), (224, 224), (37, 53), (1000, 750), (1080, 1920), (3, 2000)]:
        g = np.zeros(10, np.int32)
        harness.harness_geometry(h, w, 224, g.ctypes.data)
        rh, rw, top, left, kh, kv, row0, rows, col0, cols = g.tolist()
        assert (rh, rw) == P.resize_output_size(h, w)
        assert (top, left) == P.center_crop_box(rh, rw)
        assert 0 <= row0 and row0 + rows <= h and 0 <= col0 and col0 + cols <= w and rows >= 1 and cols >= 1


def test_golden_cases_bit_exact(harness):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))
    for i, (h, w) in enumerate(gold["cases"].tolist()):
        img = P.synthetic_image(h, w, seed=3 * h + w)
        px, u8 = run(harness, img)
        assert sha(u8) == str(gold[f"u8_sha_{i}"]), f"u8 differs from PIL for {h}x{w}"
        assert sha(px) == str(gold[f"px_sha_{i}"]), f"pixel_values differ from CLIPImageProcessor for {h}x{w}"


def test_random_shapes_and_block_sizes_vs_oracle(harness):
    rng = np.random.default_rng(11)
    shapes = [(1, 1), (2, 3), (3, 2), (224, 1), (1, 224), (223, 225), (449, 447), (31, 1500), (1500, 31)]
    shapes += [(int(rng.integers(2, 700)), int(rng.integers(2, 700))) for _ in range(8)]
    for n, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        want_u8 = P.resize_and_crop_u8(img)
        want_px = P.rescale_normalize(want_u8)
        for nthr in ((256,) if n % 3 else (1, 96, 256, 1024)):
            px, u8 = run(harness, img, nthr=nthr)
            assert np.array_equal(u8, want_u8), (h, w, nthr)
            assert np.array_equal(px, want_px), (h, w, nthr)


def test_other_output_side(harness):
    """The tiny test config of the path uses 56x56 pictures (v_image = 56)."""
    rng = np.random.default_rng(5)
    for h, w in [(100, 80), (56, 56), (30, 200)]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        px, u8 = run(harness, img, S=56)
        assert np.array_equal(u8, P.resize_and_crop_u8(img, 56))
        assert np.array_equal(px, P.clip_preprocess(img, 56))


# ---- the shipped library's host-only entry points and the Python processor (still no GPU) ------------------------------
@pytest.fixture(scope="module")
def processor_cls():
    import __graft_entry__ as ge
    ge.build()
    from visualcla.image_processing_vcla import VclaImageProcessor
    return VclaImageProcessor


def test_library_tap_tables_match_oracle(processor_cls):
    for n_in, n_out in [(640, 298), (53, 317), (224, 224), (4000, 224), (1, 9)]:
        first, count, taps = processor_cls.resample_taps(n_in, n_out)
        xmin, cnt, kk = P.resample_coeffs(n_in, n_out)
        assert np.array_equal(first, xmin) and np.array_equal(count, cnt) and np.array_equal(taps, kk)


def test_processor_surface_and_validation(processor_cls, tmp_path):
    from visualcla import _native as N
    p = processor_cls(patch_size=14)
    assert p.size["shortest_edge"] == 224 and p.crop_size == {"height": 224, "width": 224} and p.patch_size == 14
    assert (p.size["shortest_edge"] // p.patch_size) ** 2 + 1 == 257          # ref modeling_utils.py:139
    assert p.workspace_bytes(480, 640) >= 480 * 224 * 3
    with pytest.raises(N.NativeError):
        p.workspace_bytes(1, 40000)
    for bad in (dict(resample=2), dict(do_center_crop=False), dict(crop_size=200), dict(rescale_factor=1 / 256),
                dict(crop_size={"height": 224, "width": 200})):
        with pytest.raises(ValueError):
            processor_cls(**bad)
    p.save_pretrained(str(tmp_path))
    q = processor_cls.from_pretrained(str(tmp_path), patch_size=14)
    assert q.to_dict() == p.to_dict()
    # an old-style HF config (plain ints) loads too
    import json
    with open(tmp_path / "preprocessor_config.json", "w") as fh:
        json.dump({"size": 224, "crop_size": 224, "do_resize": True, "do_center_crop": True, "do_normalize": True, "resample": 3,
                   "image_mean": list(P.CLIP_MEAN), "image_std": list(P.CLIP_STD), "feature_extractor_type": "CLIPFeatureExtractor"}, fh)
    assert processor_cls.from_pretrained(str(tmp_path)).size == {"shortest_edge": 224}
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(N.NativeError):          # no CPU fallback
            p(np.zeros((8, 8, 3), np.uint8))
