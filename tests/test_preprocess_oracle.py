"""Pins oracle/clip_preprocess_oracle.py -- the restatement of the pre-processing either side of the hot path
(PIL bicubic resize -> centre crop -> 1/255 -> mean/std) -- to tests/golden/preprocess.npz, which
oracle/gen_golden_preprocess.py produced from PIL.Image.resize and transformers' PIL-backed CLIPImageProcessor.
Integer work end to end until the last float32 affine step: the bar is bit-exact."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clip_preprocess_oracle as P  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))


def test_pipeline_cases_bit_exact(gold):
    for i, (h, w) in enumerate(gold["cases"].tolist()):
        img = P.synthetic_image(h, w, seed=3 * h + w)
        assert sha(img) == str(gold[f"in_sha_{i}"]), f"picture generator drifted for case {i}"
        if f"in_{i}" in gold:
            assert np.array_equal(img, gold[f"in_{i}"])
        u8 = P.resize_and_crop_u8(img)
        assert u8.shape == (224, 224, 3) and u8.dtype == np.uint8
        assert sha(u8) == str(gold[f"u8_sha_{i}"]), f"resize+crop differs from PIL for {h}x{w}"
        if f"u8_{i}" in gold:
            assert np.array_equal(u8, gold[f"u8_{i}"])
        px = P.rescale_normalize(u8)
        assert px.shape == (3, 224, 224) and px.dtype == np.float32
        assert sha(px) == str(gold[f"px_sha_{i}"]), f"pixel_values differ from CLIPImageProcessor for {h}x{w}"


def test_plain_resizes_bit_exact(gold):
    for j, (h, w, oh, ow) in enumerate(gold["resizes"].tolist()):
        img = P.synthetic_image(h, w, seed=h + w)
        r = P.resize_bicubic_u8(img, oh, ow)
        assert sha(r) == str(gold[f"rs_sha_{j}"])
        if f"rs_{j}" in gold:
            assert np.array_equal(r, gold[f"rs_{j}"])


def test_coefficient_tables(gold):
    for tag in ("640_298", "480_224", "53_317"):
        n_in, n_out = map(int, tag.split("_"))
        xmin, cnt, kk = P.resample_coeffs(n_in, n_out)
        assert np.array_equal(xmin, gold[f"coef_{tag}_xmin"])
        assert np.array_equal(cnt, gold[f"coef_{tag}_cnt"])
        assert np.array_equal(kk, gold[f"coef_{tag}_k"])
        # every row is a partition of unity up to the 22-bit quantisation of its taps
        s = kk.astype(np.int64).sum(axis=1)
        assert np.all(np.abs(s - (1 << P.PRECISION_BITS)) <= kk.shape[1])
        assert np.all(xmin >= 0) and np.all(xmin + cnt <= n_in) and np.all(cnt >= 1)


def test_size_rules():
    # HF get_resize_output_image_size(shortest_edge) truncates the long side; the crop box floors
    assert P.resize_output_size(480, 640) == (224, 298)
    assert P.resize_output_size(640, 480) == (298, 224)
    assert P.resize_output_size(225, 223) == (226, 224)
    assert P.resize_output_size(224, 224) == (224, 224)
    assert P.resize_output_size(37, 53) == (224, 320)
    assert P.center_crop_box(224, 298) == (0, 37)
    assert P.center_crop_box(226, 224) == (1, 0)


def test_properties():
    # a constant picture stays constant through both passes (coefficients sum to one, rounding included)
    for v in (0, 1, 127, 254, 255):
        img = np.full((61, 97, 3), v, np.uint8)
        assert np.all(P.resize_bicubic_u8(img, 224, 356) == v)
        assert np.all(P.resize_bicubic_u8(img, 20, 31) == v)
    # identity resize is a no-op; channels are independent
    img = P.synthetic_image(50, 70, 1)
    assert np.array_equal(P.resize_bicubic_u8(img, 50, 70), img)
    r = P.resize_bicubic_u8(img, 33, 91)
    for c in range(3):
        assert np.array_equal(P.resize_bicubic_u8(img[:, :, c:c + 1], 33, 91)[:, :, 0], r[:, :, c])
    # flipping commutes with the (symmetric) filter
    assert np.array_equal(P.resize_bicubic_u8(img[:, ::-1].copy(), 33, 91), r[:, ::-1])
    assert np.array_equal(P.resize_bicubic_u8(img[::-1].copy(), 33, 91), r[::-1])
    # the float step is exactly the per-channel affine map of the 256 byte values
    lut = P.rescale_normalize(np.arange(256, dtype=np.uint8).reshape(256, 1, 1).repeat(3, axis=2))
    assert lut.shape == (3, 256, 1) and np.all(np.diff(lut[:, :, 0], axis=1) > 0)
    assert abs(float(lut[0, 0, 0]) - (0 - 0.48145466) / 0.26862954) < 1e-6


def test_live_against_pil():
    """Where Pillow is importable (it is in this image), also compare fresh random sizes against it directly."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(7)
    for _ in range(12):
        h, w = int(rng.integers(5, 400)), int(rng.integers(5, 400))
        oh, ow = int(rng.integers(3, 300)), int(rng.integers(3, 300))
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(P.resize_bicubic_u8(img, oh, ow), ref), (h, w, oh, ow)
