"""Generate tests/golden/preprocess.npz: the image pre-processing step run through the REAL
dependencies of the reference -- PIL.Image.resize and transformers' PIL-backed CLIPImageProcessor -- on deterministic
pictures.  Run in the authoring container only:

    python oracle/gen_golden_preprocess.py

The fixture pins oracle/clip_preprocess_oracle.py (tests/test_preprocess_oracle.py).  The script refuses to write a
fixture the restatement does not reproduce bit for bit.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import clip_preprocess_oracle as P  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "preprocess.npz")

# (h, w): up-scaling, both aspect orientations, identity, heavy down-scaling, off-by-one crop boxes
CASES = [(37, 53), (300, 200), (224, 224), (225, 223), (500, 333), (480, 640), (1000, 750)]
FULL_OUTPUT = (0, 3, 5)          # cases whose 224x224x3 result is stored whole; the rest are pinned by SHA-256
# plain PIL resizes (no crop), incl. a one-axis resize and a 2x box-like reduction
RESIZES = [((64, 64), (32, 32)), ((100, 80), (50, 200)), ((7, 5), (224, 224)), ((90, 224), (224, 224))]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from PIL import Image
    try:
        from transformers.models.clip import CLIPImageProcessorPil as PilProcessor       # transformers >= 5
    except ImportError:
        from transformers import CLIPImageProcessor as PilProcessor                      # transformers 4.x: PIL pipeline
    proc = PilProcessor()
    out = {"cases": np.asarray(CASES, np.int32), "resizes": np.asarray([a + b for a, b in RESIZES], np.int32)}
    for i, (h, w) in enumerate(CASES):
        img = P.synthetic_image(h, w, seed=3 * h + w)
        oh, ow = P.resize_output_size(h, w)
        ref_resized = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        top, left = P.center_crop_box(oh, ow)
        ref_u8 = ref_resized[top:top + 224, left:left + 224]
        ref_px = proc(images=Image.fromarray(img), return_tensors="np").pixel_values[0]
        assert np.array_equal(P.resize_and_crop_u8(img), ref_u8), f"u8 mismatch {h}x{w}"
        assert ref_px.dtype == np.float32 and np.array_equal(P.clip_preprocess(img), ref_px), f"f32 mismatch {h}x{w}"
        out[f"in_sha_{i}"] = np.asarray(sha(img))
        out[f"u8_sha_{i}"] = np.asarray(sha(ref_u8))
        if i in FULL_OUTPUT:
            out[f"u8_{i}"] = ref_u8                      # what the dependency produced, not the restatement
        out[f"px_sha_{i}"] = np.asarray(sha(ref_px))
        if h * w <= 10000:
            out[f"in_{i}"] = img                         # small inputs stored too: guards the picture generator itself
        print(f"case {i}: {h}x{w} -> resize {oh}x{ow} -> crop ({top},{left})  ok")
    for j, ((h, w), (oh, ow)) in enumerate(RESIZES):
        img = P.synthetic_image(h, w, seed=h + w)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(P.resize_bicubic_u8(img, oh, ow), ref), f"resize mismatch {h}x{w}->{oh}x{ow}"
        out[f"rs_sha_{j}"] = np.asarray(sha(ref))
        if ref.size <= 40000:
            out[f"rs_{j}"] = ref
        print(f"resize {j}: {h}x{w} -> {oh}x{ow}  ok")
    # coefficient tables of the 7B path's commonest reductions, for kernel debugging
    for tag, (n_in, n_out) in {"640_298": (640, 298), "480_224": (480, 224), "53_317": (53, 317)}.items():
        xmin, cnt, kk = P.resample_coeffs(n_in, n_out)
        out[f"coef_{tag}_xmin"], out[f"coef_{tag}_cnt"], out[f"coef_{tag}_k"] = xmin, cnt, kk
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
