"""`VclaImageProcessor`: the CLIP image pre-processing of the reference, run on the GPU.

The reference builds HF's `CLIPImageProcessor` (ref: models/visualcla/modeling_utils.py:130) and calls it on the host for
every request (:150-152, :187-189): convert RGB -> resize(shortest_edge, BICUBIC) -> center_crop -> 1/255 -> (x-mean)/std,
all in PIL/numpy.  This class has the same call surface (`proc(image, return_tensors='pt').pixel_values`,
`.size["shortest_edge"]`, `.patch_size`, `.image_mean/.image_std`, `from_pretrained`) but uploads the raw RGB bytes and runs
csrc/preprocess.cu (`vcla_preprocess_image`): Pillow's 8-bit bicubic resample reproduced bit for bit, only the cropped
window computed, result left in HBM in the dtype the vision tower wants.

Opt-in: `get_model_and_tokenizer_and_processor` returns the HF processor unless the environment
sets VCLA_GPU_PREPROCESS=1; or swap it by hand with
    model.image_processor = VclaImageProcessor.from_pretrained(vision_dir, patch_size=model.image_processor.patch_size)
There is no CPU fallback: without libvcla.so or a CUDA device the call raises.
"""
import ctypes as C
import json
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_DTYPES = {torch.float32: N.VCLA_F32, torch.float16: N.VCLA_F16, torch.bfloat16: N.VCLA_BF16}


class _Features(dict):
    """Minimal BatchFeature: dict with attribute access (`.pixel_values`) and `.to()`."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, *a, **k):
        return _Features({n: v.to(*a, **k) for n, v in self.items()})


class VclaImageProcessor:
    model_input_names = ["pixel_values"]

    def __init__(self, size=224, crop_size=None, image_mean: Sequence[float] = OPENAI_CLIP_MEAN,
                 image_std: Sequence[float] = OPENAI_CLIP_STD, resample=3, do_resize=True, do_center_crop=True,
                 do_rescale=True, rescale_factor=1 / 255, do_normalize=True, do_convert_rgb=True, patch_size=None,
                 device=None, dtype=torch.float32, **_unused):
        side = size["shortest_edge"] if isinstance(size, dict) else int(size)
        crop = crop_size if crop_size is not None else side
        if isinstance(crop, dict):
            if crop.get("height") != crop.get("width"):
                raise ValueError(f"only square crops are supported, got {crop}")
            crop = crop["height"]
        # the kernel implements exactly CLIP's pipeline; anything else would silently change the pixels
        if not (do_resize and do_center_crop and do_rescale and do_normalize):
            raise ValueError("VclaImageProcessor implements resize+center_crop+rescale+normalize only (all must be enabled)")
        if int(resample) != 3:
            raise ValueError(f"only BICUBIC (3) resampling is implemented, got resample={resample}")
        if int(crop) != side:
            raise ValueError(f"crop_size {crop} must equal size.shortest_edge {side}")
        if abs(rescale_factor - 1 / 255) > 1e-12:
            raise ValueError("rescale_factor must be 1/255")
        if dtype not in _DTYPES:
            raise ValueError(f"unsupported dtype {dtype}")
        self.size = {"shortest_edge": side}
        self.crop_size = {"height": side, "width": side}
        self.image_mean, self.image_std = [float(v) for v in image_mean], [float(v) for v in image_std]
        self.resample, self.rescale_factor, self.do_convert_rgb = 3, 1 / 255, bool(do_convert_rgb)
        self.patch_size = patch_size
        self.device = torch.device(device) if device is not None else None
        self.dtype = dtype
        self._mean = (C.c_float * 3)(*np.asarray(self.image_mean, np.float32).tolist())
        self._std = (C.c_float * 3)(*np.asarray(self.image_std, np.float32).tolist())
        self._workspace: Optional[torch.Tensor] = None
        self.launches = 0            # kernels enqueued so far (2 per picture)

    # ---- construction ------------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path: str, **kwargs) -> "VclaImageProcessor":
        """Reads `preprocessor_config.json` the way CLIPImageProcessor.from_pretrained does (ref :130)."""
        f = os.path.join(path, "preprocessor_config.json")
        if not os.path.isfile(f):
            raise EnvironmentError(f"{f} not found")
        with open(f) as fh:
            cfg = json.load(fh)
        keep = ("size", "crop_size", "image_mean", "image_std", "resample", "do_resize", "do_center_crop", "do_rescale",
                "rescale_factor", "do_normalize", "do_convert_rgb")
        args = {k: cfg[k] for k in keep if k in cfg}
        if isinstance(args.get("size"), int):                 # pre-4.25 configs: {"size": 224, "crop_size": 224}
            args["size"] = {"shortest_edge": args["size"]}
        args.update(kwargs)
        return cls(**args)

    def to_dict(self):
        return {"image_processor_type": "CLIPImageProcessor", "size": dict(self.size), "crop_size": dict(self.crop_size),
                "image_mean": list(self.image_mean), "image_std": list(self.image_std), "resample": 3, "do_resize": True,
                "do_center_crop": True, "do_rescale": True, "rescale_factor": 1 / 255, "do_normalize": True,
                "do_convert_rgb": self.do_convert_rgb}

    def save_pretrained(self, path: str):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "preprocessor_config.json"), "w") as fh:
            json.dump(self.to_dict(), fh, indent=2)

    # ---- host side -----------------------------------------------------------------------------------------------------
    def _rgb_bytes(self, image) -> torch.Tensor:
        """One picture -> contiguous uint8 (H, W, 3) torch tensor (host or device)."""
        if isinstance(image, torch.Tensor):
            t = image
        else:
            if hasattr(image, "convert"):                     # PIL.Image: HF convert_to_rgb (do_convert_rgb)
                if self.do_convert_rgb and image.mode != "RGB":
                    image = image.convert("RGB")
                image = np.array(image)                       # writable copy (PIL exposes a read-only buffer)
            t = torch.from_numpy(np.ascontiguousarray(image))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
            raise ValueError(f"expected an RGB uint8 picture of shape (H, W, 3), got {tuple(t.shape)} {t.dtype}")
        return t.contiguous()

    def workspace_bytes(self, height: int, width: int) -> int:
        n = N.load().vcla_preprocess_workspace_bytes(int(height), int(width), self.size["shortest_edge"])
        if n < 0:
            raise N.NativeError(f"preprocess: {N.load().vcla_last_error().decode()}")
        return int(n)

    def __call__(self, images, return_tensors="pt", **_unused):
        if return_tensors not in ("pt", None):
            raise ValueError("VclaImageProcessor returns device tensors only (return_tensors='pt')")
        lib = N.load()
        if not torch.cuda.is_available():
            raise N.NativeError("VclaImageProcessor needs a CUDA device (no CPU fallback)")
        pics = images if isinstance(images, (list, tuple)) else [images]
        if len(pics) == 0:
            raise ValueError("no images")
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        side = self.size["shortest_edge"]
        out = torch.empty(len(pics), 3, side, side, dtype=self.dtype, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for i, pic in enumerate(pics):
                rgb = self._rgb_bytes(pic).to(dev)
                h, w = int(rgb.shape[0]), int(rgb.shape[1])
                need = self.workspace_bytes(h, w)
                if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                    self._workspace = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)
                N.check(lib.vcla_preprocess_image(N.ptr(rgb), h, w, side, self._mean, self._std, N.ptr(self._workspace),
                                                  self._workspace.numel(), C.c_void_p(out[i].data_ptr()), _DTYPES[self.dtype],
                                                  stream), "vcla_preprocess_image")
                self.launches += 2
        return _Features(pixel_values=out)

    preprocess = __call__

    @staticmethod
    def resample_taps(in_size: int, out_size: int):
        """Pillow's tap table of one axis, as the kernels use it: (first[out], count[out], taps[out, ksize]) int32.
        Host-only (no GPU)."""
        lib = N.load()
        ks = lib.vcla_resample_taps(int(in_size), int(out_size), None, None, None, 0)
        if ks < 0:
            raise N.NativeError(lib.vcla_last_error().decode())
        first, count = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32)
        taps = np.zeros((out_size, ks), np.int32)
        rc = lib.vcla_resample_taps(int(in_size), int(out_size), first.ctypes.data, count.ctypes.data, taps.ctypes.data, ks)
        if rc != ks:
            raise N.NativeError(lib.vcla_last_error().decode())
        return first, count, taps
