"""GPU-side data path of the reference's training scripts.

`datasets/collator.py:11-22` hands a list of PIL images to HF `ViTImageProcessor` on the data-loader workers: resize to
224 x 224 (Pillow BILINEAR), rescale by 1/255, normalise with the checkpoint's mean / std, and ships fp32 tensors
(602 KB per image) to the GPU.  Here the workers only stack the raw uint8 pixels (3 KB per CIFAR image); the resize and
the normalisation run on the GPU (`odevit_preprocess_u8`, csrc/preprocess.cu) with Pillow's fixed-point arithmetic
restated exactly -- the resized uint8 image is bit-identical to `PIL.Image.resize`.

    processor = GpuImageProcessor(size=224, image_mean=[0.485, 0.456, 0.406], image_std=[0.229, 0.224, 0.225])
    collator = Collator(processor)                       # same name, same output dict as the reference's
    loader = DataLoader(dataset, collate_fn=collator.classification_collate_fn, ...)
    for data in loader:
        out = model(**data["pixel_values"], labels=data["labels"].cuda())      # train.py:40-53 unchanged
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_vp = ctypes.c_void_p


class PixelBatch(dict):
    """The processor's return value: a dict with the `.to(device)` of HF's `BatchFeature` (train.py:44 calls it)."""

    def to(self, *args, **kwargs):
        return PixelBatch({k: (v.to(*args, **kwargs) if torch.is_tensor(v) else v) for k, v in self.items()})


class GpuImageProcessor:
    """`ViTImageProcessor(do_resize, do_rescale, do_normalize)` for uint8 RGB images of ONE size per call, on the GPU.
    Call it with a uint8 tensor / array [B,H,W,3] or a list of PIL images / HWC arrays; returns {"pixel_values":
    fp32 CUDA tensor [B,3,S,S]} (a dict, so `model(**processed)` works as with the HF `BatchFeature`)."""

    def __init__(self, size: int | Tuple[int, int] | Dict[str, int] = 224, image_mean: Sequence[float] = (0.485, 0.456, 0.406),
                 image_std: Sequence[float] = (0.229, 0.224, 0.225), rescale_factor: float = 1 / 255,
                 device: Optional[torch.device | str] = None):
        if isinstance(size, dict):
            size = (int(size["height"]), int(size["width"]))
        elif isinstance(size, int):
            size = (size, size)
        self.size = (int(size[0]), int(size[1]))
        self.image_mean = tuple(float(np.float32(m)) for m in image_mean)
        self.image_std = tuple(float(np.float32(s)) for s in image_std)
        self.rescale_factor = float(np.float32(rescale_factor))
        self.device = torch.device(device) if device is not None else None
        self._tables: Dict[Tuple[int, int, str], Tuple[torch.Tensor, ...]] = {}

    @staticmethod
    def host_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
        """Pillow's BILINEAR coefficient tables (bounds [out,2], kk [out,ksize], int32) for in_size -> out_size."""
        L = _lib.lib()
        ksize = L.odevit_pil_bilinear_ksize(in_size, out_size)
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        kk = np.zeros((out_size, ksize), dtype=np.int32)
        st = L.odevit_pil_bilinear_tables(in_size, out_size, bounds.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                          kk.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        _lib.check(st, "odevit_pil_bilinear_tables")
        return bounds, kk

    def _device_tables(self, H: int, W: int, dev: torch.device):
        key = (H, W, str(dev))
        t = self._tables.get(key)
        if t is None:
            bh, kh = self.host_tables(W, self.size[1])
            bv, kv = self.host_tables(H, self.size[0])
            t = tuple(torch.from_numpy(a).to(dev) for a in (bh, kh, bv, kv))
            self._tables[key] = t
        return t

    @staticmethod
    def stack_uint8(images) -> torch.Tensor:
        """[B,H,W,3] uint8 CPU tensor from a list of PIL images / HWC uint8 arrays of one size."""
        if torch.is_tensor(images):
            return images
        if isinstance(images, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(images))
        arrs = [np.asarray(im.convert("RGB") if hasattr(im, "convert") else im, dtype=np.uint8) for im in images]
        if len({a.shape for a in arrs}) != 1:
            raise ValueError("GpuImageProcessor: the images of one call must share a size (group them by size)")
        return torch.from_numpy(np.stack(arrs))

    def __call__(self, images, return_tensors: str = "pt", return_uint8: bool = False) -> Dict[str, torch.Tensor]:
        x = self.stack_uint8(images)
        if x.dtype != torch.uint8 or x.ndim != 4 or x.shape[-1] != 3:
            raise ValueError(f"expected uint8 [B,H,W,3] RGB images, got {tuple(x.shape)} {x.dtype}")
        dev = self.device or (x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        x = x.to(dev, non_blocking=True).contiguous()
        B, H, W, _ = x.shape
        Sh, Sw = self.size
        bh, kh, bv, kv = self._device_tables(H, W, dev)
        tmp = torch.empty(B, H, Sw, 3, dtype=torch.uint8, device=dev)
        out = torch.empty(B, 3, Sh, Sw, dtype=torch.float32, device=dev)
        out_u8 = torch.empty(B, Sh, Sw, 3, dtype=torch.uint8, device=dev) if return_uint8 else None
        mean = (ctypes.c_float * 3)(*self.image_mean)
        std = (ctypes.c_float * 3)(*self.image_std)
        with torch.cuda.device(dev):
            st = _lib.lib().odevit_preprocess_u8(_vp(x.data_ptr()), B, H, W, Sh, Sw, _vp(bh.data_ptr()), _vp(kh.data_ptr()),
                                                 _vp(bv.data_ptr()), _vp(kv.data_ptr()), self.rescale_factor, mean, std,
                                                 _vp(tmp.data_ptr()), _vp(out.data_ptr()),
                                                 _vp(out_u8.data_ptr()) if out_u8 is not None else None,
                                                 _vp(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(st, "odevit_preprocess_u8")
        res = PixelBatch({"pixel_values": out})
        if return_uint8:
            res["resized_uint8"] = out_u8
        return res


class Collator:
    """datasets/collator.py:6-22 with the processor on the GPU: same constructor, same `classification_collate_fn`
    output keys.  With `defer=True` (what a multi-worker DataLoader needs: CUDA may not be touched in a forked worker)
    the collate function only stacks the uint8 pixels under "uint8_images" and `finish(batch)` runs the processor in
    the training process."""

    def __init__(self, processor: GpuImageProcessor, pad_token: Optional[int] = -100, defer: bool = False):
        self.processor = processor
        self._pad_token = pad_token
        self.defer = defer

    def classification_collate_fn(self, batch):
        pixel_values = [item[0] for item in batch]
        labels = [item[1] for item in batch]
        images = [item[0] for item in batch]
        out = {"labels": torch.tensor(labels), "raw_images": images}
        if self.defer:
            out["uint8_images"] = GpuImageProcessor.stack_uint8(pixel_values)
        else:
            out["pixel_values"] = self.processor(pixel_values, return_tensors="pt")
        return out

    def finish(self, batch):
        if "pixel_values" not in batch:
            batch["pixel_values"] = self.processor(batch.pop("uint8_images"), return_tensors="pt")
        return batch
