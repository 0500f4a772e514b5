"""The training data path of one minibatch on the device: what
``codes/data/LRHR_dataset.py:83-121`` does per sample on a DataLoader worker — MATLAB-bicubic x1/4 of the whole HR image
(``data/util.py:345-412``), random 32x32 / 128x128 crop, flip / rotate (``util.py:94-106``), BGR->RGB, HWC->CHW — as ONE
kernel launch per batch over HR images that already live in device memory (csrc/esrp_data.cu, ``esrp_lrhr_batch``).

The random decisions are drawn on the host with Python's ``random`` in the reference's order (two ``randint``, then one
``random()`` per enabled flip / rotation), so ``random.seed`` reproduces the reference's crops; the bicubic tables are the
reference's ``calculate_weights_indices`` (``util.py:219-274``) evaluated once per image size.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import random
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib


class _Job(C.Structure):
    _fields_ = [("img", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("rnd_h", C.c_int32), ("rnd_w", C.c_int32),
                ("hflip", C.c_int32), ("vflip", C.c_int32), ("rot90", C.c_int32), ("wh", C.c_void_p), ("ih", C.c_void_p),
                ("ww", C.c_void_p), ("iw", C.c_void_p), ("ph", C.c_int32), ("pw", C.c_int32), ("sym_hs", C.c_int32),
                ("sym_ws", C.c_int32)]


def bicubic_tables(in_length: int, scale: float) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """util.py:219-274 (`calculate_weights_indices`, kernel 'cubic', width 4, antialiasing on) in float32 like the
    reference's torch code: (weights [out, P], first padded index of each window [out], sym_len_s, sym_len_e)."""
    out_length = math.ceil(in_length * scale)
    kw = 4.0 / scale if scale < 1 else 4.0
    x = np.linspace(1, out_length, out_length, dtype=np.float32)
    u = (x / np.float32(scale) + np.float32(0.5 * (1 - 1 / scale))).astype(np.float32)
    left = np.floor(u - np.float32(kw / 2)).astype(np.float32)
    P = math.ceil(kw) + 2
    idx = left[:, None] + np.linspace(0, P - 1, P, dtype=np.float32)[None, :]
    d = np.abs((u[:, None] - idx) * (np.float32(scale) if scale < 1 else np.float32(1.0))).astype(np.float32)
    d2, d3 = d ** 2, d ** 3
    wts = ((1.5 * d3 - 2.5 * d2 + 1) * (d <= 1) + (-0.5 * d3 + 2.5 * d2 - 4 * d + 2) * ((d > 1) & (d <= 2))).astype(np.float32)
    if scale < 1:
        wts = np.float32(scale) * wts
    wts = (wts / wts.sum(1, keepdims=True)).astype(np.float32)
    zero_cols = (wts == 0).sum(0)
    if zero_cols[0] != 0:
        idx, wts = idx[:, 1:P - 1], wts[:, 1:P - 1]
    if zero_cols[-1] != 0:
        idx, wts = idx[:, 0:P - 2], wts[:, 0:P - 2]
    sym_s = int(-idx.min() + 1)
    sym_e = int(idx.max() - in_length)
    first = (idx[:, 0] + sym_s - 1).astype(np.int32)
    return np.ascontiguousarray(wts), np.ascontiguousarray(first), sym_s, sym_e


class LRHRBatcher:
    """``batch(images)``: a list of uint8 HWC BGR images on the device (each side >= hr_size and a multiple of `scale`)
    -> (LR [B,3,hr_size/scale,...], HR [B,3,hr_size,hr_size]) fp32 RGB in [0,1], one crop per image."""

    def __init__(self, device: torch.device, scale: int = 4, hr_size: int = 128, use_flip: bool = True, use_rot: bool = True):
        self.lib = _lib.load()
        if self.lib.esrp_sizeof_lrhr_job() != C.sizeof(_Job):
            raise RuntimeError("esrp_lrhr_job_t layout mismatch between libesrp.so and data_gpu.py")
        self.device, self.scale, self.hr_size = torch.device(device), scale, hr_size
        self.use_flip, self.use_rot = use_flip, use_rot
        self._tables: Dict[int, tuple] = {}

    def _table(self, n: int):
        t = self._tables.get(n)
        if t is None:
            w, first, s, _ = bicubic_tables(n, 1.0 / self.scale)
            t = (torch.from_numpy(w).to(self.device), torch.from_numpy(first).to(self.device), w.shape[1], s)
            self._tables[n] = t
        return t

    def draw(self, h: int, w: int):
        """The reference's draws for an h x w HR image (LRHR_dataset.py:99-100, util.py:96-98), in its order."""
        lr_size = self.hr_size // self.scale
        rnd_h = random.randint(0, max(0, h // self.scale - lr_size))
        rnd_w = random.randint(0, max(0, w // self.scale - lr_size))
        hflip = self.use_flip and random.random() < 0.5
        vflip = self.use_rot and random.random() < 0.5
        rot90 = self.use_rot and random.random() < 0.5
        return rnd_h, rnd_w, bool(hflip), bool(vflip), bool(rot90)

    def batch(self, images: Sequence[torch.Tensor], params: Sequence[tuple] = None):
        if params is None:
            params = [self.draw(im.shape[0], im.shape[1]) for im in images]
        jobs = (_Job * len(images))()
        pw_max = 0
        for j, (im, (rnd_h, rnd_w, hflip, vflip, rot90)) in enumerate(zip(images, params)):
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or im.device != self.device or not im.is_contiguous():
                raise RuntimeError("LRHRBatcher: images must be contiguous uint8 HWC (BGR) tensors on " + str(self.device))
            h, w = int(im.shape[0]), int(im.shape[1])
            if h % self.scale or w % self.scale or h < self.hr_size or w < self.hr_size:
                raise RuntimeError(f"LRHRBatcher: image {h}x{w} must be >= {self.hr_size} and a multiple of {self.scale} "
                                   "(util.modcrop / the resize of LRHR_dataset.py:86-91 happen before)")
            wh, ih, ph, shs = self._table(h)
            ww, iw, pw, sws = self._table(w)
            pw_max = max(pw_max, pw)
            jobs[j] = _Job(im.data_ptr(), h, w, rnd_h, rnd_w, int(hflip), int(vflip), int(rot90), wh.data_ptr(), ih.data_ptr(),
                           ww.data_ptr(), iw.data_ptr(), ph, pw, shs, sws)
        n = len(images)
        lr_size = self.hr_size // self.scale
        jobs_dev = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(self.device)
        lr = torch.empty((n, 3, lr_size, lr_size), dtype=torch.float32, device=self.device)
        hr = torch.empty((n, 3, self.hr_size, self.hr_size), dtype=torch.float32, device=self.device)
        xw_max = (lr_size - 1) * self.scale + pw_max + 2
        with torch.cuda.device(self.device):
            _lib.check(self.lib.esrp_lrhr_batch(jobs_dev.data_ptr(), n, self.scale, self.hr_size, xw_max, lr.data_ptr(), hr.data_ptr(),
                                                torch.cuda.current_stream(self.device).cuda_stream), "esrp_lrhr_batch")
        self._keep = jobs_dev   # (the launch reads it asynchronously)
        return lr, hr
