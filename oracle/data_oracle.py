"""TEST INFRASTRUCTURE — numpy restatement of the reference's training data path for one sample (SURVEY.md section 8f
rank 4): MATLAB-style antialiased bicubic downscale (codes/data/util.py:211-274 `cubic` / `calculate_weights_indices`,
:345-412 `imresize_np`), the random crop of LR and HR (codes/data/LRHR_dataset.py:98-105), flip / rotate augmentation
(util.py:94-106 `augment`) and BGR->RGB, HWC->CHW (LRHR_dataset.py:116-121).  Pinned against outputs of the reference's own
functions (tests/golden/make_golden_data.py -> data_path.npz).  Only tests/, smoke() and bench.py's baseline legs may import it.
"""
from __future__ import annotations

import math
import random
from typing import List, Tuple

import numpy as np


def cubic(x: np.ndarray) -> np.ndarray:
    """util.py:211-216 (float32 like the reference's torch tensors)."""
    absx = np.abs(x).astype(np.float32)
    absx2, absx3 = absx ** 2, absx ** 3
    return ((1.5 * absx3 - 2.5 * absx2 + 1) * (absx <= 1).astype(np.float32) +
            (-0.5 * absx3 + 2.5 * absx2 - 4 * absx + 2) * ((absx > 1) & (absx <= 2)).astype(np.float32)).astype(np.float32)


def calculate_weights_indices(in_length: int, out_length: int, scale: float, kernel_width: float = 4.0,
                              antialiasing: bool = True) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """util.py:219-274.  Returns (weights [out, P'], indices [out, P'] into the symmetric-padded input, sym_len_s, sym_len_e)."""
    if scale < 1 and antialiasing:
        kernel_width = kernel_width / scale
    x = np.linspace(1, out_length, out_length, dtype=np.float32)
    u = (x / np.float32(scale) + np.float32(0.5 * (1 - 1 / scale))).astype(np.float32)
    left = np.floor(u - np.float32(kernel_width / 2)).astype(np.float32)
    P = math.ceil(kernel_width) + 2
    indices = left[:, None] + np.linspace(0, P - 1, P, dtype=np.float32)[None, :]
    dist = u[:, None] - indices
    if scale < 1 and antialiasing:
        weights = np.float32(scale) * cubic(dist * np.float32(scale))
    else:
        weights = cubic(dist)
    weights = (weights / weights.sum(1, keepdims=True)).astype(np.float32)
    zero_cols = (weights == 0).sum(0)
    if zero_cols[0] != 0:
        indices, weights = indices[:, 1:P - 1], weights[:, 1:P - 1]
    if zero_cols[-1] != 0:
        indices, weights = indices[:, 0:P - 2], weights[:, 0:P - 2]
    sym_len_s = int(-indices.min() + 1)
    sym_len_e = int(indices.max() - in_length)
    indices = indices + sym_len_s - 1
    return np.ascontiguousarray(weights), np.ascontiguousarray(indices.astype(np.int64)), sym_len_s, sym_len_e


def _sym_pad(img: np.ndarray, s: int, e: int, axis: int) -> np.ndarray:
    """The reference's 'symmetric copying' (util.py:371-383): s mirrored rows in front, e behind."""
    n = img.shape[axis]
    head = np.flip(np.take(img, range(0, s), axis=axis), axis=axis)
    tail = np.flip(np.take(img, range(n - e, n), axis=axis), axis=axis) if e > 0 else np.take(img, [], axis=axis)
    return np.concatenate([head, img, tail], axis=axis)


def imresize_np(img: np.ndarray, scale: float, antialiasing: bool = True) -> np.ndarray:
    """util.py:345-412: HWC float image in, HWC float32 out; H pass first, then W pass."""
    img = img.astype(np.float32)
    in_h, in_w, _ = img.shape
    out_h, out_w = math.ceil(in_h * scale), math.ceil(in_w * scale)
    wh, ih, hs, he = calculate_weights_indices(in_h, out_h, scale, 4.0, antialiasing)
    ww, iw, ws, we = calculate_weights_indices(in_w, out_w, scale, 4.0, antialiasing)
    aug = _sym_pad(img, hs, he, 0)
    k = wh.shape[1]
    out1 = np.stack([np.tensordot(wh[i], aug[ih[i, 0]:ih[i, 0] + k], axes=(0, 0)) for i in range(out_h)], 0).astype(np.float32)
    aug = _sym_pad(out1, ws, we, 1)
    k = ww.shape[1]
    out2 = np.stack([np.tensordot(aug[:, iw[j, 0]:iw[j, 0] + k], ww[j], axes=(1, 0)) for j in range(out_w)], 1).astype(np.float32)
    return out2


def draw_sample_params(h_lr: int, w_lr: int, lr_size: int, use_flip: bool = True, use_rot: bool = True):
    """The random draws of one training sample in the reference's order: LRHR_dataset.py:99-100 (randint h, randint w),
    then util.py:96-98 (random() for hflip, vflip, rot90 — each only drawn when its option is on, `a and random() < .5`)."""
    rnd_h = random.randint(0, max(0, h_lr - lr_size))
    rnd_w = random.randint(0, max(0, w_lr - lr_size))
    hflip = use_flip and random.random() < 0.5
    vflip = use_rot and random.random() < 0.5
    rot90 = use_rot and random.random() < 0.5
    return rnd_h, rnd_w, bool(hflip), bool(vflip), bool(rot90)


def augment(img: np.ndarray, hflip: bool, vflip: bool, rot90: bool) -> np.ndarray:
    """util.py:100-104 on an HWC image."""
    if hflip:
        img = img[:, ::-1, :]
    if vflip:
        img = img[::-1, :, :]
    if rot90:
        img = img.transpose(1, 0, 2)
    return img


def lrhr_sample(img_hr_bgr01: np.ndarray, scale: int, hr_size: int, params) -> Tuple[np.ndarray, np.ndarray]:
    """LRHR_dataset.py:83-121 for one HR image (HWC BGR float [0,1], both sides >= hr_size and multiples of scale) with
    on-the-fly LR (no random_scale): whole-image bicubic, crop, augment, BGR->RGB, HWC->CHW.  Returns (LR, HR) CHW float32."""
    rnd_h, rnd_w, hflip, vflip, rot90 = params
    lr = imresize_np(img_hr_bgr01, 1.0 / scale, True)
    lr_size = hr_size // scale
    lr = lr[rnd_h:rnd_h + lr_size, rnd_w:rnd_w + lr_size, :]
    hr = img_hr_bgr01[rnd_h * scale:rnd_h * scale + hr_size, rnd_w * scale:rnd_w * scale + hr_size, :]
    lr, hr = augment(lr, hflip, vflip, rot90), augment(hr, hflip, vflip, rot90)
    to_chw = lambda a: np.ascontiguousarray(np.transpose(a[:, :, [2, 1, 0]], (2, 0, 1))).astype(np.float32)
    return to_chw(lr), to_chw(hr)
