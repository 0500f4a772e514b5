#!/usr/bin/env python
"""Fixtures for the data path, made by the REFERENCE's own functions (build container only):

    python tests/golden/make_golden_data.py   ->  tests/golden/data_path.npz

codes/data/util.py is imported unmodified; `lmdb` (imported at util.py:6, used only by the lmdb readers) is not installed
here and is replaced by an empty stand-in module for the import.  Stored: a synthetic HR image (uint8 BGR), the output of
util.imresize_np(img / 255, 1/4), the weights / indices of util.calculate_weights_indices for the two lengths, and — for
several seeds — the crop / flip / rotate decisions and the LR / HR tensors produced by the statements of
LRHR_dataset.py:98-121 (random.randint x2, util.augment, channel swap, transpose) run verbatim on those arrays.
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    sys.path.insert(0, os.path.join(REF, "codes"))
    from data import util
    rng = np.random.default_rng(77)
    # smooth-ish synthetic image with edges, 200 x 264 (multiples of 4, both >= 128)
    h, w = 200, 264
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([np.sin(xx / 9.0) * np.cos(yy / 13.0), ((xx // 24 + yy // 16) % 2) * 1.0, np.cos((xx + yy) / 17.0)], 2)
    img_u8 = np.clip((base * 0.4 + 0.5) * 255 + rng.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)
    img = img_u8.astype(np.float32) / 255.0     # util.read_img: astype(float32) / 255.
    lr_full = util.imresize_np(img, 1 / 4, True)
    out = {"img_u8": img_u8, "lr_full": lr_full.astype(np.float32)}
    for name, (n_in, n_out) in {"H": (h, h // 4), "W": (w, w // 4)}.items():
        wts, idx, s, e = util.calculate_weights_indices(n_in, n_out, 1 / 4, "cubic", 4, True)
        out["weights_" + name] = wts.numpy().astype(np.float32)
        out["indices_" + name] = idx.numpy().astype(np.int64)
        out["sym_" + name] = np.array([s, e])
    scale, HR_size = 4, 128
    seeds = [0, 1, 2, 3, 4, 5, 6, 7]
    for sd in seeds:
        random.seed(sd)
        img_HR, img_LR = img, lr_full
        # ---- LRHR_dataset.py:93-105 verbatim (use_flip = use_rot = True as in train_ESRGANplus.json) ----
        H, W, C = img_LR.shape
        LR_size = HR_size // scale
        rnd_h = random.randint(0, max(0, H - LR_size))
        rnd_w = random.randint(0, max(0, W - LR_size))
        img_LR = img_LR[rnd_h:rnd_h + LR_size, rnd_w:rnd_w + LR_size, :]
        rnd_h_HR, rnd_w_HR = int(rnd_h * scale), int(rnd_w * scale)
        img_HR = img_HR[rnd_h_HR:rnd_h_HR + HR_size, rnd_w_HR:rnd_w_HR + HR_size, :]
        state = random.getstate()
        img_LR, img_HR = util.augment([img_LR, img_HR], True, True)
        # (recover the three decisions augment drew, for the record)
        random.setstate(state)
        flags = [random.random() < 0.5, random.random() < 0.5, random.random() < 0.5]
        # ---- LRHR_dataset.py:116-121 verbatim ----
        img_HR = img_HR[:, :, [2, 1, 0]]
        img_LR = img_LR[:, :, [2, 1, 0]]
        img_HR = torch.from_numpy(np.ascontiguousarray(np.transpose(img_HR, (2, 0, 1)))).float()
        img_LR = torch.from_numpy(np.ascontiguousarray(np.transpose(img_LR, (2, 0, 1)))).float()
        out[f"s{sd}_params"] = np.array([rnd_h, rnd_w] + [int(f) for f in flags])
        out[f"s{sd}_LR"] = img_LR.numpy()
        out[f"s{sd}_HR"] = img_HR.numpy()
    out["seeds"] = np.array(seeds)
    np.savez_compressed(os.path.join(HERE, "data_path.npz"), **out)
    print("data_path.npz", os.path.getsize(os.path.join(HERE, "data_path.npz")))


if __name__ == "__main__":
    main()
