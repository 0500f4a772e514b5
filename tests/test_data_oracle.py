"""The data-path oracle (oracle/data_oracle.py) against fixtures made by the reference's own codes/data/util.py and the
statements of LRHR_dataset.py:93-121 (tests/golden/make_golden_data.py)."""
import os
import random

import numpy as np

from oracle import data_oracle as D


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "data_path.npz"))


def test_weights_and_indices_match_reference(golden_dir):
    g = _g(golden_dir)
    h, w = g["img_u8"].shape[:2]
    for name, n_in in (("H", h), ("W", w)):
        wts, idx, s, e = D.calculate_weights_indices(n_in, n_in // 4, 1 / 4, 4.0, True)
        assert (s, e) == tuple(int(v) for v in g["sym_" + name])
        assert np.array_equal(idx, g["indices_" + name])
        assert np.abs(wts - g["weights_" + name]).max() <= 1e-7


def test_imresize_matches_reference(golden_dir):
    g = _g(golden_dir)
    img = g["img_u8"].astype(np.float32) / 255.0
    lr = D.imresize_np(img, 1 / 4, True)
    assert lr.shape == g["lr_full"].shape
    assert np.abs(lr - g["lr_full"]).max() <= 2e-6


def test_sample_draws_and_tensors_match_reference(golden_dir):
    g = _g(golden_dir)
    img = g["img_u8"].astype(np.float32) / 255.0
    h, w = img.shape[0] // 4, img.shape[1] // 4
    kinds = set()
    for sd in g["seeds"]:
        random.seed(int(sd))
        params = D.draw_sample_params(h, w, 32, True, True)
        assert list(map(int, params)) == [int(v) for v in g[f"s{sd}_params"]], sd
        kinds.add(params[2:])
        lr, hr = D.lrhr_sample(img, 4, 128, params)
        assert lr.shape == (3, 32, 32) and hr.shape == (3, 128, 128)
        assert np.abs(lr - g[f"s{sd}_LR"]).max() <= 2e-6, sd
        assert np.array_equal(hr, g[f"s{sd}_HR"]), sd
    assert len(kinds) >= 4, "the seeds must exercise several flip / rotate combinations"
