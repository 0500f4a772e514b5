"""The device data path (esrganplus_b200/data_gpu.py over csrc/esrp_data.cu) against the oracle restatement of
LRHR_dataset.py:83-121 / data/util.py (pinned to the reference's own functions by tests/test_data_oracle.py) and against
the fixtures those reference functions produced."""
import os
import random

import numpy as np
import pytest
import torch

from esrganplus_b200.data_gpu import LRHRBatcher
from oracle import data_oracle as D

pytestmark = pytest.mark.gpu


def test_lrhr_batch_matches_reference_fixture(cuda_dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "data_path.npz"))
    img = torch.from_numpy(g["img_u8"]).to(cuda_dev)
    b = LRHRBatcher(cuda_dev, scale=4, hr_size=128)
    seeds = [int(s) for s in g["seeds"]]
    params = []
    for sd in seeds:
        random.seed(sd)
        p = b.draw(img.shape[0], img.shape[1])          # same host draws as the reference for the same seed
        assert list(map(int, p)) == [int(v) for v in g[f"s{sd}_params"]], sd
        params.append(p)
    lr, hr = b.batch([img] * len(seeds), params)
    for k, sd in enumerate(seeds):
        assert torch.equal(hr[k].cpu(), torch.from_numpy(g[f"s{sd}_HR"])), sd              # pure gather: bit exact
        d = (lr[k].cpu() - torch.from_numpy(g[f"s{sd}_LR"])).abs().max().item()
        assert d <= 3e-6, (sd, d)                                                          # fp32 sums in another order


@pytest.mark.parametrize("shape", [(128, 128), (132, 260), (480, 360)])
def test_lrhr_batch_matches_oracle_on_ragged_sizes(cuda_dev, shape):
    """Image borders (symmetric padding), the smallest legal image, non-square sizes; every flip / rotation combination."""
    h, w = shape
    rng = np.random.default_rng(h * w)
    imgs_np = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(8)]
    b = LRHRBatcher(cuda_dev, scale=4, hr_size=128)
    params = []
    for k in range(8):
        corner_h = [0, h // 4 - 32][k & 1]
        corner_w = [0, w // 4 - 32][(k >> 1) & 1]
        params.append((corner_h, corner_w, bool(k & 1), bool(k & 2), bool(k & 4)))
    lr, hr = b.batch([torch.from_numpy(a).to(cuda_dev) for a in imgs_np], params)
    for k in range(8):
        rl, rh = D.lrhr_sample(imgs_np[k].astype(np.float32) / 255.0, 4, 128, params[k])
        assert torch.equal(hr[k].cpu(), torch.from_numpy(rh)), k
        assert (lr[k].cpu() - torch.from_numpy(rl)).abs().max().item() <= 3e-6, k


def test_lrhr_batch_rejects_bad_images(cuda_dev):
    b = LRHRBatcher(cuda_dev)
    with pytest.raises(RuntimeError):
        b.batch([torch.zeros(100, 128, 3, dtype=torch.uint8, device=cuda_dev)])
    with pytest.raises(RuntimeError):
        b.batch([torch.zeros(128, 128, 3, dtype=torch.uint8)])
