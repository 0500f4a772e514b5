#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the REFERENCE ITSELF.

Run in the build container only (needs /root/reference; that path does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules (codes/models/modules/{block,architecture}.py and test_image/*.py) are
imported unmodified; the only intervention is an in-memory replacement of
GaussianNoise.__init__, whose original hard-codes `.to(torch.device('cuda'))` (block.py:115) and
so cannot even be constructed without an NVIDIA driver.  Weights are synthetic and deterministic
(numpy PCG64 via oracle.esrgan_oracle.synth_state_dict_*), loaded with load_state_dict(strict=True)
so the key set itself is validated by the reference.  Outputs are stored as float32 .npz.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import esrgan_oracle as O  # noqa: E402


def import_reference():
    sys.path.insert(0, os.path.join(REF, "codes"))
    import models.modules.architecture as arch
    import models.modules.block as B

    def _init(self, sigma=0.1, is_relative_detach=False):
        nn.Module.__init__(self)
        self.sigma = sigma
        self.is_relative_detach = is_relative_detach
        self.noise = torch.tensor(0, dtype=torch.float)

    B.GaussianNoise.__init__ = _init
    return arch, B


def import_reference_test_image():
    """test_image/{architecture,block}.py are top-level modules named like the codes/ ones."""
    import importlib.util
    spec_b = importlib.util.spec_from_file_location("block", os.path.join(REF, "test_image", "block.py"))
    tb = importlib.util.module_from_spec(spec_b)
    saved = sys.modules.get("block")
    sys.modules["block"] = tb
    spec_b.loader.exec_module(tb)

    def _init(self, sigma=0.1, is_relative_detach=False):
        nn.Module.__init__(self)
        self.sigma = sigma
        self.is_relative_detach = is_relative_detach
        self.noise = torch.tensor(0, dtype=torch.float)

    tb.GaussianNoise.__init__ = _init
    spec_a = importlib.util.spec_from_file_location("architecture_ti", os.path.join(REF, "test_image", "architecture.py"))
    ta = importlib.util.module_from_spec(spec_a)
    spec_a.loader.exec_module(ta)
    if saved is not None:
        sys.modules["block"] = saved
    return ta, tb


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def rng_tensor(seed, *shape, lo=0.0, hi=1.0):
    r = np.random.default_rng(seed)
    return torch.from_numpy(r.uniform(lo, hi, shape).astype("float32"))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    arch, B = import_reference()
    meta = {"torch": torch.__version__, "reference": "ncarraz/ESRGANplus @ /root/reference"}

    # ---- structural fixture: key / shape / dtype dumps ---------------------------
    g = arch.RRDBNet(3, 3, 64, 23, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA",
                     upsample_mode="upconv")
    d = arch.Discriminator_VGG_128(3, 64, norm_type="batch", act_type="leakyrelu", mode="CNA")
    struct = {
        "G_nb23_nf64": [[k, list(v.shape), str(v.dtype)] for k, v in g.state_dict().items()],
        "D_vgg128": [[k, list(v.shape), str(v.dtype)] for k, v in d.state_dict().items()],
        "G_num_params": sum(p.numel() for p in g.parameters()),
        "D_num_params": sum(p.numel() for p in d.parameters()),
        "G_repr_sha256": hashlib.sha256(str(g).encode()).hexdigest(),
        "D_repr_sha256": hashlib.sha256(str(d).encode()).hexdigest(),
    }
    with open(os.path.join(HERE, "structure.json"), "w") as f:
        json.dump(struct, f)
    del g

    # ---- RDB / RRDB (block.py:232-291) ------------------------------------------------------------
    sd_small = O.synth_state_dict_g(3, 3, 64, 1, seed=11)
    rdb = B.ResidualDenseBlock_5C(64, 3, 32, 1, True, "zero", None, "leakyrelu", "CNA")
    rdb_sd = {k[len("model.1.sub.0.RDB1."):]: v for k, v in sd_small.items() if k.startswith("model.1.sub.0.RDB1.")}
    rdb.load_state_dict(rdb_sd, strict=True)
    x = rng_tensor(101, 2, 64, 16, 16, lo=-1, hi=1)
    rdb.eval()
    with torch.no_grad():
        y_eval = rdb(x)
    # train mode with an injected noise tensor: patch forward of the noise module only for this call
    nz = torch.from_numpy(np.random.default_rng(202).standard_normal((2, 64, 16, 16)).astype("float32"))
    rdb.train()
    orig_fwd = B.GaussianNoise.forward

    def fwd_injected(self, t):
        if self.training and self.sigma != 0:
            scale = self.sigma * t.detach() if self.is_relative_detach else self.sigma * t
            return t + nz * scale
        return t

    B.GaussianNoise.forward = fwd_injected
    xg = x.clone().requires_grad_(True)
    y_train = rdb(xg)
    gy = rng_tensor(303, 2, 64, 16, 16, lo=-1, hi=1)
    y_train.backward(gy)
    B.GaussianNoise.forward = orig_fwd
    grads = {k: p.grad.detach().clone() for k, p in rdb.named_parameters()}
    np.savez_compressed(
        os.path.join(HERE, "rdb64.npz"), x=x.numpy(), y_eval=y_eval.numpy(), noise=nz.numpy(),
        y_train=y_train.detach().numpy(), gy=gy.numpy(), gx=xg.grad.numpy(),
        **{"gw_norm." + k: np.array([v.norm().item(), v.sum().item(), v.abs().max().item()], dtype="float64") for k, v in grads.items()},
        **{"gw_head." + k: v.flatten()[:64].numpy() for k, v in grads.items()},
        sd_digest=np.array(sd_digest(rdb_sd)))

    rrdb = B.RRDB(64, 3, 32, 1, True, "zero", None, "leakyrelu", "CNA")
    rrdb_sd = {k[len("model.1.sub.0."):]: v for k, v in sd_small.items() if k.startswith("model.1.sub.0.")}
    rrdb.load_state_dict(rrdb_sd, strict=True)
    rrdb.eval()
    xr = rng_tensor(102, 1, 64, 12, 20, lo=-1, hi=1)
    with torch.no_grad():
        yr = rrdb(xr)
    np.savez_compressed(os.path.join(HERE, "rrdb64.npz"), x=xr.numpy(), y_eval=yr.numpy(), sd_digest=np.array(sd_digest(rrdb_sd)))

    # ---- RRDBNet config 1: nb=1 nf=32 on one 32x32 tile (test_image/test.py plumbing) --------------
    sd_c1 = O.synth_state_dict_g(3, 3, 32, 1, seed=21)
    net = arch.RRDBNet(3, 3, 32, 1, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA", upsample_mode="upconv")
    net.load_state_dict(sd_c1, strict=True)
    net.eval()
    x1 = rng_tensor(0, 1, 3, 32, 32)
    with torch.no_grad():
        y1 = net(x1)
    # uint8 plumbing of test_image/test.py:31-40
    img_u8 = (x1[0].permute(1, 2, 0).numpy()[:, :, ::-1] * 255.0).round().astype("uint8")  # HWC BGR
    img = img_u8 * 1.0 / 255
    t_in = torch.from_numpy(np.transpose(img[:, :, [2, 1, 0]], (2, 0, 1))).float().unsqueeze(0)
    with torch.no_grad():
        out = net(t_in).data.squeeze().float().cpu().clamp_(0, 1).numpy()
    out = np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))
    out_u8 = (out * 255.0).round().astype("uint8")
    np.savez_compressed(os.path.join(HERE, "rrdbnet_c1_nb1_nf32.npz"), x=x1.numpy(), y=y1.numpy(), img_u8=img_u8, out_u8=out_u8,
                        sd_digest=np.array(sd_digest(sd_c1)))

    # same weights through the standalone test_image copy (RRDB_Net): identical keys, identical eval output
    ta, _tb = import_reference_test_image()
    net_ti = ta.RRDB_Net(3, 3, 32, 1, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA", res_scale=1, upsample_mode="upconv")
    net_ti.load_state_dict(sd_c1, strict=True)
    net_ti.eval()
    with torch.no_grad():
        y1_ti = net_ti(x1)
    meta["test_image_RRDB_Net_equals_RRDBNet_eval"] = bool(torch.equal(y1, y1_ti))
    meta["test_image_keys_equal"] = list(net_ti.state_dict().keys()) == list(net.state_dict().keys())

    # ---- RRDBNet nb=23 nf=64 (the benchmark architecture) on small tiles ---------------------------
    sd_full = O.synth_state_dict_g(3, 3, 64, 23, seed=31)
    net = arch.RRDBNet(3, 3, 64, 23, gc=32, upscale=4, norm_type=None, act_type="leakyrelu", mode="CNA", upsample_mode="upconv")
    net.load_state_dict(sd_full, strict=True)
    net.eval()
    x2 = rng_tensor(1, 1, 3, 24, 24)
    x3 = rng_tensor(2, 2, 3, 19, 37)   # ragged, not a multiple of any tile size
    with torch.no_grad():
        y2 = net(x2)
        y3 = net(x3)
    np.savez_compressed(os.path.join(HERE, "rrdbnet_nb23_nf64.npz"), x24=x2.numpy(), y24=y2.numpy(), x_ragged=x3.numpy(),
                        y_ragged=y3.numpy(), sd_digest=np.array(sd_digest(sd_full)))

    # ---- Discriminator_VGG_128 (architecture.py:87-129), eval + one train-mode forward -------------
    sd_d = O.synth_state_dict_d(3, 64, seed=41)
    d.load_state_dict(sd_d, strict=True)
    xd = rng_tensor(3, 4, 3, 128, 128)
    d.eval()
    with torch.no_grad():
        yd_eval = d(xd)
    d.train()
    xdg = xd.clone().requires_grad_(True)
    yd_train = d(xdg)
    yd_train.sum().backward()
    after = d.state_dict()
    np.savez_compressed(
        os.path.join(HERE, "dvgg128.npz"), x=xd.numpy(), y_eval=yd_eval.numpy(), y_train=yd_train.detach().numpy(),
        gx_train=xdg.grad.numpy().astype("float32"),
        **{"after." + k: v.numpy() for k, v in after.items() if "running" in k or "num_batches" in k},
        sd_digest=np.array(sd_digest(sd_d)))

    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
