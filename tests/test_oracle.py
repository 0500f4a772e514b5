"""The CPU oracle (oracle/esrgan_oracle.py) against fixtures produced by the REFERENCE ITSELF
(tests/golden/make_golden.py imports /root/reference with the GaussianNoise ctor shim).

Tolerance: the oracle issues the same ATen fp32 ops on the same operands as the reference modules, so
forward values are required to be bit-identical (atol=0) when run with the same thread count; oneDNN
may pick a different blocking on a different host, so the assertion allows 2e-6 * max|ref| there."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import esrgan_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


def _close(a: torch.Tensor, b_np, tol=2e-6):
    b = torch.from_numpy(np.asarray(b_np))
    scale = max(1.0, b.abs().max().item())
    err = (a - b).abs().max().item()
    assert err <= tol * scale, f"max|d|={err} scale={scale}"


def _rdb_sd(seed=11):
    sd = O.synth_state_dict_g(3, 3, 64, 1, seed=seed)
    return sd


def test_rdb_eval_and_injected_noise(golden_dir):
    g = _load(golden_dir, "rdb64.npz")
    sd = _rdb_sd()
    x = torch.from_numpy(g["x"])
    y = O.rdb_forward(x, sd, "model.1.sub.0.RDB1.")
    _close(y, g["y_eval"])
    nz = torch.from_numpy(g["noise"])
    xg = x.clone().requires_grad_(True)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("model.1.sub.0.RDB1.")}
    yt = O.rdb_forward(xg, sdg, "model.1.sub.0.RDB1.", training=True, noise=nz)
    _close(yt.detach(), g["y_train"])
    yt.backward(torch.from_numpy(g["gy"]))
    _close(xg.grad, g["gx"], tol=1e-5)
    for k, v in sdg.items():
        short = k[len("model.1.sub.0.RDB1."):]
        head = g["gw_head." + short]
        _close(v.grad.flatten()[:64], head, tol=1e-4)
        norm, s, mx = g["gw_norm." + short]
        assert abs(v.grad.norm().item() - norm) <= 1e-4 * max(1.0, norm)


def test_rrdb_eval(golden_dir):
    g = _load(golden_dir, "rrdb64.npz")
    sd = _rdb_sd()
    y = O.rrdb_forward(torch.from_numpy(g["x"]), sd, "model.1.sub.0.")
    _close(y, g["y_eval"])


def test_rrdbnet_config1_and_png_plumbing(golden_dir):
    g = _load(golden_dir, "rrdbnet_c1_nb1_nf32.npz")
    sd = O.synth_state_dict_g(3, 3, 32, 1, seed=21)
    x = torch.from_numpy(g["x"])
    y = O.rrdbnet_forward(x, sd, nb=1)
    _close(y, g["y"])
    # test_image/test.py:31-40 plumbing on the uint8 image
    img = g["img_u8"] * 1.0 / 255
    t = torch.from_numpy(np.transpose(img[:, :, [2, 1, 0]], (2, 0, 1))).float().unsqueeze(0)
    out = O.rrdbnet_forward(t, sd, nb=1).squeeze().clamp_(0, 1).numpy()
    out = (np.transpose(out[[2, 1, 0]], (1, 2, 0)) * 255.0).round().astype("uint8")
    assert np.abs(out.astype(int) - g["out_u8"].astype(int)).max() <= 1
    assert (out != g["out_u8"]).mean() < 1e-3


def test_rrdbnet_nb23(golden_dir):
    g = _load(golden_dir, "rrdbnet_nb23_nf64.npz")
    sd = O.synth_state_dict_g(3, 3, 64, 23, seed=31)
    assert O.n_rrdb_blocks(sd) == 23
    _close(O.rrdbnet_forward(torch.from_numpy(g["x24"]), sd, nb=23), g["y24"])
    _close(O.rrdbnet_forward(torch.from_numpy(g["x_ragged"]), sd, nb=23), g["y_ragged"])


def test_discriminator(golden_dir):
    g = _load(golden_dir, "dvgg128.npz")
    sd = O.synth_state_dict_d(3, 64, seed=41)
    x = torch.from_numpy(g["x"])
    y, _ = O.discriminator_vgg128_forward(x, sd, training=False)
    _close(y, g["y_eval"], tol=1e-5)
    xg = x.clone().requires_grad_(True)
    yt, newbuf = O.discriminator_vgg128_forward(xg, sd, training=True)
    _close(yt.detach(), g["y_train"], tol=1e-4)
    yt.sum().backward()
    _close(xg.grad, g["gx_train"], tol=1e-3)
    for k, v in newbuf.items():
        ref = g["after." + k]
        if "num_batches" in k:
            assert int(v) == int(ref)
        else:
            _close(v, ref, tol=1e-5)


def test_oracle_gradients_match_reference_fixture(golden_dir):
    """The oracle differentiated by torch autograd reproduces the gradients the REFERENCE's RRDBNet produced
    (tests/golden/make_golden_grads.py): the gradient parity tests on the GPU stand on this."""
    g = _load(golden_dir, "rrdbnet_grad_nb1_nf64.npz")
    sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict_g(3, 3, 64, 1, seed=11).items()}
    y = O.rrdbnet_forward(torch.from_numpy(g["x"]), sd, nb=1)
    _close(y.detach(), g["y"], tol=1e-5)
    (y * torch.from_numpy(g["gy"])).sum().backward()
    for k in [str(n) for n in g["names"]]:
        gr = sd[k].grad
        ref = g["norm." + k]
        assert abs(gr.norm().item() - ref[0]) <= 1e-4 * ref[0] + 1e-9, k
        assert abs(gr.abs().max().item() - ref[2]) <= 1e-3 * ref[2] + 1e-9, k
        head = torch.from_numpy(g["head." + k])
        assert (gr.flatten()[:64] - head).abs().max().item() <= 1e-3 * max(ref[2], 1e-12), k
        if "full." + k in g:
            _close(gr, g["full." + k], tol=1e-3)


def test_golden_meta(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    assert meta["test_image_RRDB_Net_equals_RRDBNet_eval"] and meta["test_image_keys_equal"]


def test_vgg_feature_oracle_matches_reference_fixture(golden_dir):
    """oracle.vgg_feature_forward against features, L1 feature loss and input gradient produced by the reference's own
    VGGFeatureExtractor (tests/golden/make_golden_vgg.py)."""
    import numpy as np
    d = np.load(os.path.join(golden_dir, "vgg_feature_34.npz"))
    sd = O.synth_state_dict_vgg(34, seed=3)
    fake = torch.from_numpy(d["fake"]).requires_grad_(True)
    fea = O.vgg_feature_forward(fake, sd)
    real_fea = O.vgg_feature_forward(torch.from_numpy(d["real"]), sd).detach()
    for got, key in ((fea.detach(), "fake_fea"), (real_fea, "real_fea")):
        ref = torch.from_numpy(d[key])
        assert (got - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() + 1e-6
    loss = torch.nn.functional.l1_loss(fea, real_fea)
    assert abs(loss.item() - float(d["loss"])) <= 1e-6
    loss.backward()
    ref = torch.from_numpy(d["dfake"])
    assert (fake.grad - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
