"""Diagnostic: input gradient of VGGFeatureExtractor truncated at several depths, every ReLU gate open (weights x 0.1, biases
+5), against the storage-precision oracle evaluated on the CPU and on CUDA (do the two break max-pool ties alike?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import esrganplus_b200 as E
from oracle import esrgan_oracle as O

dev = torch.device("cuda:0")
bf = lambda t: t.to(torch.bfloat16).float()


class R(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return bf(t)

    @staticmethod
    def backward(ctx, g):
        return bf(g)


def emulated(x, sd, fl):
    h = R.apply((x - sd["mean"]) / sd["std"])
    for idx, kind, _a, _b in O.vgg19_feature_layout(fl):
        if kind == "conv":
            h = F.conv2d(h, bf(sd[f"features.{idx}.weight"]), sd[f"features.{idx}.bias"], padding=1)
        elif kind == "relu":
            h = R.apply(F.relu(h))
        else:
            h = F.max_pool2d(h, 2, 2)
    return h


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return round(((a - b).norm() / b.norm()).item(), 5), round((torch.dot(a, b) / (a.norm() * b.norm())).item(), 6)


for fl in (0, 2, 5, 7, 10, 16, 19, 25, 28, 34):
    sd = O.synth_state_dict_vgg(34, seed=21)
    for k in sd:
        if k.endswith(".weight"):
            sd[k] = sd[k] * 0.1
        elif k.endswith(".bias"):
            sd[k] = torch.full_like(sd[k], 5.0)
    net = E.VGGFeatureExtractor(feature_layer=fl)
    net.load_state_dict({k: v for k, v in sd.items() if k in net.state_dict()}, strict=True)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(8)
    x = torch.rand(2, 3, 64, 64, generator=g)
    xg = x.to(dev).requires_grad_(True)
    fea = net(xg)
    gy = torch.randn(fea.shape, generator=g)
    (fea * gy.to(dev)).sum().backward()
    xe = x.clone().requires_grad_(True)
    fe = emulated(xe, sd, fl)
    (fe * gy).sum().backward()
    sdc = {k: v.to(dev) for k, v in sd.items()}
    xc = x.to(dev).requires_grad_(True)
    fc = emulated(xc, sdc, fl)
    (fc * gy.to(dev)).sum().backward()
    print(f"feature_layer {fl:2d}: fwd vs cpu-emu {rel(fea.detach(), fe.detach())}  grad vs cpu-emu {rel(xg.grad, xe.grad)}  "
          f"grad vs cuda-emu {rel(xg.grad, xc.grad)}  cpu-emu vs cuda-emu {rel(xe.grad, xc.grad)}", flush=True)
