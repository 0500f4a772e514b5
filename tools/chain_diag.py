#!/usr/bin/env python
"""Bisecting aid for the persistent conv chain: chain vs one launch per conv vs a torch emulation of the kernels'
storage precision (bf16 operands / stored activations, fp32 accumulation and trunk) on one shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
n, h, w, nb = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda:0")
sd = random_state_dict_g(3, 3, 64, nb, seed=41)
net = E.RRDBNet(3, 3, 64, nb); net.load_state_dict(sd); net = net.to(dev).eval()
for p in net.parameters(): p.requires_grad = False
x = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(5)).to(dev)
eng = net._engine_for(dev)


def bf(t):
    return t.to(torch.bfloat16).float()


def emulate(x):
    W = {k: v.to(dev) for k, v in sd.items()}
    cv = lambda t, k, pad=1: F.conv2d(t, bf(W[k + ".weight"]), W.get(k + ".bias"), padding=pad)
    lr = lambda t: F.leaky_relu(t, 0.2)
    fea_f = cv(bf(x), "model.0")
    tf, tb = fea_f, bf(fea_f)
    for i in range(nb):
        rr = tf
        for r in (1, 2, 3):
            p = f"model.1.sub.{i}.RDB{r}."
            x1 = bf(lr(cv(tb, p + "conv1.0")))
            x2 = bf(lr(cv(torch.cat([tb, x1], 1), p + "conv2.0")) + F.conv2d(tb, bf(W[p + "conv1x1.weight"])))
            x3 = bf(lr(cv(torch.cat([tb, x1, x2], 1), p + "conv3.0")))
            x4 = bf(lr(cv(torch.cat([tb, x1, x2, x3], 1), p + "conv4.0")) + x2)
            x5 = cv(torch.cat([tb, x1, x2, x3, x4], 1), p + "conv5.0")
            tf = 0.2 * x5 + tf
            if r == 3:
                tf = 0.2 * tf + rr
            tb = bf(tf)
    u = bf(cv(tb, f"model.1.sub.{nb}") + fea_f)
    u = bf(lr(cv(F.interpolate(u, scale_factor=2, mode="nearest"), "model.3")))
    u = bf(lr(cv(F.interpolate(u, scale_factor=2, mode="nearest"), "model.6")))
    u = bf(lr(cv(u, "model.8")))
    return cv(u, "model.10")


with torch.no_grad():
    ye = emulate(x)
    eng.set_chain(False); yp = net(x).clone(); yp2 = net(x).clone()
    eng.set_chain(True); yc = net(x).clone(); nchained = eng.num_chained_convs; nl = eng.num_launches
    yc2 = net(x).clone()
std = ye.std().item()
rel = lambda a, b: round((a - b).abs().max().item() / std, 5)
rms = lambda a, b: round((a - b).pow(2).mean().sqrt().item() / std, 6)
env = {k: v for k, v in os.environ.items() if k.startswith("ESRP_")}
print(json.dumps({"env": env, "shape": [n, h, w, nb], "chained": nchained, "launches": nl,
                  "chain_vs_emul": [rel(yc, ye), rms(yc, ye)], "plain_vs_emul": [rel(yp, ye), rms(yp, ye)],
                  "chain_vs_plain": [rel(yc, yp), rms(yc, yp)], "plain_rerun": rel(yp, yp2), "chain_rerun": rel(yc, yc2)}))
