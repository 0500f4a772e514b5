#!/usr/bin/env python
"""Forward time of RRDBNet(nb=23, nf=64, x4) on 128x128 LR tiles as a function of the batch size per call (does the
L2-resident working set of a smaller batch outweigh its higher per-launch overhead?)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, 23); net.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31)); net = net.to(dev).eval()
for p in net.parameters(): p.requires_grad = False
with torch.no_grad():
    for b in (2, 4, 8, 16, 32):
        x = torch.rand(b, 3, 128, 128, device=dev)
        for _ in range(3): net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): net(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"batch": b, "ms": round(ms, 3), "ms_per_tile": round(ms / b, 4), "out_MP_per_s": round(b * 0.262144 / ms * 1e3, 1)}))
