#!/usr/bin/env python
"""Forward ms per batch size (128x128 LR tiles, RRDBNet nb=23 nf=64), persistent conv chain on / off."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, 23); net.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31)); net = net.to(dev).eval()
for p in net.parameters(): p.requires_grad = False
eng = net._engine_for(dev)
out = []
with torch.no_grad():
    for b in (1, 2, 4, 8, 16, 32):
        x = torch.rand(b, 3, 128, 128, device=dev)
        row = {"batch": b}
        for chain in (True, False):
            eng.set_chain(chain)
            for _ in range(3): net(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 20
            e0.record()
            for _ in range(steps): net(x)
            e1.record(); torch.cuda.synchronize()
            row["chain_ms" if chain else "plain_ms"] = round(e0.elapsed_time(e1) / steps, 3)
            row["chain_launches" if chain else "plain_launches"] = eng.num_launches
        row["ms_per_tile_chain"] = round(row["chain_ms"] / b, 3); row["ms_per_tile_plain"] = round(row["plain_ms"] / b, 3)
        out.append(row)
        print(json.dumps(row), flush=True)
