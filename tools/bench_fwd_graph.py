#!/usr/bin/env python
"""Config-2 forward replayed as ONE CUDA graph (torch.cuda.graph around the module call) against the plain launch loop:
does the per-launch gap of the 354-launch stream matter?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, 23); net.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31)); net = net.to(dev).eval()
for p in net.parameters(): p.requires_grad = False
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


with torch.no_grad():
    x = torch.rand(16, 3, 128, 128, device=dev)
    for _ in range(5): y0 = net(x)
    plain = timed(lambda: net(x))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): net(x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = net(x)
    g.replay(); torch.cuda.synchronize()
    same = bool(torch.equal(y, y0))
    rounds = []
    for _ in range(4):
        rounds.append({"torch_graph_ms": round(timed(g.replay), 3), "engine_ms": round(timed(lambda: net(x)), 3)})
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("ESRP_")}, "first_engine_ms": round(plain, 3),
                  "rounds": rounds, "same_output": same}))
