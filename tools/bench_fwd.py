#!/usr/bin/env python
"""Config-2 forward only (16 x 128x128 LR tiles, RRDBNet nb=23 nf=64 x4): ms per step with CUDA events.  Used for quick
A/B runs of the ESRP_* timing switches (they are read once per process, so one process per setting)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g
dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, 23); net.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31)); net = net.to(dev).eval()
for p in net.parameters(): p.requires_grad = False
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
with torch.no_grad():
    x = torch.rand(16, 3, 128, 128, device=dev)
    for _ in range(5): net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): net(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
env = {k: v for k, v in os.environ.items() if k.startswith("ESRP_")}
print(json.dumps({"env": env, "ms": round(ms, 3), "out_MP_per_s": round(16 * 0.262144 / ms * 1e3, 1),
                  "tflops": round(9.47291947 / ms * 1e3, 1), "launches": net._engines[dev].num_launches}))
