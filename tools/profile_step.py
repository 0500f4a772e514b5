#!/usr/bin/env python
"""One config-2 forward inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:conv3x3_tc -c 12 -o gpurun_out/prof python tools/profile_step.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--tile", type=int, default=128)
ap.add_argument("--nb", type=int, default=23)
ap.add_argument("--train", action="store_true")
ap.add_argument("--bwd", action="store_true", help="profile one training forward + backward (parameters require grad)")
a = ap.parse_args()
dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, a.nb)
net.load_state_dict(random_state_dict_g(3, 3, 64, a.nb, seed=31))
net = net.to(dev)
net.train(a.train)
for p in net.parameters():
    p.requires_grad = False
x = torch.rand(a.batch, 3, a.tile, a.tile, device=dev)
if a.bwd:
    for p in net.parameters():
        p.requires_grad = True
    dy = torch.randn(a.batch, 3, 4 * a.tile, 4 * a.tile, device=dev)
    net(x).backward(dy)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(x).backward(dy)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")
    sys.exit(0)
with torch.no_grad():
    net(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
