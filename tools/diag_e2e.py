"""Diagnostic: the fp32 e2e leg of bench.py (H2D of the tiles, forward, D2H of the 50 MB result on a copy stream) timed in
blocks of 10 steps for a while, as the FIRST CUDA process on a fresh box; plus the raw pinned D2H / H2D rates."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import esrganplus_b200 as E
from esrganplus_b200.synth import random_state_dict_g

dev = torch.device("cuda:0")
net = E.RRDBNet(3, 3, 64, 23)
net.load_state_dict(random_state_dict_g(3, 3, 64, 23, seed=31))
net = net.to(dev).eval()
for p in net.parameters():
    p.requires_grad = False
x_host = torch.rand(16, 3, 128, 128).pin_memory()
y_hosts = [torch.empty(16, 3, 512, 512).pin_memory() for _ in range(2)]
copy_stream = torch.cuda.Stream(device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def rate(fn, nbytes, reps=10):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9


with torch.no_grad():
    yd = torch.empty(16, 3, 512, 512, device=dev)
    for k in range(3):
        print(f"raw D2H {rate(lambda: y_hosts[0].copy_(yd, non_blocking=True), yd.numel() * 4):.1f} GB/s, "
              f"raw H2D {rate(lambda: yd.copy_(y_hosts[1], non_blocking=True), yd.numel() * 4):.1f} GB/s", flush=True)
    xd = x_host.to(dev)
    for _ in range(5):
        net(xd)
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(10):
        net(xd)
    b.record()
    torch.cuda.synchronize()
    print(f"resident: {a.elapsed_time(b) / 10:.3f} ms/step", flush=True)
    i = 0
    for blk in range(12):
        torch.cuda.synchronize()
        a, b = ev(), ev()
        t0 = time.perf_counter()
        a.record()
        for _ in range(10):
            xd2 = x_host.to(dev, non_blocking=True)
            y = net(xd2)
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                y_hosts[i & 1].copy_(y, non_blocking=True)
                y.record_stream(copy_stream)
            i += 1
        host_ms = (time.perf_counter() - t0) * 1e3 / 10
        torch.cuda.current_stream().wait_stream(copy_stream)
        b.record()
        torch.cuda.synchronize()
        print(f"e2e block {blk}: {a.elapsed_time(b) / 10:.3f} ms/step (host issue {host_ms:.3f} ms/step)", flush=True)
    print(f"raw D2H {rate(lambda: y_hosts[0].copy_(yd, non_blocking=True), yd.numel() * 4):.1f} GB/s", flush=True)
