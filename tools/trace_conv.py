#!/usr/bin/env python
"""In-kernel timeline of CTA 0 for RDB-shaped convs (uses esrp_conv3x3_t.trace).
Prints, per role, the clock64 deltas between consecutive pipeline events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from esrganplus_b200 import conv as K

SHAPES = {
    "conv1": (64, [(0, 0)], 32, 32, 0),
    "conv3": (64, [(0, 0), (1, 0)], 32, 32, 0),
    "conv5h": (64, [(0, 0), (1, 0), (1, 64)], 32, 32, 0),
    "hr1": (64, [(0, 0)], 16, 3, 0),
    # the growth chunk read from a 64-channel tensor (a pixel's 128 bytes contiguous with its neighbours') instead of one half
    # of the 128-channel growth buffer (128 of every 256 bytes): does the strided fetch cost DRAM efficiency?
    "conv3c": (64, [(0, 0), (1, 0)], 32, 32, 0),
}
LAYOUT = int(os.environ.get("ESRP_LAYOUT", "1"))


def run(name, n=16, h=128, w=128, variant=0, iters=50):
    kc, chunks, bn, cout, aux = SHAPES[name]
    dev = "cuda"
    s0 = torch.randn(n, h, w, 64, device=dev).to(torch.bfloat16)
    s1 = torch.randn(n, h, w, 64 if name.endswith("c") else 128, device=dev).to(torch.bfloat16)
    cin = kc * len(chunks)
    wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
    wp = K.pack_conv3x3_weights(wt, kc, bn, [i * kc for i in range(len(chunks))], layout=LAYOUT)
    bias = torch.zeros(bn, device=dev)
    out = torch.zeros((n, h, w, 128), device=dev, dtype=torch.bfloat16)
    out_nchw = torch.zeros((n, cout, h, w), device=dev) if cout < 16 else None
    tr = torch.zeros(3 * 1024, dtype=torch.int64, device=dev)
    call = K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout, w_packed=wp, w_layout=LAYOUT,
                      bias=bias, act=1, out_bf16=None if cout < 16 else out, out_nchw=out_nchw, ob_c0=0, variant=variant)
    for _ in range(3):
        call.launch()
    # replay through a CUDA graph so host-side planning/launch cost is out of the picture
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        call.launch()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                call.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    flops = 2.0 * n * h * w * cout * 9 * cin
    print(f"## {name} variant={variant}: {us:.1f} us  {flops / us * 1e-6:.0f} TFLOP/s")
    if variant & 0x1F00:  # timing experiments (ESRP_DBG_*): no timeline
        return
    call.trace = tr
    call.launch()
    torch.cuda.synchronize()
    t = tr.cpu().view(3, 1024)
    if int(t[2][1023]) > 0:   # -DESRP_TRACE_FINE builds: kernel entry / role-loop start / exit of CTA 0
        c_in, g_in, c_out, g_out, c_loop = (int(t[2][i]) for i in (1023, 1022, 1021, 1020, 1019))
        print(f"  CTA 0: entry -> role loops {c_loop - c_in} cycles, entry -> exit {c_out - c_in} cycles = {g_out - g_in} ns "
              f"({(c_out - c_in) / max(1, g_out - g_in):.2f} GHz); launch-to-launch {us * 1000:.0f} ns")
        t[2][1019:] = 0
    t0 = min(int(t[r][0]) for r in range(3) if int(t[r][0]) > 0)
    print(f"== {name} variant={variant} chunks={len(chunks)} kc={kc} bn={bn}")
    for r, role in enumerate(["producer(start, then after each empty-wait)", "mma(start; per chunk: full ok, q_empty ok, mmas issued, committed)",
                              "epilogue(start; per row: q_full ok, loaded+released, math done, stores issued)"]):
        ev = [int(v) - t0 for v in t[r] if int(v) > 0]
        print(role)
        print("  abs:", ev[:40])
        print("  dlt:", [b - a for a, b in zip(ev[:-1], ev[1:])][:60])


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("v=")] or ["conv1", "conv3", "conv5h", "hr1"]
    variants = [int(a[2:]) for a in sys.argv[1:] if a.startswith("v=")] or [0]
    for nm in names:
        for v in variants:
            run(nm, variant=v)
