#!/usr/bin/env python
"""BASELINE.json config 5: RRDB conv micro-bench on [16, C, 128, 128] — channel sweep Cin = 64 .. 320 (step 32) at
Cout = 32 (the five native dense-block points are Cin = 64/96/128/160 at Cout 32 and 192 at Cout 64), for the three
operators of a conv's forward + backward:
  fwd   : esrp_conv3x3_nhwc (tcgen05 row kernel), bias + LeakyReLU + sign bits, bf16 out
  dgrad : the same kernel over esrp_pack_dgrad_weights (K = Cout-side channels, 32 output channels)
  wgrad : conv3x3_wgrad_kernel, Cin/32 units x one 64-column dY block
Each point: CUDA-graph replay of 20 launches, CUDA events; TFLOP/s counts ALGORITHMIC flops 2*9*Cin*Cout*px (padding
of Cin to the 64-channel chunk is overhead, not work); `pipe` = fraction of the measured sustained bf16 peak; `gbs` =
algorithmic bytes (operands read once + outputs written once) / time.  Prints one JSON line per point.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from esrganplus_b200 import _lib
from esrganplus_b200 import conv as K

N, H, W = 16, 128, 128
PX = N * H * W
LAYOUT = _lib.LAYOUT_ROW
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
PEAK = peaks.get("bf16_tflops_sustained", 1400.0)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters  # us


def point(cin, cout):
    dev = "cuda"
    cpad = (cin + 63) // 64 * 64
    chunks = cpad // 64
    srcs = [torch.randn(N, H, W, 64, device=dev).to(torch.bfloat16), torch.randn(N, H, W, max(64, cpad - 64), device=dev).to(torch.bfloat16)]
    ch = [(0, 0)] + [(1, 64 * i) for i in range(chunks - 1)]
    wt = torch.zeros(cout, cpad, 3, 3, device=dev)
    wt[:, :cin] = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
    out = {"cin": cin, "cout": cout, "n": N, "h": H, "w": W}
    flops = 2.0 * 9 * cin * cout * PX
    res = {}
    # forward: <= 32 output channels per launch (64 = two launches over disjoint weight rows, as in the engine)
    calls = []
    ob = torch.zeros(N, H, W, 128, device=dev, dtype=torch.bfloat16)
    mask = torch.zeros(N, H, W, 8, device=dev, dtype=torch.int16)
    of = torch.zeros(N, H, W, 64, device=dev)
    for r0 in range(0, cout, 32):
        if chunks <= 3:
            wp = K.pack_conv3x3_weights(wt, 64, 32, [64 * i for i in range(chunks)], row0=r0, rows=32, layout=LAYOUT)
            calls.append(K.ConvCall(n=N, h=H, w=W, srcs=srcs, kc=64, chunks=ch, bn=32, cout=32, w_packed=wp, w_layout=LAYOUT,
                                    bias=torch.zeros(32, device=dev), act=1, out_bf16=ob, ob_c0=r0, mask_out=mask, mask_out_c0=r0))
        else:
            # extended points (Cin > 192): the row kernel keeps <= 3 chunks of weights resident, so K runs as groups of
            # <= 3 chunks accumulating in an fp32 NHWC buffer (no activation between the partial sums)
            for g0 in range(0, chunks, 3):
                cg = list(range(g0, min(g0 + 3, chunks)))
                wp = K.pack_conv3x3_weights(wt, 64, 32, [64 * i for i in cg], row0=r0, rows=32, layout=LAYOUT)
                c = K.ConvCall(n=N, h=H, w=W, srcs=srcs, kc=64, chunks=[ch[i] for i in cg], bn=32, cout=32, w_packed=wp,
                               w_layout=LAYOUT, out_f32=of, of_c0=r0)
                if g0 > 0:
                    c.r1, c.r1_c0, c.s1 = of, r0, 1.0
                calls.append(c)
    us = timed(lambda: [c.launch() for c in calls])
    res["fwd"] = (us, flops, (cin + cout) * 2 * PX)
    # dgrad: K = cout-side channels (one 64-chunk), output = 32 input channels per launch, all cin/32 slices
    dy = torch.randn(N, H, W, 64, device=dev).to(torch.bfloat16)
    dcalls = []
    dx = torch.zeros(N, H, W, cpad, device=dev, dtype=torch.bfloat16)
    groups = [(wt, 32 * g, 1.0) for g in range((cout + 31) // 32)]
    if len(groups) == 1:
        groups.append(None)
    for s in range((cin + 31) // 32):
        wpd = K.pack_dgrad_weights(groups, 32 * s, 32, 64, 32, layout=LAYOUT)
        dcalls.append(K.ConvCall(n=N, h=H, w=W, srcs=[dy], kc=64, chunks=[(0, 0)], bn=32, cout=32, w_packed=wpd, w_layout=LAYOUT,
                                 out_bf16=dx, ob_c0=32 * s, mask_in=mask, mask_in_c0=0))
    us = timed(lambda: [c.launch() for c in dcalls])
    res["dgrad"] = (us, flops, (cin + cout) * 2 * PX)
    # wgrad: cin/32 units against one 64-column dY block
    x_all = torch.randn(N, H, W, cpad, device=dev).to(torch.bfloat16)
    units_n = (cin + 31) // 32
    acc = torch.zeros(units_n, 9, 64, 32, device=dev)
    bacc = torch.zeros(64, device=dev)
    units = [(x_all, 32 * g, dy, 0, acc[g], bacc if g == 0 else None) for g in range(units_n)]
    us = timed(lambda: K.conv3x3_wgrad(units, N, H, W, 0))
    res["wgrad"] = (us, flops, (cin + cout) * 2 * PX)
    for k, (us, fl, by) in res.items():
        out[k] = {"us": round(us, 2), "tflops": round(fl / us * 1e-6, 1), "pipe": round(fl / us * 1e-6 / PEAK, 3), "gbs": round(by / us * 1e-3, 1)}
    return out


if __name__ == "__main__":
    native = {(64, 32), (96, 32), (128, 32), (160, 32), (192, 64)}
    pts = [(c, 32) for c in range(64, 321, 32)] + [(192, 64)]
    for cin, cout in pts:
        r = point(cin, cout)
        r["native_rdb_point"] = (cin, cout) in native
        r["peak_tflops"] = PEAK
        print(json.dumps(r), flush=True)
