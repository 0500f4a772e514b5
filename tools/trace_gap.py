#!/usr/bin/env python
"""Launch-to-launch anatomy of the row kernel (needs a -DESRP_TRACE_FINE build: `make -C esrganplus_b200/csrc EXTRA=-DESRP_TRACE_FINE`):
%globaltimer at entry and exit of EVERY CTA for two consecutive dependent launches of the same conv.  Prints, in ns relative
to the first CTA entry of launch A: the spread of CTA entries / exits of A, the entries of B, and per SM slot the time between
A's exit and B's entry (programmatic dependent launch lets B's CTA start as soon as A's CTA on that SM is gone)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from esrganplus_b200 import conv as K

SHAPES = {"conv1": [(0, 0)], "conv3": [(0, 0), (1, 0)], "conv5h": [(0, 0), (1, 0), (1, 64)]}


def run(name, variant):
    dev = "cuda"
    n, h, w, kc, bn, cout = 16, 128, 128, 64, 32, 32
    chunks = SHAPES[name]
    s0 = torch.randn(n, h, w, 64, device=dev).to(torch.bfloat16)
    s1 = torch.randn(n, h, w, 128, device=dev).to(torch.bfloat16)
    wt = torch.randn(cout, kc * len(chunks), 3, 3, device=dev) * 0.02
    wp = K.pack_conv3x3_weights(wt, kc, bn, [i * kc for i in range(len(chunks))], layout=1)
    bias = torch.zeros(bn, device=dev)
    out = torch.zeros((n, h, w, 128), device=dev, dtype=torch.bfloat16)
    trs = [torch.zeros(3 * 1024, dtype=torch.int64, device=dev) for _ in range(2)]
    calls = [K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout, w_packed=wp, w_layout=1, bias=bias,
                        act=1, out_bf16=out, ob_c0=0, variant=variant, trace=t) for t in [None] + trs]
    for _ in range(5):
        calls[0].launch()
    torch.cuda.synchronize()
    for c in (calls[0], calls[0], calls[0], calls[1], calls[2], calls[0], calls[0]):
        c.launch()
    torch.cuda.synchronize()
    a, b = (t.cpu().view(3, 1024) for t in trs)
    g = 148
    a_in, a_out, b_in, b_out = a[0][512:512 + g], a[1][512:512 + g], b[0][512:512 + g], b[1][512:512 + g]
    t0 = int(a_in.min())
    f = lambda t: (int(t.min()) - t0, int(t.median()) - t0, int(t.max()) - t0)
    print(f"## {name} variant={variant} (ns relative to A's first CTA entry; min / median / max over {g} CTAs)")
    print("  A entry", f(a_in), " A exit", f(a_out))
    print("  B entry", f(b_in), " B exit", f(b_out))
    # CTA -> SM is not recorded; pair by order: the k-th CTA of A to exit frees the slot the k-th CTA of B to enter takes
    ae, be = torch.sort(a_out).values, torch.sort(b_in).values
    d = (be - ae)
    print("  B's k-th entry minus A's k-th exit:", (int(d.min()), int(d.median()), int(d.max())),
          "; B first entry - A last exit:", int(b_in.min()) - int(a_out.max()),
          "; CTA residency A:", f(a_out - a_in + t0), "; launch period (median entry to median entry):", int(b_in.median()) - int(a_in.median()))
    if os.environ.get("TRACE_GAP_DUMP"):
        U, hh = n * h, h
        res = (a_out - a_in).tolist()
        for c in range(g):
            u0, u1 = U * c // g, U * (c + 1) // g
            segs = (u1 - 1) // hh - u0 // hh + 1
            print(f"    cta {c:3d} rows {u1 - u0} segs {segs} first_row_in_image {u0 % hh:3d} residency {res[c]} ns  entry {int(a_in[c]) - t0} exit {int(a_out[c]) - t0}")


if __name__ == "__main__":
    for nm in [a for a in sys.argv[1:] if not a.startswith("v=")] or ["conv1", "conv3", "conv5h"]:
        for v in [int(a[2:]) for a in sys.argv[1:] if a.startswith("v=")] or [8192]:
            run(nm, v)
