#!/usr/bin/env python
"""GPU probe for the tcgen05 conv kernel: runs every (kc, bn, variant) combination against a
torch fp32 conv2d on bf16-rounded operands, each case in its own subprocess with a timeout (a
trap poisons the CUDA context; a hang must not take the box down), then times RDB-shaped convs.

Usage on the GPU box:  python tools/probe_conv.py --all --out gpurun_out/probe.jsonl
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ref_conv(x_nhwc_list, chunks, kc, w, bias, act):
    """fp32 conv2d over the logical concat of the chunks (bf16-rounded operands)."""
    import torch
    import torch.nn.functional as F
    xs = [x_nhwc_list[si][..., c0:c0 + kc] for (si, c0) in chunks]
    x = torch.cat(xs, dim=3).float().permute(0, 3, 1, 2).contiguous()
    wq = w.to(torch.bfloat16).float()
    y = F.conv2d(x, wq, bias, padding=1)
    if act:
        y = F.leaky_relu(y, 0.2)
    return y  # NCHW fp32


def case_correct(kc, bn, variant, n=2, h=20, w=27, fused=False):
    import torch
    from esrganplus_b200 import conv as K
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234 + kc + bn + variant)
    dev = "cuda"
    c0t, c1t = 64, 128
    s0 = (torch.randn(n, h, w, c0t, device=dev)).to(torch.bfloat16)
    s1 = (torch.randn(n, h, w, c1t, device=dev)).to(torch.bfloat16)
    if kc == 64:
        chunks = [(0, 0), (1, 64)]
    else:
        chunks = [(0, 0), (0, 32), (1, 32)]
    cin = kc * len(chunks)
    cout = bn if bn >= 32 else 3
    wt = torch.randn(cout, cin, 3, 3, device=dev) * (1.0 / (cin * 9) ** 0.5)
    bias = torch.randn(cout, device=dev)
    lc0 = [i * kc for i in range(len(chunks))]
    wp = K.pack_conv3x3_weights(wt, kc, bn, lc0)
    bias_p = torch.zeros(bn, device=dev)
    bias_p[:cout] = bias
    ref = _ref_conv([s0, s1], chunks, kc, wt, bias, act=1)
    res = {}
    if cout < 16:
        out = torch.full((n, cout, h, w), float("nan"), device=dev)
        call = K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout,
                          w_packed=wp, bias=bias_p, act=1, out_nchw=out, variant=variant)
        call.launch()
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        res["max_err"] = err
        res["ref_max"] = ref.abs().max().item()
        res["ok"] = bool(err < 2e-3 * max(1.0, res["ref_max"]))
        return res
    if not fused:
        out = torch.full((n, h, w, 192), float("nan"), device=dev, dtype=torch.bfloat16)
        call = K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout,
                          w_packed=wp, bias=bias_p, act=1, out_bf16=out, ob_c0=64, variant=variant)
        call.launch()
        torch.cuda.synchronize()
        got = out[..., 64:64 + cout].float().permute(0, 3, 1, 2)
        err = (got - ref).abs().max().item()
        res["max_err"] = err
        res["ref_max"] = ref.abs().max().item()
        untouched = torch.isnan(out[..., :64].float()).all().item() and torch.isnan(out[..., 64 + cout:].float()).all().item()
        res["untouched_ok"] = bool(untouched)
        res["ok"] = bool(err < 0.02 * max(1.0, res["ref_max"]) and untouched)
        return res
    # fused epilogue: aux 1x1 over chunk 0, r1 (fp32), r2 (bf16), fp32 + bf16 outputs
    wa = torch.randn(cout, kc, 1, 1, device=dev) * (1.0 / kc ** 0.5)
    wap = K.pack_conv1x1_weights(wa, kc, bn, [0])
    r1 = torch.randn(n, h, w, 64, device=dev)
    r2 = torch.randn(n, h, w, 64, device=dev).to(torch.bfloat16)
    out_b = torch.zeros((n, h, w, 64), device=dev, dtype=torch.bfloat16)
    out_f = torch.zeros((n, h, w, 64), device=dev)
    call = K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout,
                      w_packed=wp, bias=bias_p, act=1, s0=0.5, w_aux=wap, aux_chunks=1,
                      r1=r1, r1_c0=0, s1=0.25, r2=r2, r2_c0=0, s2=0.2,
                      out_bf16=out_b, out_f32=out_f, variant=variant)
    call.launch()
    torch.cuda.synchronize()
    import torch.nn.functional as F
    x0 = s0[..., 0:kc].float().permute(0, 3, 1, 2)
    aux = F.conv2d(x0, wa.to(torch.bfloat16).float())
    v = 0.5 * ref + aux + 0.25 * r1[..., :cout].permute(0, 3, 1, 2)
    v = 0.2 * v + r2[..., :cout].float().permute(0, 3, 1, 2)
    got_f = out_f[..., :cout].permute(0, 3, 1, 2)
    got_b = out_b[..., :cout].float().permute(0, 3, 1, 2)
    res["max_err"] = (got_f - v).abs().max().item()
    res["max_err_bf16"] = (got_b - v).abs().max().item()
    res["ref_max"] = v.abs().max().item()
    res["ok"] = bool(res["max_err"] < 2e-3 * max(1.0, res["ref_max"]) and res["max_err_bf16"] < 0.02 * max(1.0, res["ref_max"]))
    return res


def case_layout():
    import torch
    from esrganplus_b200 import conv as K
    x = torch.randn(2, 3, 19, 23, device="cuda")
    y = K.nchw_f32_to_nhwc_bf16(x, 32)
    ref = torch.zeros(2, 19, 23, 32, device="cuda", dtype=torch.bfloat16)
    ref[..., :3] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    ok1 = torch.equal(y, ref)
    z = torch.randn(2, 9, 11, 64, device="cuda").to(torch.bfloat16)
    u = K.upsample2x_nhwc_bf16(z)
    uref = z.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ok2 = torch.equal(u, uref)
    b = K.nhwc_bf16_to_nchw_f32(z, 64)
    ok3 = torch.equal(b, z.float().permute(0, 3, 1, 2).contiguous())
    return {"nchw_to_nhwc": bool(ok1), "upsample": bool(ok2), "nhwc_to_nchw": bool(ok3), "ok": bool(ok1 and ok2 and ok3)}


RDB_SHAPES = {
    # name: (kc, chunks[(src,c0)], bn, cout, aux_chunks)
    "conv1": (64, [(0, 0)], 32, 32, 0),
    "conv2": (32, [(0, 0), (0, 32), (1, 0)], 32, 32, 2),
    "conv3": (64, [(0, 0), (1, 0)], 32, 32, 0),
    "conv4": (32, [(0, 0), (0, 32), (1, 0), (1, 32), (1, 64)], 32, 32, 0),
    "conv5": (64, [(0, 0), (1, 0), (1, 64)], 64, 64, 0),
    "conv64": (64, [(0, 0)], 64, 64, 0),
}


def case_time(name, variant, n=16, h=128, w=128, iters=20):
    import torch
    from esrganplus_b200 import conv as K
    kc, chunks, bn, cout, aux = RDB_SHAPES[name]
    dev = "cuda"
    s0 = torch.randn(n, h, w, 64, device=dev).to(torch.bfloat16)
    s1 = torch.randn(n, h, w, 128, device=dev).to(torch.bfloat16)
    cin = kc * len(chunks)
    wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
    wp = K.pack_conv3x3_weights(wt, kc, bn, [i * kc for i in range(len(chunks))])
    wap = None
    if aux:
        wa = torch.randn(cout, kc * aux, 1, 1, device=dev) * 0.02
        wap = K.pack_conv1x1_weights(wa, kc, bn, [i * kc for i in range(aux)])
    bias = torch.zeros(bn, device=dev)
    out = torch.zeros((n, h, w, 128), device=dev, dtype=torch.bfloat16)
    call = K.ConvCall(n=n, h=h, w=w, srcs=[s0, s1], kc=kc, chunks=chunks, bn=bn, cout=cout, w_packed=wp,
                      bias=bias, act=1, w_aux=wap, aux_chunks=aux, out_bf16=out, ob_c0=0, variant=variant)
    for _ in range(3):
        call.launch()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call.launch()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000.0 / iters
    flops = 2.0 * n * h * w * cout * (9 * cin + (kc * aux))
    return {"us": us, "tflops": flops / us * 1e-6, "ok": True}


def run_case(spec):
    kind = spec["kind"]
    if kind == "layout":
        return case_layout()
    if kind == "correct":
        return case_correct(spec["kc"], spec["bn"], spec["variant"], fused=spec.get("fused", False),
                            n=spec.get("n", 2), h=spec.get("h", 20), w=spec.get("w", 27))
    if kind == "time":
        return case_time(spec["name"], spec["variant"])
    raise ValueError(kind)


def all_specs(quick=False):
    specs = [{"kind": "layout"}]
    # variant bits: 1 = aligned (3 boxes), 2 = MT1
    for variant in (3, 1, 2, 0):
        for kc in (64, 32):
            for bn in (32, 64, 16):
                specs.append({"kind": "correct", "kc": kc, "bn": bn, "variant": variant})
    for variant in (3, 0):
        for kc in (64, 32):
            specs.append({"kind": "correct", "kc": kc, "bn": 32, "variant": variant, "fused": True})
    specs.append({"kind": "correct", "kc": 64, "bn": 64, "variant": 0, "n": 3, "h": 64, "w": 48})
    specs.append({"kind": "correct", "kc": 32, "bn": 32, "variant": 0, "n": 1, "h": 5, "w": 7})
    if not quick:
        for variant in (0, 2):
            for name in RDB_SHAPES:
                specs.append({"kind": "time", "name": name, "variant": variant})
    return specs


def worker(specs):
    for i, spec in specs:
        try:
            res = run_case(spec)
        except Exception as e:  # a CUDA error poisons the context: report and stop this worker
            print("RESULT " + json.dumps({"i": i, "ok": False, "error": repr(e)[:600]}), flush=True)
            return
        res["i"] = i
        print("RESULT " + json.dumps(res), flush=True)


def main():
    import select
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--worker", type=str, default=None, help="JSON list of [index, spec] (internal)")
    ap.add_argument("--out", type=str, default="gpurun_out/probe.jsonl")
    ap.add_argument("--timeout", type=int, default=120, help="seconds of silence before a worker is killed")
    args = ap.parse_args()
    if args.worker:
        worker(json.loads(args.worker))
        return
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    specs = list(enumerate(all_specs(args.quick)))
    results = {}
    pending = specs
    while pending:
        pr = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", json.dumps(pending)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
        last = time.time()
        first = True
        reported_error = False
        while True:
            limit = args.timeout + (180 if first else 0)  # first import of torch can take a minute
            r, _, _ = select.select([pr.stdout], [], [], 1.0)
            if r:
                line = pr.stdout.readline()
                if not line:
                    break
                if line.startswith("RESULT "):
                    res = json.loads(line[7:])
                    reported_error = "error" in res
                    results[res.pop("i")] = res
                    last = time.time()
                    first = False
            elif pr.poll() is not None:
                break
            elif time.time() - last > limit:
                pr.kill()
                break
        try:
            pr.wait(timeout=10)
        except subprocess.TimeoutExpired:
            pr.kill()
        err = pr.stderr.read()[-800:] if pr.stderr else ""
        rest = [(i, sp) for (i, sp) in pending if i not in results]
        if rest and not reported_error:
            # the first unfinished case crashed or hung the worker
            i, sp = rest[0]
            results[i] = {"ok": False, "crashed": True, "rc": pr.returncode, "stderr": err}
            rest = rest[1:]
        pending = rest
    fails = 0
    with open(args.out, "w") as f:
        for i, sp in specs:
            res = results.get(i, {"ok": False, "missing": True})
            res["spec"] = sp
            if not res.get("ok"):
                fails += 1
            f.write(json.dumps(res) + "\n")
            print(json.dumps(res), flush=True)
    print(f"probe done: {fails} failing cases of {len(specs)}")


if __name__ == "__main__":
    main()
