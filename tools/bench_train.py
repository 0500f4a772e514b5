#!/usr/bin/env python
"""Config 4 of BASELINE.json: ESRGAN+ GAN train step (RRDBNet nb=23 nf=64 G + Discriminator_VGG_128 D, no perceptual),
128x128 HR crops.  Prints one JSON line: imgs/s (device time, max over ranks) plus a host-vs-device breakdown.

  python tools/bench_train.py [--batch 32] [--steps 10] [--warmup 3] [--nb 23] [--phases]
Under torchrun the global batch is split across ranks (strong scaling) unless --weak.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nb", type=int, default=23)
    ap.add_argument("--weak", action="store_true")
    ap.add_argument("--phases", action="store_true", help="time G fwd / G bwd / D separately (adds syncs)")
    ap.add_argument("--perceptual", action="store_true", help="add the VGG19 feature loss of the shipped recipe (feature_weight 1)")
    args = ap.parse_args()
    import torch
    import esrganplus_b200 as E
    from esrganplus_b200.autograd import data_parallel
    from esrganplus_b200.gan_step import GanTrainStep
    from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g, random_state_dict_vgg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    bs = args.batch if args.weak else max(1, args.batch // world)
    netG = E.RRDBNet(3, 3, 64, args.nb)
    netG.load_state_dict(random_state_dict_g(3, 3, 64, args.nb, seed=31, scale=0.1, zero_bias=True))  # ~ kaiming x 0.1 (networks.py:104)
    netD = E.Discriminator_VGG_128(3, 64)
    netD.load_state_dict(random_state_dict_d(3, 64, seed=32))
    netG, netD = netG.to(dev).train(), netD.to(dev).train()
    if world > 1:
        data_parallel(netG)
        data_parallel(netD)
    netF = None
    if args.perceptual:
        netF = E.VGGFeatureExtractor()
        netF.load_state_dict(random_state_dict_vgg(34, seed=33))
        netF = netF.to(dev).eval()
    step = GanTrainStep(netG, netD, netF=netF)
    g = torch.Generator().manual_seed(rank)
    lr = torch.rand(bs, 3, 32, 32, generator=g).to(dev)
    hr = torch.rand(bs, 3, 128, 128, generator=g).to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step.step(lr, hr)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step.step(lr, hr)
    e1.record()
    t_host = time.perf_counter() - t0
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    out = {"metric": "gan_train_imgs_per_sec", "value": bs * world * args.steps / (ms * 1e-3), "unit": "imgs/s", "n_gpus": world,
           "batch_per_gpu": bs, "steps": args.steps, "ms_per_step": ms / args.steps,
           "host_issue_ms_per_step": t_host / args.steps * 1e3, "nb": args.nb,
           "scaling": "weak" if args.weak else "strong", "finite": bool(torch.isfinite(step.log["l_d_real"]).item()),
           "perceptual": bool(args.perceptual)}
    if netF is not None:   # the feature extractor on its own: forward, and forward + input gradient (12.7 GFLOP / image / pass)
        def tf(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        with torch.no_grad():
            f_ms = tf(lambda: netF(hr))
        hrg = hr.clone().requires_grad_(True)
        fb_ms = tf(lambda: netF(hrg).sum().backward())
        out["vgg19"] = {"fwd_ms": round(f_ms, 3), "fwd_bwd_ms": round(fb_ms, 3), "fwd_tflops": round(12.7e9 * bs / (f_ms * 1e-3) / 1e12, 1),
                        "fwd_bwd_tflops": round(2 * 12.7e9 * bs / (fb_ms * 1e-3) / 1e12, 1)}
    if args.phases:
        def timed(fn):
            torch.cuda.synchronize()
            a = time.perf_counter()
            r = fn()
            host = time.perf_counter() - a
            torch.cuda.synchronize()
            return r, (time.perf_counter() - a) * 1e3, host * 1e3
        ph = {}
        for p in netD.parameters():
            p.requires_grad = False
        fake, ph["g_fwd_ms"], ph["g_fwd_host_ms"] = timed(lambda: netG(lr))
        pred, ph["d_fwd_graph_ms"], ph["d_fwd_graph_host_ms"] = timed(lambda: netD(fake))
        _, ph["d_fwd_nograd_ms"], ph["d_fwd_nograd_host_ms"] = timed(lambda: netD(hr))
        _, ph["d_dgrad_plus_g_bwd_ms"], ph["d_dgrad_plus_g_bwd_host_ms"] = timed(lambda: pred.sum().backward())
        _, ph["optG_ms"], _ = timed(step.optimizer_G.step)
        _, ph["g_fwd_after_update_ms"], ph["g_fwd_after_update_host_ms"] = timed(lambda: netG(lr))
        for p in netD.parameters():
            p.requires_grad = True
        pr, ph["d_fwd_train_ms"], ph["d_fwd_train_host_ms"] = timed(lambda: netD(hr))
        _, ph["d_bwd_full_ms"], ph["d_bwd_full_host_ms"] = timed(lambda: pr.sum().backward())
        _, ph["optD_ms"], _ = timed(step.optimizer_D.step)
        eng = netG._engines[dev]
        ph["g_launches_fwd_bwd"] = list(eng.train_launches())
        out["phases"] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in ph.items()}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
