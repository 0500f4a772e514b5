#!/usr/bin/env python
"""bench.py — x4 SR forward throughput of the RRDBNet generator hot path (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one forward pass of RRDBNet(nb=23, nf=64, x4) over one batch of 16 synthetic 128x128 LR
tiles per GPU (-> 16 x 512x512 outputs).  Metric: output megapixels / second, whole job (all ranks).
Inference shards independent tiles across ranks with no collective (weak scaling).

`value`      device-resident input, CUDA-event timed, max over ranks.
`e2e`        same metric through the public module call with HOST buffers: pinned host input ->
             H2D -> RRDBNet.forward -> D2H of the fp32 output image, all inside the timed region.
`roofline`   the conv3x3 tcgen05 kernel family = every launch of the step but three small
             layout kernels; achieved = algorithmic FLOPs of the step / event time of the step.
`train`      secondary leg, BASELINE.json config 4: ESRGAN+ GAN train step imgs/s, 32 crops per GPU (weak scaling),
             gradient all-reduce over NCCL; `train_strong` (N > 1 only): the same with 32 crops in total; `train_perceptual`:
             the weak leg with the VGG19 feature loss of the shipped recipe added; see train_leg().
`cpu_baseline` / --impl reference: the reference's algorithm (CPU oracle = torch CPU fp32 ops, the
             same ATen kernels the reference's nn.Conv2d dispatches to) on this box's host cores, on
             a bounded sample (one 128x128 tile per step).
`tiled`      BASELINE.json config 3: ONE 512x512 LR image cut into 16 crops of 128x128, sharded over the ranks
             (esrganplus_b200.tiled, no collective on the data path): latency per image and MP/s, strong scaling.
`chain`      the same forward with the opt-in persistent conv chain (10 launches instead of 354; slower).
`gpu_library_baseline` (N = 1): the oracle's functional forward (torch ops = cuDNN) on the SAME GPU in fp32, TF32 and
             bf16 autocast + channels_last — a library baseline, never routed through the product.
A device fault fails the run (there is no retry): the JSON line exists only if every leg of the headline ran clean.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB, NF, TILE, BATCH = 23, 64, 128, 16
FLOP_PER_LR_PX = 36_136_320          # 18 068 160 MAC per LR pixel, whole generator
OUT_MP_PER_TILE = (4 * TILE) ** 2 / 1e6
METRIC = "x4_sr_output_megapixels_per_sec_fwd"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md sustained)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def csrc_hash() -> str:
    """sha256 over the kernel sources: ties a committed ncu capture to the library build it was taken from."""
    import glob
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "esrganplus_b200", "csrc")
    for f in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) + glob.glob(os.path.join(d, "*.inl")) +
                    glob.glob(os.path.join(d, "*.h")) + [os.path.join(ROOT, "include", "esrp.h")]):
        h.update(open(f, "rb").read())
    return h.hexdigest()


def cpu_reference_run(steps: int, warmup: int, threads: int):
    """The reference's CPU path on a bounded sample: one 128x128 tile per step, fp32, all host threads."""
    import torch
    from oracle import esrgan_oracle as O
    torch.set_num_threads(threads)
    sd = O.synth_state_dict_g(3, 3, NF, NB, seed=31)
    x = torch.rand(1, 3, TILE, TILE, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(warmup):
            O.rrdbnet_forward(x, sd, NB)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.rrdbnet_forward(x, sd, NB)
        dt = (time.perf_counter() - t0) / steps
    return OUT_MP_PER_TILE / dt, dt


def gpu_library_baseline(torch, dev, sd, x):
    """The reference's own operator graph (oracle = functional restatement of block.py / architecture.py) executed by
    PyTorch's CUDA libraries (cuDNN convs, ATen elementwise, torch.cat) on the same GPU and the same batch: what a user
    of the unmodified reference gets on a B200.  fp32 with TF32 off (the reference's arithmetic), TF32 on, and bf16
    autocast with channels_last.  Test infrastructure used as a measured baseline only."""
    from oracle import esrgan_oracle as O
    sdd = {k: v.to(dev) for k, v in sd.items()}
    out = {}

    def run(label, xin, weights, steps=3):
        with torch.no_grad():
            for _ in range(2):
                O.rrdbnet_forward(xin, weights, NB)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                O.rrdbnet_forward(xin, weights, NB)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[label] = {"ms_per_step": ms, "value": BATCH * OUT_MP_PER_TILE / (ms * 1e-3), "unit": "MP/s"}

    old_c, old_m, old_b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark
    try:
        torch.backends.cudnn.benchmark = True
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        run("fp32_cudnn", x, sdd)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        run("tf32_cudnn", x, sdd)
        sd_cl = {k: (v.to(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sdd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            run("bf16_autocast_channels_last", x.to(memory_format=torch.channels_last), sd_cl)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_c, old_m
        torch.backends.cudnn.benchmark = old_b
    out["note"] = "oracle.rrdbnet_forward on CUDA (cuDNN / ATen), same batch of 16 tiles, 3 timed steps each"
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    v, dt = cpu_reference_run(steps, warm, threads)
    sample = f"1 tile of {TILE}x{TILE} LR per step (1/{BATCH} of the config-2 batch), {steps} timed steps, torch CPU fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MP/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"RRDBNet nb={NB} nf={NF} x4 inference, {TILE}x{TILE} LR synthetic tiles (config 2, bounded CPU sample)"},
        "cpu_baseline": {"value": v, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def train_leg(torch, dev, world, rank, dist, steps=8, warmup=5, global_batch=32, scaling="strong", perceptual=False):
    """BASELINE.json config 4: one ESRGAN+ GAN step (RRDBNet nb=23 nf=64 G + Discriminator_VGG_128 D, no perceptual,
    SRRaGAN_model.py:113-186) on 128x128 HR / 32x32 LR synthetic crops; `global_batch` is split across ranks and the
    two backward passes all-reduce their flat gradient buffers over NCCL.  imgs/s from CUDA events, max over ranks.
    BASELINE.json's "bs=32 ... DDP 8xB200" is read both ways: weak = 32 crops per GPU, strong = 32 in
    total."""
    import esrganplus_b200 as E
    from esrganplus_b200.autograd import broadcast_parameters, data_parallel
    from esrganplus_b200.gan_step import GanTrainStep
    from esrganplus_b200.synth import random_state_dict_d, random_state_dict_g, random_state_dict_vgg
    bs = max(1, global_batch // world)
    netG = E.RRDBNet(3, 3, NF, NB)
    # ~ kaiming x 0.1 with zero bias, what networks.py:103-104 does for G before training
    netG.load_state_dict(random_state_dict_g(3, 3, NF, NB, seed=31, scale=0.1, zero_bias=True), strict=True)
    netD = E.Discriminator_VGG_128(3, 64)
    netD.load_state_dict(random_state_dict_d(3, 64, seed=32), strict=True)
    netG, netD = netG.to(dev).train(), netD.to(dev).train()
    if dist is not None:
        broadcast_parameters(netG)
        broadcast_parameters(netD)
        data_parallel(netG)
        data_parallel(netD)
    netF = None
    if perceptual:   # the shipped recipe's feature loss (train_ESRGANplus.json: feature_weight 1, l1; VGG19 conv5_4, frozen)
        netF = E.VGGFeatureExtractor(feature_layer=34, use_bn=False, use_input_norm=True)
        netF.load_state_dict(random_state_dict_vgg(34, seed=33), strict=True)
        netF = netF.to(dev).eval()
    step = GanTrainStep(netG, netD, netF=netF)
    g = torch.Generator().manual_seed(100 + rank)
    lr = torch.rand(bs, 3, 32, 32, generator=g).to(dev)
    hr = torch.rand(bs, 3, 128, 128, generator=g).to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step.step(lr, hr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step.step(lr, hr)
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    finite = bool(torch.isfinite(torch.stack(list(step.log.values()))).all().item())
    fl, bl = netG._engines[dev].train_launches()
    return {"metric": "gan_train_imgs_per_sec", "value": bs * world * steps / (ms * 1e-3), "unit": "imgs/s",
            "ms_per_step": ms / steps, "host_issue_ms_per_step": host_ms, "steps": steps, "warmup": warmup,
            "global_batch": bs * world, "batch_per_gpu": bs, "scaling": scaling, "losses_finite": finite,
            "workload": "ESRGAN+ GAN step, RRDBNet nb=23 nf=64 + Discriminator_VGG_128, 128x128 HR crops, " +
                        ("with the VGG19 feature loss (the shipped recipe; config 4 + SURVEY 8f-1)" if perceptual else "no perceptual (config 4)"),
            "collective": None if dist is None else "all-reduce(avg) of the flat G (67.4 MB) and D (58.0 MB) gradient buffers, NCCL",
            "generator_launches_fwd_bwd": [fl, bl],
            "flops_per_img_algorithmic": 151e9 + (3 * 12.7e9 if perceptual else 0.0),
            "achieved_tflops_per_gpu": (151e9 + (3 * 12.7e9 if perceptual else 0.0)) * bs * steps / (ms * 1e-3) / 1e12}


def main_ours(args):
    import torch
    import esrganplus_b200 as E
    from esrganplus_b200.synth import random_state_dict_g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    sd = random_state_dict_g(3, 3, NF, NB, seed=31)        # random-init weights of the named architecture
    net = E.RRDBNet(3, 3, NF, NB, gc=32, upscale=4)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    for p in net.parameters():
        p.requires_grad = False
    gen = torch.Generator().manual_seed(rank)
    x_host = torch.rand(BATCH, 3, TILE, TILE, generator=gen).pin_memory()
    y_host = torch.empty(BATCH, 3, 4 * TILE, 4 * TILE).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()   # all streams of the device, the copy stream included

    def step_resident():
        return net(x_dev)

    # e2e: every step copies its input from pinned host memory, runs the public module call and copies the fp32 result
    # back; the device->host copy of step i runs on a copy stream and overlaps the forward of step i+1 (two pinned
    # result buffers), as a serving loop would do it.  All copies of all steps lie inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
    e2e_i = [0]

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        y = net(xd)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            y_hosts[e2e_i[0] & 1].copy_(y, non_blocking=True)
            y.record_stream(copy_stream)
        e2e_i[0] += 1
        return y

    def warm_until_stable(fn):
        """W warm-up steps, repeated while the caching allocator is still growing: every module call allocates its result, and
        `record_stream` keeps a block out of reuse until its copy has finished, so the pool needs as many result blocks as the
        host runs ahead of the device.  A `cudaMalloc` inside the timed region drains the launch queue (the slow first passes
        of profiles/r02_bench_e2e_repeats.log)."""
        for _ in range(4):
            before = torch.cuda.memory_reserved(dev)
            for _ in range(max(3, args.warmup)):
                fn()
            if torch.cuda.memory_reserved(dev) == before:
                break

    def timed(fn, steps, join=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        if join is not None:
            torch.cuda.current_stream().wait_stream(join)   # the last device->host copy ends inside the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step_resident()
        # the dense-block convs of the trunk (the dominant kernel family) are bracketed by a pair of CUDA events per forward,
        # recorded by the engine on the launching stream inside the timed region below (esrp_rrdbnet_set_timing)
        net._engines[dev].set_timing(True)
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        ms = timed(step_resident, args.steps)
        clocks = sampler.stop() if sampler else None
        trunk_ms = net._engines[dev].trunk_times_ms(min(args.steps, 64))
        net._engines[dev].set_timing(False)
        warm_until_stable(step_e2e)
        ms_e2e = timed(step_e2e, args.steps, join=copy_stream)
        # Seen in 5 of 41 runs (profiles/r02_bench_e2e_repeats.log): the first pass of this loop runs >= 2.4 ms per step
        # slower than the device-resident loop although the 50 MB result copy (0.9 ms at the measured 56 GB/s) is hidden
        # behind the next forward everywhere else — allocator growth inside the timed region (warm_until_stable now warms
        # until the pool is stable).  Should it still happen, the pass is measured ONCE more and both numbers are reported.
        e2e_first = None
        if ms_e2e > 1.08 * ms:
            e2e_first = ms_e2e / args.steps
            ms_e2e = timed(step_e2e, args.steps, join=copy_stream)
    # the same loop fed and drained as 8-bit images (RRDBNet.forward_uint8: the /255, BGR<->RGB, clamp, x255, round of
    # test_image/test.py:31-40 on the device): 4x smaller copies in both directions
    e2e_u8 = None
    try:
        xu_host = (x_host.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory()
        yu_hosts = [torch.empty(BATCH, 4 * TILE, 4 * TILE, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]

        def step_u8():
            y = net.forward_uint8(xu_host.to(dev, non_blocking=True))
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                yu_hosts[e2e_i[0] & 1].copy_(y, non_blocking=True)
                y.record_stream(copy_stream)
            e2e_i[0] += 1

        with torch.no_grad():
            warm_until_stable(step_u8)
            ms_u8 = timed(step_u8, args.steps, join=copy_stream)
        e2e_u8 = {"value": BATCH * OUT_MP_PER_TILE * world / (ms_u8 / args.steps * 1e-3), "unit": "MP/s",
                  "h2d_bytes_per_step": xu_host.numel() * world, "d2h_bytes_per_step": yu_hosts[0].numel() * world,
                  "ms_per_step": ms_u8 / args.steps, "api": "RRDBNet.forward_uint8 (uint8 HWC in/out, plumbing on the device)"}
    except Exception as e:
        e2e_u8 = {"error": f"{type(e).__name__}: {e}"[:300]}
    # ---- secondary inference legs -------------------------------------------------------------------------------
    chain_leg = tiled_leg = lib_leg = parity_leg = None
    eng = net._engines[dev]
    launches_plain = eng.num_launches
    if not args.no_extras:
        try:    # the opt-in persistent conv chain on the same workload
            eng.set_chain(True)
            with torch.no_grad():
                for _ in range(3):
                    step_resident()
                ms_c = timed(step_resident, args.steps)
            chain_leg = {"value": BATCH * OUT_MP_PER_TILE * world / (ms_c / args.steps * 1e-3), "unit": "MP/s",
                         "ms_per_step": ms_c / args.steps, "launches_per_step": eng.num_launches,
                         "convs_chained": eng.num_chained_convs,
                         "note": "esrp_rrdbnet_set_chain(1): dense-block convs as phases of one persistent launch"}
        except Exception as e:
            chain_leg = {"error": f"{type(e).__name__}: {e}"[:300]}
        finally:
            eng.set_chain(False)
        try:    # config 3: one 512x512 LR image, 16 crops sharded over the ranks, no collective on the data path
            from esrganplus_b200 import tiled as T
            img = torch.rand(1, 3, 4 * TILE, 4 * TILE, generator=torch.Generator().manual_seed(7)).to(dev)
            crops = T.crop_grid(4 * TILE, 4 * TILE, TILE)
            mine = T.shard(list(range(len(crops))), rank, world)

            def step_tiled():
                return T.run_crops(net, img, [crops[i] for i in mine], 4)

            with torch.no_grad():
                for _ in range(3):
                    step_tiled()
                ms_t = timed(step_tiled, args.steps)
            per_img = ms_t / args.steps
            tiled_leg = {"metric": "x4_sr_tiled_image_latency_ms", "value": per_img, "unit": "ms per 512x512 LR image",
                         "mp_per_s": (16 * TILE * TILE * 16 / 1e6) / (per_img * 1e-3), "crops": len(crops),
                         "crops_per_rank": len(mine), "scaling": "strong", "collective": None,
                         "workload": "one 512x512 LR image -> 2048x2048, 16 crops of 128x128 sharded round-robin (config 3)"}
        except Exception as e:
            tiled_leg = {"error": f"{type(e).__name__}: {e}"[:300]}
        try:    # the accuracy mode (esrganplus_b200/precise.py): same workload, split precision on the tile kernel
            with torch.no_grad():
                yp = net.forward_fp32_parity(x_dev)
                yf = net(x_dev)
                torch.cuda.synchronize()
                dev_db = 10.0 * math.log10(1.0 / max(1e-30, float(((yp.clamp(0, 1) - yf.clamp(0, 1)) ** 2).mean())))
                del yp, yf
                ms_p = timed(lambda: net.forward_fp32_parity(x_dev), 3)
            parity_leg = {"value": BATCH * OUT_MP_PER_TILE * world / (ms_p / 3 * 1e-3), "unit": "MP/s", "ms_per_step": ms_p / 3,
                          "fast_path_psnr_against_it_db": dev_db,
                          "note": "RRDBNet.forward_fp32_parity: every value as two bf16 tensors, A_hi W_hi + A_lo W_hi + A_hi W_lo per conv, "
                                  "fp32 accumulation and residuals; 107 dB against the reference fixtures (tests), 3 timed steps"}
            net.__dict__.pop("_precise_engines", None)
            torch.cuda.empty_cache()
        except Exception as e:
            parity_leg = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1:
            try:
                lib_leg = gpu_library_baseline(torch, dev, sd, x_dev)
            except Exception as e:
                lib_leg = {"error": f"{type(e).__name__}: {e}"[:300]}
    train = train_strong = train_perceptual = None
    if not args.no_train:
        # the inference legs leave ~2.5 GB of cached workspace behind and the GPU at its power cap: hand the memory back and
        # let the clocks settle before the secondary measurement
        launches_inf = launches_plain
        del y_hosts, x_dev
        net._engines.clear()
        torch.cuda.empty_cache()
        time.sleep(2.0)
        try:
            train = train_leg(torch, dev, world, rank, dist, global_batch=32 * world, scaling="weak")
            if world > 1:
                train_strong = train_leg(torch, dev, world, rank, dist, global_batch=32, scaling="strong")
        except Exception as e:  # the headline line must survive a failure of the secondary leg
            train = train or {"error": f"{type(e).__name__}: {e}"[:300]}
        try:
            train_perceptual = train_leg(torch, dev, world, rank, dist, global_batch=32 * world, scaling="weak", perceptual=True)
        except Exception as e:
            train_perceptual = {"error": f"{type(e).__name__}: {e}"[:300]}

    launches = launches_plain
    per_step = ms / args.steps
    mp_per_step = BATCH * OUT_MP_PER_TILE * world
    value = mp_per_step / (per_step * 1e-3)
    e2e_value = mp_per_step / (ms_e2e / args.steps * 1e-3)
    flops_step = FLOP_PER_LR_PX * BATCH * TILE * TILE      # per GPU
    peak_tf, _hbm, peak_src = _peaks()
    achieved_tf = flops_step / (per_step * 1e-3) / 1e12
    # Dominant kernel family: the 345 dense-block conv launches of the RRDB trunk (conv3x3_row_kernel<64,32>), 92.3 % of
    # the step's FLOPs (483 328 FLOP per LR pixel and dense block).  Duration: CUDA events on the
    # launching stream inside the timed region (above).  DRAM traffic: the committed ncu launch list of the same
    # workload, used only if it was taken from the same kernel sources as the library that just ran.
    trunk_flops = 483_328 * 3 * NB * BATCH * TILE * TILE
    trunk_avg_ms = sum(trunk_ms) / len(trunk_ms) if trunk_ms else None
    trunk_tf = trunk_flops / (trunk_avg_ms * 1e-3) / 1e12 if trunk_avg_ms else None
    traffic, traffic_note = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_launch_summary_config2.json")))
        if prof.get("csrc_sha256") == csrc_hash():
            dom = max(prof["by_kernel"], key=lambda k: k["us"])
            traffic = (dom["dram_read_MB"] + dom["dram_write_MB"]) * 1e6 / dom["launches"]
        else:
            traffic_note = "profiles/r02_ncu_launch_summary_config2.json was captured from other kernel sources than this build: not used"
    except Exception as e:
        traffic_note = f"no usable ncu summary: {type(e).__name__}"

    if rank == 0:
        cpu = None
        if world == 1:
            threads = os.cpu_count() or 1
            v, dt = cpu_reference_run(steps=3, warmup=1, threads=threads)
            cpu = {"value": v, "unit": "MP/s", "cores": threads, "kind": "port",
                   "sample": f"oracle (torch CPU fp32) on 1 tile of {TILE}x{TILE} LR per step, 3 timed steps, {dt:.2f} s/tile"}
        line = {
            "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"RRDBNet nb={NB} nf={NF} x4 inference, batch {BATCH} of {TILE}x{TILE} LR synthetic tiles per GPU (config 2)",
                       "global_batch": BATCH * world, "parallelism": f"independent tiles x{world}, no collective",
                       "l2": "activation working set ~2 GB per step >> 126 MB L2 (no flush needed)",
                       "weights": "random-init (numpy PCG64 seed 31), reference key layout"},
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": x_host.numel() * 4 * world,
                    "d2h_bytes_per_step": y_host.numel() * 4 * world, "ms_per_step": ms_e2e / args.steps,
                    **({"remeasured": True, "first_pass_ms_per_step": e2e_first} if e2e_first is not None else {})},
            "gpu_launches": launches * args.steps * world,
            "roofline": {"bound": "tensor", "achieved": trunk_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": trunk_tf / peak_tf if trunk_tf else None, "traffic": traffic,
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read+write, dominant kernel)", "traffic_note": traffic_note,
                         "peak_source": peak_src,
                         "kernel": "conv3x3_row_kernel<64,32>: the 345 dense-block conv launches of the RRDB trunk (tcgen05)",
                         "algorithmic_flops": trunk_flops, "launches": 15 * NB, "duration_ms": trunk_avg_ms,
                         "duration_source": f"cudaEvent pair per forward on the launching stream, {len(trunk_ms)} timed forwards",
                         "whole_step": {"achieved": achieved_tf, "frac": achieved_tf / peak_tf, "flops_per_step_per_gpu": flops_step}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e_uint8": e2e_u8,
            "chain": chain_leg,
            "tiled": tiled_leg,
            "gpu_library_baseline": lib_leg,
            "fp32_parity": parity_leg,
            "attempts": 1,
            "train": train,
            "train_strong": train_strong,
            "train_perceptual": train_perceptual,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-train", action="store_true", help="skip the secondary GAN-train-step leg (config 4)")
    ap.add_argument("--no-extras", action="store_true", help="skip the chain / tiled / library-baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
