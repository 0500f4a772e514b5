#!/usr/bin/env python
"""Copies / condenses this round's GPU outputs from gpurun_out/ (scratch) into profiles/ (tracked) under r02_ names.
Large logs (compute-sanitizer synccheck, the NVML poller) are condensed to what DESIGN.md quotes.  Re-runnable."""
import collections
import json
import os
import re
import shutil
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def copy(src, dst):
    s = os.path.join(SRC, src)
    if os.path.exists(s):
        shutil.copyfile(s, os.path.join(DST, dst))
        print("copied", src, "->", dst)
    else:
        print("MISSING", src)


def cat(srcs, dst, header=""):
    with open(os.path.join(DST, dst), "w") as out:
        if header:
            out.write(header.rstrip() + "\n")
        for title, s in srcs:
            p = os.path.join(SRC, s)
            if not os.path.exists(p):
                print("MISSING", s)
                continue
            out.write(f"## {title} ({s})\n")
            out.write(open(p, errors="replace").read().rstrip() + "\n")
    print("wrote", dst)


def synccheck_summary(logs, dst):
    """Group compute-sanitizer synccheck reports by (message, kernel, source line, barrier address) with the blocks hit."""
    out = {}
    for title, name in logs:
        p = os.path.join(SRC, name)
        if not os.path.exists(p):
            print("MISSING", name)
            continue
        groups = collections.OrderedDict()
        cur = None
        for line in open(p, errors="replace"):
            line = line.rstrip()
            m = re.match(r"========= (Barrier error detected\..*|.*[Ee]rror.*detected.*)$", line)
            if m and "ERROR SUMMARY" not in line:
                cur = {"msg": m.group(1).strip(), "block": None, "addr": None, "frames": []}
                continue
            if cur is None:
                continue
            m = re.search(r"by thread \((\d+),\d+,\d+\) in block \((\d+),", line)
            if m:
                cur["thread"], cur["block"] = int(m.group(1)), int(m.group(2))
            m = re.search(r"shared address (0x[0-9a-f]+)", line)
            if m:
                cur["addr"] = m.group(1)
            m = re.search(r"Device Frame: (?:void )?([\w:]+)[<(].* in ([\w.]+:\d+)", line)
            if m:
                cur["frames"].append(f"{m.group(1)} {m.group(2)}")
            if line.strip() == "=========":
                key = (cur["msg"], cur["addr"], tuple(cur["frames"][-1:]))
                g = groups.setdefault(key, {"count": 0, "blocks": set(), "warps": set()})
                g["count"] += 1
                g["blocks"].add(cur["block"])
                if "thread" in cur:
                    g["warps"].add(cur["thread"] // 32)
                cur = None
        text = open(p, errors="replace").read()
        tail = [l for l in text.splitlines() if "ERROR SUMMARY" in l or re.search(r"\d+ (passed|failed)", l)]
        out[title] = {"log": name, "printed_reports": [
            {"message": k[0], "barrier_shared_address": k[1], "kernel_frame": list(k[2]), "reports": v["count"],
             "blocks": sorted(b for b in v["blocks"] if b is not None), "warps": sorted(v["warps"])} for k, v in groups.items()],
            "tail": tail[-4:]}
    json.dump(out, open(os.path.join(DST, dst), "w"), indent=1)
    print("wrote", dst)


def nvml_summary(name, dst):
    p = os.path.join(SRC, name)
    if not os.path.exists(p):
        print("MISSING", name)
        return
    mhz, watts, reasons = [], [], collections.Counter()
    for line in open(p, errors="replace"):
        parts = [x.strip() for x in line.split(",")]
        if len(parts) < 3:
            continue
        try:
            mhz.append(float(parts[0].split()[0]))
            watts.append(float(parts[1].split()[0]))
        except ValueError:
            continue
        reasons[", ".join(parts[2:])] += 1
    busy = [m for m, w in zip(mhz, watts) if w > 300]
    json.dump({"log": name, "what": "nvidia-smi polled every 20 ms beside the stress runs (tools/gpu_stress_nvml.sh)",
               "samples": len(mhz), "samples_above_300W": len(busy),
               "sm_mhz_median_above_300W": statistics.median(busy) if busy else None,
               "sm_mhz_min_max": [min(mhz), max(mhz)] if mhz else None, "power_w_max": max(watts) if watts else None,
               "throttle_reason_columns": dict(reasons.most_common(8))}, open(os.path.join(DST, dst), "w"), indent=1)
    print("wrote", dst)


def main():
    os.makedirs(DST, exist_ok=True)
    # --- robustness
    copy("n_stress.jsonl", "r02_stress_100_forward_processes_with_nvml_poller.jsonl")
    copy("n_stress_bench.jsonl", "r02_stress_30_bench_processes_with_nvml_poller.jsonl")
    nvml_summary("n_nvml_poll.log", "r02_stress_nvml_poller_summary.json")
    cat([(f"full GPU suite, run {i}", f"n_pytest_{i}.log") for i in range(1, 8)] + [("stress driver output", "n_call.log")],
        "r02_stress_seven_full_gpu_suites.log", "# seven consecutive `pytest -m gpu` runs + the stress driver's own log (tools/gpu_stress_nvml.sh)")
    cat([("racecheck, row kernel, alternating issuers (row_alt 2)", "n_sanitizer_racecheck_alt2.log"),
         ("racecheck, row kernel, split issuers (row_alt 0)", "n_sanitizer_racecheck_alt0.log")],
        "r02_compute_sanitizer_racecheck_row_kernel.log")
    synccheck_summary([("row kernel, alternating issuers, first run", "n_sanitizer_synccheck_alt2.log"),
                       ("row kernel, split issuers, first run", "n_sanitizer_synccheck_alt0.log"),
                       ("row kernel alt 2, tok[] moved by 8 bytes (ESRP_SYNCCHECK_PAD)", "p_sync_row_alt2_padded.log"),
                       ("row kernel alt 2 after the parity-trick-free waits", "p_sync_row_alt2.log"),
                       ("row kernel, pre-arrived barriers, tmem_holder on its own line, no PDL", "r_sync_row.log"),
                       ("tile kernel", "p_sync_tile.log"),
                       ("chain kernel", "q_sync_chain.log")], "r02_compute_sanitizer_synccheck_summary.json")
    # --- conv chain: measurements behind the negative result
    cat([("chain vs launch-per-conv, config 2 forward (ESRP_NO_CHAIN=1 = plain)", "l_ab.jsonl"), ("earlier protocol states", "k_ab.jsonl"),
         ("dbg switches: 2 = no neighbour-flag wait, 266 = no flag wait + no TMA + no epilogue", "j_call.log")],
        "r02_chain_ab_config2.log")
    copy("m_batch.jsonl", "r02_chain_vs_plain_batch_sizes.jsonl")
    for tag in ("dbg0", "dbg2", "dbg266"):
        copy(f"k_trace_{tag}.json", f"r02_chain_timeline_{tag}.json")
    copy("k_diag.jsonl", "r02_chain_parity_vs_storage_precision.jsonl")
    # --- solver / training
    copy("q_grad_metrics.log", "r02_gradient_parity_metrics.log")
    cat([("32 crops / 4 crops per step, native solver (tools/bench_train.py)", "s_train.log"), ("before flat gradients", "q_train.log"),
         ("host profile of the GAN step", "s_prof_host.log")], "r02_train_step_timing.log")
    cat([("two models in one process, before the fix (NaN on the first discriminator pass of the second model)", "u_train_a.log"),
         ("same with switches", "u_train_b.log"), ("after the fix", "w_call.log"), ("NaN-primed allocator, single model", "v_poison.log")],
        "r02_discriminator_lazy_pack_race.log")
    # --- small experiments
    cat([("turn token handed over before the row's MMAs (both issuers in flight): slower, one parity failure", "x_ab.jsonl"),
         ("config-2 forward as one torch CUDA graph vs the engine's stream launches, alternating in one process", "x_graph_ab2.jsonl")],
        "r02_experiments_issue_overlap_and_graph.log")
    copy("t_bench2.json", "r02_bench_line_2gpu_before_pack_race_fix.json")
    copy("s_bench.json", "r02_bench_line_1gpu.json")


if __name__ == "__main__":
    main()
