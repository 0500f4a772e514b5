#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/ (read here in the build container, no GPU needed).

  python tools/summarize_ncu.py list  gpurun_out/launches.csv  "<what was profiled>"  > profiles/rNN_..._summary.json
  python tools/summarize_ncu.py full  gpurun_out/prof_raw.csv                          > profiles/rNN_ncu_full_....json
`list`: CSV log of `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`.
`full`: CSV of `ncu -i prof.ncu-rep --page raw --csv`.
"""
import csv
import json
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def rows_of(path):
    lines = open(path, newline="").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def to_float(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def unit_scale(unit, metric):
    u = (unit or "").lower()
    if metric.startswith("gpu__time"):
        return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
    if "bytes" in metric:
        return {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
    return 1.0


def do_list(path, what):
    per = OrderedDict()
    for r in rows_of(path):
        e = per.setdefault(r["ID"], {"kernel": r["Kernel Name"][:100], "us": 0.0, "rd": 0.0, "wr": 0.0})
        m, v = r["Metric Name"], to_float(r["Metric Value"])
        if v is None:
            continue
        v *= unit_scale(r.get("Metric Unit"), m)
        if m.startswith("gpu__time"):
            e["us"] = v
        elif m == "dram__bytes_read.sum":
            e["rd"] = v
        elif m == "dram__bytes_write.sum":
            e["wr"] = v
    launches = list(per.values())
    tot = sum(l["us"] for l in launches)
    by = OrderedDict()
    for l in launches:
        k = by.setdefault(l["kernel"], {"kernel": l["kernel"], "launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
        k["launches"] += 1
        k["us"] += l["us"]
        k["dram_read_MB"] += l["rd"]
        k["dram_write_MB"] += l["wr"]
    for k in by.values():
        k["share"] = k["us"] / tot if tot else 0.0
    import hashlib, glob, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "esrganplus_b200", "csrc")
    hh = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) + glob.glob(os.path.join(d, "*.inl")) +
                    glob.glob(os.path.join(d, "*.h")) + [os.path.join(root, "include", "esrp.h")]):
        hh.update(open(f, "rb").read())
    out = {"step": what, "csrc_sha256": hh.hexdigest(), "launches": len(launches), "total_us": tot, "dram_read_MB": sum(l["rd"] for l in launches),
           "dram_write_MB": sum(l["wr"] for l in launches), "by_kernel": sorted(by.values(), key=lambda k: -k["us"]),
           "sequence_us": [round(l["us"], 2) for l in launches[:160]],
           "sequence_kernels": [l["kernel"][:48] for l in launches[:160]]}
    print(json.dumps(out, indent=1))


def do_full(path):
    per = OrderedDict()
    rows = rows_of(path)
    for r in rows:
        # --page raw --csv: one row per launch, one column per metric (units in the second header row)
        if "Kernel Name" in r and r.get("ID", "").isdigit():
            e = {"kernel": r["Kernel Name"][:100]}
            for k in KEEP:
                if k in r:
                    e[k] = r[k]
            per[r["ID"]] = e
    print(json.dumps(list(per.values()), indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "list":
        do_list(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        do_full(sys.argv[2])
