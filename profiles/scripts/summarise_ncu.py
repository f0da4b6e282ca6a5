#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python profiles/scripts/summarise_ncu.py launches gpurun_out/launches_r01.csv            -> per-kernel time shares (stdout, markdown)
  python profiles/scripts/summarise_ncu.py full gpurun_out/k2_full_r01.ncu-rep [...]       -> key counters per captured launch (json lines)
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
    "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "sass__inst_executed_shared_loads", "smsp__inst_executed.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "smsp__cycles_active.avg", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "msecond": 1e-3,
        "usecond": 1e-6, "nsecond": 1e-9}


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("xtd::", "")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    tot = OrderedDict()
    for r in rows:
        k = short(r[4])
        ns = float(r[-1].replace(",", ""))
        d = tot.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += ns
    total = sum(v[1] for v in tot.values())
    print(f"| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.2f} | {100 * ns / total:.2f} % |")
    print(f"| all ({len(rows)} launches) | | {total / 1e6:.2f} | |")


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            rec = OrderedDict(report=p.split("/")[-1])
            for h, u, v in zip(hdr, units, vals):
                if h == "Kernel Name":
                    rec["kernel"] = short(v)
                elif h in ("Grid Size", "Block Size"):
                    rec[h] = v
                elif h in KEYS:
                    try:
                        x = float(v.replace(",", ""))
                    except ValueError:
                        continue
                    if u in UNIT and ("bytes" in h or "time" in h):
                        x *= UNIT[u]
                        u = "byte" if "bytes" in h else "s"
                    rec[h] = x if not u else [x, u]
            print(json.dumps(rec))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
