#!/usr/bin/env python
"""One C3 step from an ncu launch list taken inside bench.py's profiler range (BIDI_NCU_RANGE=1,
metrics gpu__time_duration.sum, smsp__inst_executed.sum, dram__bytes_read.sum, dram__bytes_write.sum):
a markdown table of the launches of ONE step and the DRAM traffic per agent-step as JSON.

usage: python tools/steady_launches.py launches.csv agents [traffic.json] > profiles/xxx.md
The capture holds the launches of two steps (the range brackets `--steps 2`); the second step is used."""
import csv
import glob
import hashlib
import io
import json
import os
import sys
from collections import OrderedDict


def kernel_source_sha():
    """sha256 over the CUDA sources of the library: bench.py marks `roofline.traffic` stale when the build it runs
    was compiled from other sources than the ones this capture was taken on."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bidinet_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(root, "*.cu")) + glob.glob(os.path.join(root, "*.cuh")) + glob.glob(os.path.join(root, "*.inl")) + glob.glob(os.path.join(root, "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def main():
    path, agents = sys.argv[1], int(sys.argv[2])
    lines = [l for l in open(path) if l.startswith('"')]
    launches = OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        d = launches.setdefault(int(row["ID"]), {"name": row["Kernel Name"].split("(")[0].replace("void ", "").replace("bidi::", "")})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        elif row["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[row["Metric Name"]] = v
    ls = list(launches.values())
    ls = ls[len(ls) // 2:]  # the second of the two captured steps
    T = sum(l["gpu__time_duration.sum"] for l in ls)
    WF, BC = "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
    banks = all(WF in l and BC in l for l in ls)  # the shared-memory bank-conflict counters, when the capture has them
    print("| # | kernel | us | warp instructions | DRAM read MB | DRAM written MB | TB/s | share |" + (" shared-memory wavefronts | of them bank-conflict replays |" if banks else ""))
    print("|---:|---|---:|---:|---:|---:|---:|---:|" + ("---:|---:|" if banks else ""))
    tot_b = col_b = col_t = 0.0
    col_n = 0
    for k, l in enumerate(ls):
        t, rd, wr = l["gpu__time_duration.sum"], l["dram__bytes_read.sum"], l["dram__bytes_write.sum"]
        tot_b += rd + wr
        if l["name"].startswith("k_columns"):
            col_b += rd + wr; col_t += t; col_n += 1
        print("| %d | %s | %.1f | %d | %.1f | %.1f | %.2f | %.1f%% |" % (k, l["name"], t, l["smsp__inst_executed.sum"], rd / 1e6, wr / 1e6,
                                                                       (rd + wr) / t / 1e6, 100 * t / T)
              + (" %d | %.1f%% |" % (l[WF], 100.0 * l[BC] / max(1.0, l[WF])) if banks else ""))
    print("| | **step** | %.1f | | | %.1f total | %.2f | 100%% |" % (T, tot_b / 1e6, tot_b / T / 1e6))
    print()
    print("k_columns*: %d launches per step, %.1f MB of DRAM traffic per step = %.3f MB per agent-step, %.1f%% of the step's time."
          % (col_n, col_b / 1e6, col_b / agents / 1e6, 100 * col_t / T))
    if len(sys.argv) > 3:
        json.dump({"columns": {"dram_bytes_per_agent_step": col_b / agents, "launches_per_step": col_n, "agents": agents,
                               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, k_columns*, one steady-state step (tools/steady_launches.py)"},
                   "whole_step": {"dram_bytes_per_agent_step": tot_b / agents},
                   "kernel_source_sha": kernel_source_sha()}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
