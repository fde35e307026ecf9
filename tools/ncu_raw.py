#!/usr/bin/env python
"""Key metrics per kernel instance from `ncu -i rep --page raw --csv`.
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_raw.py [--md]"""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smem/blk"), ("sm__warps_active.avg.per_cycle_active", "warps/SM"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%")]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    md = "--md" in sys.argv
    head = [n + ("" if not units[i] else " [" + units[i] + "]") for i, n in cols]
    print(("| " + " | ".join(head) + " |") if md else "\t".join(head))
    if md:
        print("|" + "---|" * len(head))
    for r in rows[2:]:
        vals = [r[i].split("(")[0] if n == "kernel" else r[i] for i, n in cols]
        print(("| " + " | ".join(vals) + " |") if md else "\t".join(vals))


if __name__ == "__main__":
    main()
