#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel instance from `ncu --page source --csv`.
usage: ncu -i rep --page source --csv [--kernel-name regex:..] > src.csv; python tools/ncu_hot.py src.csv [instance] [top]"""
import csv
import sys


def main(path, inst=0, top=25):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    starts.append(len(rows))
    s, e = starts[inst], starts[inst + 1]
    hdr = rows[s + 1]
    body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in body)
    inst_exec = sum(int(r[ix["Instructions Executed"]]) for r in body)
    print("kernel:", rows[s][1], "| instance", inst, "| samples", tot, "| warp-instructions", inst_exec, "| SASS lines", len(body))
    agg = {c: sum(int(r[ix[c]]) for r in body) for c in stall_cols}
    print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, tot)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(body)), key=lambda k: -int(body[k][ix["# Samples"]]))[:top]
    for k in sorted(order):
        r = body[k]
        st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("%5d %5.1f%% exec=%9s  %-62s %s" % (k, 100.0 * int(r[ix["# Samples"]]) / max(1, tot), r[ix["Instructions Executed"]],
                                              r[ix["Source"]].strip()[:62], " ".join("%s:%d" % (n, v) for v, n in st if v)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
