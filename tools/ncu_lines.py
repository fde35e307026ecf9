#!/usr/bin/env python
"""Per-source-line warp-stall samples and executed instructions of one kernel.

Joins the SASS page of an ncu report (`ncu -i rep --page source --csv --kernel-name ... > sass.csv`)
with the line table of the same build (`cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt`):
instructions are matched in order, the n-th SASS row of the page is the n-th instruction of the function.

usage: ncu_lines.py sass.csv dis.txt <mangled function name> [min_percent] [instance]"""
import csv
import re
import sys
from collections import defaultdict


def line_table(dis, fn):
    out, inside, cur = [], False, ("?", 0)
    for ln in open(dis):
        if ln.startswith(".text."):
            inside = ln.strip() == ".text." + fn + ":"
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    return out


def main():
    sass, dis, fn = sys.argv[1:4]
    minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    rows = list(csv.reader(open(sass)))
    inst = int(sys.argv[5]) if len(sys.argv) > 5 else 0   # which kernel instance of the page
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = his[inst]
    end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    rows = rows[:end]
    hdr = rows[hi]
    si, ii, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[hi + 1:] if len(r) > ii]
    lt = line_table(dis, fn)
    if len(lt) != len(body):
        print("warning: %d SASS rows in the report, %d instructions in the disassembly" % (len(body), len(lt)))
    agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    ts = ti = 0.0
    for k, r in enumerate(body):
        key = lt[k] if k < len(lt) else ("?", 0)
        s, n = float(r[si] or 0), float(r[ii] or 0)
        agg[key][0] += s
        agg[key][1] += n
        for j in stall:
            agg[key][2][hdr[j]] += float(r[j] or 0)
        ts += s
        ti += n
    print("samples %d, warp instructions %d" % (ts, ti))
    for key in sorted(agg):
        s, n, st = agg[key]
        if 100 * s / ts >= minp or 100 * n / ti >= minp:
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print("%-24s %5d  %5.1f%% samples %5.1f%% instr   %s" % (key[0], key[1], 100 * s / ts, 100 * n / ti,
                  " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in top if v > 0)))


if __name__ == "__main__":
    main()
