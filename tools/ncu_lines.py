#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to source lines.

  ncu -i rep.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all lib.so ; nvdisasm -g -c x.cubin > dis.txt
  python tools/ncu_lines.py sass.csv dis.txt <mangled kernel substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst = [(r[1].strip(), float(r[ie] or 0), float(r[ns] or 0)) for r in rows[2:] if len(r) > ie]
    # source line of every instruction of the kernel, in order
    lines, cur, infn = [], ("?", 0), False
    for ln in open(dis):
        if ln.startswith(".text."):
            infn = kern in ln
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    print("sass instructions: ncu %d, nvdisasm %d" % (len(inst), len(lines)))
    n = min(len(inst), len(lines))
    agg = defaultdict(lambda: [0.0, 0.0, 0])
    for i in range(n):
        a = agg[lines[i]]
        a[0] += inst[i][1]
        a[1] += inst[i][2]
        a[2] += 1
    tot = sum(a[0] for a in agg.values())
    tots = sum(a[1] for a in agg.values())
    print("total warp instructions %.3g, samples %d" % (tot, tots))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% inst %5.1f%% stall  sass=%3d  %s:%d" % (100 * a[0] / tot, 100 * a[1] / max(1, tots), a[2], f, l))


if __name__ == "__main__":
    main()
