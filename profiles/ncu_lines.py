#!/usr/bin/env python
"""Per-source-line roll-up of an ncu report (no GPU needed).

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g <cubin> > dis.txt
    python profiles/ncu_lines.py sass.csv dis.txt <mangled kernel substring> [top]

ncu's CSV source page lists SASS instructions (address, executed count, stall
samples); nvdisasm -g tags every SASS offset with its file:line.  Joined on the
offset, summed per line (of the innermost inlined function).
"""
import collections
import csv
import re
import sys


def load_lines(dis, kernel):
    cur, lines, inside = None, {}, False
    for ln in open(dis):
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return lines


def main():
    sass, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = load_lines(dis, kernel)
    rows = list(csv.reader(open(sass)))
    hdr = rows[1]
    ia, ie, it, isamp = (hdr.index("Address"), hdr.index("Instructions Executed"),
                         hdr.index("Thread Instructions Executed"), hdr.index("# Samples"))
    base = int(rows[2][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        key = lines.get(off, (("?", 0), ""))[0]
        v = (int(r[ie]), int(r[it]), int(r[isamp]))
        for i in range(3):
            agg[key][i] += v[i]
            tot[i] += v[i]
    print("total: %d warp-instr, %d thread-instr (%.2f threads/instr), %d samples"
          % (tot[0], tot[1], tot[1]/max(tot[0], 1), tot[2]))
    byfile = collections.defaultdict(lambda: [0, 0, 0])
    for k, v in agg.items():
        for i in range(3):
            byfile[k[0] if k else "?"][i] += v[i]
    for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  %-24s %5.1f%% instr  %5.1f%% samples  %.1f thr/instr"
              % (k, 100.*v[0]/tot[0], 100.*v[2]/max(tot[2], 1), v[1]/max(v[0], 1)))
    print("top lines:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-24s:%-4d %5.2f%% instr  %5.2f%% samples  %.1f thr/instr"
              % (k[0], k[1], 100.*v[0]/tot[0], 100.*v[2]/max(tot[2], 1), v[1]/max(v[0], 1)))


if __name__ == "__main__":
    main()
