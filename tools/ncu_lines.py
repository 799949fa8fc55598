#!/usr/bin/env python3
"""Attribute ncu warp-stall samples of a kernel to source lines.

ncu's CSV source page is SASS-only; this joins it with `nvdisasm -g` line info of the same cubin.
usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict


def main():
    rep, lib, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    off2line = {}
    for f in os.listdir(tmp):
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        inside, line = False, None
        for ln in dis.split("\n"):
            if ln.startswith(".text."):
                inside = kname in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
            if m:
                off2line[int(m.group(1), 16)] = line
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    isrc = hdr.index("Source") if "Source" in hdr else None
    ops = defaultdict(int)
    per = defaultdict(lambda: [0, 0, defaultdict(int)])
    started = False
    for r in rows:
        if not r or r[0] == "Address" or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        ln = off2line.get(a - base)
        e = per[ln]
        e[0] += int(r[isamp] or 0)
        e[1] += int(r[iinst] or 0)
        if isrc is not None:
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
            if m:
                ops[m.group(1)] += int(r[iinst] or 0)
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                e[2][hdr[i]] += v
    tot = sum(e[0] for e in per.values())
    toti = sum(e[1] for e in per.values())
    print("total samples %d, warp instructions %d" % (tot, toti))
    for ln, e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(e[2].items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%%  inst %5.1f%%  %-28s %s" % (100.0 * e[0] / max(tot, 1), 100.0 * e[1] / max(toti, 1), ln,
                                                 " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in st)))


    if ops:
        print("\nexecuted warp instructions by opcode (top 30):")
        for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:30]:
            print("  %-12s %6.2f%%  %d" % (k, 100.0 * v / max(toti, 1), v))
    print("\ntop source lines by executed instructions:")
    for ln, e in sorted(per.items(), key=lambda kv: -kv[1][1])[:25]:
        print("  inst %5.1f%%  samples %5.1f%%  %s" % (100.0 * e[1] / max(toti, 1), 100.0 * e[0] / max(tot, 1), ln))


if __name__ == "__main__":
    main()
