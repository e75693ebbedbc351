#!/usr/bin/env python
"""SASS of one kernel of libfs2.so: static opcode histogram, the lines that prove the sm_100a features
(UBLKCP = bulk TMA copy, USETMAXREG = setmaxnreg warp specialisation, SYNCS = mbarrier), spill traffic
(STL / LDL) and, per role of the warp-specialised update kernel, the instruction counts of its loops.

    python scripts/sass_report.py [--lib fast_slam_b200/libfs2.so] [--kernel fs2_update_ws_kernel] [--dump FILE]

Static counts only (no GPU needed); the executed counts are in the ncu summaries next to this file's output.
"""
import argparse
import collections
import re
import subprocess
import sys


def disassemble(lib, kernel):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    lines, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
    return lines


def opcode(txt):
    t = re.sub(r"^@!?U?P\d+\s+", "", txt)
    return t.split()[0].split(".")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="fast_slam_b200/libfs2.so")
    ap.add_argument("--kernel", default="fs2_update_ws_kernel")
    ap.add_argument("--dump", default=None, help="write the full listing (address, instruction) here")
    a = ap.parse_args()
    ins = disassemble(a.lib, a.kernel)
    if not ins:
        sys.exit("kernel %s not found in %s" % (a.kernel, a.lib))
    if a.dump:
        with open(a.dump, "w") as f:
            for i, (ad, t) in enumerate(ins):
                f.write("%5d %05x  %s\n" % (i, ad, t))
    print("# %s in %s: %d SASS instructions" % (a.kernel, a.lib, len(ins)))
    hist = collections.Counter(opcode(t) for _, t in ins)
    print("# static opcode histogram (top 30)")
    for k, v in hist.most_common(30):
        print("%-12s %5d" % (k, v))
    print("# sm_100a feature lines")
    for i, (ad, t) in enumerate(ins):
        if re.search(r"UBLKCP|USETMAXREG|SYNCS|ELECT|FMUL2|FFMA2|FADD2", t):
            print("%5d %05x  %s" % (i, ad, t))
    spills = [(i, ad, t) for i, (ad, t) in enumerate(ins) if re.search(r"\b(STL|LDL)\b", opcode(t))]
    print("# local-memory (spill) instructions: %d" % len(spills))
    for i, ad, t in spills:
        print("%5d %05x  %s" % (i, ad, t))
    # backward branches = loops: report [target, branch] ranges and their lengths
    addr_to_idx = {ad: i for i, (ad, _) in enumerate(ins)}
    loops = []
    for i, (ad, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= ad and tgt in addr_to_idx:
                loops.append((addr_to_idx[tgt], i))
    print("# loops (backward branches): first..last instruction index, length")
    for lo, hi in sorted(set(loops)):
        print("%5d..%5d  %4d" % (lo, hi, hi - lo + 1))


if __name__ == "__main__":
    main()
