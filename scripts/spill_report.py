#!/usr/bin/env python
"""Local-memory (spill) accesses of fs2_update_ws_kernel by warp role, from the SASS of a built object / library.
usage: python scripts/spill_report.py [path to .so or .o]   (needs cuobjdump)"""
import re
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "fast_slam_b200/libfs2.so"
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur = None
funcs = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        funcs[cur].append(line)
for name, lines in funcs.items():
    if "fs2_update_ws_kernel" not in name:
        continue
    role = "prologue"
    stats = {}
    for l in lines:
        if "USETMAXREG.DEALLOC" in l:
            role = "screener"
        elif "USETMAXREG.TRY_ALLOC" in l:
            role = "applier"
        d = stats.setdefault(role, {"n": 0, "LDL": 0, "STL": 0, "UR": 0})
        d["n"] += 1
        if re.search(r"\bLDL", l):
            d["LDL"] += 1
        if re.search(r"\bSTL", l):
            d["STL"] += 1
        if re.search(r"\bUR\d+", l):
            d["UR"] += 1
    print(name[:40], {k: v for k, v in stats.items()})
