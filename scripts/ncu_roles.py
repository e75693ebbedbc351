#!/usr/bin/env python
"""Warp-stall samples of fs2_update_ws_kernel split by warp role (screener / applier), from an ncu report taken with
--set full --import-source on.  usage: python scripts/ncu_roles.py report.ncu-rep [bucket]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size"]
print(d.get("Kernel Name", "")[:70])
for k in keys:
    print("  %-62s %s" % (k, d.get(k)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = rows[2:]
ci = {n: i for i, n in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
role_of = []
role = "prologue"
for r in data:
    if "USETMAXREG.DEALLOC" in r[1]:
        role = "screener"
    elif "USETMAXREG.TRY_ALLOC" in r[1]:
        role = "applier"
    role_of.append(role)
tot = collections.Counter()
inst = collections.Counter()
st = collections.defaultdict(collections.Counter)
for r, ro in zip(data, role_of):
    tot[ro] += int(r[ci["# Samples"]])
    inst[ro] += int(r[ci["Instructions Executed"]])
    for s in stalls:
        st[ro][s[6:]] += int(r[ci[s]])
for ro in ("prologue", "screener", "applier"):
    print("%-9s samples %7d  warp-instr %.1fM  %s" % (ro, tot[ro], inst[ro] / 1e6, st[ro].most_common(7)))
if bucket:
    for lo in range(0, len(data), bucket):
        hi = min(lo + bucket, len(data))
        t = sum(int(r[ci["# Samples"]]) for r in data[lo:hi])
        if t < 200:
            continue
        c = collections.Counter()
        for r in data[lo:hi]:
            for s in stalls:
                c[s[6:]] += int(r[ci[s]])
        print(lo, hi, role_of[lo], "samples", t, "instr %.1fM" % (sum(int(r[ci["Instructions Executed"]]) for r in data[lo:hi]) / 1e6), c.most_common(4))
top = sorted(range(len(data)), key=lambda k: -int(data[k][ci["# Samples"]]))[:30]
for k in sorted(top):
    r = data[k]
    print("%5d %-9s %-64s %6s %10s %s" % (k, role_of[k], r[1].strip()[:64], r[ci["# Samples"]], r[ci["Instructions Executed"]],
                                            max(stalls, key=lambda s: int(r[ci[s]]))[6:]))
