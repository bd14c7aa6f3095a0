#!/usr/bin/env python
"""Aggregate the source page of an .ncu-rep by device function (development aid):
    python tools/ncu_by_function.py rep.ncu-rep [transport.cuh of the profiled build]"""
import bisect, csv, io, re, subprocess, sys
rep = sys.argv[1]
srcf = sys.argv[2] if len(sys.argv) > 2 else "opengate_b200/csrc/transport.cuh"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if r and r[0].isdigit() and hdr and len(r) == len(hdr):
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        g = lambda k: int(d.get(k) or 0)
        agg.append((cur, int(r[0]), g("# Samples"), g("Instructions Executed"), g("stall_no_inst"), g("stall_long_sb"), g("stall_short_sb"), g("stall_wait"), g("Thread Instructions Executed")))
src = open(srcf).read().split("\n")
funcs = []
for i, l in enumerate(src):
    if re.match(r"\s*(__device__|__global__)", l):
        txt = l + " " + (src[i + 1] if i + 1 < len(src) else "")
        m = re.search(r"\b(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)", "", txt))
        funcs.append((i + 1, m.group(1)))
starts = [f[0] for f in funcs]
ts = sum(a[2] for a in agg) or 1; ti = sum(a[3] for a in agg) or 1
byf = {}
for f, ln, smp, ins, ni, lsb, ssb, wt, tins in agg:
    name = f
    if f == "transport.cuh":
        i = bisect.bisect_right(starts, ln) - 1
        name = funcs[i][1] if i >= 0 else "?"
    a = byf.setdefault(name, [0] * 7)
    for k, v in enumerate((smp, ins, ni, lsb, ssb, wt, tins)): a[k] += v
print(f"{'function':26s} smp%  ins%  lanes | no_inst long_sb short_sb wait  (% of all samples)")
for k, v in sorted(byf.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"{k:26s} {100*v[0]/ts:5.1f} {100*v[1]/ti:5.1f} {v[6]/max(v[1],1):5.1f} | {100*v[2]/ts:5.1f} {100*v[3]/ts:7.1f} {100*v[4]/ts:7.1f} {100*v[5]/ts:5.1f}")
