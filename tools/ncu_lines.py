"""Aggregate an ncu source-page CSV by source line: executed instructions and stall samples.
usage: python tools/ncu_lines.py <file.ncu-rep> [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"; hdr = None; agg = collections.OrderedDict(); tot_i = tot_s = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].rsplit("/", 1)[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or not r: continue
    if r[0] != "":  # source line summary row
        key = f"{fname}:{r[0]}"
        inst, smp = int(r[iI] or 0), int(r[iS] or 0)
        a = agg.setdefault(key, [0, 0, r[1].strip()[:90]])
        a[0] += inst; a[1] += smp; tot_i += inst; tot_s += smp
print(f"total instructions {tot_i:,}  samples {tot_s:,}")
print("--- by instructions executed")
for k, (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*i/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% smp  {k:28s} {src}")
