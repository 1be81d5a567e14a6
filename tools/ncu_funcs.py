"""Aggregate an ncu source-page CSV by enclosing device function of abr_step.cuh / abr_kernels.cuh.
usage: python tools/ncu_funcs.py <file.ncu-rep> [world_steps]"""
import csv, io, subprocess, sys, collections, re
from pathlib import Path
rep = sys.argv[1]; ws = float(sys.argv[2]) if len(sys.argv) > 2 else None
root = Path(__file__).resolve().parents[1] / "ambersim_b200/csrc"
starts = {}
for fn in ("abr_step.cuh", "abr_kernels.cuh"):
    lst = []
    for n, line in enumerate((root / fn).read_text().splitlines(), 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?(?:__global__|__device__)[^(]*?\b(\w+)\(", line)
        if m: lst.append((n, m.group(1)))
    starts[fn] = lst
def func_of(fn, line):
    name = "?"
    for n, f in starts.get(fn, []):
        if n <= line: name = f
        else: break
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"; hdr = None; agg = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].rsplit("/", 1)[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or not r or r[0] == "": continue
    f = func_of(fname, int(r[0])) if fname in starts else fname
    agg[f] += int(r[iI] or 0); smp[f] += int(r[iS] or 0); tot += int(r[iI] or 0); tots += int(r[iS] or 0)
print(f"total warp-instructions {tot:,}" + (f"  = {tot/ws:,.0f} per world-step" if ws else ""))
for f, v in agg.most_common(40):
    print(f"{100*v/tot:5.1f}% inst {100*smp[f]/max(tots,1):5.1f}% smp " + (f"{v/ws:8.0f}/ws  " if ws else "") + f)
