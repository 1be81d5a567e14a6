"""Executed-instruction and stall-sample breakdown of a limb-kernel ncu report by source stage.
usage: python tools/ncu_stages.py <file.ncu-rep>"""
import collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ambersim_b200", "csrc")
src = open(os.path.join(root, "abr_limb.cuh")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*// -{20,} (.*)", l)
    if m: marks.append((i, m.group(1)))
    m = re.match(r"^(?:template <[^>]*>\s*)?(?:__device__ __forceinline__|__device__ __noinline__|__global__) .*?(\w+)\(", l)
    if m: marks.append((i, "fn " + m.group(1)))
marks.sort()
def stage(ln):
    name = "?"
    for i, n in marks:
        if i <= ln: name = n
        else: break
    return name
def _i(x):
    try: return int(x)
    except ValueError: return 0
fname = "?"; hdr = None; agg = collections.Counter(); smp = collections.Counter(); noi = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].rsplit("/", 1)[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iN = hdr.index("stall_no_inst"); continue
    if hdr is None or not r or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    k = stage(ln) if fname == "abr_limb.cuh" else fname
    agg[k] += _i(r[iI]); smp[k] += _i(r[iS]); noi[k] += _i(r[iN])
tot = sum(agg.values()); tots = sum(smp.values())
print(f"executed warp-instructions {tot:,}; stall samples {tots:,} (no_instruction {sum(noi.values()):,})")
for k, v in agg.most_common(32): print(f"{100*v/tot:5.1f}% inst {100*smp[k]/max(tots,1):5.1f}% samples {100*noi[k]/max(smp[k],1):5.1f}% of them no_inst   {k}")
