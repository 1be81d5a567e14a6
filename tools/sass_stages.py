"""Static SASS instruction histogram of a limb kernel by source stage.
usage: python tools/sass_stages.py <object.o> [kernel-substring]"""
import collections, re, subprocess, sys, tempfile, os
obj = os.path.abspath(sys.argv[1]); pat = sys.argv[2] if len(sys.argv) > 2 else "rollout"
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur = None; cnt = collections.Counter(); infunc = False; ops = collections.Counter()
for line in dis.splitlines():
    if line.startswith(".text."): infunc = pat in line
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if infunc and m: cnt[cur] += 1; ops[m.group(1).split(".")[0]] += 1
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
agg = collections.Counter()
for (f, ln), c in cnt.items(): agg[stage(ln) if f == "abr_limb.cuh" else f] += c
tot = sum(cnt.values()); print("static instructions:", tot)
for k, v in agg.most_common(30): print(f"{v:6d} {100*v/tot:5.1f}%  {k}")
print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(25)))
