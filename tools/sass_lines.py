"""Attribute SASS instruction counts of a kernel to source lines (needs -lineinfo).
usage: python tools/sass_lines.py <cubin> <kernel-substring> [top]"""
import collections, re, subprocess, sys
cubin, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
cnt = collections.Counter(); infn = False; key = "?"; total = 0
for line in out.splitlines():
    if line.lstrip().startswith(".section"):
        infn = ".text." in line and pat in line
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        key = f"{m.group(1).rsplit('/',1)[-1]}:{m.group(2)}"; continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        cnt[key] += 1; total += 1
print("total", total)
for k, v in cnt.most_common(top): print(v, k)
