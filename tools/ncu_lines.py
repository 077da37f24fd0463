"""Per-source-line executed instructions and stall samples of one kernel: joins `nvdisasm -g` line info of the cubin
with the SASS page of an .ncu-rep by instruction order.
usage: python tools/ncu_lines.py report.ncu-rep libtopopt_cuda.so <mangled-name-substring> [topN]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, sub = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
syms = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"\.text\.(\S+)", syms)))
cands = [n for n in names if sub in n]
assert cands, "no kernel matches"
fun = cands[0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
lines = []
cur = ("?", 0)
inside = False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        inside = fun in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
print("kernel", fun, "sass instrs: nvdisasm", len(lines), "ncu", len(data))
n = min(len(lines), len(data))
ex = collections.Counter()
sm = collections.Counter()
for k in range(n):
    ex[lines[k]] += int(data[k][ix["Instructions Executed"]] or 0)
    sm[lines[k]] += int(data[k][ix["# Samples"]] or 0)
tot_e, tot_s = sum(ex.values()), sum(sm.values())
print(f"total executed {tot_e}  samples {tot_s}")
srcs = {}
for (f, l), v in sorted(ex.items(), key=lambda kv: -kv[1])[:topn]:
    if f not in srcs:
        for root in ("topopt.jl_b200/csrc",):
            p = os.path.join(root, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{f}:{l:<5d} exec {100 * v / tot_e:5.1f}%  samples {100 * sm[(f, l)] / max(1, tot_s):5.1f}%  {text}")
