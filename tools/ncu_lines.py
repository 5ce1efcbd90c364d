#!/usr/bin/env python
"""Per-source-line profile: join an ncu SASS source page with nvdisasm -g line markers.

usage: python tools/ncu_lines.py report.ncu-rep lib.so 'kernelILb1ELb1ELb0' [top_n]
Prints, per CUDA source line, warp-level instructions executed and stall samples (share of the kernel)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kpat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, line, infn = None, None, False
    for l in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            cur = m.group(1); infn = kpat in cur; line = None
            if infn: lines_of[cur] = {}
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
        if m:
            lines_of[cur][int(m.group(1), 16)] = line
if not lines_of:
    sys.exit("kernel pattern not found in " + lib)
kname = sorted(lines_of)[0]
lmap = lines_of[kname]
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(rows[hi])}
body = [r for r in rows[hi + 1:] if len(r) >= len(rows[hi])]
base = int(body[0][ix["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in body:
    off = int(r[ix["Address"]], 16) - base
    key = lmap.get(off)
    n, t, s = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    for a in (agg[key], tot):
        a[0] += n; a[1] += t; a[2] += s
src = {}
print(f"{kname[-60:]}: warp-inst {tot[0]:,} samples {tot[2]:,}")
for key, (n, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", key[0])
        if path not in src and os.path.exists(path):
            src[path] = open(path).read().splitlines()
        if path in src and key[1] - 1 < len(src[path]):
            text = src[path][key[1] - 1].strip()[:90]
    print(f"{str(key[1]) if key else '?':>5} {100*n/tot[0]:6.2f}% inst  {100*s/max(tot[2],1):6.2f}% stall  thr {t/max(n,1):4.1f} | {text}")
