#!/usr/bin/env python
"""Stage-level profile of the march kernel: join an ncu SASS source page with nvdisasm's INLINE line chains and
attribute every executed instruction to a stage of the march (where in cloud_march.cu / cloud_march_ray.inl the call site lies).

usage: python tools/ncu_stages.py report.ncu-rep lib.so|cloud_march.cu.o 'cloud_march_kernelILb1ELb1ELb0ELb1E' [csrc directory of that build]

ncu_lines.py answers "which source line"; helpers such as dot / mad3 / mixg are inlined everywhere, so their lines collect a third of
the kernel.  Here the chain  `line 49 inlined at line 68 inlined at line 385 inlined at line 763 ...`  is walked from the innermost
frame outwards until a line falls inside a named stage (the table below, line ranges found by function name in the source)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kpat = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "project-marshmallow_b200", "csrc")   # directory of the sources `lib` was built from
SOURCES = {f: open(os.path.join(CSRC, f)).read().splitlines() for f in ("cloud_march.cu", "cloud_march_ray.inl") if os.path.exists(os.path.join(CSRC, f))}


def find(pat, start=None):
    """(file, line) of the first line matching pat; start = (file, line) restricts the search to that file from that line on"""
    for f, lines in SOURCES.items():
        if start and f != start[0]:
            continue
        for i in range(start[1] - 1 if start else 0, len(lines)):
            if re.search(pat, lines[i]):
                return f, i + 1
    raise SystemExit("pattern not found in the march sources: " + pat)


def body(pat):
    """(file, first, last) line of the function whose signature matches pat (brace matching from its first '{')"""
    f, a = find(pat)
    lines, depth, seen = SOURCES[f], 0, False
    for i in range(a - 1, len(lines)):
        for ch in lines[i]:
            if ch == "{": depth += 1; seen = True
            elif ch == "}": depth -= 1
        if seen and depth == 0:
            return f, a, i + 1
    raise SystemExit("unbalanced braces after " + pat)


def span(first, last):
    assert first[0] == last[0]
    return first[0], first[1], last[1]


def before(pos):
    return pos[0], pos[1] - 1


# stages, innermost-first priority: the first frame (walking outwards) that lies in one of these ranges names the instruction
trip = body(r"void warp_trip\(")
ct = body(r"float cloudTest\(")
trip0, trip1, ct0, ct1 = (trip[0], trip[1]), (trip[0], trip[2]), (ct[0], ct[1]), (ct[0], ct[2])
t_call_test = find(r"density = cloudTest<MARCH_HW", trip0)
t_ballot = find(r"litMask = __ballot_sync", trip0)
t_tail = find(r"if \(r\.alive\) \{", t_ballot)
c_fetch = find(r"Fetch3<HW, P2> dn\(P\.tex\[TEX_LOWRES\]", ct0)
c_blend = find(r"float layerDensity = blendLayers", ct0)
c_cov = find(r"float k = clampg\(REMAP_C\(gmin", ct0)
c_ero = find(r"float2 nzw = dn", ct0)
STAGES = [
    ("light sample, relaxed arithmetic (lightSampleFast)", body(r"float lightSampleFast\(")),
    ("light samples: dealing (item, sample) pairs through shared memory", body(r"float warpSharedLightSamples\(")),
    ("lit term: Beer / in-scatter / phase (litTerm)", body(r"float litTerm\(")),
    ("det_powf (binary64 pow of the coverage bias)", body(r"float det_powf\(")),
    ("cloudHiRes (curl + hi-res fetch, erosion, remap)", body(r"float cloudHiRes\(")),
    ("cloudTest: height gradients + all-zero early out", span(ct0, before(c_fetch))),
    ("cloudTest: low-res + placement fetch (shell projection, coordinates)", span(c_fetch, before(c_blend))),
    ("cloudTest: layer blend, density gate", span(c_blend, before(c_cov))),
    ("cloudTest: coverage bias (remap, pow call)", span(c_cov, before(c_ero))),
    ("cloudTest: erosion, exact early-out, two remaps", span(c_ero, ct1)),
    ("trip: position, shell projection, height, wind offset", span(trip0, t_call_test)),
    ("trip: hit / miss state machine (CC:426-437, 468-474)", span((t_call_test[0], t_call_test[1] + 1), before(t_ballot))),
    ("trip: lit-lane ballot, transmittance update", span(t_ballot, before(t_tail))),
    ("trip: termination tests, t += step (CC:476-482, 408)", span(t_tail, trip1)),
    ("ray set-up: sky colour (atmosphereColorPhysical)", body(r"v3 atmosphereColorPhysical\(")),
    ("ray set-up: shell intersections (raySphereT)", body(r"float raySphereT\(")),
    ("ray set-up: night background", body(r"v3 nightBackground\(")),
    ("ray set-up: ray, sun disk, phase function", body(r"void ray_setup\(")),
    ("ray finish: composite", body(r"float4 ray_finish\(")),
    ("pixel addressing / store", body(r"bool dispatch_pixel\(")),
    ("pixel addressing / store", body(r"void store_pixel\(")),
    ("kernel: prologue, loop vote", body(r"void .*cloud_march_kernel\(")),
]


def stage_of(chain):
    for f, line in chain:                                    # innermost frame first
        if f not in SOURCES:
            continue
        for name, (sf, a, b) in STAGES:
            if sf == f and a <= line <= b:
                return name
    return "other (libm slow paths, CUDA headers)"


tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
chains = {}
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or chains:
        continue
    out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur, pending, fresh = False, [], [], True
    for l in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            infn = kpat in m.group(1) and "fma" not in m.group(1) and not chains
            cur, pending = [], []
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
        if m:
            if pending:
                cur, pending = pending, []
            chains[int(m.group(1), 16)] = cur
if not chains:
    sys.exit("kernel pattern not found in " + lib)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(rows[hi])}
data = [r for r in rows[hi + 1:] if len(r) >= len(rows[hi])]
base = int(data[0][ix["Address"]], 16)
agg = collections.OrderedDict((name, [0, 0, 0]) for name, _ in STAGES)
agg["other (libm slow paths, CUDA headers)"] = [0, 0, 0]
tot = [0, 0, 0]
votes = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    n, t, s = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    st = stage_of(chains.get(off, []))
    for a in (agg[st], tot):
        a[0] += n; a[1] += t; a[2] += s
    if "VOTE.ANY" in r[ix["Source"]] and st.startswith("kernel"):
        votes = max(votes, n)                                # the loop-closing vote executes once per warp-trip
print(f"{kpat}: {tot[0]:,} warp-instructions, {tot[1]/tot[0]:.1f} threads per instruction, {tot[2]:,} stall samples, {votes:,} warp-trips")
print(f"{'stage':76s} {'inst %':>7s} {'stall %':>8s} {'thr/inst':>8s} {'inst/warp-trip':>15s}")
for name, (n, t, s) in agg.items():
    if n:
        print(f"{name:76s} {100*n/tot[0]:7.2f} {100*s/max(tot[2],1):8.2f} {t/n:8.1f} {n/max(votes,1):15.1f}")
print(f"{'total':76s} {100.0:7.2f} {100.0:8.2f} {tot[1]/tot[0]:8.1f} {tot[0]/max(votes,1):15.1f}")
