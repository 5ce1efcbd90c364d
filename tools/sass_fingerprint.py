#!/usr/bin/env python
"""Fingerprint of the compiled kernels: one md5 per kernel over its SASS instruction stream (addresses and encodings stripped).

    python tools/sass_fingerprint.py                 # print the fingerprint of the current in-tree build
    python tools/sass_fingerprint.py --write FILE    # record it
    python tools/sass_fingerprint.py --check FILE    # compare the current build with a recorded one

profiles/r02pow_sass_fingerprint.json is the build that was validated on the B200 in the round's last GPU calls (profiles/r02pow_pytest.log: 180 GPU tests,
r02pow_bench_C3_n1.json, r02pow_sanitizer_*.log).  Refactorings made after the GPU budget was spent (per-ray / per-pixel source moved into headers that the CPU
suite also compiles for the host) were accepted only with `--check` green: the shipped kernels are byte for byte the validated ones."""
import hashlib, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.environ.get("MM_BUILD_DIR") or os.path.join(ROOT, "project-marshmallow_b200", "build")     # MM_BUILD_DIR: the objects of another checkout


def fingerprint():
    out = {}
    for obj in sorted(f for f in os.listdir(BUILD) if f.endswith(".cu.o")):
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
        name, h, n = None, None, 0
        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if name:
                    out[name] = {"md5": h.hexdigest(), "instructions": n}
                demangled = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
                name, h, n = obj[:-5] + ": " + demangled, hashlib.md5(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
            if m and name:
                h.update(m.group(1).encode()); n += 1
        if name:
            out[name] = {"md5": h.hexdigest(), "instructions": n}
    return out


if __name__ == "__main__":
    fp = fingerprint()
    if len(sys.argv) == 3 and sys.argv[1] == "--write":
        json.dump(fp, open(sys.argv[2], "w"), indent=1, sort_keys=True)
        print(f"{len(fp)} kernels recorded in {sys.argv[2]}")
    elif len(sys.argv) == 3 and sys.argv[1] == "--check":
        ref = json.load(open(sys.argv[2]))
        bad = [k for k in sorted(set(ref) | set(fp)) if ref.get(k) != fp.get(k)]
        for k in bad:
            print("DIFFERS:", k, ref.get(k), fp.get(k))
        print(f"{len(fp) - len(bad)} of {len(fp)} kernels identical to {sys.argv[2]}")
        sys.exit(1 if bad else 0)
    else:
        for k, v in fp.items():
            print(v["md5"], v["instructions"], k)
