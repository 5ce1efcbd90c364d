"""Build-quality guards that need no GPU: the sm_100a code in the product library is inspected with cuobjdump.

The march is instruction-issue bound and its occupancy is set by registers per thread: the texture-unit variants of
K1 / K1s / K1p must fit 8 blocks of 128 threads per SM (<= 64 registers), the FP32-sampler variants 7 (<= 72), nothing may spill
(LOCAL 0; the small STACK belongs to the out-of-line libm slow paths of ray set-up), the default variant must sample with the texture
unit and the FP32-sampler variant must not, and the library must carry sm_100a code only."""
import re
import shutil
import subprocess


CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _usage(mm):
    out = subprocess.run([CUOBJDUMP, "-res-usage", mm.library_path()], capture_output=True, text=True, check=True).stdout
    fns = {}
    for name, res in re.findall(r"Function (\S+):\n\s+(REG:.*)", out):
        fns[name] = {k: int(v) for k, v in re.findall(r"([A-Z]+)(?:\[0\])?:(\d+)", res)}
    return out, fns


def _variants(fns, kernel, *flags):
    """mangled names of `kernel<flags..., ...>` instantiations (bool template arguments appear as Lb0E / Lb1E)"""
    pat = re.compile(r"\d+" + kernel + "I" + "".join(f"Lb{int(f)}E" for f in flags))
    return {n: r for n, r in fns.items() if pat.search(n)}


def test_library_carries_sm100a_code_only(mm):
    out = subprocess.run([CUOBJDUMP, "-lelf", mm.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_march_kernels_fit_their_register_budget_and_do_not_spill(mm):
    _, fns = _usage(mm)
    assert len(fns) > 100                                            # both arithmetic builds of every variant
    for kernel, hw_budget, fp32_budget in (("cloud_march_kernel", 64, 72), ("cloud_march_split_kernel", 64, 80), ("cloud_march_persistent_kernel", 64, 72)):
        hw = _variants(fns, kernel, 1, 1)                            # <MARCH_HW, LIGHT_HW, ...>: the default sampler mode
        fp32 = _variants(fns, kernel, 0)                             # the march filters in FP32 (HYBRID, EXACT)
        assert hw and fp32, kernel
        for n, r in hw.items():
            assert r["REG"] <= hw_budget, (n, r)
        for n, r in fp32.items():
            assert r["REG"] <= fp32_budget, (n, r)
        for n, r in {**hw, **fp32}.items():
            assert r["LOCAL"] == 0 and r["SHARED"] <= 8 * 1024, (n, r)   # 8 blocks per SM must not be limited by shared memory
            production = re.search(kernel + r"ILb[01]ELb[01]ELb0E", n)   # <.., .., CNT = false, ..>: the variants without diagnostic counters
            assert r["STACK"] <= (48 if production else 96), (n, r)      # call frames of the cold out-of-line functions; the counter variants spill a little


def test_default_variant_samples_with_the_texture_unit(mm):
    _, fns = _usage(mm)
    default = [n for n in _variants(fns, "cloud_march_kernel", 1, 1, 0, 1) if "fma" not in n]
    exact = [n for n in _variants(fns, "cloud_march_kernel", 0, 0, 0, 1) if "fma" not in n]
    assert len(default) == 1 and len(exact) == 1, (default, exact)

    def sass(fn):
        return subprocess.run([CUOBJDUMP, "-sass", "-fun", fn, mm.library_path()], capture_output=True, text=True, check=True).stdout

    d, e = sass(default[0]), sass(exact[0])
    assert len(re.findall(r"\bTEX\b", d)) >= 8                       # march: placement, low-res, curl, hi-res; the same four in the light samples
    assert not re.findall(r"\bTEX\b", e)                             # FILTER_EXACT never touches the texture unit ...
    assert "LDG.E.128.CONSTANT" in e and re.search(r"\bFFMA2\b", e)  # ... it loads pair-major footprints and lerps on packed FP32
    # the uncontracted contract: the decision path is FMUL + FADD, so they outnumber FFMA in the default build
    assert len(re.findall(r"\bFMUL\b", d)) > len(re.findall(r"\bFFMA\b", d))
    # no local-memory traffic inside the march loop (the region between the loop's first and last warp vote)
    lines = d.splitlines()
    votes = [i for i, l in enumerate(lines) if "VOTE.ANY" in l]
    assert len(votes) >= 2
    loop = "\n".join(lines[votes[0]:votes[-1] + 1])
    assert "STL" not in loop and "LDL" not in loop
