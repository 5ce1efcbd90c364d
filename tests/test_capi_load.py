"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/marshmallow.h declares, and
fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(mm):
    lib = mm.load_library()
    names = mm.exported_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/marshmallow.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", mm.library_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert not [s for s in exported if s.startswith("om_")], "the product library must not contain oracle symbols"


def test_product_does_not_link_or_import_the_oracle(mm):
    ldd = subprocess.run(["ldd", mm.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
    import _pkg
    for root, _, files in os.walk(_pkg.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_binding" not in text, f"{f} references the oracle"


def test_product_has_no_host_build_of_the_kernels(mm):
    """The kernel source is written so that tests/host_build can compile it for the host (MM_HOST_BUILD) and hold it to the oracle; the product must never
    do that: no build script or binding defines the macro, and the library carries neither the test harnesses' entry points nor host copies of the per-ray functions."""
    import _pkg
    for root, _, files in os.walk(_pkg.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp")) or f in ("Makefile",):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "define MM_HOST_BUILD" not in text and "-DMM_HOST_BUILD" not in text, f"{f} turns the host build on"
    out = subprocess.run(["nm", "-C", mm.library_path()], capture_output=True, text=True).stdout
    assert not [l for l in out.splitlines() if " hb_" in l or " hm_" in l]
    assert "cloudTest" not in out and "reproject_texel" not in out and "god_ray_alpha" not in out, "host copies of device functions in the product library"


def test_header_is_plain_c(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "marshmallow.h"\nint main(void){ return sizeof(mm_ctx*) > 0 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_no_gpu_means_loud_failure_not_fallback(mm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mm.MarshmallowError) as e:
        mm.ComputeShader(0, (64, 36))
    assert "no CPU fallback" in str(e.value)


def test_null_arguments_are_rejected(mm):
    lib = mm.load_library()
    assert lib.mm_create(0, None) == -1
    assert lib.mm_destroy(None) == -1
    assert lib.mm_set_uniforms(None, None, None, None, None) == -1
    assert lib.mm_dispatch(None, 0, 0, 1, 1, None) == -1
    assert lib.mm_host_sky(0.25, 0.25, 10.0, 2.0, 0.005, 0.8, None, 0.0, 0, None, None) == -1
    assert lib.mm_version().startswith(b"marshmallow-b200")


def test_host_mirror_uniform_block_sizes(mm):
    sun, sky = mm.host_sky(0.25, 0.25)
    cam = mm.host_camera((0, 1, 1), 0.0, 0.0)
    assert cam.nbytes == 160 and sun.nbytes == 116 and sky.nbytes == 52     # Shader.h:24-29, SkyManager.h:8-36
    sm = mm.SkyManager()
    sm.rebuildSkyFromNewSun(0.25, 0.25)
    sm.setTime(3.0)
    assert sm.getSky()[11] == 3.0 and sm.getSun().nbytes == 116
