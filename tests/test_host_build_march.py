"""The march kernel's own per-ray SOURCE, checked on the CPU.  csrc/cloud_march_ray.inl -- ray set-up, cloudTest with its exact early-outs and the pow filter,
cloudHiRes, the FP32 sampler, the relaxed light sample, litTerm, the composite -- is compiled for the host (tests/host_build/march_host.cu, -DMM_HOST_BUILD; the
kernels' SASS is unaffected) and driven ray by ray; texture-unit fetches go to the oracle's bit-exact model of the unit.  Gates as in tests/test_march_parity_gpu.py:
the decision path bit for bit (loop trips, fetch and lit counters, the alpha channel), colours within the parity gate -- for the three sampler modes and both
arithmetic definitions.  Test infrastructure only: the product never runs on the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_build", "march_host.cu")
CSRC = os.path.join(os.path.dirname(HERE), "project-marshmallow_b200", "csrc")
FILTER_EXACT, FILTER_HW, FILTER_HYBRID = 0, 1, 2


class _SamplerCtx(C.Structure):
    _fields_ = [("scene", C.c_void_p), ("filter", C.c_int)]


def _build(fma, extra=(), tag=""):
    lib = os.path.join(HERE, "host_build", ("libmarch_host_fma" if fma else "libmarch_host") + tag + ".so")
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("cloud_march_ray.inl", "common.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                        "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-diag-suppress", "177", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                        "--cudart", "static", f"-DMM_FMA={int(fma)}", *extra, "-o", lib, SRC], check=True)
    l = C.CDLL(lib)
    l.hm_march.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p] * 4
    l.hm_det_powf.restype, l.hm_det_powf.argtypes = C.c_float, [C.c_float, C.c_float]
    l.hm_cloud_shadow.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    assert l.hm_arith() == int(fma)
    return l


@pytest.fixture(scope="module")
def host_march():
    return {False: _build(False), True: _build(True)}


def _host_frame(lib, sc, W, H, filt, counters, S, om_filter, nightsky=None):
    tex = [sc["textures"]["placement"], nightsky, sc["textures"]["curl"], sc["textures"]["lowres"], sc["textures"]["hires"]]
    tex = [None if t is None else np.ascontiguousarray(t, np.uint8) for t in tex]
    ptrs = (C.c_void_p * 5)(*[None if t is None else t.ctypes.data for t in tex])
    dims = np.zeros((5, 3), np.int32)
    for i, t in enumerate(tex):
        if t is not None:
            dims[i] = (t.shape[1], t.shape[0], 1) if t.ndim == 3 else (t.shape[2], t.shape[1], t.shape[0])
    cam, sun, sky = (np.ascontiguousarray(x, np.float32) for x in (sc["cam"], sc["sun"], sc["sky"]))
    out = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros((H, W, 4), np.uint32)
    ctx = _SamplerCtx(S.s, om_filter)
    rc = lib.hm_march(ob._p(cam), ob._p(sun), ob._p(sky), C.cast(ptrs, C.c_void_p), ob._p(dims), filt, int(counters), W, H,
                      C.cast(ob.lib().om_sample_callback, C.c_void_p), C.cast(C.byref(ctx), C.c_void_p), ob._p(out), ob._p(cnt))
    assert rc == 0
    return out, cnt


CASES = [("C1", 96, 54, {}), ("C3", 96, 54, {}), ("C2b", 64, 36, {}), ("C5", 64, 36, {}), ("C1", 61, 35, {"wind": (1.0, 0.05, 1.0), "time": 37.5})]


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("filt", [FILTER_EXACT, FILTER_HW, FILTER_HYBRID])
@pytest.mark.parametrize("name,W,H,over", CASES)
def test_march_source_decisions_equal_the_oracle(host_march, mm, assets, name, W, H, over, filt, fma):
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    om_filter = ob.OM_FILTER_TEXUNIT if filt == FILTER_HW else ob.OM_FILTER_FP32     # the definition the MARCH samples with (test_march_parity_gpu.py)
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=om_filter, arith=ob.OM_ARITH_FMA if fma else ob.OM_ARITH_IEEE)
    want, want_cnt = S.march(W, H)
    # with counters: the light samples take the exact path too, so EVERY counter must match; HYBRID's light samples use the other sampler, so only FP32 / HW here
    if filt != FILTER_HYBRID:
        got, cnt = _host_frame(host_march[fma], sc, W, H, filt, True, S, om_filter)
        assert np.array_equal(cnt, want_cnt), "loop trips / fetches / lit steps differ from the oracle"
        assert np.array_equal(got[..., 3].view(np.uint32), want[..., 3].view(np.uint32)), "alpha differs"
        rep = ob.parity_report(want, got)
        assert rep["pass"] and rep["max_abs_diff_8bit"] <= 1, rep
    # without counters: the production variant (relaxed light samples in the texture-unit modes): decisions and alpha exact, colours inside the gate
    got, _ = _host_frame(host_march[fma], sc, W, H, filt, False, S, ob.OM_FILTER_TEXUNIT if filt != FILTER_EXACT else ob.OM_FILTER_FP32)
    assert np.array_equal(got[..., 3].view(np.uint32), want[..., 3].view(np.uint32)), "alpha (accumulated density) differs: a march decision changed"
    rep = ob.parity_report(want, got)
    assert rep["pass"] and rep["max_abs_diff_8bit"] <= 2, rep
    S.close()


def test_night_frame_source(host_march, mm, assets):
    W, H = 64, 36
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H, elevation=0.75)
    star = scenes.shipped_night_sky()
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ob.OM_FILTER_FP32, nightsky=star)
    want, want_cnt = S.march(W, H)
    got, cnt = _host_frame(host_march[False], sc, W, H, FILTER_EXACT, True, S, ob.OM_FILTER_FP32, nightsky=star)
    assert np.array_equal(cnt, want_cnt)
    rep = ob.parity_report(want, got)
    assert rep["pass"], rep
    S.close()


def test_det_powf_source_equals_the_oracle(host_march):
    rng = np.random.default_rng(5)
    l = ob.lib()
    for fma, ref in ((False, l.om_det_powf), (True, l.om_det_powf_fma)):
        for x, y in zip(rng.uniform(1e-6, 1.0, 4000).astype(np.float32), rng.uniform(0.8, 1.0, 4000).astype(np.float32)):
            a, b = host_march[fma].hm_det_powf(float(x), float(y)), ref(float(x), float(y))
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (fma, x, y, a, b)


@pytest.mark.parametrize("name,W,H", [("C5", 96, 54), ("C3", 96, 54), ("C1", 96, 54)])
def test_pow_filter_changes_no_bit(host_march, mm, assets, name, W, H):
    """The exact work elimination in front of det_powf (MM_POW_FILTER) at source level: the same build with and without it produces the
    same frame in every bit of every channel, and the same counters.  (C5 is the storm: coverage 0.9 everywhere, every march trip past the gate calls the pow.)"""
    plain = _build(False, extra=("-DMM_POW_FILTER=0",), tag="_nopowfilter")
    sc = scenes.make_scene(mm, name, assets, W=W, H=H)
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ob.OM_FILTER_TEXUNIT)
    for filt, om_filter in ((FILTER_EXACT, ob.OM_FILTER_FP32), (FILTER_HW, ob.OM_FILTER_TEXUNIT)):
        for counters in (True, False):
            a, ca = _host_frame(host_march[False], sc, W, H, filt, counters, S, om_filter)
            b, cb = _host_frame(plain, sc, W, H, filt, counters, S, om_filter)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb), (filt, counters)
    S.close()


@pytest.mark.parametrize("name", ["C3", "C2"])
def test_march_source_on_a_larger_frame(host_march, mm, assets, name):
    """83 k rays per scene and sampler: the rarer paths (rays grazing the horizon with 250 trips, the 10-miss rule, saturated accumulations)"""
    W, H = 384, 216
    sc = scenes.make_scene(mm, name, assets, W=W, H=H)
    for filt, om_filter in ((FILTER_HW, ob.OM_FILTER_TEXUNIT), (FILTER_EXACT, ob.OM_FILTER_FP32)):
        S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=om_filter)
        want, want_cnt = S.march(W, H)
        got, cnt = _host_frame(host_march[False], sc, W, H, filt, True, S, om_filter)
        assert want_cnt[..., 0].max() >= 200                 # the long rays are in the frame
        assert np.array_equal(cnt, want_cnt)
        assert np.array_equal(got[..., 3].view(np.uint32), want[..., 3].view(np.uint32))
        assert ob.parity_report(want, got)["pass"]
        S.close()


@pytest.mark.parametrize("name,over", [("C1", {}), ("C3", {}), ("C1", {"wind": (1.0, 0.05, 1.0), "time": 37.5}), ("C5", {})])
def test_cloud_shadow_source_is_bit_exact(host_march, mm, assets, name, over):
    """K7 (model.frag:240-283): the per-point source against om_cloud_shadow, densities and texture() counts bit for bit, both sampler definitions"""
    sc = scenes.make_scene(mm, name, assets, W=64, H=36, **over)
    rng = np.random.default_rng(11)
    pos = (rng.uniform(-1.0, 1.0, (20000, 3)) * np.array([30000.0, 400.0, 30000.0])).astype(np.float32)
    cam, sun, sky = (np.ascontiguousarray(x, np.float32) for x in (sc["cam"], sc["sun"], sc["sky"]))
    pl, lo = np.ascontiguousarray(sc["textures"]["placement"], np.uint8), np.ascontiguousarray(sc["textures"]["lowres"], np.uint8)
    for filt, om_filter in ((FILTER_EXACT, ob.OM_FILTER_FP32), (FILTER_HW, ob.OM_FILTER_TEXUNIT)):
        S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=om_filter)
        want, want_nf = S.cloud_shadow(pos, want_fetches=True)
        got, nf = np.empty(len(pos), np.float32), np.empty(len(pos), np.uint32)
        ctx = _SamplerCtx(S.s, om_filter)
        rc = host_march[False].hm_cloud_shadow(ob._p(cam), ob._p(sun), ob._p(sky), ob._p(pl), pl.shape[1], pl.shape[0], ob._p(lo), lo.shape[0], filt, ob._p(pos), len(pos),
                                               C.cast(ob.lib().om_sample_callback, C.c_void_p), C.cast(C.byref(ctx), C.c_void_p), ob._p(got), ob._p(nf))
        assert rc == 0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and np.array_equal(nf, want_nf)
        assert (want > 0).mean() > 0.02 or name == "C5"      # the positions do reach cloud
        S.close()
