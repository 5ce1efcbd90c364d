"""The reference's OWN shader text, executed on the CPU, against the oracle's restatement of it.

oracle/Makefile rewrites /root/reference/.../Shaders/compute-clouds.comp lexically into C++ (oracle/glsl_to_cpp.py: float suffixes,
`.xyz` -> `.xyz()`, parameter qualifiers, uniform blocks -> structs; no expression is touched) and compiles it inside the GLSL
environment of oracle/glsl_env.h into oracle/_ref/libref_cc.so; the shader text never enters the repository.  Control flow,
constants, argument orders (incl. the swapped heightBiasCoverage arguments), operator order and every quirk are then the
reference's.  The environment supplies the language (vector types and built-ins with the contract's definitions) and routes
texture() to the oracle's sampler.  Bar: every channel of every pixel bit-identical, total fetch counts identical.

Needs oracle/_ref/libref_cc.so (built wherever /root/reference is present; the built file travels with the tree).
"""
import ctypes as C
import os

import numpy as np
import pytest

import scenes

REF_CC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_cc.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_CC), reason="oracle/_ref/libref_cc.so not built (reference tree absent)")


class _SamplerCtx(C.Structure):
    _fields_ = [("scene", C.c_void_p), ("filter", C.c_int)]


REF_CC_FMA = os.path.join(os.path.dirname(REF_CC), "libref_cc_fma.so")


def _run_reference_shader(oracle, S, filt, sc, ids, arith="ieee"):
    """arith "fma": the same shader text compiled in glsl_env_fma.h, whose operators apply the lexical contraction rule"""
    lib = C.CDLL(REF_CC_FMA if arith == "fma" else REF_CC)
    ref_cc_run = lib.ref_cc_fma_run if arith == "fma" else lib.ref_cc_run
    ref_cc_run.argtypes = [C.c_void_p] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.full((1080, 1920, 4), -7.0, np.float32)
    written = np.zeros((1080, 1920), np.uint8)
    fetches = (C.c_ulonglong * 2)()
    ctx = _SamplerCtx(S.s, filt)
    cam, sun, sky = (np.ascontiguousarray(sc[k], np.float32) for k in ("cam", "sun", "sky"))
    ids = np.ascontiguousarray(ids, np.uint32)
    rc = ref_cc_run(oracle._p(cam), oracle._p(sun), oracle._p(sky), C.cast(oracle.lib().om_sample_callback, C.c_void_p), C.byref(ctx),
                        oracle._p(ids), len(ids), oracle._p(out), oracle._p(written), fetches)
    assert rc == 0
    return out, written.astype(bool), (int(fetches[0]), int(fetches[1]))


CASES = [("C1", 5, "fp32", {}), ("C3", 0, "fp32", {}), ("C5b", 15, "fp32", {}), ("C2b", 9, "texunit", {}),
         ("C1", 3, "fp32", dict(time=123.5, wind=(0.7, 0.05, -1.3))), ("C1", 6, "fp32", dict(elevation=0.75)),
         ("C1", 11, "texunit", {}), ("C1", 2, "texunit", dict(elevation=0.9, yaw=0.7, pitch=-0.6))]


@pytest.mark.parametrize("arith", ["ieee", "fma"])
@pytest.mark.parametrize("name,phase,sampler,over", CASES)
def test_oracle_equals_the_reference_shader_text(mm, oracle, assets, name, phase, sampler, over, arith):
    """Both arithmetic definitions: one rounding per operator (cloud_march_oracle.c vs glsl_env.h) and the lexical fused-multiply-add
    rule (cloud_march_oracle_fma.c, restated by hand, vs glsl_env_fma.h, which applies it mechanically to the shader text)."""
    W, H = 1920, 1080                                  # the shader hard-codes its extent (CC:283-285)
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, pixel_phase=phase, **over)
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    filt = oracle.OM_FILTER_TEXUNIT if sampler == "texunit" else oracle.OM_FILTER_FP32
    # libm pow on both sides: the shader's pow() is the language's; the oracle's deterministic pow is pinned separately
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=filt, pow_mode=oracle.OM_POW_LIBM, nightsky=night,
                     arith=oracle.OM_ARITH_FMA if arith == "fma" else oracle.OM_ARITH_IEEE)
    # one reference dispatch (phase `phase`), ten 4-row bands of the frame: 4800 rays from top to below the horizon
    want, wcnt = S.march(W, H, mode=oracle.OM_PHASE16, row_begin=0, row_stride=27, row_block=4, out=np.full((H, W, 4), -7.0, np.float32))
    ys, xs = np.nonzero((want != -7.0).any(axis=-1))
    assert len(ys) == 4800 and (xs % 4 == phase % 4).all() and (ys % 4 == phase // 4).all()
    got, written, fetches = _run_reference_shader(oracle, S, filt, sc, np.stack([xs // 4, ys // 4], 1), arith)
    assert np.array_equal(written, (want != -7.0).any(axis=-1)), "the shader wrote a different set of pixels"
    bad = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
    assert not bad.any(), (int(bad.sum()), got[bad][:3], want[bad][:3])
    assert fetches == (int(wcnt[..., 1].sum()), int(wcnt[..., 2].sum())), "texture() call counts differ"
    assert wcnt[..., 0].sum() > 100000                 # the comparison marched clouds, not only sky


def test_the_deterministic_pow_only_moves_last_bits(mm, oracle, assets):
    """The one place the oracle departs from `the language's pow` on purpose: heightBiasCoverage uses det_powf so that a GPU can
    reproduce it bit for bit.  Against the reference shader text run with libm's powf the frames differ only where that pow's last
    bit flips a threshold: a small fraction of pixels."""
    W, H = 1920, 1080
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H, pixel_phase=5)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], pow_mode=oracle.OM_POW_DET)
    want, _ = S.march(W, H, mode=oracle.OM_PHASE16, row_begin=0, row_stride=27, row_block=4, out=np.full((H, W, 4), -7.0, np.float32))
    ys, xs = np.nonzero((want != -7.0).any(axis=-1))
    got, _, _ = _run_reference_shader(oracle, S, oracle.OM_FILTER_FP32, sc, np.stack([xs // 4, ys // 4], 1))
    rep = oracle.parity_report(want[ys, xs][None], got[ys, xs][None])
    assert rep["frac_within_1"] > 0.995 and rep["alpha_identical_frac"] > 0.995, rep


# ---------------------------------------------------------------------------------------------------------------------------
# The other restated shaders -- reproject.comp, god-ray.frag, radialBlur.frag, tonemap.frag and the shadow march inside model.frag
# -- compiled the same way into oracle/_ref/libref_passes.so.
REF_PASSES = os.path.join(os.path.dirname(REF_CC), "libref_passes.so")


def _passes():
    lib = C.CDLL(REF_PASSES)
    vp, i32 = C.c_void_p, C.c_int
    lib.ref_reproject.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.ref_god_ray.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.ref_radial_blur.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.ref_tonemap.argtypes = [vp, i32, i32, vp]
    lib.ref_cloud_shadow.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp]
    return lib


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.skipif(not os.path.exists(REF_PASSES), reason="oracle/_ref/libref_passes.so not built (reference tree absent)")
@pytest.mark.parametrize("W,H", [(160, 90), (97, 61)])
def test_reproject_oracle_equals_the_reference_shader_text(mm, oracle, W, H):
    rng = np.random.default_rng(W)
    src = (rng.random((H, W, 4)) * 20).astype(np.float32)
    lib = _passes()
    cams = [(mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD), mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD)),
            (mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD), mm.host_camera((3.0, 1.0, 2.0), -np.pi / 2 + 0.01, -19 * scenes.DEG2RAD)),
            (mm.host_camera((5, 2, -3), 0.4, -0.6), mm.host_camera((4, 2, -3.5), 0.47, -0.55))]
    for cam, prev in cams:
        want = oracle.reproject(cam, prev, src)
        got = np.empty_like(src)
        assert lib.ref_reproject(oracle._p(cam), oracle._p(prev), oracle._p(src), W, H, oracle._p(got)) == 0
        assert _same_bits(want, got)


@pytest.mark.skipif(not os.path.exists(REF_PASSES), reason="oracle/_ref/libref_passes.so not built (reference tree absent)")
@pytest.mark.parametrize("name,over", [("C1", {}), ("C3", {}), ("C1", dict(yaw=-1.2, pitch=-0.5)), ("C1", dict(elevation=0.75))])
def test_post_chain_oracle_equals_the_reference_shader_texts(mm, oracle, assets, name, over):
    W, H = 160, 90
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    img, _ = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], nightsky=night).march(W, H, counters=False)
    cam, sun = np.ascontiguousarray(sc["cam"], np.float32), np.ascontiguousarray(sc["sun"], np.float32)
    lib = _passes()
    fb1, fb2, fb3 = np.empty_like(img), np.empty_like(img), np.empty_like(img)
    want1 = oracle.god_ray(cam, sun, img)
    assert lib.ref_god_ray(oracle._p(cam), oracle._p(sun), oracle._p(img), W, H, oracle._p(fb1)) == 0
    assert _same_bits(want1, fb1), "god-ray.frag"
    want2 = oracle.radial_blur(cam, sun, want1)
    assert lib.ref_radial_blur(oracle._p(cam), oracle._p(sun), oracle._p(want1), W, H, oracle._p(fb2)) == 0
    assert _same_bits(want2, fb2), "radialBlur.frag"
    want3 = oracle.tonemap_present(want2)
    assert lib.ref_tonemap(oracle._p(want2), W, H, oracle._p(fb3)) == 0
    quant = np.floor(255.0 * np.clip(np.nan_to_num(fb3, nan=0.0), 0.0, 1.0) + 0.5).astype(np.uint8)     # the swapchain's UNORM8 store
    assert np.array_equal(quant, want3), "tonemap.frag"


@pytest.mark.skipif(not os.path.exists(REF_PASSES), reason="oracle/_ref/libref_passes.so not built (reference tree absent)")
@pytest.mark.parametrize("name,sampler,over", [("C1", "fp32", {}), ("C5", "texunit", {}), ("C3", "fp32", {}),
                                               ("C1", "texunit", dict(time=60.0, wind=(0.7, 0.05, -1.3))), ("C1", "fp32", dict(pitch=0.3, yaw=0.4))])
def test_cloud_shadow_oracle_equals_model_frag(mm, oracle, assets, name, sampler, over):
    """model.frag's main() is run whole for every position; its accumDensity (which the shader only folds into the fragment colour) is
    read back through a reference the rewrite puts in place of the variable's declaration (glsl_to_cpp.py --probe)."""
    sc = scenes.make_scene(mm, name, assets, **over)
    filt = oracle.OM_FILTER_TEXUNIT if sampler == "texunit" else oracle.OM_FILTER_FP32
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=filt, pow_mode=oracle.OM_POW_LIBM)
    rng = np.random.default_rng(17)
    pos = np.concatenate([rng.uniform(-50, 50, (4000, 3)), rng.uniform(-20000, 20000, (4000, 3)) * np.float32([1, 0.02, 1])]).astype(np.float32)
    want, nf = S.cloud_shadow(pos, want_fetches=True)
    got = np.empty(len(pos), np.float32)
    fetches = (C.c_ulonglong * 2)()
    ctx = _SamplerCtx(S.s, filt)
    cam, sun, sky = (np.ascontiguousarray(sc[k], np.float32) for k in ("cam", "sun", "sky"))
    assert _passes().ref_cloud_shadow(oracle._p(cam), oracle._p(sun), oracle._p(sky), C.cast(oracle.lib().om_sample_callback, C.c_void_p),
                                      C.byref(ctx), oracle._p(pos), len(pos), oracle._p(got), fetches) == 0
    assert _same_bits(want, got)
    assert int(fetches[0]) + int(fetches[1]) == int(nf.sum())
    assert (want > 0).mean() > 0.03
