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


def _run_reference_shader(oracle, S, filt, sc, ids):
    ref = C.CDLL(REF_CC)
    ref.ref_cc_run.argtypes = [C.c_void_p] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.full((1080, 1920, 4), -7.0, np.float32)
    written = np.zeros((1080, 1920), np.uint8)
    fetches = (C.c_ulonglong * 2)()
    ctx = _SamplerCtx(S.s, filt)
    cam, sun, sky = (np.ascontiguousarray(sc[k], np.float32) for k in ("cam", "sun", "sky"))
    ids = np.ascontiguousarray(ids, np.uint32)
    rc = ref.ref_cc_run(oracle._p(cam), oracle._p(sun), oracle._p(sky), C.cast(oracle.lib().om_sample_callback, C.c_void_p), C.byref(ctx),
                        oracle._p(ids), len(ids), oracle._p(out), oracle._p(written), fetches)
    assert rc == 0
    return out, written.astype(bool), (int(fetches[0]), int(fetches[1]))


CASES = [("C1", 5, "fp32", {}), ("C3", 0, "fp32", {}), ("C5b", 15, "fp32", {}), ("C2b", 9, "texunit", {}),
         ("C1", 3, "fp32", dict(time=123.5, wind=(0.7, 0.05, -1.3))), ("C1", 6, "fp32", dict(elevation=0.75)),
         ("C1", 11, "texunit", {}), ("C1", 2, "texunit", dict(elevation=0.9, yaw=0.7, pitch=-0.6))]


@pytest.mark.parametrize("name,phase,sampler,over", CASES)
def test_oracle_equals_the_reference_shader_text(mm, oracle, assets, name, phase, sampler, over):
    W, H = 1920, 1080                                  # the shader hard-codes its extent (CC:283-285)
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, pixel_phase=phase, **over)
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    filt = oracle.OM_FILTER_TEXUNIT if sampler == "texunit" else oracle.OM_FILTER_FP32
    # libm pow on both sides: the shader's pow() is the language's; the oracle's deterministic pow is pinned separately
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=filt, pow_mode=oracle.OM_POW_LIBM, nightsky=night)
    # one reference dispatch (phase `phase`), ten 4-row bands of the frame: 4800 rays from top to below the horizon
    want, wcnt = S.march(W, H, mode=oracle.OM_PHASE16, row_begin=0, row_stride=27, row_block=4, out=np.full((H, W, 4), -7.0, np.float32))
    ys, xs = np.nonzero((want != -7.0).any(axis=-1))
    assert len(ys) == 4800 and (xs % 4 == phase % 4).all() and (ys % 4 == phase // 4).all()
    got, written, fetches = _run_reference_shader(oracle, S, filt, sc, np.stack([xs // 4, ys // 4], 1))
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
