"""Known-answer tests that pin the CPU oracle's helpers to compute-clouds.comp, derived by hand from the GLSL
(the reference has no tests of its own).  Expected values are computed here in numpy float32 from
the formulas as written at the cited lines, so a transcription slip in the C oracle shows up."""
import ctypes as C

import numpy as np
import pytest

f32 = np.float32


def test_remap_matches_glsl_formula(oracle):            # CC:65-71
    l = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(200):
        v, a, b, c, d = (f32(x) for x in rng.uniform(-2, 2, 5))
        want = c + (((v - a) / (b - a)) * (d - c))
        assert f32(l.om_remap(v, a, b, c, d)) == f32(want)
        wc = min(max(want, c), d) if c <= d else None
        if wc is not None:
            assert f32(l.om_remapClamped(v, a, b, c, d)) == f32(wc)


def test_remap_clamped_division_by_zero_quirk_q6(oracle):   # CC:69-71 with oldMin == oldMax (CC:227,248,250)
    l = oracle.lib()
    assert l.om_remapClamped(f32(0.5), f32(1.0), f32(1.0), f32(0.0), f32(1.0)) == 0.0      # -inf -> 0
    assert l.om_remapClamped(f32(1.5), f32(1.0), f32(1.0), f32(0.0), f32(1.0)) == 1.0      # +inf -> 1
    assert l.om_remapClamped(f32(1.0), f32(1.0), f32(1.0), f32(0.0), f32(1.0)) == 0.0      # 0/0 = NaN -> lo


def test_hg_phase(oracle):                               # CC:73-77
    l = oracle.lib()
    for c, g in ((1.0, 0.6), (0.0, 0.6), (-1.0, 0.89), (0.3, 0.8)):
        want = 0.07957747154594767 * ((1 - g * g) / (1 - 2 * g * c + g * g) ** 1.5)
        assert l.om_hgPhase(f32(c), f32(g)) == pytest.approx(want, rel=2e-6)
    assert l.om_hgPhase(f32(1.0), f32(0.6)) == pytest.approx(0.07957747 * 0.64 / 0.4 ** 3, rel=1e-5)


def test_ray_sphere_intersection_keeps_quirk_q1(oracle):  # CC:147-177
    """From (0,1,1) straight up through the inner shell (centre y=-995000, 'w' = diameter 2e6): the true distance
    is 4999, but the shader measures from the translated+scaled origin (~(0, 0.4975, 0)), giving ~4999.5."""
    l = oracle.lib()
    ro = np.array([0, 1, 1], f32); rd = np.array([0, 1, 0], f32)
    sph = np.array([0, -995000.0, 1, 2000000.0], f32)
    t = C.c_float()
    ok = l.om_raySphereIntersection(ro.ctypes.data, rd.ctypes.data, sph.ctypes.data, C.byref(t))
    assert ok == 1
    assert t.value == pytest.approx(4999.5025, abs=0.07)      # float32 granularity at 1e6 is 0.0625
    assert abs(t.value - 4999.0) > 0.3                        # i.e. NOT the true distance
    # a ray that misses (camera far outside, pointing away) is invalid and reports t = 0 (our definition)
    ro2 = np.array([0, 5e6, 0], f32)
    assert l.om_raySphereIntersection(ro2.ctypes.data, rd.ctypes.data, sph.ctypes.data, C.byref(t)) == 0 and t.value == 0.0


def test_cloud_layer_density(oracle):                     # CC:193-204
    l = oracle.lib()

    def ref(h, t):
        h = f32(min(max(h, 0), 1)); t = f32(t)
        rm = lambda v, a, b, c, d: f32(c) + (((f32(v) - f32(a)) / (f32(b) - f32(a))) * (f32(d) - f32(c)))
        cu = max(f32(0), rm(h, 0, .2, 0, 1) * rm(h, .7, .9, 1, 0))
        sc = max(f32(0), rm(h, 0, .2, 0, 1) * rm(h, .2, .7, 1, 0))
        st = max(f32(0), rm(h, 0, .1, 0, 1) * rm(h, .2, .3, 1, 0))
        mix = lambda x, y, a: x * (f32(1) - a) + y * a
        cl = lambda x: f32(min(max(x, f32(0)), f32(1)))
        return mix(mix(st, sc, cl(t * f32(2))), mix(sc, cu, cl((t - f32(.5)) * f32(2))), t)

    for h in (0.0, 0.05, 0.15, 0.25, 0.5, 0.69, 0.8, 0.9, 0.95, 1.0):
        for t in (0.0, 0.25, 0.5, 0.75, 1.0):
            assert f32(l.om_cloudLayerDensity(f32(h), f32(t))) == pytest.approx(float(ref(h, t)), rel=1e-6, abs=1e-7)
    # identically zero at and above h = 0.9 for every cloud type (used by the kernel's exact early-out)
    for h in np.linspace(0.9, 1.0, 50):
        for t in np.linspace(0, 1, 11):
            assert l.om_cloudLayerDensity(f32(h), f32(t)) == 0.0


def test_height_bias_coverage_and_det_pow(oracle):        # CC:206-208
    l = oracle.lib()
    # exponent: clamp(remap(height, .7, .8, 1, .8), .8, 1): 1 below .7, .8 above .8
    assert l.om_heightBiasCoverage(f32(0.37), f32(0.5)) == f32(0.37)          # pow(x, 1) == x exactly
    assert l.om_heightBiasCoverage(f32(0.37), f32(0.85)) == pytest.approx(0.37 ** 0.8, rel=1e-6)
    assert l.om_heightBiasCoverage(f32(0.0), f32(0.85)) == 0.0
    rng = np.random.default_rng(5)
    x = rng.random(20000).astype(f32); y = rng.uniform(0.8, 1.0, 20000).astype(f32)
    got = np.array([l.om_det_powf(a, b) for a, b in zip(x, y)], f32)
    exact = np.power(x.astype(np.float64), y.astype(np.float64)).astype(f32)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - exact.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp == 0).mean() > 0.99


def test_sampler_texel_centres_edges_and_wrap(oracle, assets, mm):
    """Sampler semantics (Texture.cpp:29-52, 315-338): texel centres at (i+0.5)/N return the texel, REPEAT wraps,
    half-way points average, coordinates are (u,v,w) = (x,y,z) with slice index = z."""
    import scenes
    sc = scenes.make_scene(mm, "C1", assets)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
    low = assets["lowres"]
    for (x, y, z) in ((0, 0, 0), (5, 17, 99), (127, 127, 127)):
        uvw = np.array([[(x + .5) / 128, (y + .5) / 128, (z + .5) / 128]], f32)
        got = S.sample(oracle.OM_TEX_LOWRES, oracle.OM_FILTER_FP32, uvw)[0]
        assert np.array_equal(got, (low[z, y, x].astype(f32) * f32(1.0 / 255.0)).astype(f32))
        wrapped = S.sample(oracle.OM_TEX_LOWRES, oracle.OM_FILTER_FP32, uvw + f32(3.0))[0]
        assert np.array_equal(got, wrapped)
    # midway between x = 127 and x = 0 (wrap) on slice 3, row 9
    mid = S.sample(oracle.OM_TEX_LOWRES, oracle.OM_FILTER_FP32, np.array([[1.0, 9.5 / 128, 3.5 / 128]], f32))[0]
    want = (low[3, 9, 127].astype(f32) + low[3, 9, 0].astype(f32)) * f32(0.5) * f32(1 / 255.0)
    assert np.allclose(mid, want, rtol=1e-6)
    pl = assets["placement"]
    got = S.sample(oracle.OM_TEX_PLACEMENT, oracle.OM_FILTER_FP32, np.array([[10.5 / 512, 20.5 / 512, 0]], f32))[0]
    assert np.array_equal(got, (pl[20, 10].astype(f32) * f32(1 / 255.0)).astype(f32))


def test_tonemap_map(oracle):                            # tonemap.frag:11-28
    img = np.array([[[0, 0.5, 50.2, 0.3], [1e-3, 5.0, 700.0, 1.7]]], f32)
    out = oracle.tonemap_rgba8(img)

    def uc2(x):
        return ((x * (0.15 * x + 0.05) + 0.004) / (x * (0.15 * x + 0.5) + 0.06)) - 0.02 / 0.3

    def tm(x):
        return int(np.floor(255 * min(max((uc2(0.7 * x) / uc2(50.2)) ** (1 / 2.2), 0), 1) + 0.5))

    assert out[0, 0, 0] == 0 and out[0, 0, 1] == tm(0.5) and out[0, 0, 3] == 77      # 255 * 0.3f = 76.50000113 -> round half up
    assert out[0, 1, 1] == tm(5.0) and out[0, 1, 2] == 255 and out[0, 1, 3] == 255


def test_march_statistics_match_survey_probe(oracle, assets, mm):
    """Workload pins from SURVEY 8d (C1-like view): ~60 loop trips and ~3-4 lit steps per pixel, max trips > 200
    (quirk Q3: the two `continue`s skip the step counter), rays below the horizon do no march at all."""
    import scenes
    sc = scenes.make_scene(mm, "C1", assets)
    img, cnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(320, 180)
    assert 50 < cnt[..., 0].mean() < 75 and cnt[..., 0].max() > 200
    assert 2 < cnt[..., 3].mean() < 5
    assert np.isfinite(img).all() and img[..., 3].min() >= 0 and img[..., 3].max() <= 1
    assert (cnt[..., 1] == cnt[..., 2]).all()              # every cloudTest / cloudHiRes does one 2D + one 3D fetch
    sc2 = scenes.make_scene(mm, "C1", assets, pitch=+30 * scenes.DEG2RAD)     # looking 30 degrees DOWN: all rays killed (CC:351)
    img2, cnt2 = oracle.Scene(sc2["textures"], sc2["cam"], sc2["sun"], sc2["sky"]).march(64, 36)
    assert cnt2.sum() == 0 and (img2[..., :3] > 0).all()


def test_windowed_replay_model_equals_the_plain_loop(mm, oracle, assets):
    """The construction behind the kernel's ray-split mode (G consecutive trips evaluated up front, the loop body replayed over
    them, the window closed by the first event) as a CPU model inside the oracle: bit-identical frames and counters for every G."""
    import scenes
    lib = oracle.lib()
    for name, filt in (("C1", oracle.OM_FILTER_TEXUNIT), ("C5b", oracle.OM_FILTER_FP32), ("C3", oracle.OM_FILTER_TEXUNIT)):
        sc = scenes.make_scene(mm, name, assets, W=96, H=54)
        S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=filt)
        lib.om_set_window(1)
        ref, rcnt = S.march(96, 54)
        try:
            for g in (2, 4, 8, 11, 32):
                lib.om_set_window(g)
                img, cnt = S.march(96, 54)
                assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)) and np.array_equal(cnt, rcnt), (name, g)
        finally:
            lib.om_set_window(1)
