"""GPU parity tests proper: the CUDA march (through the C-ABI) against the CPU oracle on the same inputs.

Bars (north star / SURVEY 8d): after the reference tonemap to RGBA8, max |diff| <= 2 per channel and
>= 99.9 % of pixels within 1; for FILTER_EXACT additionally ZERO branch flips (per-pixel loop-trip,
fetch and lit-step counts identical) and bit-identical alpha (alpha depends only on the decision path).
"""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

_FILTERS = {"exact": ("MM_FILTER_EXACT", "OM_FILTER_FP32"), "hybrid": ("MM_FILTER_HYBRID", "OM_FILTER_FP32"), "hw": ("MM_FILTER_HW", "OM_FILTER_TEXUNIT")}


def _render(mm, sc, filter_mode, mode=0, rows=(0, 1, 1), counters=True, trips=0):
    cs = mm.ComputeShader(0, (sc["W"], sc["H"]), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    cs.enableCounters(counters)
    cs.setFilterMode(filter_mode)
    cs.setLanesPerRay(trips)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    cs.dispatch(mode, *rows)
    cs.synchronize()
    img = cs.readOutput()
    cnt = cs.readCounters() if counters else None
    cs.close()
    return img, cnt


@pytest.mark.parametrize("name,W,H", [("C1", 320, 180), ("C3", 320, 180), ("C2b", 256, 144), ("C5", 256, 144), ("C5b", 256, 144)])
def test_exact_mode_matches_oracle(mm, oracle, assets, name, W, H):
    sc = scenes.make_scene(mm, name, assets, W=W, H=H)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
    ref, rcnt = S.march(W, H)
    img, cnt = _render(mm, sc, mm.MM_FILTER_EXACT)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    print(name, rep)
    assert rep["branch_flip_pixels"] == 0
    assert rep["counter_mismatch_pixels"] == 0
    assert rep["alpha_identical_frac"] == 1.0
    assert rep["max_abs_diff_8bit"] <= 1
    assert rep["frac_within_1"] == 1.0
    # raw HDR floats: only shading transcendentals (exp/pow) may differ, by a few ulp
    rel = np.abs(ref - img) / np.maximum(np.abs(ref), 1e-6)
    assert rel.max() < 1e-4


def test_exact_mode_with_wind_and_time(mm, oracle, assets):
    sc = scenes.make_scene(mm, "C1", assets, W=192, H=108, time=123.5, wind=(0.7, 0.05, -1.3))
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(192, 108)
    img, cnt = _render(mm, sc, mm.MM_FILTER_EXACT)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    assert rep["counter_mismatch_pixels"] == 0 and rep["max_abs_diff_8bit"] <= 1, rep


def test_hybrid_mode_decisions_match_oracle(mm, oracle, assets):
    """HYBRID filters the light-cone samples in hardware: decisions (trip counts, alpha) stay exact."""
    sc = scenes.make_scene(mm, "C1", assets, W=320, H=180)
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(320, 180)
    img, cnt = _render(mm, sc, mm.MM_FILTER_HYBRID)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    print("hybrid", rep)
    assert rep["branch_flip_pixels"] == 0
    assert rep["alpha_identical_frac"] == 1.0
    assert rep["pass"], rep


def test_hw_sampler_is_the_texunit_model(mm, oracle, assets):
    """The live texture unit against the oracle's integer model of it (OM_FILTER_TEXUNIT): bit-identical on fresh
    random coordinates for the four march textures and for a non-power-of-two 2D texture (star-map-like)."""
    rng = np.random.default_rng(41)
    star = rng.integers(0, 256, (27, 50, 4), dtype=np.uint8)
    tex = {"placement": assets["placement"], "curl": assets["curl"], "lowres": assets["lowres"], "hires": assets["hires"]}
    S = oracle.Scene(tex, np.zeros(40, np.float32), np.zeros(29, np.float32), np.zeros(13, np.float32), nightsky=star)
    cs = mm.ComputeShader(0, (8, 8), placement=tex["placement"], curl=tex["curl"], lowRes=tex["lowres"], hiRes=tex["hires"], nightSky=star)
    for slot, lo, hi in ((mm.MM_TEX_PLACEMENT, -2, 2), (mm.MM_TEX_CURL, -20, 20), (mm.MM_TEX_LOWRES, -4, 4), (mm.MM_TEX_HIRES, -80, 80),
                         (mm.MM_TEX_NIGHTSKY, -3, 3)):
        uvw = rng.uniform(lo, hi, (200000, 3)).astype(np.float32)
        uvw[:64] = np.float32([[i / 128.0, i / 64.0 - 0.25, 1.0 - i / 32.0] for i in range(64)])   # texel centres / edges
        got = cs.sample(slot, mm.MM_FILTER_HW, uvw)
        want = S.sample(slot, oracle.OM_FILTER_TEXUNIT, uvw)
        bad = (got.view(np.uint32) != want.view(np.uint32)).any(axis=1)
        assert not bad.any(), (slot, int(bad.sum()), uvw[bad][:4], got[bad][:4], want[bad][:4])
    cs.close()


@pytest.mark.parametrize("name,W,H", [("C1", 320, 180), ("C3", 320, 180), ("C2b", 256, 144), ("C5", 256, 144), ("C5b", 256, 144)])
def test_hw_mode_matches_texunit_oracle(mm, oracle, assets, name, W, H):
    """FILTER_HW marches with the hardware sampler.  Against the oracle filtering with the bit-exact model of that
    sampler -- the reference shader as it runs on this GPU -- every decision is identical: zero branch flips,
    identical fetch counters, bit-identical alpha; RGB differs only through the shading transcendentals."""
    sc = scenes.make_scene(mm, name, assets, W=W, H=H)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT)
    ref, rcnt = S.march(W, H)
    img, cnt = _render(mm, sc, mm.MM_FILTER_HW)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    print(name, "hw vs texunit oracle", rep)
    assert rep["branch_flip_pixels"] == 0
    assert rep["counter_mismatch_pixels"] == 0
    assert rep["alpha_identical_frac"] == 1.0
    assert rep["max_abs_diff_8bit"] <= 1 and rep["frac_within_1"] == 1.0
    rel = np.abs(ref - img) / np.maximum(np.abs(ref), 1e-6)
    assert rel.max() < 1e-4
    # production variant (no counters): light-cone samples take the relaxed-arithmetic path
    img2, _ = _render(mm, sc, mm.MM_FILTER_HW, counters=False)
    rep2 = oracle.parity_report(ref, img2)
    print(name, "hw production variant", rep2)
    assert rep2["alpha_identical_frac"] == 1.0
    assert rep2["max_abs_diff_8bit"] <= 2 and rep2["frac_within_1"] >= 0.999


@pytest.mark.parametrize("name,W,H,over", [("C1", 320, 180, {}), ("C3", 256, 144, {}), ("C5b", 200, 113, {}),
                                           ("C1", 192, 108, dict(time=123.5, wind=(0.7, 0.05, -1.3))), ("C1", 97, 61, dict(elevation=0.75))])
@pytest.mark.parametrize("mode", ["hw", "exact", "hybrid"])
def test_lanes_per_ray_change_nothing(mm, oracle, assets, name, W, H, over, mode):
    """mm_set_lanes_per_ray(2|4|8) runs the ray-split kernel: G consecutive loop trips of a ray evaluated side by side and
    replayed through the loop's state machine.  It is a scheduling choice: the image (every bit of it, colour included), the
    counters and therefore every decision equal the one-thread-per-ray kernel, and all equal the oracle's counters."""
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    kfilter, ofilter = getattr(mm, _FILTERS[mode][0]), getattr(oracle, _FILTERS[mode][1])
    out = {}
    for lanes in (1, 2, 4, 8):
        for counters in (True, False):
            cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                                  lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
            cs.allocOutput()
            cs.enableCounters(counters)
            cs.setFilterMode(kfilter)
            cs.setLanesPerRay(lanes)
            img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
            out[lanes, counters] = (img, cs.readCounters() if counters else None)
            cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])       # one reference-style phase dispatch too
            cs.dispatch(mm.MM_PHASE16)
            cs.dispatch(mm.MM_FULL, 1, 3, 2)                                     # and a row-sharded partial frame on top of it
            cs.synchronize()
            out[lanes, counters, "p16"] = cs.readOutput()
            cs.close()
    for lanes in (2, 4, 8):
        for counters in (True, False):
            assert np.array_equal(out[1, counters][0].view(np.uint32), out[lanes, counters][0].view(np.uint32)), (lanes, counters)
            assert np.array_equal(out[1, counters, "p16"].view(np.uint32), out[lanes, counters, "p16"].view(np.uint32)), (lanes, counters)
        assert np.array_equal(out[1, True][1], out[lanes, True][1]), lanes
    _, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ofilter, nightsky=night).march(W, H)
    cols = [0, 3] if mode == "hybrid" else [0, 1, 2, 3]     # hybrid's light-cone samples are filtered differently from its oracle:
    assert np.array_equal(out[8, True][1][..., cols], rcnt[..., cols])      # their fetch counts may differ, trips and lit steps may not


@pytest.mark.parametrize("W,H", [(1, 1), (5, 3), (33, 9)])
def test_lanes_per_ray_on_ragged_extents(mm, assets, W, H):
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    frames = [_render(mm, sc, mm.MM_FILTER_HW, counters=False, trips=lanes)[0] for lanes in (1, 2, 4, 8)]
    for f in frames[1:]:
        assert np.array_equal(f.view(np.uint32), frames[0].view(np.uint32))


def test_hw_mode_against_float_filter_oracle_reports_tail(mm, oracle, assets):
    """The same frame under the two sampler definitions (hardware 8-bit weights vs exact binary32 weights) differs
    by a tail of pixels whose rays flip a threshold (the march is chaotic there, SURVEY 7 hard part 2): two
    conformant Vulkan implementations differ the same way.  Reported and bounded, not hidden."""
    sc = scenes.make_scene(mm, "C1", assets, W=320, H=180)
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(320, 180)
    img, cnt = _render(mm, sc, mm.MM_FILTER_HW)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    print("hw vs float-filter oracle", rep)
    assert rep["frac_within_1"] > 0.97
    assert rep["branch_flip_pixels"] < 0.05 * 320 * 180


def test_phase16_writes_only_its_pixels(mm, oracle, assets):
    """MM_PHASE16 reproduces one reference dispatch (CC:291-301): pixels == offset mod 4x4, rest untouched."""
    W, H = 128, 72
    for phase in (0, 5, 15):
        sc = scenes.make_scene(mm, "C1", assets, W=W, H=H, pixel_phase=phase)
        S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
        sentinel = np.full((H, W, 4), -7.0, np.float32)
        ref, _ = S.march(W, H, mode=oracle.OM_PHASE16, out=sentinel.copy())
        import torch
        t = torch.from_numpy(sentinel.copy()).cuda()
        cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                              lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
        cs.bindOutput(t.data_ptr())
        cs.setFilterMode(mm.MM_FILTER_EXACT)
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
        cs.dispatch(mm.MM_PHASE16)
        cs.synchronize()
        img = t.cpu().numpy()
        cs.close()
        written = (img != -7.0).any(axis=-1)
        ys, xs = np.nonzero(written)
        assert (xs % 4 == phase % 4).all() and (ys % 4 == phase // 4).all()
        assert written.sum() == (W // 4) * (H // 4)
        assert (written == (ref != -7.0).any(axis=-1)).all()
        rep = oracle.parity_report(ref, img)
        assert rep["max_abs_diff_8bit"] <= 1, rep


def test_row_partition_union_equals_full_frame(mm, assets):
    """Multi-GPU sharding contract: the union of the N row-cyclic partitions is byte-identical to one full dispatch."""
    import torch
    W, H = 200, 117      # deliberately ragged: not multiples of the tile or of the row block
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    full, _ = _render(mm, sc, mm.MM_FILTER_EXACT, counters=False)
    for n, block in ((2, 1), (3, 2), (8, 4)):
        t = torch.full((H, W, 4), -7.0, dtype=torch.float32, device="cuda")
        cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                              lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
        cs.bindOutput(t.data_ptr())
        cs.setFilterMode(mm.MM_FILTER_EXACT)
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
        for r in range(n):
            cs.dispatch(mm.MM_FULL, r, n, block)
        cs.synchronize()
        img = t.cpu().numpy()
        assert np.array_equal(img.view(np.uint32), full.view(np.uint32)), (n, block)
        # the boustrophedon order of the same partition (MM_ROWS_SNAKE): each rank writes exactly owned_rows(..., snake=True)
        t.fill_(-7.0)
        for r in range(n):
            cs.dispatch(mm.MM_FULL | mm.MM_ROWS_SNAKE, r, n, block)
            cs.synchronize()
            if r == 0:
                rows = np.nonzero((t.cpu().numpy() != -7.0).any(axis=(1, 2)))[0]
                assert np.array_equal(rows, mm.multigpu.owned_rows(H, 0, n, block, snake=True)), (n, block)
        img = t.cpu().numpy()
        cs.close()
        assert np.array_equal(img.view(np.uint32), full.view(np.uint32)), (n, block, "snake")


def test_render_to_host_end_to_end(mm, oracle, assets):
    sc = scenes.make_scene(mm, "C1", assets, W=160, H=90)
    cs = mm.ComputeShader(0, (160, 90), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])          # a fresh context filters with the texture unit (MM_FILTER_HW)
    rgba8 = cs.tonemapRGBA8()
    cs.close()
    ref, _ = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT).march(160, 90, counters=False)
    rep = oracle.parity_report(ref, img)
    assert rep["alpha_identical_frac"] == 1.0 and rep["max_abs_diff_8bit"] <= 2 and rep["frac_within_1"] >= 0.999, rep
    # K4 tonemap kernel vs the oracle's tonemap of the same device image
    d = np.abs(rgba8.astype(int) - oracle.tonemap_rgba8(img).astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.999


def test_missing_state_fails_loudly(mm, assets):
    cs = mm.ComputeShader(0, (64, 36))
    cs.allocOutput()
    with pytest.raises(mm.MarshmallowError):
        cs.dispatch()          # no uniforms, no textures
    cs.close()


def test_det_pow_bit_exact_on_gpu(mm, oracle):
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.random(200000, dtype=np.float32), np.float32([0, 1, 1e-30, 1e-6, 0.5, 0.999999])])
    y = np.concatenate([rng.uniform(0.8, 1.0, 200000).astype(np.float32), np.float32([0.8, 0.9, 0.85, 1.0, 0.8, 0.95])])
    cs = mm.ComputeShader(0, (8, 8))
    got = cs.detPow(x, y)
    cs.close()
    want = np.array([oracle.lib().om_det_powf(float(a), float(b)) for a, b in zip(x[:20000], y[:20000])], np.float32)
    assert np.array_equal(got[:20000].view(np.uint32), want.view(np.uint32))
    tail = np.array([oracle.lib().om_det_powf(float(a), float(b)) for a, b in zip(x[-6:], y[-6:])], np.float32)
    assert np.array_equal(got[-6:].view(np.uint32), tail.view(np.uint32))
    # and it is a pow: within 1 ulp of the correctly rounded value almost everywhere
    exact = np.power(x.astype(np.float64), y.astype(np.float64)).astype(np.float32)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - exact.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1


def test_exact_sampler_bit_exact_on_gpu(mm, oracle, assets):
    rng = np.random.default_rng(3)
    sc = scenes.make_scene(mm, "C1", assets)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
    cs = mm.ComputeShader(0, (8, 8), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    for slot in (mm.MM_TEX_PLACEMENT, mm.MM_TEX_CURL, mm.MM_TEX_LOWRES, mm.MM_TEX_HIRES):
        uvw = rng.uniform(-3, 60, (50000, 3)).astype(np.float32)
        uvw[:64] = np.float32([[i / 128.0, i / 64.0 - 0.25, 1.0 - i / 32.0] for i in range(64)])   # texel centres / edges
        got = cs.sample(slot, mm.MM_FILTER_EXACT, uvw)
        want = S.sample(slot, oracle.OM_FILTER_FP32, uvw)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), slot
        hw = cs.sample(slot, mm.MM_FILTER_HW, uvw)
        assert np.abs(hw - want).max() < 1.0 / 256 + 1e-3      # 8-bit weights: close, not equal
    cs.close()


def test_exact_divide_by_constant_is_the_ieee_quotient(mm):
    """The march divides by literal constants through a 3-instruction exact sequence (csrc/cloud_march.cu,
    div_const).  Exhaustive: every binary32 dividend, every constant in use, zero mismatching bit patterns."""
    cs = mm.ComputeShader(0, (8, 8))
    which, seen = 0, []
    while True:
        r = cs.selftestDiv(which)
        if r is None:
            break
        c, bad = r
        seen.append(c)
        assert bad == 0, f"div_const(x, {c!r}) differs from x / {c!r} for {bad} dividends"
        which += 1
    cs.close()
    assert len(seen) >= 15          # 12 constants (incl. the 0.3 of `stepSize /= 0.3`) + sqrt + rcp + remapClampedTo1


@pytest.mark.parametrize("name,mode", [("C2", "hw"), ("C3", "hw"), ("C2", "hybrid"), ("C2", "exact"), ("C3", "hybrid")])
def test_full_size_baseline_configs_pass_the_parity_gate(mm, oracle, assets, name, mode):
    """BASELINE configs at their FULL size (1920x1080 noon, 3840x2160 sunset): the north-star gate -- max <= 2/255
    per channel, >= 99.9 % of pixels within 1/255 after the reference tonemap -- plus zero branch flips."""
    sc = scenes.make_scene(mm, name, assets)
    W, H = sc["W"], sc["H"]
    kfilter, ofilter = getattr(mm, _FILTERS[mode][0]), getattr(oracle, _FILTERS[mode][1])
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ofilter).march(W, H)
    img, cnt = _render(mm, sc, kfilter)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    print(name, mode, rep)
    assert rep["branch_flip_pixels"] == 0
    assert rep["alpha_identical_frac"] == 1.0
    assert rep["max_abs_diff_8bit"] <= 2 and rep["frac_within_1"] >= 0.999
    # the production variant (no fetch counters; in HYBRID its light-cone samples use the relaxed arithmetic path)
    img2, _ = _render(mm, sc, kfilter, counters=False)
    rep2 = oracle.parity_report(ref, img2)
    print(name, mode, "production variant", rep2)
    assert rep2["alpha_identical_frac"] == 1.0            # alpha = f(accumulated density): every decision still exact
    assert rep2["max_abs_diff_8bit"] <= 2 and rep2["frac_within_1"] >= 0.999


def test_size_independent_properties_at_8k(mm, assets):
    """C5 (7680x4320 storm) is too large for the oracle in test time; check properties instead: the frame is a
    pure function of its inputs (two renders are bit-identical), a row-cyclic 8-way sharding of it equals the
    single dispatch bit for bit, rays below the horizon carry the sky colour with untouched alpha seed, and every
    value is finite with alpha in [0,1]."""
    import torch
    sc = scenes.make_scene(mm, "C5", assets)
    W, H = sc["W"], sc["H"]
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.setFilterMode(mm.MM_FILTER_HYBRID)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    a = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    b = torch.full((H, W, 4), -7.0, dtype=torch.float32, device="cuda")
    cs.bindOutput(a.data_ptr())
    cs.dispatch(mm.MM_FULL)
    cs.synchronize()
    cs.bindOutput(b.data_ptr())
    for r in range(8):
        cs.dispatch(mm.MM_FULL, r, 8, 2)
    cs.synchronize()
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    assert torch.isfinite(a).all()
    assert float(a[..., 3].min()) >= 0.0 and float(a[..., 3].max()) <= 1.0
    below = a[H - 64:]                     # the bottom rows look below the horizon in this view: no cloud, alpha = seed
    assert float(below[..., :3].min()) > 0.0
    cs.close()


@pytest.mark.parametrize("W,H", [(1, 1), (17, 5), (33, 9), (130, 70)])
def test_ragged_and_tiny_extents(mm, oracle, assets, W, H):
    """Extents that are not multiples of the 16x8 block or the 4x4 phase grid, down to a single pixel."""
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
    ref, rcnt = S.march(W, H)
    img, cnt = _render(mm, sc, mm.MM_FILTER_EXACT)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    assert rep["counter_mismatch_pixels"] == 0 and rep["max_abs_diff_8bit"] <= 1, rep
    # one reference-style phase dispatch on the same ragged extent
    sc5 = scenes.make_scene(mm, "C1", assets, W=W, H=H, pixel_phase=6)
    S5 = oracle.Scene(sc5["textures"], sc5["cam"], sc5["sun"], sc5["sky"])
    sentinel = np.full((H, W, 4), -7.0, np.float32)
    ref5, _ = S5.march(W, H, mode=oracle.OM_PHASE16, out=sentinel.copy())
    import torch
    t = torch.from_numpy(sentinel.copy()).cuda()
    cs = mm.ComputeShader(0, (W, H), placement=sc5["textures"]["placement"], curl=sc5["textures"]["curl"],
                          lowRes=sc5["textures"]["lowres"], hiRes=sc5["textures"]["hires"])
    cs.bindOutput(t.data_ptr())
    cs.setFilterMode(mm.MM_FILTER_EXACT)
    cs.updateUniformBuffers(sc5["cam"], None, sc5["sky"], sc5["sun"])
    cs.dispatch(mm.MM_PHASE16)
    cs.synchronize()
    got = t.cpu().numpy()
    cs.close()
    assert ((got != -7.0).any(-1) == (ref5 != -7.0).any(-1)).all()
    assert oracle.parity_report(ref5, got)["max_abs_diff_8bit"] <= 1


def test_bad_arguments_are_rejected_on_gpu(mm, assets):
    cs = mm.ComputeShader(0, (64, 36), placement=assets["placement"], curl=assets["curl"], lowRes=assets["lowres"], hiRes=assets["hires"])
    cs.allocOutput()
    sun, sky = mm.host_sky(0.25, 0.25)
    cs.updateUniformBuffers(mm.host_camera((0, 1, 1), 0.0, 0.0), None, sky, sun)
    for bad in ((7, 0, 1, 1), (0, -1, 1, 1), (0, 0, 0, 1), (0, 0, 1, 0)):
        with pytest.raises(mm.MarshmallowError):
            cs.dispatch(*bad)
    with pytest.raises(mm.MarshmallowError):
        cs.uploadTexture(mm.MM_TEX_LOWRES, assets["placement"])      # a 2D texture into a 3D sampler slot
    with pytest.raises(mm.MarshmallowError):
        cs.setFilterMode(9)
    for bad_lanes in (3, 16, -1):
        with pytest.raises(mm.MarshmallowError):
            cs.setLanesPerRay(bad_lanes)
    with pytest.raises(mm.MarshmallowError):
        cs.bindHostMirror(np.zeros((36, 64, 4), np.float32))          # pageable host memory cannot be a kernel-side mirror
    cs.dispatch()                                                     # still usable after the errors
    cs.synchronize()
    cs.close()


def test_render_to_host_pinned_and_pageable_agree(mm, assets):
    """mm_render_to_host fuses the device->host transfer into the kernel for page-locked destinations; the frame must be
    bit-identical to the copy path (pageable destination) and to the device image."""
    import torch
    sc = scenes.make_scene(mm, "C1", assets, W=200, H=113)
    cs = mm.ComputeShader(0, (200, 113), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    pageable = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    pinned_t = torch.full((113, 200, 4), -7.0, dtype=torch.float32).pin_memory()
    pinned = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=pinned_t.numpy())
    device = cs.readOutput()
    cs.close()
    assert np.array_equal(pinned.view(np.uint32), pageable.view(np.uint32))
    assert np.array_equal(pinned.view(np.uint32), device.view(np.uint32))


@pytest.mark.parametrize("rest", ["fused", "copy", "copy2", None])
@pytest.mark.parametrize("pitch_deg", [0.0, 50.0])
def test_host_mirror_paths_deliver_the_same_frame(mm, assets, rest, pitch_deg, monkeypatch):
    """The mirrored dispatch has three ways to get a row into host memory (stores fused into the kernel, the copy engine behind
    one launch, the copy engine behind two launches; mm_dispatch picks by the kind of dispatch, MM_E2E_REST forces one).  Every one
    must deliver the device image bit for bit, for whole frames and for shares of a sharded frame, with and without rows below
    the horizon (a camera pitched up has none)."""
    import torch
    if rest is None:
        monkeypatch.delenv("MM_E2E_REST", raising=False)
    else:
        monkeypatch.setenv("MM_E2E_REST", rest)
    W, H = 211, 173
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H, pitch=np.radians(-20.0 + pitch_deg))
    cam = sc["cam"]
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    cs.updateUniformBuffers(cam, None, sc["sky"], sc["sun"])
    pinned = torch.full((H, W, 4), -7.0, dtype=torch.float32).pin_memory()
    cs.bindHostMirror(pinned.numpy())
    cs.dispatch(mm.MM_FULL)
    cs.synchronize()
    whole = cs.readOutput()
    assert np.array_equal(pinned.numpy().view(np.uint32), whole.view(np.uint32))
    for begin, stride in ((1, 3), (2, 4), (7, 8)):
        pinned.fill_(-7.0)
        cs.dispatch(mm.MM_FULL, begin, stride, 8)
        cs.synchronize()
        rows = mm.multigpu.owned_rows(H, begin, stride, 8)
        other = np.setdiff1d(np.arange(H), rows)
        got = pinned.numpy()
        assert (got[other] == -7.0).all()
        assert np.array_equal(got[rows].view(np.uint32), whole[rows].view(np.uint32)), (rest, begin, stride)
    cs.close()


def test_bound_host_mirror_receives_every_dispatch(mm, assets):
    """mm_bind_host_mirror: every dispatch also stores its pixels into a page-locked host frame (what each rank of a sharded
    frame does with the shared host frame); the host frame equals the device image bit for bit, for full and partial dispatches."""
    import torch
    W, H = 203, 117
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    pinned = torch.full((H, W, 4), -7.0, dtype=torch.float32).pin_memory()
    cs.bindHostMirror(pinned.numpy())
    cs.dispatch(mm.MM_FULL, 1, 3, 8)                   # one rank's share of a 3-way sharded frame
    cs.synchronize()
    part = pinned.numpy().copy()
    rows = mm.multigpu.owned_rows(H, 1, 3, 8)
    other = np.setdiff1d(np.arange(H), rows)
    assert (part[other] == -7.0).all() and (part[rows] != -7.0).any(axis=-1).all()
    cs.dispatch(mm.MM_FULL)
    cs.synchronize()
    assert np.array_equal(pinned.numpy().view(np.uint32), cs.readOutput().view(np.uint32))
    cs.bindHostMirror(None)
    pinned.fill_(-7.0)
    cs.dispatch(mm.MM_FULL)
    cs.synchronize()
    assert (pinned.numpy() == -7.0).all()              # unbound: the host frame is left alone
    cs.close()


def test_randomised_scenes_match_their_oracles(mm, oracle, assets):
    """Seeded random cameras, suns (day and night), times, winds and placements, small frames: the default mode against the
    texture-unit-model oracle and the FP32 mode against the binary32 oracle -- counters and alpha bit for bit."""
    rng = np.random.default_rng(2024)
    night_map = scenes.synthetic_night_sky()
    for i in range(12):
        over = dict(yaw=float(rng.uniform(-np.pi, np.pi)), pitch=float(-rng.uniform(0.02, 1.2)), elevation=float(rng.uniform(0.0, 1.0)),
                    azimuth=float(rng.uniform(0.0, 1.0)), time=float(rng.uniform(0, 500)),
                    wind=tuple(float(x) for x in rng.uniform(-1.5, 1.5, 3)), pos=tuple(float(x) for x in rng.uniform(-500, 500, 3)))
        if i % 3 == 2:
            over["placement"] = (int(rng.integers(40, 255)), int(rng.integers(0, 256)))
        W, H = int(rng.integers(40, 140)), int(rng.integers(24, 80))
        sc = scenes.make_scene(mm, "C1", assets, W=W, H=H, **over)
        night = night_map if sc["sun"][5] < 0 else None
        for mode in ("hw", "exact"):
            kfilter, ofilter = getattr(mm, _FILTERS[mode][0]), getattr(oracle, _FILTERS[mode][1])
            ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ofilter, nightsky=night).march(W, H)
            cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                                  lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
            cs.allocOutput()
            cs.enableCounters(True)
            cs.setFilterMode(kfilter)
            cs.setLanesPerRay(int(rng.choice([0, 1, 2, 4, 8])))
            img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
            cnt = cs.readCounters()
            cs.close()
            rep = oracle.parity_report(ref, img, rcnt, cnt)
            assert rep["counter_mismatch_pixels"] == 0 and rep["alpha_identical_frac"] == 1.0 and rep["max_abs_diff_8bit"] <= 1, (i, mode, over, rep)


def test_texture_peak_microbenchmark_reports_a_plausible_ceiling(mm, assets):
    """mm_measure_tex_peak: L1-resident filtered fetches.  A B200's nominal ceiling is 148 SM x 4 quads/clk x 1.965 GHz =
    1.16 Tquad/s; the measurement must be a positive rate of that order (and fail loudly on an unbound slot)."""
    cs = mm.ComputeShader(0, (8, 8), curl=assets["curl"], hiRes=assets["hires"])
    ms2, q2 = cs.measureTexPeak(mm.MM_TEX_CURL, 2048)
    ms3, q3 = cs.measureTexPeak(mm.MM_TEX_HIRES, 2048)
    print("tex peak: bilinear %.1f Gquad/s (%.3f ms), trilinear %.1f Gquad/s (%.3f ms)" % (q2 / 1e9, ms2, q3 / 1e9, ms3))
    assert ms2 > 0 and ms3 > 0
    assert 2e10 < q2 < 1e13 and 2e10 < q3 < 1e13
    with pytest.raises(mm.MarshmallowError):
        cs.measureTexPeak(mm.MM_TEX_LOWRES, 2048)
    cs.close()


def test_dispatch_multi_assembles_one_frame_from_several_contexts(mm, assets):
    """mm_dispatch_multi: n contexts of one process (here three on the same GPU, bound to the same image) each march their
    partition; the frame equals a single context's full dispatch bit for bit."""
    import torch
    W, H = 210, 119
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    full, _ = _render(mm, sc, mm.MM_FILTER_HW, counters=False)
    t = torch.full((H, W, 4), -7.0, dtype=torch.float32, device="cuda")
    shaders = []
    for _ in range(3):
        cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                              lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
        cs.bindOutput(t.data_ptr())
        cs.setFilterMode(mm.MM_FILTER_HW)
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
        shaders.append(cs)
    mm.dispatchMulti(shaders, mm.MM_FULL, 8)
    for cs in shaders:
        cs.synchronize()
    img = t.cpu().numpy()
    for cs in shaders:
        cs.close()
    assert np.array_equal(img.view(np.uint32), full.view(np.uint32))


@pytest.mark.parametrize("mode", ["hw", "exact"])
def test_night_frame_with_the_shipped_star_map(mm, oracle, assets, mode):
    """The night branch (CC:365-384, 491-493) with the star map the application binds: Textures/NightSky/nightSky_noOrange.png,
    1920x1080 -- not a power of two, so REPEAT wraps by modulo in the FP32 sampler and through the 21-bit coordinate of the texture
    unit model (VERDICT r1 missing 5: night was only tested with a synthetic 96x64 map)."""
    night = scenes.shipped_night_sky()
    assert night.shape == (1080, 1920, 4)
    W, H = 480, 270
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H, elevation=0.75, time=40.0)
    assert sc["sun"][5] < 0                                                   # sun.direction.y < 0: night
    kfilter, ofilter = getattr(mm, _FILTERS[mode][0]), getattr(oracle, _FILTERS[mode][1])
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ofilter, nightsky=night).march(W, H)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
    cs.allocOutput()
    cs.enableCounters(True)
    cs.setFilterMode(kfilter)
    img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cnt = cs.readCounters()
    cs.close()
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    print("night, shipped star map,", mode, rep)
    assert rep["counter_mismatch_pixels"] == 0 and rep["alpha_identical_frac"] == 1.0 and rep["max_abs_diff_8bit"] <= 1, rep
    assert int(rcnt[..., 1].max()) > 0 and float(ref[..., :3].max()) > 0.0


@pytest.mark.parametrize("name", ["C5", "C5b"])
def test_native_8k_rows_against_the_oracle(mm, oracle, assets, name):
    """BASELINE config 5 at its NATIVE 7680x4320: every 64th row of the frame against the oracle -- counters and alpha bit for bit,
    RGBA8 within 1 -- next to the size-independent properties above.  C5b is the measured worst case."""
    import torch
    sc = scenes.make_scene(mm, name, assets)
    W, H = sc["W"], sc["H"]
    rows = np.arange(0, H, 64)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT)
    ref, rcnt = S.march(W, H, row_begin=0, row_stride=64, row_block=1)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    out = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    cs.bindOutput(out.data_ptr())
    cs.enableCounters(True)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    cs.dispatch(mm.MM_FULL)
    cs.synchronize()
    img = out[::64].cpu().numpy()
    cnt = cs.readCounters()[::64]
    cs.close()
    rep = oracle.parity_report(ref[rows], img, rcnt[rows], cnt)
    print(name, "native 8K, every 64th row", rep)
    assert rep["branch_flip_pixels"] == 0 and rep["counter_mismatch_pixels"] == 0 and rep["alpha_identical_frac"] == 1.0
    assert rep["max_abs_diff_8bit"] <= 1 and rep["frac_within_1"] == 1.0
    assert int(rcnt[rows][..., 0].sum()) > 1000000
