"""K1p -- persistent warps pulling pixel tiles from a dynamic queue (mm_set_scheduler) -- is a scheduling choice: every image bit
and every work counter must equal the static-grid kernel (K1), for every refill threshold, sampler mode, dispatch mode and
partition; plus the argument checks the advisor asked for around mm_dispatch."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _ctx(mm, sc, W, H, night=None):
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
    cs.allocOutput()
    return cs


@pytest.mark.parametrize("name,W,H,over", [("C1", 320, 180, {}), ("C3", 256, 144, {}), ("C5b", 200, 113, {}),
                                           ("C1", 97, 61, dict(elevation=0.75)), ("C1", 33, 9, {}), ("C1", 1, 1, {})])
@pytest.mark.parametrize("mode", ["hw", "exact", "hybrid"])
def test_persistent_scheduler_changes_nothing(mm, oracle, assets, name, W, H, over, mode):
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    kfilter = {"hw": mm.MM_FILTER_HW, "exact": mm.MM_FILTER_EXACT, "hybrid": mm.MM_FILTER_HYBRID}[mode]
    out = {}
    for sched, refill in ((mm.MM_SCHED_STATIC, 0), (mm.MM_SCHED_PERSISTENT, 32), (mm.MM_SCHED_PERSISTENT, 16), (mm.MM_SCHED_PERSISTENT, 8)):
        for counters in (True, False):
            cs = _ctx(mm, sc, W, H, night)
            cs.enableCounters(counters)
            cs.setFilterMode(kfilter)
            cs.setLanesPerRay(1)
            cs.setScheduler(sched, refill)
            img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
            cnt = cs.readCounters() if counters else None
            cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
            cs.dispatch(mm.MM_PHASE16)                                           # a reference-style phase dispatch ...
            cs.dispatch(mm.MM_FULL, 1, 3, 2)                                     # ... and a row-sharded partial frame on top of it
            cs.synchronize()
            out[sched, refill, counters] = (img, cnt, cs.readOutput())
            cs.close()
    base = out[mm.MM_SCHED_STATIC, 0, True]
    for key, (img, cnt, p16) in out.items():
        ref = out[mm.MM_SCHED_STATIC, 0, key[2]]
        assert np.array_equal(img.view(np.uint32), ref[0].view(np.uint32)), key
        assert np.array_equal(p16.view(np.uint32), ref[2].view(np.uint32)), key
        if cnt is not None:
            assert np.array_equal(cnt, base[1]), key


def test_persistent_scheduler_matches_the_oracle_at_1080p(mm, oracle, assets):
    """The default dispatch (K1p, a tile at a time) on the full-size C2 frame against the texture-unit-model oracle."""
    sc = scenes.make_scene(mm, "C2", assets)
    W, H = sc["W"], sc["H"]
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT).march(W, H)
    for refill in (32, 8):
        cs = _ctx(mm, sc, W, H)
        cs.enableCounters(True)
        cs.setScheduler(mm.MM_SCHED_PERSISTENT, refill)
        img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
        cnt = cs.readCounters()
        cs.close()
        rep = oracle.parity_report(ref, img, rcnt, cnt)
        assert rep["branch_flip_pixels"] == 0 and rep["counter_mismatch_pixels"] == 0 and rep["alpha_identical_frac"] == 1.0, (refill, rep)
        assert rep["max_abs_diff_8bit"] <= 1, (refill, rep)


@pytest.mark.parametrize("lanes", [0, 1, 4, 8])
def test_non_power_of_two_march_textures_with_any_lane_setting(mm, oracle, assets, lanes):
    """a non-power-of-two march texture in an FP32-sampler mode runs the generic one-lane kernel; the lane
    count must be decided before the block rows are planned, or half the rows of a small frame are never written."""
    W, H = 96, 54                                       # far below one wave: lanes 0 would pick the 8-lane split
    rng = np.random.default_rng(5)
    placement = rng.integers(0, 256, (50, 27, 4), dtype=np.uint8)       # 27 x 50: REPEAT wrap by modulo
    placement[..., 2] = rng.integers(100, 256, (50, 27))
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    sc["textures"] = dict(sc["textures"], placement=placement)
    ref, rcnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(W, H)
    for sched in (mm.MM_SCHED_STATIC, mm.MM_SCHED_PERSISTENT):
        import torch
        t = torch.full((H, W, 4), -7.0, dtype=torch.float32, device="cuda")
        cs = mm.ComputeShader(0, (W, H), placement=placement, curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
        cs.bindOutput(t.data_ptr())
        cs.enableCounters(True)
        cs.setFilterMode(mm.MM_FILTER_EXACT)
        cs.setLanesPerRay(lanes)
        cs.setScheduler(sched)
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
        cs.dispatch()
        cs.synchronize()
        img, cnt = t.cpu().numpy(), cs.readCounters()
        cs.close()
        assert (img != -7.0).any(axis=-1).all(), "rows left unwritten"
        rep = oracle.parity_report(ref, img, rcnt, cnt)
        assert rep["counter_mismatch_pixels"] == 0 and rep["alpha_identical_frac"] == 1.0 and rep["max_abs_diff_8bit"] <= 1, (lanes, sched, rep)


def test_dispatch_argument_checks(mm, assets):
    """the pixel phase of MM_PHASE16 must be 0..15; 0 <= row_begin < row_stride in both modes and in the planner."""
    sc = scenes.make_scene(mm, "C1", assets, W=64, H=36)
    cs = _ctx(mm, sc, 64, 36)
    for bad_phase in (-1.0, 16.0, float("nan"), -1e9):
        sun = sc["sun"].copy()
        sun[11] = bad_phase
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sun)
        with pytest.raises(mm.MarshmallowError):
            cs.dispatch(mm.MM_PHASE16)
        cs.dispatch(mm.MM_FULL)                          # the full-frame mode never reads the phase
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    for mode in (mm.MM_FULL, mm.MM_PHASE16, mm.MM_FULL | mm.MM_ROWS_SNAKE):
        for begin, stride in ((2, 2), (5, 3), (1, 1)):
            with pytest.raises(mm.MarshmallowError):
                cs.dispatch(mode, begin, stride, 2)
    for bad in ((9, 0), (2, 40), (1, 7)):
        with pytest.raises(mm.MarshmallowError):
            cs.setScheduler(*bad)
    cs.synchronize()
    cs.close()
    with pytest.raises(mm.MarshmallowError):
        mm.plan_block_rows(sc["cam"], 36, mm.MM_FULL, 3, 3, 2, 8)


def test_default_mode_keeps_no_float_copies(mm, assets):
    """the default (texture-unit) mode needs only the cudaArrays; the 16x larger pair-major float copies are
    built on the first dispatch of an FP32-sampler mode.  Observable as device memory: ~80 MB appear only then."""
    import torch
    sc = scenes.make_scene(mm, "C1", assets, W=64, H=36)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    cs = _ctx(mm, sc, 64, 36)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    cs.dispatch()
    cs.synchronize()
    used_hw = free0 - torch.cuda.mem_get_info()[0]
    cs.setFilterMode(mm.MM_FILTER_EXACT)
    cs.dispatch()
    cs.synchronize()
    used_exact = free0 - torch.cuda.mem_get_info()[0]
    cs.close()
    assert used_hw < 48 << 20, used_hw
    assert used_exact - used_hw > 60 << 20, (used_hw, used_exact)
