"""K1x2 -- two neighbouring rays per thread on packed FP32 instructions (mm_set_scheduler(MM_SCHED_PACKED)) -- must produce K1's frame bit
for bit: every component of a packed operation is the scalar kernel's IEEE operation.  K1 itself is pinned to the oracle by
tests/test_march_parity_gpu.py and tests/test_arith_fma_gpu.py (counters, alpha, RGBA8), so equality with K1 carries those guarantees over;
the full-size frames are also compared with the oracle directly."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _frames(mm, sc, W, H, sched, arith, night=None):
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
    cs.allocOutput()
    cs.setArithmetic(arith)
    cs.setLanesPerRay(1)
    cs.setScheduler(sched)
    full = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    cs.dispatch(mm.MM_PHASE16)                                           # a reference-style phase dispatch ...
    cs.dispatch(mm.MM_FULL, 1, 3, 2)                                     # ... and a row-sharded partial frame on top of it
    cs.dispatch(mm.MM_FULL | mm.MM_ROWS_SNAKE, 0, 3, 8)
    cs.synchronize()
    mixed = cs.readOutput()
    cs.close()
    return full, mixed


@pytest.mark.parametrize("name,W,H,over", [("C1", 320, 180, {}), ("C3", 256, 144, {}), ("C5b", 200, 113, {}), ("C2b", 131, 77, {}),
                                           ("C1", 192, 108, dict(time=123.5, wind=(0.7, 0.05, -1.3))), ("C1", 97, 61, dict(elevation=0.75)),
                                           ("C1", 33, 9, {}), ("C1", 2, 1, {}), ("C1", 1, 1, {})])
@pytest.mark.parametrize("arith", ["ieee", "fma"])
def test_packed_kernel_equals_k1(mm, assets, name, W, H, over, arith):
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    a = mm.MM_ARITH_FMA if arith == "fma" else mm.MM_ARITH_IEEE
    want = _frames(mm, sc, W, H, mm.MM_SCHED_STATIC, a, night)
    got = _frames(mm, sc, W, H, mm.MM_SCHED_PACKED, a, night)
    for w, g, what in zip(want, got, ("full frame", "phase + partitions")):
        bad = (w.view(np.uint32) != g.view(np.uint32)).any(axis=-1)
        assert not bad.any(), (what, int(bad.sum()), np.argwhere(bad)[:4].tolist(), w[bad][:2], g[bad][:2])


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_packed_kernel_full_size(mm, oracle, assets, name):
    sc = scenes.make_scene(mm, name, assets)
    W, H = sc["W"], sc["H"]
    want, _ = _frames(mm, sc, W, H, mm.MM_SCHED_STATIC, mm.MM_ARITH_IEEE)
    got, _ = _frames(mm, sc, W, H, mm.MM_SCHED_PACKED, mm.MM_ARITH_IEEE)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    ref, _ = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT).march(W, H, counters=False)
    rep = oracle.parity_report(ref, got)
    assert rep["alpha_identical_frac"] == 1.0 and rep["max_abs_diff_8bit"] <= 2 and rep["frac_within_1"] >= 0.999, rep


def test_packed_request_falls_back_where_it_does_not_apply(mm, assets):
    """The packed kernel exists for the texture-unit mode without counters; any other request runs K1 and the bits are the same anyway."""
    sc = scenes.make_scene(mm, "C1", assets, W=96, H=54)
    for filt in (mm.MM_FILTER_EXACT, mm.MM_FILTER_HYBRID, mm.MM_FILTER_HW):
        outs = []
        for sched in (mm.MM_SCHED_STATIC, mm.MM_SCHED_PACKED):
            cs = mm.ComputeShader(0, (96, 54), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                                  lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
            cs.allocOutput()
            cs.enableCounters(True)
            cs.setFilterMode(filt)
            cs.setScheduler(sched)
            img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
            outs.append((img, cs.readCounters()))
            cs.close()
        assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32)) and np.array_equal(outs[0][1], outs[1][1])
