"""The exact pow filter of the march (cloud_march.cu, cloudTest, MM_POW_FILTER) as a model in the oracle: the same predicate decides, from
an estimate of coverage = h^k, the outcome of CC:247-250 without the deterministic pow whenever it can.  The oracle evaluates the exact path
as always and compares every prediction with it bit for bit (pow_filter_class); a single mismatch would mean the kernel's shortcut can
change a frame.  CPU only; the GPU side of the claim is the unchanged set of bit-exact parity tests (tests/test_march_parity_gpu.py)."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import scenes


def _classes(name, assets, W, H, filter_mode):
    sc = scenes.scene_from_config(name, assets)
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=filter_mode)
    lib = ob.lib()
    buf = np.zeros((H, W, 256), np.uint8)
    mism = C.c_ulonglong.in_dll(lib, "om_powclass_mismatches")
    before = mism.value
    lib.om_set_powclass_buffer.argtypes = [C.c_void_p]
    lib.om_set_powclass_buffer(buf.ctypes.data)
    try:
        _, cnt = S.march(W, H)
    finally:
        lib.om_set_powclass_buffer(None)
        S.close()
    return buf & 0x7f, cnt, mism.value - before          # bit 7 is another diagnostic (tools/pow_filter_bound.py)


@pytest.mark.parametrize("name,filter_mode", [("C1", ob.OM_FILTER_TEXUNIT), ("C1", ob.OM_FILTER_FP32), ("C3", ob.OM_FILTER_TEXUNIT),
                                              ("C2", ob.OM_FILTER_FP32), ("C2b", ob.OM_FILTER_TEXUNIT), ("C5", ob.OM_FILTER_TEXUNIT), ("C5b", ob.OM_FILTER_FP32)])
def test_filter_predictions_equal_the_exact_path(assets, name, filter_mode):
    W, H = 240, 135
    cls, cnt, mismatches = _classes(name, assets, W, H, filter_mode)
    assert mismatches == 0
    trips = int(cnt[..., 0].sum())
    n = {c: int((cls == c).sum()) for c in (1, 2, 3)}
    calls = sum(n.values())
    assert trips > 100_000
    if name in ("C2b", "C5b"):                               # constant coverage 0.5 <= 0.7: k == 1 everywhere, the pow is never called
        assert calls == 0
        return
    assert calls > 0.03 * trips                              # the pow is on the path of a real share of the trips ...
    assert n[1] > 0 and n[2] > 0 and n[3] > 0                # ... every class occurs ...
    assert n[3] < 0.5 * calls                                # ... and the exact pow is needed for fewer than half of the calls


def test_filter_skips_whole_warps(assets):
    """K1 is warp-synchronous: a warp pays for the out-of-line pow when any lane of its 8x4 tile needs it in that iteration."""
    W, H = 480, 64
    cls, cnt, mismatches = _classes("C3", assets, W, H, ob.OM_FILTER_TEXUNIT)
    assert mismatches == 0
    tiles = cls.reshape(H // 4, 4, W // 8, 8, 256).transpose(0, 2, 1, 3, 4).reshape(-1, 32, 256)
    today, filtered = (tiles > 0).any(axis=1).sum(), (tiles == 3).any(axis=1).sum()
    assert today > 0 and filtered < 0.7 * today
