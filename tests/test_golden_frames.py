"""Committed golden frames (tools/make_golden_frames.py): the oracle must keep reproducing them (CPU), and the
CUDA march must match them (GPU) -- decisions and alpha bit for bit, colour within the RGBA8 gate."""
import ast
import glob
import os

import numpy as np
import pytest

import scenes

FRAMES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frames", "*.npz")))
# the same cases filtered by the oracle's bit-exact model of the B200 texture unit: targets of MM_FILTER_HW
FRAMES_TEXUNIT = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frames_texunit", "*.npz")))


def _scene(mm, assets, g):
    name, W, H, over = ast.literal_eval(str(g["config"]))
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    assert np.array_equal(sc["cam"], g["cam"]) and np.array_equal(sc["sun"], g["sun"]) and np.array_equal(sc["sky"], g["sky"])
    night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
    return sc, W, H, night


def test_there_are_golden_frames():
    assert len(FRAMES) >= 5 and len(FRAMES_TEXUNIT) >= 5


def _ids(paths):
    return [os.path.basename(os.path.dirname(p)) + "/" + os.path.basename(p) for p in paths]


@pytest.mark.parametrize("path", FRAMES + FRAMES_TEXUNIT, ids=_ids(FRAMES + FRAMES_TEXUNIT))
def test_oracle_reproduces_golden(mm, oracle, assets, path):
    g = np.load(path)
    sc, W, H, night = _scene(mm, assets, g)
    filt = oracle.OM_FILTER_TEXUNIT if "frames_texunit" in path else oracle.OM_FILTER_FP32
    img, cnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], nightsky=night, filter_mode=filt).march(W, H)
    assert np.array_equal(cnt, g["counters"].astype(np.uint32))
    assert np.array_equal(img[..., 3], g["rgba32f"][..., 3])
    d = np.abs(oracle.tonemap_rgba8(img).astype(int) - g["rgba8"].astype(int))
    assert d.max() <= 1                     # libm may differ by an ulp between hosts: shading only


@pytest.mark.gpu
@pytest.mark.parametrize("path", FRAMES, ids=[os.path.basename(p) for p in FRAMES])
@pytest.mark.parametrize("mode", ["exact", "hybrid"])
def test_cuda_matches_golden(mm, oracle, assets, path, mode):
    g = np.load(path)
    sc, W, H, night = _scene(mm, assets, g)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
    cs.allocOutput()
    cs.enableCounters(True)
    cs.setFilterMode(mm.MM_FILTER_EXACT if mode == "exact" else mm.MM_FILTER_HYBRID)
    img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cnt = cs.readCounters()
    cs.close()
    assert np.array_equal(cnt[..., 0], g["counters"][..., 0]), "loop-trip counts differ: a march decision flipped"
    assert np.array_equal(cnt[..., 3], g["counters"][..., 3])
    assert np.array_equal(img[..., 3], g["rgba32f"][..., 3]), "alpha depends only on the decision path and must be bit-exact"
    d = np.abs(oracle.tonemap_rgba8(img).astype(int) - g["rgba8"].astype(int)).max(axis=-1)
    assert d.max() <= (1 if mode == "exact" else 2) and (d <= 1).mean() >= 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("path", FRAMES_TEXUNIT, ids=_ids(FRAMES_TEXUNIT))
def test_cuda_hw_sampler_matches_texunit_golden(mm, oracle, assets, path):
    """MM_FILTER_HW (texture-unit filtering, the default) against the oracle frames filtered by the texture-unit model."""
    g = np.load(path)
    sc, W, H, night = _scene(mm, assets, g)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"], nightSky=night)
    cs.allocOutput()
    cs.enableCounters(True)
    cs.setFilterMode(mm.MM_FILTER_HW)
    img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cnt = cs.readCounters()
    cs.enableCounters(False)                       # production variant: relaxed-arithmetic light-cone samples
    img2 = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cs.close()
    assert np.array_equal(cnt, g["counters"].astype(np.uint32)), "a march decision or fetch count differs"
    for im, tol in ((img, 1), (img2, 2)):
        assert np.array_equal(im[..., 3], g["rgba32f"][..., 3]), "alpha depends only on the decision path and must be bit-exact"
        d = np.abs(oracle.tonemap_rgba8(im).astype(int) - g["rgba8"].astype(int)).max(axis=-1)
        assert d.max() <= tol and (d <= 1).mean() >= 0.999
