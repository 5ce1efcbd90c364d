"""GPU noise-texture builds: K2 curl noise against the reference's shipped golden vector, K3 volumes
against the CPU statement of the same generator (byte-identical) and the shipped volumes' statistics."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_curl_noise_equals_shipped_golden_vector(mm, assets):
    cs = mm.ComputeShader(0, (8, 8))
    got = cs.buildCurlNoise()
    cs.close()
    assert got.shape == (128, 128, 4)
    assert (got[..., 3] == 255).all()
    assert np.array_equal(got, assets["curl"]), f"{(got != assets['curl']).sum()} bytes differ from Textures/CurlNoiseFBM.tga"


def test_curl_noise_is_bound_to_the_sampler_slot(mm, assets):
    """mm_build_curl_noise binds MM_TEX_CURL: sampling it equals sampling an upload of the shipped texture."""
    rng = np.random.default_rng(0)
    uvw = rng.uniform(-2, 2, (4096, 3)).astype(np.float32)
    a = mm.ComputeShader(0, (8, 8))
    a.buildCurlNoise(want_copy=False)
    sa = a.sample(mm.MM_TEX_CURL, mm.MM_FILTER_EXACT, uvw)
    a.close()
    b = mm.ComputeShader(0, (8, 8), curl=assets["curl"])
    sb = b.sample(mm.MM_TEX_CURL, mm.MM_FILTER_EXACT, uvw)
    b.close()
    assert np.array_equal(sa, sb)


def test_noise_volumes_byte_identical_to_cpu_statement(mm, oracle):
    cs = mm.ComputeShader(0, (8, 8))
    for seed in (0, 12345678901234567):
        low, hi = cs.buildNoiseVolumes(seed)
        rlow, rhi = oracle.build_noise_volumes(seed)
        assert np.array_equal(hi, rhi), (seed, (hi != rhi).sum())
        assert np.array_equal(low, rlow), (seed, (low != rlow).sum())
    cs.close()


def test_noise_volumes_tile_and_feed_the_march(mm, oracle, assets):
    import scenes
    cs = mm.ComputeShader(0, (96, 54), placement=assets["placement"], curl=assets["curl"])
    low, hi = cs.buildNoiseVolumes(0)
    # seamless tiling: wrap-around neighbour differences look like interior neighbour differences
    for v in (low, hi):
        f = v[..., :3].astype(np.float32)
        for ax in range(3):
            interior = np.abs(np.diff(f, axis=ax)).mean()
            seam = np.abs(np.take(f, 0, axis=ax) - np.take(f, -1, axis=ax)).mean()
            assert seam < 1.5 * interior
    assert (hi[..., 3] == 0).all()
    # the generated volumes are bound: a frame renders with them and matches the oracle given the same bytes
    sc = scenes.make_scene(mm, "C1", assets, W=96, H=54)
    cs.allocOutput()
    cs.enableCounters(True)
    img = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cnt = cs.readCounters()
    cs.close()
    tex = dict(sc["textures"], lowres=low, hires=hi)
    # a fresh context marches with the hardware sampler -> the oracle filters with its texture-unit model
    ref, rcnt = oracle.Scene(tex, sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT).march(96, 54)
    rep = oracle.parity_report(ref, img, rcnt, cnt)
    assert rep["counter_mismatch_pixels"] == 0 and rep["max_abs_diff_8bit"] <= 1, rep
