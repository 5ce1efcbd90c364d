"""K3 has no reference arithmetic to follow (the volumes are a Houdini bake; parity unpinned).  What can be pinned is
statistical: the generator's channels must look like the shipped volumes'  and must tile.  Run on the CPU
statement of the generator; the GPU build is byte-identical to it (tests/test_noise_gpu.py)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def volumes(oracle):
    return oracle.build_noise_volumes(0)


def test_integer_hash_known_answers(oracle):
    l = oracle.lib()
    assert l.om_noise_hash(0, 0, 0, 0) == 0                      # fmix32(0) == 0
    vals = {l.om_noise_hash(x, y, z, 1) for x in range(8) for y in range(8) for z in range(8)}
    assert len(vals) == 512                                      # no collisions on a small lattice
    bits = np.array([[(l.om_noise_hash(x, 3, 5, 9) >> b) & 1 for b in range(32)] for x in range(4096)])
    assert np.all(np.abs(bits.mean(0) - 0.5) < 0.05)             # every output bit is balanced


def test_channel_statistics_match_shipped_volumes(volumes, assets):
    low, hi = volumes
    for ours, ref, chans in ((low, assets["lowres"], 4), (hi, assets["hires"], 3)):
        a = ours.reshape(-1, 4)[:, :chans] / 255.0
        b = ref.reshape(-1, 4)[:, :chans] / 255.0
        assert np.allclose(a.mean(0), b.mean(0), atol=0.01)
        assert np.allclose(a.std(0), b.std(0), atol=0.01)
    R = low[..., 0].ravel() / 255.0
    assert np.allclose(np.percentile(R, [25, 50, 75]), [0.435, 0.506, 0.576], atol=0.02)      # SURVEY 8c percentiles
    assert (R > 0.3).mean() > 0.95                                                           # shipped: 97.2 %
    fbm = (0.625 * low[..., 1] + 0.25 * low[..., 2] + 0.125 * low[..., 3]) / 255.0           # the erosion CC:247 builds
    assert fbm.mean() == pytest.approx(0.686, abs=0.01) and fbm.std() == pytest.approx(0.071, abs=0.01)
    assert (hi[..., 3] == 0).all()                                                           # shipped hi-res alpha is identically 0


def test_volumes_tile_seamlessly(volumes):
    for v in volumes:
        f = v[..., :3].astype(np.float32)
        for ax in range(3):
            interior = np.abs(np.diff(f, axis=ax)).mean()
            seam = np.abs(np.take(f, 0, axis=ax) - np.take(f, -1, axis=ax)).mean()
            assert seam < 1.5 * interior


def test_seed_changes_the_volumes_not_their_statistics(oracle, volumes):
    low7, hi7 = oracle.build_noise_volumes(7)
    assert (low7 != volumes[0]).mean() > 0.9
    assert np.allclose(low7.reshape(-1, 4).mean(0), volumes[0].reshape(-1, 4).mean(0), atol=3.0)
