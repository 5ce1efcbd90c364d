"""Cloud shadows: the mesh shader's 6-step march (model.frag:240-283) as a standalone pass.

CPU half: properties and hand-derived known answers of the oracle (oracle/cloud_march_oracle.c: om_cloud_shadow).
GPU half: the CUDA pass against the oracle, bit for bit, under both sampler definitions."""
import numpy as np
import pytest

import scenes


def _points(seed=0, n=20000):
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-50, 50, (n, 3)),                                   # mesh-sized scene around the origin
                           rng.uniform(-20000, 20000, (n, 3)) * np.float32([1, 0.02, 1]),    # terrain-sized ground patch
                           np.float32([[0, 0, 0], [0, 1, 1], [1e6, 0, 0], [0, -10, 0]])]).astype(np.float32)


CASES = [("C1", {}), ("C3", {}), ("C5", {}), ("C1", dict(time=60.0, wind=(0.7, 0.05, -1.3))), ("C1", dict(pitch=0.3, yaw=0.4))]


def test_shadow_oracle_properties(mm, oracle, assets):
    pos = _points()
    seen_any = False
    for name, over in CASES:
        sc = scenes.make_scene(mm, name, assets, **over)
        S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
        d, nf = S.cloud_shadow(pos, want_fetches=True)
        assert np.isfinite(d).all() and d.min() >= 0.0 and d.max() <= 1.0
        # model.frag:269-272: a density above 0.99 is stored as exactly 1 and ends the march
        assert not ((d > 0.99) & (d < 1.0)).any()
        assert ((nf == 12) | (d == 1.0)).all() and (nf % 2 == 0).all() and nf.max() <= 12
        assert np.array_equal(d, S.cloud_shadow(pos, nthreads=1))                           # a pure function of its inputs
        seen_any |= bool((d > 0).mean() > 0.03)
    assert seen_any


def test_shadow_is_zero_without_coverage_and_uses_the_mesh_shaders_constants(mm, oracle, assets):
    sc = scenes.make_scene(mm, "C1", assets)
    pos = _points(1, 4000)
    tex = dict(sc["textures"])
    tex["placement"] = scenes.constant_placement(0, 0)             # cloud type 0 -> stratus only; still nonzero near the shell base
    lo = np.zeros_like(assets["lowres"])                            # R = 0 <= 0.3: remapClamped(...) = 0 -> density 0 everywhere
    tex["lowres"] = lo
    d = oracle.Scene(tex, sc["cam"], sc["sun"], sc["sky"]).cloud_shadow(pos)
    assert (d == 0).all()
    # S3: the exponent floor is 0.6 (CC uses 0.8): coverage = h^k with k = clamp(remap(min(.85, r), .7, .8, 1, .6), .6, 1)
    k = oracle.lib().om_remap(0.85, 0.7, 0.8, 1.0, 0.6)
    assert abs(k - 0.4) < 1e-6                                      # below the floor -> clamped to 0.6 by the shader


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "hw"])
@pytest.mark.parametrize("name,over", CASES)
def test_cuda_cloud_shadow_matches_oracle(mm, oracle, assets, name, over, mode):
    sc = scenes.make_scene(mm, name, assets, **over)
    pos = _points(5)
    ofilt = oracle.OM_FILTER_TEXUNIT if mode == "hw" else oracle.OM_FILTER_FP32
    ref, rnf = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=ofilt).cloud_shadow(pos, want_fetches=True)
    cs = mm.ComputeShader(0, (8, 8), placement=sc["textures"]["placement"], lowRes=sc["textures"]["lowres"])
    cs.setFilterMode(mm.MM_FILTER_HW if mode == "hw" else mm.MM_FILTER_EXACT)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    got, nf = cs.cloudShadow(pos, want_fetches=True)
    cs.close()
    bad = got.view(np.uint32) != ref.view(np.uint32)
    assert not bad.any(), (int(bad.sum()), pos[bad][:3], got[bad][:3], ref[bad][:3])
    assert np.array_equal(nf, rnf)
    assert (ref > 0).mean() > 0.03                                  # the comparison saw clouds, not only clear sky


@pytest.mark.gpu
def test_cloud_shadow_device_arrays_and_errors(mm, oracle, assets):
    import torch
    sc = scenes.make_scene(mm, "C5", assets)
    pos = _points(9, 3000)
    cs = mm.ComputeShader(0, (8, 8), placement=sc["textures"]["placement"], lowRes=sc["textures"]["lowres"])
    with pytest.raises(mm.MarshmallowError):
        cs.cloudShadow(pos)                                        # no uniforms yet
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    cs.setFilterMode(mm.MM_FILTER_HW)
    host = cs.cloudShadow(pos)
    dpos = torch.from_numpy(pos).cuda()
    dout = torch.full((len(pos),), -1.0, dtype=torch.float32, device="cuda")
    cs.cloudShadowDevice(dpos.data_ptr(), len(pos), dout.data_ptr())
    cs.synchronize()
    assert np.array_equal(dout.cpu().numpy(), host)
    assert cs.cloudShadow(np.zeros((0, 3), np.float32)).shape == (0,)
    cs.close()
    bare = mm.ComputeShader(0, (8, 8))
    bare.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    with pytest.raises(mm.MarshmallowError):
        bare.cloudShadow(pos)                                      # textures not bound
    bare.close()
