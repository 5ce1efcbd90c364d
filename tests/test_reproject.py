"""Reprojection pass (reproject.comp): oracle properties on CPU, CUDA vs oracle bit-exactness and the
engine's reproject + 1/16-phase cadence on GPU."""
import numpy as np
import pytest

import scenes


def _cams(mm, dyaw=0.01, dpitch=-0.004, dpos=(30.0, 2.0, -20.0)):
    prev = mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD)
    cur = mm.host_camera((0 + dpos[0], 1 + dpos[1], 1 + dpos[2]), -np.pi / 2 + dyaw, -20 * scenes.DEG2RAD + dpitch)
    return cur, prev


def test_static_camera_reprojects_onto_itself(mm, oracle):
    rng = np.random.default_rng(0)
    src = rng.random((54, 96, 4), dtype=np.float32)
    cam = mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD)
    out = oracle.reproject(cam, cam, src)
    # same camera: every tap lands on (or within a rounding of) the pixel itself
    close = np.isclose(out, src, rtol=1e-5, atol=1e-6).all(axis=-1)
    assert close.mean() > 0.98
    assert np.array_equal(out[..., 3][close], src[..., 3][close])         # alpha is a single tap, never averaged


def test_moving_camera_shifts_the_image(mm, oracle):
    src = np.zeros((72, 128, 4), np.float32)
    src[:, 64:] = 1.0                                                      # a vertical edge
    cur, prev = _cams(mm, dyaw=0.05, dpitch=0.0, dpos=(0, 0, 0))
    out = oracle.reproject(cur, prev, src)
    edge_src = np.argmax(src[36, :, 0] > 0.5)
    edge_out = np.argmax(out[36, :, 0] > 0.5)
    assert edge_out != edge_src                                            # the edge moved with the yaw
    assert np.isfinite(out).all() and out.min() >= 0 and out.max() <= 1


@pytest.mark.gpu
def test_cuda_reproject_is_bit_exact(mm, oracle):
    import torch
    rng = np.random.default_rng(1)
    for (W, H) in ((96, 54), (201, 113), (1920, 1080)):
        src = (rng.random((H, W, 4), dtype=np.float32) * 40).astype(np.float32)
        cur, prev = _cams(mm)
        want = oracle.reproject(cur, prev, src)
        sun, sky = mm.host_sky(0.25, 0.25)
        cs = mm.ComputeShader(0, (W, H))
        t_src = torch.from_numpy(src).cuda()
        t_dst = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        cs.bindOutput(t_dst.data_ptr())
        cs.bindPrevious(t_src.data_ptr())
        cs.updateUniformBuffers(cur, prev, sky, sun)
        cs.dispatchReproject()
        cs.synchronize()
        got = t_dst.cpu().numpy()
        if (W, H) == (1920, 1080):
            ms = cs.lastKernelMs()
            print(f"reproject 1080p: {ms:.3f} ms, {W * H * 32 / ms / 1e6:.0f} GB/s algorithmic (32 B/pixel)")
        cs.close()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (W, H)


@pytest.mark.gpu
def test_reproject_needs_previous_camera_and_distinct_images(mm):
    import torch
    cs = mm.ComputeShader(0, (64, 36))
    t = torch.zeros((36, 64, 4), dtype=torch.float32, device="cuda")
    cs.bindOutput(t.data_ptr())
    sun, sky = mm.host_sky(0.25, 0.25)
    cam = mm.host_camera((0, 1, 1), 0.0, 0.0)
    cs.updateUniformBuffers(cam, None, sky, sun)
    cs.bindPrevious(t.data_ptr())
    with pytest.raises(mm.MarshmallowError):
        cs.dispatchReproject()                      # no previous camera block
    cs.updateUniformBuffers(cam, cam, sky, sun)
    with pytest.raises(mm.MarshmallowError):
        cs.dispatchReproject()                      # previous == target
    cs.close()


@pytest.mark.gpu
def test_engine_cadence_reproject_plus_phase16(mm, oracle, assets):
    """The reference frame (VulkanApplication.cpp:1053-1071): reproject the previous image, then re-march the 1/16 of
    the pixels selected by sun.color.a, ping-pong.  Five frames with a drifting camera, CUDA vs oracle."""
    import torch
    W, H = 128, 72
    tex = {k: assets[k] for k in ("placement", "curl", "lowres", "hires")}
    cs = mm.ComputeShader(0, (W, H), placement=tex["placement"], curl=tex["curl"], lowRes=tex["lowres"], hiRes=tex["hires"])
    cs.setFilterMode(mm.MM_FILTER_EXACT)
    imgs = [torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
    ref_prev = np.zeros((H, W, 4), np.float32)
    prev_cam = mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD)
    for frame in range(5):
        cam = mm.host_camera((3.0 * frame, 1, 1 + 2.0 * frame), -np.pi / 2 + 0.002 * frame, -20 * scenes.DEG2RAD)
        sun, sky = mm.host_sky(0.25, 0.25, time=2.0 * frame, pixel_phase=(frame + 1) % 16)       # VulkanApplication.cpp:376,384
        # oracle frame
        ref = oracle.reproject(cam, prev_cam, ref_prev)
        S = oracle.Scene(tex, cam, sun, sky)
        ref, _ = S.march(W, H, mode=oracle.OM_PHASE16, counters=False, out=ref)
        # CUDA frame
        dst, src = imgs[frame % 2], imgs[(frame + 1) % 2]
        cs.bindOutput(dst.data_ptr())
        cs.bindPrevious(src.data_ptr())
        cs.updateUniformBuffers(cam, prev_cam, sky, sun)
        cs.dispatchReproject()
        cs.dispatch(mm.MM_PHASE16)
        cs.synchronize()
        got = dst.cpu().numpy()
        rep = oracle.parity_report(ref, got)
        assert np.array_equal(got[..., 3], ref[..., 3]), frame                  # alpha: decision path + single taps, bit-exact
        assert rep["max_abs_diff_8bit"] <= 1 and rep["frac_within_1"] == 1.0, (frame, rep)
        ref_prev, prev_cam = got.copy(), cam        # feed the CUDA image forward so shading ulps do not accumulate in the comparison
    cs.close()
