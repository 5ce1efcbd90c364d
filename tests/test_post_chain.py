"""Post chain: god-ray.frag -> radialBlur.frag -> tonemap.frag.

CPU half: known answers of the oracle restatement (oracle/post_chain_oracle.c), derived by hand from the GLSL.
GPU half: each CUDA pass against the oracle -- RGBA32F framebuffers bit for bit (only + - * / sqrt and fused lerps are
involved), the UNORM8 present within 1 (pow) -- and the fused two-kernel chain against the three passes, byte for byte.
"""
import numpy as np
import pytest

import scenes


def _frame(mm, assets, oracle, name="C1", W=160, H=90, **over):
    """A cloud image to feed the chain: the oracle's march of a small frame (alpha carries the sun disk / ambient seed)."""
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    img, _ = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(W, H, counters=False)
    return sc, img


def test_sun_screen_position_is_the_projected_sun(mm, oracle, assets):
    sc = scenes.make_scene(mm, "C1", assets)
    V = sc["cam"][:16].reshape(4, 4).T.astype(np.float64)          # column-major blocks
    P = sc["cam"][16:32].reshape(4, 4).T.astype(np.float64)
    clip = P @ V @ sc["sun"][:4].astype(np.float64)
    want = clip[:2] / clip[3]
    got = oracle.sun_screen_position(sc["cam"], sc["sun"])
    assert np.allclose(got, want, rtol=2e-5, atol=2e-5), (got, want)
    # C1 looks toward the sun's azimuth, 20 degrees up, with the sun at the zenith: the sun is above the top edge
    # (framebuffer y grows downward, so "above" is y < -1) and horizontally centred
    assert abs(got[0]) < 1e-3 and got[1] < -1.0


def test_god_ray_of_constant_alpha_is_the_geometric_series(oracle, mm, assets):
    sc = scenes.make_scene(mm, "C1", assets)
    W, H, a = 40, 24, np.float32(0.6)
    src = np.zeros((H, W, 4), np.float32)
    src[..., :3] = np.random.default_rng(1).random((H, W, 3), dtype=np.float32)
    src[..., 3] = a
    out = oracle.god_ray(sc["cam"], sc["sun"], src)
    assert np.array_equal(out[..., :3], src[..., :3])              # rgb passes through (god-ray.frag:75)
    acc, decay = np.float32(a * np.float32(0.5)), np.float32(1.0)
    for _ in range(8):                                             # god-ray.frag:58-73 with every tap equal to a
        acc = np.float32(acc + np.float32(np.float32(a * np.float32(0.5)) * np.float32(np.float32(0.125) * decay)))
        decay = np.float32(decay * np.float32(0.99))
    assert np.array_equal(out[..., 3], np.full((H, W), np.float32(acc * np.float32(0.9))))


def test_night_frames_pass_colour_through_with_alpha_one(oracle, mm, assets):
    sc = scenes.make_scene(mm, "C1", assets, elevation=0.75)       # sun below the horizon: sun.direction.y < 0
    assert sc["sun"][5] < 0
    src = np.random.default_rng(2).random((18, 32, 4), dtype=np.float32)
    for fn in (oracle.god_ray, oracle.radial_blur):
        out = fn(sc["cam"], sc["sun"], src)
        assert np.array_equal(out[..., :3], src[..., :3]) and (out[..., 3] == 1.0).all()


def test_radial_blur_of_constant_alpha_adds_sun_light(oracle, mm, assets):
    sc = scenes.make_scene(mm, "C1", assets)
    W, H, a = 48, 27, 0.25
    src = np.zeros((H, W, 4), np.float32)
    src[..., :3] = 0.5
    src[..., 3] = a
    out = oracle.radial_blur(sc["cam"], sc["sun"], src)
    light = sc["sun"][8:11].astype(np.float64) * float(sc["sun"][28]) * (1.1 * a)      # radialBlur.frag:57-62
    assert np.allclose(out[..., :3], light + 0.25, rtol=1e-5)
    assert (out[..., 3] == 1.0).all()


def test_present_is_the_uncharted2_curve_with_vignette(oracle):
    rng = np.random.default_rng(3)
    W, H = 64, 36
    src = (rng.random((H, W, 4)) * np.float64([60, 60, 60, 1])).astype(np.float32)
    src[0, 0, :3] = 0.0
    got = oracle.tonemap_present(src)
    x = 0.7 * src[..., :3].astype(np.float64)
    uc2 = lambda v: ((v * (0.15 * v + 0.05) + 0.004) / (v * (0.15 * v + 0.5) + 0.06)) - 0.02 / 0.3
    col = np.power(np.maximum(uc2(x) / uc2(50.2), 0.0), 1 / 2.2)
    u = (np.arange(W) + 0.5) / W - 0.5
    v = (np.arange(H) + 0.5) / H - 0.5
    vig = (u[None, :] ** 2 + v[:, None] ** 2)[..., None]
    col = col * (1 - vig) + np.float64([0.1, 0.05, 0.13]) * vig
    want = np.floor(255 * np.clip(col, 0, 1) + 0.5)
    assert np.abs(got[..., :3].astype(int) - want).max() <= 1
    assert (got[..., 3] == 255).all()
    bgra = oracle.tonemap_present(src, bgra=True)
    assert np.array_equal(bgra[..., [2, 1, 0, 3]], got)


# ---------------------------------------------------------------------------------------------------------------- GPU
def _dev(t):
    import torch
    return torch.from_numpy(np.ascontiguousarray(t)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name,W,H,over", [("C1", 160, 90, {}), ("C3", 200, 113, {}), ("C1", 97, 61, {"yaw": -1.2, "pitch": -0.5}),
                                           ("C1", 64, 36, {"elevation": 0.75}), ("C1", 1, 1, {}), ("C1", 33, 7, {})])
def test_cuda_passes_match_the_oracle_and_the_fused_chain_matches_the_passes(mm, oracle, assets, name, W, H, over):
    import torch
    sc, img = _frame(mm, assets, oracle, name, W, H, **over)
    cam, sun = sc["cam"], sc["sun"]
    ref1 = oracle.god_ray(cam, sun, img)
    ref2 = oracle.radial_blur(cam, sun, ref1)
    ref3 = oracle.tonemap_present(ref2)
    cs = mm.ComputeShader(0, (W, H))
    src = _dev(img)
    fb1, fb2 = torch.empty_like(src), torch.empty_like(src)
    out8, fused8 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda"), torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    cs.godRay(cam, sun, src.data_ptr(), fb1.data_ptr())
    cs.radialBlur(cam, sun, fb1.data_ptr(), fb2.data_ptr())
    cs.tonemapPresent(fb2.data_ptr(), out8.data_ptr())
    cs.postChain(cam, sun, src.data_ptr(), fused8.data_ptr())
    cs.synchronize()
    g1, g2, g3, gf = fb1.cpu().numpy(), fb2.cpu().numpy(), out8.cpu().numpy(), fused8.cpu().numpy()
    cs.close()
    assert np.array_equal(g1.view(np.uint32), ref1.view(np.uint32)), "god-ray framebuffer differs from the oracle"
    assert np.array_equal(g2.view(np.uint32), ref2.view(np.uint32)), "radial-blur framebuffer differs from the oracle"
    assert np.abs(g3.astype(int) - ref3.astype(int)).max() <= 1 and (g3 == ref3).mean() > 0.99
    assert np.array_equal(gf, g3), "the fused chain must produce the bytes of the three passes"


@pytest.mark.gpu
def test_post_chain_bgra_pitch_and_errors(mm, oracle, assets):
    import torch
    W, H = 70, 40
    sc, img = _frame(mm, assets, oracle, "C1", W, H)
    cs = mm.ComputeShader(0, (W, H))
    src = _dev(img)
    rgba = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    wide = torch.full((H, W + 10, 4), 7, dtype=torch.uint8, device="cuda")           # padded rows: pitch > 4*W
    cs.postChain(sc["cam"], sc["sun"], src.data_ptr(), rgba.data_ptr())
    cs.postChain(sc["cam"], sc["sun"], src.data_ptr(), wide.data_ptr(), dst_pitch=(W + 10) * 4, bgra=True)
    cs.synchronize()
    a, b = rgba.cpu().numpy(), wide.cpu().numpy()
    assert np.array_equal(b[:, :W][..., [2, 1, 0, 3]], a) and (b[:, W:] == 7).all()
    with pytest.raises(mm.MarshmallowError):
        cs.godRay(sc["cam"], sc["sun"], src.data_ptr(), src.data_ptr())                # in place is a hazard: rejected
    with pytest.raises(mm.MarshmallowError):
        cs.postChain(sc["cam"], sc["sun"], src.data_ptr(), rgba.data_ptr(), dst_pitch=W * 4 - 4)
    with pytest.raises(mm.MarshmallowError):
        cs.radialBlur(sc["cam"], sc["sun"], src.data_ptr() + 4, rgba.data_ptr())       # misaligned source
    cs.close()


@pytest.mark.gpu
def test_full_frame_march_then_post_chain_at_1080p(mm, oracle, assets):
    """The headless frame the chain exists for: cloud march (hardware sampler) -> post chain, 1920x1080, against the
    oracle running the same four stages on the CPU."""
    import torch
    sc = scenes.make_scene(mm, "C2", assets)
    W, H = sc["W"], sc["H"]
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    cs.setFilterMode(mm.MM_FILTER_HW)
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    out8 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    cs.dispatch()
    cs.postChain(sc["cam"], sc["sun"], cs.out_ptr, out8.data_ptr(), src_pitch=cs.out_pitch)
    cs.synchronize()
    got = out8.cpu().numpy()
    hdr = cs.readOutput()
    cs.close()
    ref = oracle.tonemap_present(oracle.radial_blur(sc["cam"], sc["sun"], oracle.god_ray(sc["cam"], sc["sun"], hdr)))
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.99
    march_ref, _ = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], filter_mode=oracle.OM_FILTER_TEXUNIT).march(W, H, counters=False)
    ref_all = oracle.tonemap_present(oracle.radial_blur(sc["cam"], sc["sun"], oracle.god_ray(sc["cam"], sc["sun"], march_ref)))
    d = np.abs(got.astype(int) - ref_all.astype(int)).max(axis=-1)
    assert d.max() <= 2 and (d <= 1).mean() >= 0.999
