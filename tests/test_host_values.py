"""Host-side value producers (mm_host_sky / mm_host_camera, CPU only) against hand-derived values of the
reference formulas (SkyManager.cpp:15-70, camera.cpp:27-39,179-195, camera.h:72)."""
import os

import numpy as np
import pytest


def test_sun_at_zenith(mm):
    sun, sky = mm.host_sky(0.25, 0.25)
    d = sun[4:7]
    assert d[1] == pytest.approx(1.0, abs=1e-6) and abs(d[0]) < 1e-6 and abs(d[2]) < 1e-6
    assert sun[28] == pytest.approx(1000 * (1 - np.exp(-(1.6110731557 - 0.0) / 1.5)), rel=1e-5)      # 658.4
    assert np.allclose(sun[8:11], 1.0)                         # white at dir.y >= 1/13
    assert np.allclose(sun[0:3], 400000.0 * d, rtol=1e-6) and sun[3] == 1.0
    assert np.allclose(sun[16:19], d)                          # directionBasis column 1 = direction (CC:324)
    basis = sun[12:24].reshape(3, 4)[:, :3]
    assert np.allclose(basis @ basis.T, np.eye(3), atol=1e-5)  # orthonormal TBN
    assert sky[12] == pytest.approx(0.8) and np.allclose(sky[8:11], [1, 0.05, 1]) and sky[11] == 0.0
    # betaR = RAYLEIGH_TOTAL * (rayleigh - 1 + sunFade), sunFade = 1 - clamp(1 - exp(400000/450000), 0, 1) = 1
    assert np.allclose(sky[0:3], np.array([5.804542996261093e-6, 1.3562911419845635e-5, 3.0265902468824876e-5]) * 2.0, rtol=1e-5)
    assert np.allclose(sky[4:7], 0.434 * (0.2 * 10 * 10e-18) * np.array([1.839991851443397, 2.779802391966052, 4.079047954386109]) * 0.005, rtol=1e-5)


def test_low_sun_colour_and_intensity(mm):
    sun, _ = mm.host_sky(0.008, 0.25)                          # config C3
    dy = sun[5]
    assert dy == pytest.approx(np.sin(2 * np.pi * 0.008), rel=1e-3)
    t = min(max(dy * 13, 0), 1)
    assert np.allclose(sun[8:11], (1 - t) * np.array([2.0, 0.33922, 0.0431]) + t, rtol=1e-5)
    assert sun[28] == pytest.approx(1000 * max(0, 1 - np.exp(-(1.6110731557 - np.arccos(dy)) / 1.5)), rel=1e-4)
    assert sun[6] > 0.99                                       # towards +z


def test_night_sun(mm):
    sun, _ = mm.host_sky(0.75, 0.25)
    assert sun[5] < 0 and sun[28] == 2.0 and np.allclose(sun[8:11], [0.8, 0.9, 1.0])
    assert sun[17] == pytest.approx(-sun[5], abs=1e-6)          # basis negated below the horizon
    assert sun[1] > 0                                          # location flipped above the horizon


def test_pixel_phase_and_time(mm):
    sun, sky = mm.host_sky(0.25, 0.25, time=12.5, pixel_phase=21)
    assert sun[11] == 5.0 and sky[11] == 12.5                  # (a+1)%16 bookkeeping lives in sun.color.a (VA:384)


def test_camera_block(mm):
    cam = mm.host_camera((0, 1, 1), -np.pi / 2, -20 * 0.01745)
    view = cam[:16].reshape(4, 4).T                            # row-major view of the column-major block
    R = view[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
    fwd = R[2]                                                 # third row = m_forward; the camera looks along -forward
    assert np.allclose(-fwd, [0, np.sin(20 * 0.01745), np.cos(20 * 0.01745)], atol=1e-6)
    assert np.allclose(cam[32:36], [0, 1, 1, 1])
    assert cam[36] == pytest.approx(1920 / 1080) and cam[37] == pytest.approx(np.tan(0.5 * 0.01745 * 45), rel=1e-6)
    assert np.allclose(view[:3, 3], -R @ np.array([0, 1, 1]), atol=1e-6)


def test_block_rows_run_most_expensive_first_and_the_horizon_row_is_not_mistaken_for_free(mm):
    """mm_dispatch's cost order (host-only planner): rays just above the horizon are the longest of the frame, rows entirely
    below it are free.  The block row that STRADDLES the horizon must run first -- judged by its middle ray it looked free and
    ran last, which cost the rank owning it 15 % of its share of a 4K frame."""
    import scenes
    cfg = scenes.CONFIGS["C3"]
    cam = mm.host_camera(cfg["pos"], cfg["yaw"], cfg["pitch"], 45.0, 1920.0 / 1080.0)
    H = 2160

    def elevation(py):
        spy = 2.0 * py / H - 1.0
        d = np.array([-cam[2] - spy * cam[37] * cam[1], -cam[6] - spy * cam[37] * cam[5], -cam[10] - spy * cam[37] * cam[9]], np.float64)
        return d[1] / np.linalg.norm(d)
    horizon = next(py for py in range(H) if elevation(py) < 0)            # first row below the horizon
    for rank in range(8):
        rows = mm.multigpu.owned_rows(H, rank, 8, 8)
        order = mm.plan_block_rows(cam, H, mm.MM_FULL, rank, 8, 8, 8)
        assert sorted(order.tolist()) == list(range(len(rows) // 8))
        tops = np.array([rows[8 * b] for b in order])                     # first image row of each block row, in execution order
        above = tops < horizon
        assert above[: above.sum()].all(), "a block row with rays above the horizon was scheduled after a free one"
        assert (np.diff(tops[: above.sum()]) < 0).all(), "above the horizon: nearest to the horizon (most expensive) first"
    # rank 0 of 8 owns the straddling block row at 4K (rows 1536-1543 around the horizon row): it is the FIRST to run
    rows0 = mm.multigpu.owned_rows(H, 0, 8, 8)
    first = mm.plan_block_rows(cam, H, mm.MM_FULL, 0, 8, 8, 8)[0]
    assert rows0[8 * first] < horizon <= rows0[8 * first + 7]


REF_HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_host.so")


@pytest.mark.skipif(not os.path.exists(REF_HOST), reason="oracle/_ref/libref_host.so not built (reference tree absent)")
def test_host_sky_equals_the_reference_skymanager(mm):
    """mm_host_sky against the reference's OWN SkyManager.cpp, compiled verbatim from /root/reference (oracle/Makefile; a stub header
    stands in for the Vulkan names SkyManager.h mentions): both uniform blocks byte for byte -- sun position, basis (incl. the
    night-time negation), colour ramp, intensity, Rayleigh / Mie coefficients, wind and time -- over the app's own values and 2000
    random (elevation, azimuth, wind, time, phase) tuples.  The reference never initialises its `turbidity` member; the shim
    constructs the object on storage holding the intended value (oracle/ref_host_shim.cpp)."""
    import ctypes as C
    ref = C.CDLL(REF_HOST)
    ref.ref_host_sky.argtypes = [C.c_float] * 3 + [C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    cases = [(0.25, 0.25, (1, 0.05, 1), 0.0, 0), (0.008, 0.25, (1, 0.05, 1), 0.0, 3), (0.75, 0.25, (0.7, 0.05, -1.3), 123.5, 15),
             (0.5, 0.5, (0, 0, 0), 1.0, 1), (0.0, 0.25, (1, 1, 1), 2.0, 2), (0.5, 0.25, (1, 0.05, 1), 0.0, 0)]
    cases += [(float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), tuple(float(x) for x in rng.uniform(-2, 2, 3)),
               float(rng.uniform(0, 1000)), int(rng.integers(0, 16))) for _ in range(2000)]
    for e, a, w, t, phase in cases:
        for turbidity in (10.0, 3.5):
            sun, sky = mm.host_sky(e, a, w, t, phase, turbidity)
            rsun, rsky, wind = np.zeros(29, np.float32), np.zeros(13, np.float32), np.asarray(w, np.float32)
            assert ref.ref_host_sky(e, a, turbidity, wind.ctypes.data, t, phase, rsun.ctypes.data, rsky.ctypes.data) == 0
            assert np.array_equal(sun.view(np.uint32), rsun.view(np.uint32)), (e, a, sun, rsun)
            assert np.array_equal(sky.view(np.uint32), rsky.view(np.uint32)), (e, a, sky, rsky)


@pytest.mark.skipif(not os.path.exists(REF_HOST), reason="oracle/_ref/libref_host.so not built (reference tree absent)")
def test_host_camera_equals_the_reference_camera(mm):
    """mm_host_camera against the reference's OWN Camera class (camera.h / camera.cpp compiled from /root/reference), driven to a
    yaw / pitch the way the app's mouse handler does (camera.cpp:147-155) and read out exactly as VulkanApplication.cpp:362-368 fills
    UniformCameraObject: view (incl. the signs of its zeros), glm::perspective with the y flip, position, aspect, tan(fov/2) --
    all 160 bytes."""
    import ctypes as C
    ref = C.CDLL(REF_HOST)
    ref.ref_host_camera.argtypes = [C.c_void_p] + [C.c_float] * 5 + [C.c_void_p]
    rng = np.random.default_rng(0)
    cases = [((0, 1, 1), -np.pi / 2, -20 * 0.01745, 45.0), ((0, 1, 1), -np.pi / 2, -10 * 0.01745, 45.0), ((3, 1, 2), -np.pi / 2 + 0.01, -0.3, 45.0),
             ((0, 0, 0), 0.0, 0.0, 45.0), ((1, 2, 3), 1.0, 1.5607, 60.0)]
    cases += [(tuple(float(x) for x in rng.uniform(-500, 500, 3)), float(rng.uniform(-np.pi, np.pi)), float(rng.uniform(-1.5, 1.5)),
               float(rng.uniform(20, 90))) for _ in range(3000)]
    for pos, yaw, pitch, fov in cases:
        cam = mm.host_camera(pos, yaw, pitch, fov, np.float32(1920.0) / np.float32(1080.0))
        want, p = np.zeros(40, np.float32), np.asarray(pos, np.float32)
        assert ref.ref_host_camera(p.ctypes.data, yaw, pitch, fov, 1920.0, 1080.0, want.ctypes.data) == 0
        assert np.array_equal(cam.view(np.uint32), want.view(np.uint32)), (pos, yaw, pitch, fov)


def test_block_row_plans_are_permutations_for_every_mode_and_block_height(mm):
    import scenes
    rng = np.random.default_rng(3)
    for _ in range(200):
        cam = mm.host_camera(tuple(rng.uniform(-100, 100, 3)), float(rng.uniform(-3, 3)), float(rng.uniform(-1.4, 1.4)), 45.0, 1920.0 / 1080.0)
        H = int(rng.integers(1, 2200))
        stride = int(rng.integers(1, 9))
        begin = int(rng.integers(0, stride))
        block = int(rng.choice([1, 2, 4, 8, 16]))
        bh = int(rng.choice([2, 4, 8]))
        for mode in (mm.MM_FULL, mm.MM_FULL | mm.MM_ROWS_SNAKE, mm.MM_PHASE16):
            order = mm.plan_block_rows(cam, H, mode, begin, stride, block, bh)
            owned = (H + 3) // 4 if mode == mm.MM_PHASE16 else len(mm.multigpu.owned_rows(H, begin, stride, block, bool(mode & mm.MM_ROWS_SNAKE)))
            if mode != mm.MM_PHASE16:
                owned = -(-owned // block) * block          # the kernel enumerates whole row blocks of the partition
            assert sorted(order.tolist()) == list(range(-(-owned // bh))), (H, stride, begin, block, bh, mode)


def test_config_files_match_the_scene_definitions(mm, assets):
    """configs/*.json (tools/make_configs.py) carry the same frozen inputs as tests/scenes.py, down to the uniform blocks' floats."""
    import json
    import scenes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in scenes.CONFIGS:
        doc = json.load(open(os.path.join(root, "configs", name + ".json")))
        sc = scenes.make_scene(mm, name, assets)
        assert (doc["width"], doc["height"]) == (sc["W"], sc["H"])
        blocks = doc["uniform_blocks_f32"]
        for key, arr in (("UniformCameraObject_160B", sc["cam"]), ("UniformSunObject_116B", sc["sun"]), ("UniformSkyObject_52B", sc["sky"])):
            assert np.array_equal(np.asarray(blocks[key], np.float32), arr), (name, key)
