"""The frozen synthetic inputs of SURVEY.md section 8(d): configs C1..C5 (test/bench infrastructure).

No RNG is involved in the march: a frame is a pure function of the three uniform blocks and the four
textures.  Uniform values come from the library's host-side SkyManager/Camera mirrors
(mm_host_sky / mm_host_camera, CPU only).
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "tests", "golden", "assets")
DEG2RAD = 0.01745  # camera.h:15


def load_assets():
    out = {}
    for key, name in (("placement", "CloudPlacement"), ("curl", "CurlNoiseFBM"), ("lowres", "lowResCloudShape"), ("hires", "hiResCloudShape")):
        out[key] = np.load(os.path.join(ASSETS, name + ".npz"))["rgba8"]
    out["manifest"] = json.load(open(os.path.join(ASSETS, "MANIFEST.json")))
    return out


def shipped_night_sky():
    """The star map the application binds, Textures/NightSky/nightSky_noOrange.png (VulkanApplication.cpp:255-256): 1920x1080 RGBA8."""
    return np.load(os.path.join(ASSETS, "nightSky_noOrange.npz"))["rgba8"]


def constant_placement(r, b, size=64):
    """BASELINE's 'coverage' knob realised as a placement texture (SURVEY finding 3): R = coverage, B = cloud type."""
    t = np.zeros((size, size, 4), np.uint8)
    t[..., 0], t[..., 2], t[..., 3] = r, b, 255
    return t


def synthetic_night_sky(w=96, h=64):
    """A small non-power-of-two star map for the night branch (the shipped 1920x1080 nightSky_noOrange.png is 8 MB
    decoded and is not needed to exercise CC:365-384; REPEAT wrap on non-power-of-two extents is)."""
    rng = np.random.default_rng(42)
    t = np.zeros((h, w, 4), np.uint8)
    t[..., :3] = rng.integers(0, 40, (h, w, 3))
    stars = rng.random((h, w)) > 0.97
    t[stars, :3] = rng.integers(150, 256, (int(stars.sum()), 3))
    t[..., 3] = 255
    return t


CONFIGS = {
    # name: (W, H, camera pos, yaw, pitch(rad, negative looks up), elevation, azimuth, wind xyz, time, placement)
    "C1": dict(W=320, H=180, pos=(0.0, 1.0, 1.0), yaw=-np.pi / 2, pitch=-20 * DEG2RAD, elevation=0.25, azimuth=0.25,
               wind=(1.0, 0.05, 1.0), time=0.0, placement="shipped"),
    "C2": dict(W=1920, H=1080, pos=(0.0, 1.0, 1.0), yaw=-np.pi / 2, pitch=-20 * DEG2RAD, elevation=0.25, azimuth=0.25,
               wind=(1.0, 0.05, 1.0), time=0.0, placement="shipped"),
    "C2b": dict(W=1920, H=1080, pos=(0.0, 1.0, 1.0), yaw=-np.pi / 2, pitch=-20 * DEG2RAD, elevation=0.25, azimuth=0.25,
                wind=(1.0, 0.05, 1.0), time=0.0, placement=(128, 255)),
    "C3": dict(W=3840, H=2160, pos=(0.0, 1.0, 1.0), yaw=-np.pi / 2, pitch=-10 * DEG2RAD, elevation=0.008, azimuth=0.25,
               wind=(1.0, 0.05, 1.0), time=0.0, placement="shipped"),
    "C5": dict(W=7680, H=4320, pos=(0.0, 1.0, 1.0), yaw=-np.pi / 2, pitch=-20 * DEG2RAD, elevation=0.25, azimuth=0.25,
               wind=(1.0, 0.05, 1.0), time=0.0, placement=(230, 255)),
    "C5b": dict(W=7680, H=4320, pos=(0.0, 1.0, 1.0), yaw=-np.pi / 2, pitch=-20 * DEG2RAD, elevation=0.25, azimuth=0.25,
                wind=(1.0, 0.05, 1.0), time=0.0, placement=(128, 128)),
}


def make_scene(mm, name, assets, W=None, H=None, time=None, pixel_phase=0, **over):
    """-> dict(W, H, cam, sun, sky, textures{placement,curl,lowres,hires}).  W/H override keeps the same view
    (aspect stays 16:9 as in camera.h:31), so any config can be rendered at a resolution the oracle finishes quickly."""
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    if W:
        cfg["W"], cfg["H"] = W, H
    if time is not None:
        cfg["time"] = time
    cam = mm.host_camera(cfg["pos"], cfg["yaw"], cfg["pitch"], 45.0, 1920.0 / 1080.0)
    sun, sky = mm.host_sky(cfg["elevation"], cfg["azimuth"], cfg["wind"], cfg["time"], pixel_phase)
    tex = {k: assets[k] for k in ("curl", "lowres", "hires")}
    tex["placement"] = assets["placement"] if cfg["placement"] == "shipped" else constant_placement(*cfg["placement"])
    return dict(name=name, W=cfg["W"], H=cfg["H"], cam=cam, sun=sun, sky=sky, textures=tex)


def scene_from_config(name, assets):
    """The same scene dict as make_scene, built from configs/<name>.json alone (frozen uniform blocks; no product library involved):
    what the CPU reference arm of bench.py binds."""
    doc = json.load(open(os.path.join(ROOT, "configs", name + ".json")))
    blocks = doc["uniform_blocks_f32"]
    tex = {k: assets[k] for k in ("curl", "lowres", "hires")}
    pl = doc["placement"]
    tex["placement"] = assets["placement"] if isinstance(pl, str) else constant_placement(pl["constant_R_coverage"], pl["constant_B_type"])
    return dict(name=name, W=doc["width"], H=doc["height"], cam=np.asarray(blocks["UniformCameraObject_160B"], np.float32),
                sun=np.asarray(blocks["UniformSunObject_116B"], np.float32), sky=np.asarray(blocks["UniformSkyObject_52B"], np.float32), textures=tex)
