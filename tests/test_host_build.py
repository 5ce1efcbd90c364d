"""The product's kernel SOURCE, checked on the CPU: the per-pixel functions of K5 (csrc/reproject_pixel.h) and K6 (csrc/post_chain_pixel.h) are
__host__ __device__; tests/host_build/aux_host.cu compiles them for the host with the product's flags (no contraction) and runs them pixel by
pixel, and the results are compared with the oracle under the same gates as the GPU tests (tests/test_post_chain.py, tests/test_reproject.py):
RGBA32F framebuffers bit for bit, the presented bytes within one step.  Test infrastructure only -- the product never runs on the CPU -- but it
lets `-m "not gpu"` catch a change of the kernels' arithmetic or addressing before a GPU box sees it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_build", "aux_host.cu")
LIB = os.path.join(HERE, "host_build", "libaux_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "project-marshmallow_b200", "csrc")


@pytest.fixture(scope="module")
def hb():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("post_chain_pixel.h", "reproject_pixel.h", "curl_noise_pixel.h", "curl_table.h", "noise_volume_pixel.h", "tonemap_pixel.h", "common.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                        "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-diag-suppress", "177", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "--cudart", "static",
                        "-o", LIB, SRC], check=True)
    lib = C.CDLL(LIB)
    f, i, p = C.c_float, C.c_int, C.c_void_p
    lib.hb_god_ray.argtypes = [p, i, i, f, f, f, p]
    lib.hb_radial_blur.argtypes = [p, i, i, f, f, f, p, p]
    lib.hb_present.argtypes = [p, i, i, i, p]
    lib.hb_post_chain.argtypes = [p, i, i, f, f, f, p, i, p]
    lib.hb_reproject.argtypes = [p, p, p, i, i, p]
    lib.hb_curl_noise.argtypes = [p]
    lib.hb_noise_volumes.argtypes = [C.c_uint32, i, i, p, p]
    lib.hb_tonemap.argtypes = [p, C.c_size_t, p]
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _sun_uniforms(oracle, cam, sun):
    """what capi.cu's post_params hands the kernels: the sun's screen position, sun.direction.y, sun.color.xyz * sun.intensity (binary32 products)"""
    xy = oracle.sun_screen_position(cam, sun)
    rgb = (sun[8:11].astype(np.float32) * np.float32(sun[28])).astype(np.float32)
    return float(xy[0]), float(xy[1]), float(sun[5]), np.ascontiguousarray(rgb)


@pytest.mark.parametrize("name,W,H,over", [("C1", 160, 90, {}), ("C3", 200, 113, {}), ("C1", 97, 61, {"yaw": -1.2, "pitch": -0.5}),
                                           ("C1", 64, 36, {"elevation": 0.75}), ("C1", 1, 1, {}), ("C1", 33, 7, {})])
def test_post_chain_source_matches_the_oracle(hb, mm, oracle, assets, name, W, H, over):
    sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
    img, _ = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(W, H, counters=False)
    cam, sun = sc["cam"], sc["sun"]
    sx, sy, sdy, rgb = _sun_uniforms(oracle, cam, sun)
    ref1 = oracle.god_ray(cam, sun, img)
    ref2 = oracle.radial_blur(cam, sun, ref1)
    ref3 = oracle.tonemap_present(ref2)
    g1, g2 = np.empty_like(img), np.empty_like(img)
    g3, gf = np.empty((H, W, 4), np.uint8), np.empty((H, W, 4), np.uint8)
    assert hb.hb_god_ray(_ptr(img), W, H, sx, sy, sdy, _ptr(g1)) == 0
    assert hb.hb_radial_blur(_ptr(g1), W, H, sx, sy, sdy, _ptr(rgb), _ptr(g2)) == 0
    assert hb.hb_present(_ptr(g2), W, H, 0, _ptr(g3)) == 0
    assert hb.hb_post_chain(_ptr(img), W, H, sx, sy, sdy, _ptr(rgb), 0, _ptr(gf)) == 0
    assert np.array_equal(g1.view(np.uint32), ref1.view(np.uint32)), "god-ray framebuffer differs from the oracle"
    assert np.array_equal(g2.view(np.uint32), ref2.view(np.uint32)), "radial-blur framebuffer differs from the oracle"
    assert np.abs(g3.astype(int) - ref3.astype(int)).max() <= 1 and (g3 == ref3).mean() > 0.99
    assert np.array_equal(gf, g3), "the fused chain must produce the bytes of the three passes"


def test_present_source_bgra_swaps_red_and_blue(hb):
    rng = np.random.default_rng(3)
    src = (rng.random((9, 17, 4), dtype=np.float32) * 20).astype(np.float32)
    a, b = np.empty((9, 17, 4), np.uint8), np.empty((9, 17, 4), np.uint8)
    assert hb.hb_present(_ptr(src), 17, 9, 0, _ptr(a)) == 0 and hb.hb_present(_ptr(src), 17, 9, 1, _ptr(b)) == 0
    assert np.array_equal(a[..., [2, 1, 0, 3]], b)


@pytest.mark.parametrize("W,H", [(96, 54), (201, 113), (33, 7), (1, 1)])
@pytest.mark.parametrize("moving", [True, False])
def test_reproject_source_is_bit_exact(hb, mm, oracle, W, H, moving):
    rng = np.random.default_rng(1)
    src = (rng.random((H, W, 4), dtype=np.float32) * 40).astype(np.float32)
    prev = mm.host_camera((0, 1, 1), -np.pi / 2, -20 * scenes.DEG2RAD)
    cur = mm.host_camera((30.0, 3.0, -19.0), -np.pi / 2 + 0.01, -20 * scenes.DEG2RAD - 0.004) if moving else prev
    want = oracle.reproject(cur, prev, src)
    got = np.zeros_like(src)
    cur32, prev32 = np.ascontiguousarray(cur, np.float32), np.ascontiguousarray(prev, np.float32)
    assert hb.hb_reproject(_ptr(cur32), _ptr(prev32), _ptr(src), W, H, _ptr(got)) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_curl_noise_source_reproduces_the_shipped_texture(hb, assets, oracle):
    """K2: the kernel source of curl_noise.cu, run on the CPU, equals Textures/CurlNoiseFBM.tga -- the one golden vector the reference ships for this path
     -- byte for byte, like the GPU build (tests/test_noise_gpu.py) and the oracle's restatement."""
    got = np.zeros((128, 128, 4), np.uint8)
    assert hb.hb_curl_noise(_ptr(got)) == 0
    assert np.array_equal(got, assets["curl"])
    assert np.array_equal(got, oracle.generate_curl_noise())


@pytest.mark.parametrize("seed", [0, 12345])
def test_noise_volume_source_equals_the_cpu_statement(hb, oracle, seed):
    """K3 (our own generator; parity unpinned, no reference arithmetic exists): the kernel source on the CPU equals oracle/noise_volume_oracle.c byte for byte --
    the 32^3 volume whole, every 21st slice of the 128^3 one."""
    low, hi = oracle.build_noise_volumes(seed)
    got_low, got_hi = np.zeros_like(low), np.zeros_like(hi)
    assert hb.hb_noise_volumes(seed, 3, 21, _ptr(got_low), _ptr(got_hi)) == 0
    assert np.array_equal(got_hi, hi)
    assert np.array_equal(got_low[3::21], low[3::21])


def test_tonemap_source_is_the_parity_gates_map(hb, oracle):
    """K4 (tonemap.frag:11-28 without the vignette): the kernel source on the CPU against om_tonemap_rgba8 -- here both sides call the same C library's powf, so
    the bytes are equal (on the GPU CUDA's powf may move a value by one step)."""
    rng = np.random.default_rng(9)
    src = np.concatenate([rng.random((4000, 4), dtype=np.float32) * 60.0, rng.random((4000, 4), dtype=np.float32) * 2.0 - 0.5,
                          np.array([[0, 0, 0, 0], [50.2, 50.2, 50.2, 1], [1e9, -1.0, np.nan, 2.0]], np.float32)]).astype(np.float32)
    got = np.zeros((len(src), 4), np.uint8)
    assert hb.hb_tonemap(_ptr(src), len(src), _ptr(got)) == 0
    want = oracle.tonemap_rgba8(src.reshape(1, -1, 4)).reshape(-1, 4)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and (got == want).mean() > 0.999
