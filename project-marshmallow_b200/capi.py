"""ctypes binding of include/marshmallow.h.  Plain pointers and sizes only.

There is no CPU fallback: if libmarshmallow_b200.so is missing, load_library() raises, and every
compute entry point returns MM_ERR_CUDA (raised as MarshmallowError) when no sm_100 GPU is usable.
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "marshmallow.h")

MM_FULL, MM_PHASE16 = 0, 1
MM_ROWS_SNAKE = 0x100
MM_FILTER_EXACT, MM_FILTER_HW, MM_FILTER_HYBRID = 0, 1, 2
MM_SCHED_AUTO, MM_SCHED_STATIC, MM_SCHED_PERSISTENT, MM_SCHED_PACKED = 0, 1, 2, 3
MM_ARITH_IEEE, MM_ARITH_FMA = 0, 1
MM_TEX_PLACEMENT, MM_TEX_NIGHTSKY, MM_TEX_CURL, MM_TEX_LOWRES, MM_TEX_HIRES = range(5)


class MarshmallowError(RuntimeError):
    """Mirrors the reference's error behaviour: every failure is a runtime error (main.cpp:8-14)."""

    def __init__(self, code, msg):
        super().__init__(f"marshmallow error {code}: {msg}")
        self.code = code


def library_path():
    """MM_LIBRARY overrides the in-tree build (used by tools/ to time build variants of the same source)."""
    return os.environ.get("MM_LIBRARY") or os.path.join(HERE, "libmarshmallow_b200.so")


def exported_symbols():
    """Every function name include/marshmallow.h declares with MM_API."""
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MM_API\s+[\w\s\*]+?\b(mm_\w+)\s*\(", text)))


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise MarshmallowError(-2, f"{path} is not built (run `python project-marshmallow_b200/build.py`); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
    fp = C.POINTER(C.c_float)
    sig = {
        "mm_create": (i32, [i32, C.POINTER(vp)]),
        "mm_destroy": (i32, [vp]),
        "mm_last_error": (C.c_char_p, [vp]),
        "mm_version": (C.c_char_p, []),
        "mm_upload_tex2d": (i32, [vp, i32, vp, i32, i32]),
        "mm_upload_tex3d": (i32, [vp, i32, vp, i32, i32, i32]),
        "mm_build_curl_noise": (i32, [vp, vp]),
        "mm_build_noise_volumes": (i32, [vp, u64, vp, vp]),
        "mm_set_uniforms": (i32, [vp, vp, vp, vp, vp]),
        "mm_bind_output_linear": (i32, [vp, vp, sz, i32, i32]),
        "mm_bind_output_external_fd": (i32, [vp, i32, sz, i32, i32]),
        "mm_bind_output_external_buffer_fd": (i32, [vp, i32, sz, sz, sz, i32, i32]),
        "mm_import_semaphore_fd": (i32, [vp, i32, i32, C.POINTER(i32)]),
        "mm_wait_semaphore": (i32, [vp, i32, u64, vp]),
        "mm_signal_semaphore": (i32, [vp, i32, u64, vp]),
        "mm_release_semaphore": (i32, [vp, i32]),
        "mm_enable_peer": (i32, [vp, vp]),
        "mm_alloc_output": (i32, [vp, i32, i32, C.POINTER(vp), C.POINTER(sz)]),
        "mm_set_filter_mode": (i32, [vp, i32]),
        "mm_set_lanes_per_ray": (i32, [vp, i32]),
        "mm_set_scheduler": (i32, [vp, i32, i32]),
        "mm_set_arithmetic": (i32, [vp, i32]),
        "mm_dispatch": (i32, [vp, i32, i32, i32, i32, vp]),
        "mm_dispatch_multi": (i32, [C.POINTER(vp), i32, i32, i32, C.POINTER(vp)]),
        "mm_synchronize": (i32, [vp]),
        "mm_plan_block_rows": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, C.POINTER(i32)]),
        "mm_bind_previous_linear": (i32, [vp, vp, sz]),
        "mm_dispatch_reproject": (i32, [vp, vp]),
        "mm_render_to_host": (i32, [vp, vp, vp, vp, i32, vp]),
        "mm_tonemap_rgba8": (i32, [vp, vp, i32, vp]),
        "mm_host_register": (i32, [vp, vp, sz]),
        "mm_host_unregister": (i32, [vp, vp]),
        "mm_bind_host_mirror": (i32, [vp, vp]),
        "mm_god_ray": (i32, [vp, vp, vp, vp, sz, vp, sz, i32, i32, vp]),
        "mm_radial_blur": (i32, [vp, vp, vp, vp, sz, vp, sz, i32, i32, vp]),
        "mm_tonemap_present": (i32, [vp, vp, sz, vp, sz, i32, i32, i32, vp]),
        "mm_post_chain": (i32, [vp, vp, vp, vp, sz, vp, sz, i32, i32, i32, vp]),
        "mm_cloud_shadow": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "mm_enable_counters": (i32, [vp, i32]),
        "mm_read_counters": (i32, [vp, vp]),
        "mm_read_output": (i32, [vp, vp]),
        "mm_last_kernel_ms": (i32, [vp, fp]),
        "mm_sample": (i32, [vp, i32, i32, vp, i32, vp]),
        "mm_det_pow": (i32, [vp, vp, vp, i32, vp]),
        "mm_measure_tex_peak": (i32, [vp, i32, i32, fp, C.POINTER(C.c_double)]),
        "mm_selftest_div": (i32, [vp, i32, fp, C.POINTER(C.c_uint64)]),
        "mm_alloc_device": (i32, [vp, sz, C.POINTER(vp)]),
        "mm_free_device": (i32, [vp, vp]),
        "mm_ipc_get_handle": (i32, [vp, vp, vp]),
        "mm_ipc_open_handle": (i32, [vp, vp, C.POINTER(vp)]),
        "mm_ipc_close_handle": (i32, [vp, vp]),
        "mm_host_sky": (i32, [C.c_float] * 6 + [vp, C.c_float, i32, vp, vp]),
        "mm_host_camera": (i32, [vp, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def host_sky(elevation, azimuth, wind=(1.0, 0.05, 1.0), time=0.0, pixel_phase=0, turbidity=10.0, rayleigh=2.0,
             mie=0.005, mie_directional=0.8):
    """SkyManager values (SkyManager.cpp:15-70) -> (sun116 bytes, sky52 bytes).  CPU only."""
    lib = load_library()
    sun = np.zeros(29, np.float32)
    sky = np.zeros(13, np.float32)
    w = np.asarray(wind, np.float32)
    rc = lib.mm_host_sky(elevation, azimuth, turbidity, rayleigh, mie, mie_directional, _ptr(w), time, pixel_phase, _ptr(sun), _ptr(sky))
    if rc:
        raise MarshmallowError(rc, "mm_host_sky")
    return sun, sky


def plan_block_rows(cam, H, mode=MM_FULL, row_begin=0, row_stride=1, row_block=1, block_h=8):
    """Execution order of the block rows of a dispatch, most expensive first (host-only; mm_plan_block_rows)."""
    lib = load_library()
    cam = np.ascontiguousarray(cam, np.float32)
    order = np.zeros(4096, np.uint16)
    n = C.c_int()
    rc = lib.mm_plan_block_rows(_ptr(cam), H, mode, row_begin, row_stride, row_block, block_h, _ptr(order), C.byref(n))
    if rc:
        raise MarshmallowError(rc, "mm_plan_block_rows")
    return order[:n.value].copy()


def host_camera(position, yaw, pitch, fov_deg=45.0, aspect=1920.0 / 1080.0):
    """Camera values (camera.cpp:27-39,179-195) -> camera160 bytes.  CPU only."""
    lib = load_library()
    cam = np.zeros(40, np.float32)
    p = np.asarray(position, np.float32)
    rc = lib.mm_host_camera(_ptr(p), yaw, pitch, fov_deg, aspect, _ptr(cam))
    if rc:
        raise MarshmallowError(rc, "mm_host_camera")
    return cam
