"""ctypes binding of oracle/liboracle.so and oracle/_ref/libref_curl.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_CURL_SO = os.path.join(ROOT, "oracle", "_ref", "libref_curl.so")

OM_TEX_PLACEMENT, OM_TEX_NIGHTSKY, OM_TEX_CURL, OM_TEX_LOWRES, OM_TEX_HIRES = range(5)
OM_FILTER_FP32, OM_FILTER_FIX8, OM_FILTER_TEXUNIT = 0, 1, 2
OM_POW_DET, OM_POW_LIBM = 0, 1
OM_FULL, OM_PHASE16 = 0, 1
OM_ARITH_IEEE, OM_ARITH_FMA = 0, 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(ORACLE_SO)
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        l.om_scene_create.restype = vp
        l.om_scene_destroy.argtypes = [vp]
        l.om_scene_set_texture.argtypes = [vp, i32, vp, i32, i32, i32]
        l.om_scene_set_uniforms.argtypes = [vp, vp, vp, vp]
        l.om_scene_set_modes.argtypes = [vp, i32, i32]
        l.om_scene_set_arith.argtypes = [vp, i32]
        l.om_march.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32]
        l.om_sample.argtypes = [vp, i32, i32, vp, i32, vp]
        l.om_tonemap_rgba8.argtypes = [vp, C.c_size_t, vp]
        for name, n in (("om_det_powf", 2), ("om_det_powf_fma", 2), ("om_hgPhase", 2), ("om_remap", 5), ("om_remapClamped", 5),
                        ("om_cloudLayerDensity", 2), ("om_heightBiasCoverage", 2), ("om_curl_hash", 3)):
            fn = getattr(l, name)
            fn.restype, fn.argtypes = f32, [f32] * n
        l.om_curl_hash_index.restype, l.om_curl_hash_index.argtypes = i32, [f32] * 3
        l.om_raySphereIntersection.argtypes = [vp, vp, vp, vp]
        l.om_reproject.argtypes = [vp, vp, vp, i32, i32, vp]
        l.om_cloud_shadow.argtypes = [vp, vp, i32, vp, vp, i32]
        l.om_sun_screen_position.argtypes = [vp, vp, vp]
        l.om_god_ray.argtypes = [vp, vp, vp, i32, i32, vp]
        l.om_radial_blur.argtypes = [vp, vp, vp, i32, i32, vp]
        l.om_tonemap_present.argtypes = [vp, i32, i32, i32, vp]
        l.om_generate_curl_noise.argtypes = [vp]
        l.om_build_noise_volumes.argtypes = [C.c_uint64, vp, vp]
        l.om_noise_hash.restype, l.om_noise_hash.argtypes = C.c_uint32, [C.c_uint32] * 4
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Scene:
    def __init__(self, assets, cam, sun, sky, filter_mode=OM_FILTER_FP32, pow_mode=OM_POW_DET, nightsky=None, arith=OM_ARITH_IEEE):
        self.l = lib()
        self.s = self.l.om_scene_create()
        self._keep = []
        for slot, key in ((OM_TEX_PLACEMENT, "placement"), (OM_TEX_CURL, "curl"), (OM_TEX_LOWRES, "lowres"), (OM_TEX_HIRES, "hires")):
            self.set_texture(slot, assets[key])
        if nightsky is not None:
            self.set_texture(OM_TEX_NIGHTSKY, nightsky)
        self.set_uniforms(cam, sun, sky)
        self.set_modes(filter_mode, pow_mode)
        assert self.l.om_scene_set_arith(self.s, arith) == 0

    def set_texture(self, slot, a):
        a = np.ascontiguousarray(a, np.uint8)
        if a.ndim == 3:
            h, w, _ = a.shape
            d = 1
        else:
            d, h, w, _ = a.shape
        assert self.l.om_scene_set_texture(self.s, slot, _p(a), w, h, d) == 0

    def set_uniforms(self, cam, sun, sky):
        cam, sun, sky = (np.ascontiguousarray(x, np.float32) for x in (cam, sun, sky))
        assert self.l.om_scene_set_uniforms(self.s, _p(cam), _p(sun), _p(sky)) == 0

    def set_modes(self, filter_mode, pow_mode):
        assert self.l.om_scene_set_modes(self.s, filter_mode, pow_mode) == 0

    def march(self, W, H, mode=OM_FULL, row_begin=0, row_stride=1, row_block=1, counters=True, nthreads=0, out=None):
        if out is None:
            out = np.zeros((H, W, 4), np.float32)
        cnt = np.zeros((H, W, 4), np.uint32) if counters else None
        rc = self.l.om_march(self.s, mode, W, H, row_begin, row_stride, row_block, _p(out), _p(cnt), nthreads)
        assert rc == 0, rc
        return out, cnt

    def sample(self, slot, filter_mode, uvw):
        uvw = np.ascontiguousarray(uvw, np.float32).reshape(-1, 3)
        out = np.empty((uvw.shape[0], 4), np.float32)
        assert self.l.om_sample(self.s, slot, filter_mode, _p(uvw), uvw.shape[0], _p(out)) == 0
        return out

    def cloud_shadow(self, positions, want_fetches=False, nthreads=0):
        """model.frag:240-283 for an array of world positions -> accumDensity per point"""
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        out = np.empty(pos.shape[0], np.float32)
        nf = np.empty(pos.shape[0], np.uint32) if want_fetches else None
        assert self.l.om_cloud_shadow(self.s, _p(pos), pos.shape[0], _p(out), _p(nf), nthreads) == 0
        return (out, nf) if want_fetches else out

    def close(self):
        if self.s:
            self.l.om_scene_destroy(self.s)
            self.s = None

    def __del__(self):
        self.close()


def tonemap_rgba8(img):
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty(img.shape, np.uint8)
    lib().om_tonemap_rgba8(_p(img), img.size // 4, _p(out))
    return out


def reproject(cam, cam_prev, src):
    cam, cam_prev = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(cam_prev, np.float32)
    src = np.ascontiguousarray(src, np.float32)
    H, W, _ = src.shape
    dst = np.empty_like(src)
    assert lib().om_reproject(_p(cam), _p(cam_prev), _p(src), W, H, _p(dst)) == 0
    return dst


def sun_screen_position(cam, sun):
    cam, sun = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(sun, np.float32)
    out = np.zeros(2, np.float32)
    lib().om_sun_screen_position(_p(cam), _p(sun), _p(out))
    return out


def _post(fn, cam, sun, src):
    cam, sun = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(sun, np.float32)
    src = np.ascontiguousarray(src, np.float32)
    H, W, _ = src.shape
    dst = np.empty_like(src)
    assert fn(_p(cam), _p(sun), _p(src), W, H, _p(dst)) == 0
    return dst


def god_ray(cam, sun, src):
    return _post(lib().om_god_ray, cam, sun, src)


def radial_blur(cam, sun, src):
    return _post(lib().om_radial_blur, cam, sun, src)


def tonemap_present(src, bgra=False):
    src = np.ascontiguousarray(src, np.float32)
    H, W, _ = src.shape
    dst = np.empty((H, W, 4), np.uint8)
    assert lib().om_tonemap_present(_p(src), W, H, int(bgra), _p(dst)) == 0
    return dst


def generate_curl_noise():
    out = np.zeros((128, 128, 4), np.uint8)
    lib().om_generate_curl_noise(_p(out))
    return out


def build_noise_volumes(seed=0):
    low = np.zeros((128, 128, 128, 4), np.uint8)
    hi = np.zeros((32, 32, 32, 4), np.uint8)
    lib().om_build_noise_volumes(seed, _p(low), _p(hi))
    return low, hi


def have_ref_curl():
    return os.path.exists(REF_CURL_SO)


def ref_generate_curl_noise():
    """Runs the REFERENCE's own GenerateCurlNoise (compiled verbatim into oracle/_ref) and decodes its TGA."""
    from PIL import Image
    l = C.CDLL(REF_CURL_SO)
    l.ref_generate_curl_noise_tga.argtypes = [C.c_char_p]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "curl.tga")
        l.ref_generate_curl_noise_tga(path.encode())
        return np.array(Image.open(path).convert("RGBA"), np.uint8)


def parity_report(ref_img, test_img, ref_cnt=None, test_cnt=None):
    """The parity gate of SURVEY 8d: both RGBA32F images through the same tonemap -> RGBA8."""
    a, b = tonemap_rgba8(ref_img).astype(np.int32), tonemap_rgba8(test_img).astype(np.int32)
    d = np.abs(a - b).max(axis=-1)
    rep = {
        "max_abs_diff_8bit": int(d.max()),
        "frac_within_1": float((d <= 1).mean()),
        "frac_identical_8bit": float((d == 0).mean()),
        "float_identical_frac": float((ref_img == test_img).all(axis=-1).mean()),
        "alpha_identical_frac": float((ref_img[..., 3] == test_img[..., 3]).mean()),
    }
    # raw HDR floats: relative error percentiles over all channels
    rel = (np.abs(ref_img.astype(np.float64) - test_img.astype(np.float64)) / np.maximum(np.abs(ref_img.astype(np.float64)), 1e-6)).ravel()
    rep["rel_err_p50"], rep["rel_err_p99"], rep["rel_err_max"] = (float(x) for x in (np.percentile(rel, 50), np.percentile(rel, 99), rel.max()))
    if ref_cnt is not None and test_cnt is not None:
        rep["branch_flip_pixels"] = int((ref_cnt[..., 0] != test_cnt[..., 0]).sum())
        rep["counter_mismatch_pixels"] = int((ref_cnt != test_cnt).any(axis=-1).sum())
    rep["pass"] = rep["max_abs_diff_8bit"] <= 2 and rep["frac_within_1"] >= 0.999
    return rep
