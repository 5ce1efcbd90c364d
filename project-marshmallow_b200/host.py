"""Python mirror of the reference's ComputeShader interface (Shader.h:286-377) over the C-ABI.

    reference (C++/Vulkan)                                   here
    ComputeShader(device, ..., out, outPrev, placement,      ComputeShader(device, extent, placement,
                  nightSky, curl, lowRes, hiRes)                           nightSky, curl, lowRes, hiRes)
    updateUniformBuffers(cam, camPrev, sky, sun)             updateUniformBuffers(cam, camPrev, sky, sun)
    bindShader(cmdBuf); vkCmdDispatch; vkQueueSubmit         dispatch(mode, stream=...)
    backgroundTexture (VkImage rgba32f)                      output pointer / readOutput()

Errors raise MarshmallowError (the reference throws std::runtime_error).  No torch types cross the
boundary: device pointers and streams are passed as integers.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import MarshmallowError, _ptr


class SkyManager:
    """Value producer mirroring SkyManager (SkyManager.h:51-77): rebuildSkyFromNewSun / setTime / getSun / getSky."""

    def __init__(self, turbidity=10.0):
        self.elevation, self.azimuth = float(np.float32(np.pi) / 4), float(np.float32(np.pi) / 8)
        self.wind = (1.0, 0.05, 1.0)
        self.time = 0.0
        self.turbidity = turbidity
        self.pixel_phase = 0

    def rebuildSkyFromNewSun(self, elevation, azimuth):
        self.elevation, self.azimuth = elevation, azimuth

    def setWindDirection(self, d):
        self.wind = tuple(d)

    def setTime(self, t):
        self.time = t

    def _build(self):
        return capi.host_sky(self.elevation, self.azimuth, self.wind, self.time, self.pixel_phase, self.turbidity)

    def getSun(self):
        return self._build()[0]

    def getSky(self):
        return self._build()[1]


class Camera:
    """Value producer mirroring Camera (camera.h): position + yaw/pitch -> UniformCameraObject bytes."""

    def __init__(self, position, yaw, pitch, fov=45.0, aspect=1920.0 / 1080.0):
        self.position, self.yaw, self.pitch, self.fov, self.aspect = tuple(position), yaw, pitch, fov, aspect

    def getUniform(self):
        return capi.host_camera(self.position, self.yaw, self.pitch, self.fov, self.aspect)


def dispatchMulti(shaders, mode=capi.MM_FULL, row_block=8):
    """One frame sharded over several ComputeShader contexts of this process (mm_dispatch_multi): context i marches partition i."""
    lib = capi.load_library()
    arr = (C.c_void_p * len(shaders))(*[s._ctx.value for s in shaders])
    rc = lib.mm_dispatch_multi(arr, len(shaders), mode, row_block, None)
    if rc:
        bad = next((s for s in shaders if lib.mm_last_error(s._ctx)), shaders[0])
        raise MarshmallowError(rc, lib.mm_last_error(bad._ctx).decode())


class ComputeShader:
    def __init__(self, device, extent, placement=None, nightSky=None, curl=None, lowRes=None, hiRes=None):
        self._lib = capi.load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.mm_create(int(device), C.byref(self._ctx))
        if rc:
            raise MarshmallowError(rc, self._lib.mm_last_error(None).decode())
        self.width, self.height = int(extent[0]), int(extent[1])
        self.out_ptr, self.out_pitch = None, None
        for slot, tex in ((capi.MM_TEX_PLACEMENT, placement), (capi.MM_TEX_NIGHTSKY, nightSky), (capi.MM_TEX_CURL, curl),
                          (capi.MM_TEX_LOWRES, lowRes), (capi.MM_TEX_HIRES, hiRes)):
            if tex is not None:
                self.uploadTexture(slot, tex)

    # ---- plumbing
    def _check(self, rc):
        if rc:
            raise MarshmallowError(rc, self._lib.mm_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            self._lib.mm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- textures (Texture::initFromFile / Texture3D::initFromFile)
    def uploadTexture(self, slot, rgba8):
        a = np.ascontiguousarray(rgba8, np.uint8)
        if a.ndim == 3:
            h, w, _ = a.shape
            self._check(self._lib.mm_upload_tex2d(self._ctx, slot, _ptr(a), w, h))
        else:
            d, h, w, _ = a.shape
            self._check(self._lib.mm_upload_tex3d(self._ctx, slot, _ptr(a), w, h, d))

    def buildCurlNoise(self, want_copy=True):
        out = np.zeros((128, 128, 4), np.uint8) if want_copy else None
        self._check(self._lib.mm_build_curl_noise(self._ctx, _ptr(out) if want_copy else None))
        return out

    def buildNoiseVolumes(self, seed=0, want_copy=True):
        low = np.zeros((128, 128, 128, 4), np.uint8) if want_copy else None
        hi = np.zeros((32, 32, 32, 4), np.uint8) if want_copy else None
        self._check(self._lib.mm_build_noise_volumes(self._ctx, seed, _ptr(low) if want_copy else None, _ptr(hi) if want_copy else None))
        return low, hi

    # ---- output image (descriptor set 0)
    def allocOutput(self):
        p, pitch = C.c_void_p(), C.c_size_t()
        self._check(self._lib.mm_alloc_output(self._ctx, self.width, self.height, C.byref(p), C.byref(pitch)))
        self.out_ptr, self.out_pitch = p.value, pitch.value
        return self.out_ptr, self.out_pitch

    def bindOutput(self, device_ptr, pitch_bytes=None):
        pitch = pitch_bytes or self.width * 16
        self._check(self._lib.mm_bind_output_linear(self._ctx, C.c_void_p(int(device_ptr)), pitch, self.width, self.height))
        self.out_ptr, self.out_pitch = int(device_ptr), pitch

    def bindOutputExternalBufferFd(self, fd, alloc_bytes, offset_bytes=0, pitch_bytes=None):
        """external (Vulkan-exported, linearly laid out) memory as the output image; the library owns the fd on success"""
        pitch = pitch_bytes or self.width * 16
        self._check(self._lib.mm_bind_output_external_buffer_fd(self._ctx, int(fd), alloc_bytes, offset_bytes, pitch, self.width, self.height))
        self.out_ptr, self.out_pitch = None, pitch

    def importSemaphoreFd(self, fd, timeline=False):
        slot = C.c_int()
        self._check(self._lib.mm_import_semaphore_fd(self._ctx, int(fd), int(timeline), C.byref(slot)))
        return slot.value

    def waitSemaphore(self, slot, value=0, stream=None):
        self._check(self._lib.mm_wait_semaphore(self._ctx, slot, value, self._stream(stream)))

    def signalSemaphore(self, slot, value=0, stream=None):
        self._check(self._lib.mm_signal_semaphore(self._ctx, slot, value, self._stream(stream)))

    def releaseSemaphore(self, slot):
        self._check(self._lib.mm_release_semaphore(self._ctx, slot))

    def enablePeer(self, other):
        self._check(self._lib.mm_enable_peer(self._ctx, other._ctx))

    # ---- host frame shared by several ranks: every dispatch also stores its pixels there (D2H fused into the kernel)
    def hostRegister(self, array):
        self._check(self._lib.mm_host_register(self._ctx, _ptr(array), array.nbytes))

    def hostUnregister(self, array):
        self._check(self._lib.mm_host_unregister(self._ctx, _ptr(array)))

    def bindHostMirror(self, array_or_none):
        self._check(self._lib.mm_bind_host_mirror(self._ctx, _ptr(array_or_none) if array_or_none is not None else None))

    # ---- multi-GPU: CUDA-IPC export / import of the output image
    def allocDevice(self, nbytes):
        p = C.c_void_p()
        self._check(self._lib.mm_alloc_device(self._ctx, nbytes, C.byref(p)))
        return p.value

    def freeDevice(self, device_ptr):
        self._check(self._lib.mm_free_device(self._ctx, C.c_void_p(int(device_ptr))))

    def ipcGetHandle(self, device_ptr):
        h = np.zeros(64, np.uint8)
        self._check(self._lib.mm_ipc_get_handle(self._ctx, C.c_void_p(int(device_ptr)), _ptr(h)))
        return h.tobytes()

    def ipcOpenHandle(self, handle):
        h = np.frombuffer(handle, np.uint8).copy()
        p = C.c_void_p()
        self._check(self._lib.mm_ipc_open_handle(self._ctx, _ptr(h), C.byref(p)))
        return p.value

    def ipcCloseHandle(self, device_ptr):
        self._check(self._lib.mm_ipc_close_handle(self._ctx, C.c_void_p(int(device_ptr))))

    # ---- per frame
    def updateUniformBuffers(self, cam, camPrev, sky, sun):
        cam, sky, sun = (np.ascontiguousarray(x, np.float32) for x in (cam, sky, sun))
        assert cam.nbytes == 160 and sun.nbytes == 116 and sky.nbytes == 52
        prev = _ptr(np.ascontiguousarray(camPrev, np.float32)) if camPrev is not None else None
        self._check(self._lib.mm_set_uniforms(self._ctx, _ptr(cam), prev, _ptr(sun), _ptr(sky)))

    def setFilterMode(self, mode):
        self._check(self._lib.mm_set_filter_mode(self._ctx, mode))

    def setArithmetic(self, arith):
        """MM_ARITH_IEEE (default: one rounding per operator) or MM_ARITH_FMA (the lexical contraction rule); a different DEFINITION"""
        self._check(self._lib.mm_set_arithmetic(self._ctx, arith))

    def setLanesPerRay(self, lanes):
        """0 = per dispatch (default), 1, 2, 4, 8: scheduling only, results are identical"""
        self._check(self._lib.mm_set_lanes_per_ray(self._ctx, lanes))

    def setScheduler(self, scheduler=capi.MM_SCHED_AUTO, refill_lanes=0):
        """static grid (K1) or persistent warps pulling tiles from a dynamic queue (K1p); refill_lanes 32/16/8: scheduling only"""
        self._check(self._lib.mm_set_scheduler(self._ctx, scheduler, refill_lanes))

    def dispatch(self, mode=capi.MM_FULL, row_begin=0, row_stride=1, row_block=1, stream=None):
        """stream: None -> the context's own stream; an integer cudaStream_t otherwise (0, the legacy default
        stream that torch calls its default stream, is passed as cudaStreamLegacy = 1)."""
        sp = None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1)
        self._check(self._lib.mm_dispatch(self._ctx, mode, row_begin, row_stride, row_block, sp))

    def bindPrevious(self, device_ptr, pitch_bytes=None):
        """descriptor set 1 (backgroundTexturePrev): the previous frame's image, read by the reprojection pass"""
        self._check(self._lib.mm_bind_previous_linear(self._ctx, C.c_void_p(int(device_ptr)), pitch_bytes or self.width * 16))

    def dispatchReproject(self, stream=None):
        sp = None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1)
        self._check(self._lib.mm_dispatch_reproject(self._ctx, sp))

    def synchronize(self):
        self._check(self._lib.mm_synchronize(self._ctx))

    def lastKernelMs(self):
        ms = C.c_float()
        self._check(self._lib.mm_last_kernel_ms(self._ctx, C.byref(ms)))
        return ms.value

    def renderToHost(self, cam, sky, sun, mode=capi.MM_FULL, out=None):
        """End-to-end call: host uniforms in, host image out."""
        cam, sky, sun = (np.ascontiguousarray(x, np.float32) for x in (cam, sky, sun))
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        self._check(self._lib.mm_render_to_host(self._ctx, _ptr(cam), _ptr(sun), _ptr(sky), mode, _ptr(out)))
        return out

    def readOutput(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        self._check(self._lib.mm_read_output(self._ctx, _ptr(out)))
        return out

    def readOutputInto(self, out):
        self._check(self._lib.mm_read_output(self._ctx, _ptr(out)))
        return out

    def tonemapRGBA8(self):
        out = np.empty((self.height, self.width, 4), np.uint8)
        self._check(self._lib.mm_tonemap_rgba8(self._ctx, _ptr(out), 0, None))
        return out

    # ---- post chain (PostProcessShader x3: god-ray.frag, radialBlur.frag, tonemap.frag); device pointers as integers
    @staticmethod
    def _stream(stream):
        return None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1)

    def godRay(self, cam, sun, src_ptr, dst_ptr, extent=None, src_pitch=None, dst_pitch=None, stream=None):
        w, h = extent or (self.width, self.height)
        cam, sun = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(sun, np.float32)
        self._check(self._lib.mm_god_ray(self._ctx, _ptr(cam), _ptr(sun), C.c_void_p(int(src_ptr)), src_pitch or w * 16,
                                         C.c_void_p(int(dst_ptr)), dst_pitch or w * 16, w, h, self._stream(stream)))

    def radialBlur(self, cam, sun, src_ptr, dst_ptr, extent=None, src_pitch=None, dst_pitch=None, stream=None):
        w, h = extent or (self.width, self.height)
        cam, sun = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(sun, np.float32)
        self._check(self._lib.mm_radial_blur(self._ctx, _ptr(cam), _ptr(sun), C.c_void_p(int(src_ptr)), src_pitch or w * 16,
                                             C.c_void_p(int(dst_ptr)), dst_pitch or w * 16, w, h, self._stream(stream)))

    def tonemapPresent(self, src_ptr, dst8_ptr, extent=None, src_pitch=None, dst_pitch=None, bgra=False, stream=None):
        w, h = extent or (self.width, self.height)
        self._check(self._lib.mm_tonemap_present(self._ctx, C.c_void_p(int(src_ptr)), src_pitch or w * 16, C.c_void_p(int(dst8_ptr)),
                                                 dst_pitch or w * 4, w, h, int(bgra), self._stream(stream)))

    def postChain(self, cam, sun, src_ptr, dst8_ptr, extent=None, src_pitch=None, dst_pitch=None, bgra=False, stream=None):
        """cloud image (RGBA32F) -> swapchain bytes: god rays, radial blur, tone map + vignette in two kernels"""
        w, h = extent or (self.width, self.height)
        cam, sun = np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(sun, np.float32)
        self._check(self._lib.mm_post_chain(self._ctx, _ptr(cam), _ptr(sun), C.c_void_p(int(src_ptr)), src_pitch or w * 16,
                                            C.c_void_p(int(dst8_ptr)), dst_pitch or w * 4, w, h, int(bgra), self._stream(stream)))

    # ---- cloud shadows (the mesh shader's 6-step march, model.frag:240-283) for an array of world positions
    def cloudShadow(self, positions, want_fetches=False):
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        out = np.empty(pos.shape[0], np.float32)
        nf = np.empty(pos.shape[0], np.uint32) if want_fetches else None
        self._check(self._lib.mm_cloud_shadow(self._ctx, _ptr(pos), pos.shape[0], 0, _ptr(out), _ptr(nf) if want_fetches else None, None))
        return (out, nf) if want_fetches else out

    def cloudShadowDevice(self, pos_ptr, n, out_ptr, stream=None):
        self._check(self._lib.mm_cloud_shadow(self._ctx, C.c_void_p(int(pos_ptr)), int(n), 1, C.c_void_p(int(out_ptr)), None, self._stream(stream)))

    # ---- diagnostics
    def enableCounters(self, on=True):
        self._check(self._lib.mm_enable_counters(self._ctx, int(on)))

    def readCounters(self):
        out = np.empty((self.height, self.width, 4), np.uint32)
        self._check(self._lib.mm_read_counters(self._ctx, _ptr(out)))
        return out

    def sample(self, slot, filter_mode, uvw):
        uvw = np.ascontiguousarray(uvw, np.float32).reshape(-1, 3)
        out = np.empty((uvw.shape[0], 4), np.float32)
        self._check(self._lib.mm_sample(self._ctx, slot, filter_mode, _ptr(uvw), uvw.shape[0], _ptr(out)))
        return out

    def measureTexPeak(self, slot, iters=4096):
        """-> (ms, bilinear-quad operations per second) of an L1-resident filtered-fetch microbenchmark on `slot`"""
        ms, q = C.c_float(), C.c_double()
        self._check(self._lib.mm_measure_tex_peak(self._ctx, slot, iters, C.byref(ms), C.byref(q)))
        return ms.value, q.value

    def selftestDiv(self, which):
        c, bad = C.c_float(), C.c_uint64()
        rc = self._lib.mm_selftest_div(self._ctx, which, C.byref(c), C.byref(bad))
        if rc == -1:
            return None
        self._check(rc)
        return c.value, bad.value

    def detPow(self, x, y):
        x, y = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32)
        out = np.empty_like(x)
        self._check(self._lib.mm_det_pow(self._ctx, _ptr(x), _ptr(y), x.size, _ptr(out)))
        return out
