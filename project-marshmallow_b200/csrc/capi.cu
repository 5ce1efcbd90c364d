// capi.cu -- the C-ABI of include/marshmallow.h: context, texture residency, uniforms, dispatch.
//
// What the reference does with Vulkan objects in ComputeShader (Shader.h:286-377, Shader.cpp:633-992)
// and Texture/Texture3D (Texture.cpp) is done here with CUDA objects:
//   descriptor set 2 samplers  -> per-slot {uchar4 cudaArray + texture object, pair-major float copy (cloud_march.cu)}
//   4 uniform buffers + memcpy -> MarchParams passed by value as a __grid_constant__ kernel argument
//   descriptor set 0 image     -> pitch-linear float4 pointer (own, caller's or a peer GPU's) or a
//                                 surface object over imported Vulkan memory
//   vkCmdDispatch + submit     -> one kernel launch on the caller's stream
// Memory plan per context: ~83 MB of read-only texture data (low-res volume: 8 MB cudaArray + 64 MB
// pair-major float copy; placement 1 + 8 MB; curl, hi-res ~2 MB), mostly L2-resident on B200 (126 MB),
// HBM behind it; output 16 B/pixel.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/marshmallow.h"
#include "common.h"
#include "curl_table.h"

using namespace mm;

struct TexSlot {
    cudaArray_t array = nullptr;
    cudaTextureObject_t obj = 0;
    float4 *pairs = nullptr;      // FP32-sampler pair-major copy, built on first use by ensure_pairs (never in MM_FILTER_HW)
    int w = 0, h = 0, d = 0;
    bool is3d = false;
};

struct mm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t side = nullptr;   // copy stream of the split end-to-end dispatch (forked from / joined to the dispatch stream)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool out_is_local = false;     // the bound linear image lives on this context's device (not a peer's)
    float *e2e_scratch = nullptr;  // local staging image of the split end-to-end dispatch when the bound image is a peer's
    size_t e2e_scratch_bytes = 0;
    bool timed = false;
    TexSlot tex[TEX_COUNT];
    float cam[40], cam_prev[40], sun[29], sky[13];
    bool have_uniforms = false, have_prev_camera = false;
    const float *prev_image = nullptr;   // descriptor set 1 (resultImagePrev / sourceImage), caller-owned
    size_t prev_pitch = 0;
    float *out = nullptr;          // bound output (may be external)
    float *own_out = nullptr;      // allocation owned by the context
    size_t pitch = 0;
    int W = 0, H = 0;
    cudaExternalMemory_t ext_mem = nullptr;
    cudaMipmappedArray_t ext_mip = nullptr;
    cudaSurfaceObject_t surf = 0;
    void *ext_linear = nullptr;    // mm_bind_output_external_buffer_fd: imported memory mapped as a linear buffer
    cudaExternalSemaphore_t sems[8] = {};   // mm_import_semaphore_fd
    float *mirror = nullptr;       // set only for the duration of one mm_render_to_host dispatch
    float *host_mirror = nullptr;  // mm_bind_host_mirror: device view of a page-locked host frame every dispatch also stores into
    uint32_t *counters = nullptr;
    bool counters_on = false;
    int filter = FILTER_HW;        // production default: texture-unit filtering (parity: the oracle's texture-unit model)
    float *scratch = nullptr;      // curl-noise scratch
    uchar4 *stage = nullptr;       // upload staging (device)
    size_t stage_bytes = 0;
    float *post_plane = nullptr;   // god-ray alpha plane of mm_post_chain
    size_t post_plane_bytes = 0;
    int lanes_per_ray = 0;         // mm_set_lanes_per_ray: 0 = chosen per dispatch, 1, 2, 4, 8
    int arith = MM_ARITH_IEEE;     // mm_set_arithmetic: one rounding per operator, or the contracted (fused multiply-add) definition
    int scheduler = MM_SCHED_AUTO; // mm_set_scheduler: static grid (K1) or persistent warps with a dynamic queue (K1p)
    int refill = 32;               // K1p: dead lanes that trigger a refill (32 = a whole tile at a time)
    unsigned *queue = nullptr;     // K1p work-queue counter (device)
    int sm_count = 148;
    char err[512];
};

static char g_create_err[512] = "";

static int fail(mm_ctx *c, int code, const char *fmt, ...) {
    char *dst = c ? c->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            cudaGetLastError();   /* a reported error must not resurface in a later, unrelated call */ \
            return fail(ctx, MM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                 \
        }                                                                                           \
    } while (0)

#ifndef MM_SPLIT_WAVES
#define MM_SPLIT_WAVES 3.0
#endif

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

extern "C" {

const char *mm_version(void) { return "marshmallow-b200 0.1 (sm_100a)"; }

const char *mm_last_error(const mm_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int mm_create(int device, mm_ctx **out) {
    mm_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, MM_ERR_ARG, "mm_create: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MM_ERR_CUDA, "mm_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, MM_ERR_ARG, "mm_create: device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, MM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MM_ERR_CUDA, "mm_create: device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major, prop.minor);
    ctx = new (std::nothrow) mm_ctx();
    if (!ctx) return fail(nullptr, MM_ERR_ARG, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->err[0] = 0;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->queue, 2 * sizeof(unsigned));
    if (e != cudaSuccess) {
        fail(nullptr, MM_ERR_CUDA, "mm_create: %s", cudaGetErrorString(e));
        delete ctx;
        return MM_ERR_CUDA;
    }
    *out = ctx;
    return MM_OK;
}

static void free_slot(TexSlot &s) {
    if (s.obj) cudaDestroyTextureObject(s.obj);
    if (s.array) cudaFreeArray(s.array);
    if (s.pairs) cudaFree(s.pairs);
    s = TexSlot();
}

static void release_output(mm_ctx *ctx) {
    if (ctx->surf) { cudaDestroySurfaceObject(ctx->surf); ctx->surf = 0; }
    if (ctx->ext_mip) { cudaFreeMipmappedArray(ctx->ext_mip); ctx->ext_mip = nullptr; }
    if (ctx->ext_linear) { cudaFree(ctx->ext_linear); ctx->ext_linear = nullptr; }
    if (ctx->ext_mem) { cudaDestroyExternalMemory(ctx->ext_mem); ctx->ext_mem = nullptr; }
    if (ctx->own_out) { cudaFree(ctx->own_out); ctx->own_out = nullptr; }
    if (ctx->counters) { cudaFree(ctx->counters); ctx->counters = nullptr; }
    ctx->out = nullptr; ctx->pitch = 0; ctx->W = ctx->H = 0; ctx->out_is_local = false;
}

int mm_destroy(mm_ctx *ctx) {
    if (!ctx) return MM_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < TEX_COUNT; i++) free_slot(ctx->tex[i]);
    release_output(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->post_plane) cudaFree(ctx->post_plane);
    if (ctx->queue) cudaFree(ctx->queue);
    if (ctx->e2e_scratch) cudaFree(ctx->e2e_scratch);
    for (auto &sem : ctx->sems) if (sem) { cudaDestroyExternalSemaphore(sem); sem = nullptr; }
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MM_OK;
}

// Make the texels in device buffer `src` (uchar4, [z][y][x]) resident in slot `slot`: a cudaArray with
// the reference's sampler state for hardware filtering, and the pair-major float copy for exact filtering.
static int bind_texels(mm_ctx *ctx, int slot, const uchar4 *src, int w, int h, int d, bool is3d) {
    TexSlot &s = ctx->tex[slot];
    free_slot(s);
    s.w = w; s.h = h; s.d = d; s.is3d = is3d;
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    if (is3d) {
        CU(cudaMalloc3DArray(&s.array, &fmt, make_cudaExtent(w, h, d)));
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(const_cast<uchar4 *>(src), (size_t)w * 4, w, h);
        cp.dstArray = s.array;
        cp.extent = make_cudaExtent(w, h, d);
        cp.kind = cudaMemcpyDeviceToDevice;
        CU(cudaMemcpy3DAsync(&cp, ctx->stream));
    } else {
        CU(cudaMallocArray(&s.array, &fmt, w, h));
        CU(cudaMemcpy2DToArrayAsync(s.array, 0, 0, src, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = s.array;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;   // REPEAT (Texture.cpp:36-38, 322-324)
    td.filterMode = cudaFilterModeLinear;                                              // LINEAR (Texture.cpp:33-34, 319-320)
    td.readMode = cudaReadModeNormalizedFloat;                                         // RGBA8_UNORM (Texture.h:29,85)
    td.normalizedCoords = 1;
    CU(cudaCreateTextureObject(&s.obj, &rd, &td, nullptr));
    CU(cudaStreamSynchronize(ctx->stream));
    return MM_OK;
}

static int ensure_stage(mm_ctx *ctx, size_t bytes);

// The pair-major binary32 copy of a slot (FP32-sampler modes only; the default texture-unit mode never reads it): built on
// first use from the texels resident in the slot's cudaArray, ~16x the bytes of the texture (64 MB for the 128^3 volume).
static int ensure_pairs(mm_ctx *ctx, int slot) {
    TexSlot &s = ctx->tex[slot];
    if (s.pairs || !s.array) return MM_OK;
    size_t n = (size_t)s.w * s.h * s.d;
    int rc = ensure_stage(ctx, n * 4);
    if (rc) return rc;
    if (s.is3d) {
        cudaMemcpy3DParms cp = {};
        cp.srcArray = s.array;
        cp.dstPtr = make_cudaPitchedPtr(ctx->stage, (size_t)s.w * 4, s.w, s.h);
        cp.extent = make_cudaExtent(s.w, s.h, s.d);
        cp.kind = cudaMemcpyDeviceToDevice;
        CU(cudaMemcpy3DAsync(&cp, ctx->stream));
    } else {
        CU(cudaMemcpy2DFromArrayAsync(ctx->stage, (size_t)s.w * 4, s.array, 0, 0, (size_t)s.w * 4, s.h, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CU(cudaMalloc(&s.pairs, n * 2 * sizeof(float4)));
    CU(launch_pack_pairs(ctx->stage, s.pairs, s.w, s.h, s.d, slot == MM_TEX_PLACEMENT, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MM_OK;
}

static int ensure_stage(mm_ctx *ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return MM_OK;
    if (ctx->stage) { cudaFree(ctx->stage); ctx->stage = nullptr; ctx->stage_bytes = 0; }
    CU(cudaMalloc(&ctx->stage, bytes));
    ctx->stage_bytes = bytes;
    return MM_OK;
}

static int upload(mm_ctx *ctx, int slot, const uint8_t *rgba8, int w, int h, int d, bool is3d) {
    if (!ctx) return MM_ERR_ARG;
    if (!rgba8 || w <= 0 || h <= 0 || d <= 0) return fail(ctx, MM_ERR_ARG, "upload: bad texture arguments");
    if (slot < 0 || slot >= TEX_COUNT) return fail(ctx, MM_ERR_ARG, "upload: slot %d out of range", slot);
    bool want3d = (slot == MM_TEX_LOWRES || slot == MM_TEX_HIRES);
    if (want3d != is3d) return fail(ctx, MM_ERR_ARG, "upload: slot %d is a %s sampler", slot, want3d ? "3D" : "2D");
    CU(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)w * h * d * 4;
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->stage, rgba8, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return bind_texels(ctx, slot, ctx->stage, w, h, d, is3d);
}

int mm_upload_tex2d(mm_ctx *ctx, int slot, const uint8_t *rgba8, int w, int h) { return upload(ctx, slot, rgba8, w, h, 1, false); }
int mm_upload_tex3d(mm_ctx *ctx, int slot, const uint8_t *rgba8, int w, int h, int d) { return upload(ctx, slot, rgba8, w, h, d, true); }

int mm_build_curl_noise(mm_ctx *ctx, uint8_t *out_host) {
    if (!ctx) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    const size_t N = 128 * 128, TBL = 26 * 26 * 26;
    if (!ctx->scratch) CU(cudaMalloc(&ctx->scratch, (15 * N + 8) * sizeof(float) + TBL));
    unsigned char *dtable = reinterpret_cast<unsigned char *>(ctx->scratch + 15 * N + 8);
    static unsigned char htable[26 * 26 * 26];
    static std::once_flag htable_once;
    std::call_once(htable_once, build_curl_gradient_table, htable);
    CU(cudaMemcpyAsync(dtable, htable, TBL, cudaMemcpyHostToDevice, ctx->stream));
    int rc = ensure_stage(ctx, N * 4);
    if (rc) return rc;
    CU(launch_curl_noise(ctx->stage, ctx->scratch, dtable, ctx->stream));
    if (out_host) CU(cudaMemcpyAsync(out_host, ctx->stage, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return bind_texels(ctx, MM_TEX_CURL, ctx->stage, 128, 128, 1, false);
}

int mm_build_noise_volumes(mm_ctx *ctx, uint64_t seed64, uint8_t *out_low, uint8_t *out_hi) {
    if (!ctx) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    const size_t NL = (size_t)128 * 128 * 128, NH = (size_t)32 * 32 * 32;
    int rc = ensure_stage(ctx, (NL + NH) * 4);
    if (rc) return rc;
    uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
    uchar4 *low = ctx->stage, *hi = ctx->stage + NL;
    CU(launch_noise_volumes(seed, low, hi, ctx->stream));
    if (out_low) CU(cudaMemcpyAsync(out_low, low, NL * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_hi) CU(cudaMemcpyAsync(out_hi, hi, NH * 4, cudaMemcpyDeviceToHost, ctx->stream));
    rc = bind_texels(ctx, MM_TEX_LOWRES, low, 128, 128, 128, true);
    if (rc) return rc;
    return bind_texels(ctx, MM_TEX_HIRES, hi, 32, 32, 32, true);
}

int mm_set_uniforms(mm_ctx *ctx, const void *camera160, const void *camera_prev160, const void *sun116, const void *sky52) {
    if (!ctx) return MM_ERR_ARG;
    if (!camera160 || !sun116 || !sky52) return fail(ctx, MM_ERR_ARG, "mm_set_uniforms: null uniform block");
    // UniformCameraObjectPrev is declared but never read by the march (CC:19-23); the reprojection pass reads it
    ctx->have_prev_camera = camera_prev160 != nullptr;
    if (camera_prev160) memcpy(ctx->cam_prev, camera_prev160, 160);
    memcpy(ctx->cam, camera160, 160);
    memcpy(ctx->sun, sun116, 116);
    memcpy(ctx->sky, sky52, 52);
    ctx->have_uniforms = true;
    return MM_OK;
}

static int size_counters(mm_ctx *ctx) {
    if (ctx->counters) { cudaFree(ctx->counters); ctx->counters = nullptr; }
    if (ctx->counters_on && ctx->W > 0) {
        CU(cudaMalloc(&ctx->counters, (size_t)ctx->W * ctx->H * 16));
        CU(cudaMemset(ctx->counters, 0, (size_t)ctx->W * ctx->H * 16));
    }
    return MM_OK;
}

int mm_bind_output_linear(mm_ctx *ctx, float *dptr, size_t pitch, int w, int h) {
    if (!ctx) return MM_ERR_ARG;
    if (!dptr || w <= 0 || h <= 0 || pitch < (size_t)w * 16 || (pitch & 15) || ((uintptr_t)dptr & 15))
        return fail(ctx, MM_ERR_ARG, "mm_bind_output_linear: need a 16-byte aligned pointer and pitch >= 16*w");
    CU(cudaSetDevice(ctx->device));
    release_output(ctx);
    ctx->out = dptr; ctx->pitch = pitch; ctx->W = w; ctx->H = h;
    {   // a peer GPU's image (multi-GPU gather fused into the stores) must not be copied from by THIS context's copy engine
        cudaPointerAttributes attr;
        ctx->out_is_local = cudaPointerGetAttributes(&attr, dptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice && attr.device == ctx->device;
        cudaGetLastError();
    }
    return size_counters(ctx);
}

int mm_alloc_output(mm_ctx *ctx, int w, int h, float **dptr_out, size_t *pitch_out) {
    if (!ctx) return MM_ERR_ARG;
    if (w <= 0 || h <= 0) return fail(ctx, MM_ERR_ARG, "mm_alloc_output: bad size");
    CU(cudaSetDevice(ctx->device));
    release_output(ctx);
    CU(cudaMalloc(&ctx->own_out, (size_t)w * h * 16));
    ctx->out = ctx->own_out; ctx->pitch = (size_t)w * 16; ctx->W = w; ctx->H = h; ctx->out_is_local = true;
    if (dptr_out) *dptr_out = ctx->out;
    if (pitch_out) *pitch_out = ctx->pitch;
    return size_counters(ctx);
}

// Vulkan interop: the engine exports the VkDeviceMemory behind backgroundTexture
// (VK_FORMAT_R32G32B32A32_SFLOAT, optimal tiling, STORAGE|SAMPLED; VulkanApplication.cpp:247-250) as an
// opaque fd; it is imported here as a one-level mipmapped array and written through a surface object.
int mm_bind_output_external_fd(mm_ctx *ctx, int fd, size_t alloc_bytes, int w, int h) {
    if (!ctx) return MM_ERR_ARG;
    if (fd < 0 || alloc_bytes < (size_t)w * h * 16 || w <= 0 || h <= 0) return fail(ctx, MM_ERR_ARG, "mm_bind_output_external_fd: bad arguments");
    CU(cudaSetDevice(ctx->device));
    release_output(ctx);
    cudaExternalMemoryHandleDesc hd = {};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = alloc_bytes;
    CU(cudaImportExternalMemory(&ctx->ext_mem, &hd));
    cudaExternalMemoryMipmappedArrayDesc md = {};
    md.offset = 0;
    md.formatDesc = cudaCreateChannelDesc<float4>();
    md.extent = make_cudaExtent(w, h, 0);
    md.flags = cudaArraySurfaceLoadStore;
    md.numLevels = 1;
    CU(cudaExternalMemoryGetMappedMipmappedArray(&ctx->ext_mip, ctx->ext_mem, &md));
    cudaArray_t level0;
    CU(cudaGetMipmappedArrayLevel(&level0, ctx->ext_mip, 0));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = level0;
    CU(cudaCreateSurfaceObject(&ctx->surf, &rd));
    ctx->W = w; ctx->H = h;
    return size_counters(ctx);
}

// The same import for memory the engine laid out LINEARLY (a VkBuffer, or a VkImage created with VK_IMAGE_TILING_LINEAR whose row
// pitch the engine reads from vkGetImageSubresourceLayout): mapped as a plain device buffer and written with ordinary stores, so
// everything that works on mm_bind_output_linear (peer stores, host mirror, reprojection) works on it.  Also the one interop path
// that can be exercised without a Vulkan loader: CUDA's own virtual-memory allocations export POSIX fds (tests/test_interop_gpu.py).
int mm_bind_output_external_buffer_fd(mm_ctx *ctx, int fd, size_t alloc_bytes, size_t offset_bytes, size_t pitch_bytes, int w, int h) {
    if (!ctx) return MM_ERR_ARG;
    if (fd < 0 || w <= 0 || h <= 0 || pitch_bytes < (size_t)w * 16 || (pitch_bytes & 15) || (offset_bytes & 15) ||
        alloc_bytes < offset_bytes + pitch_bytes * (size_t)(h - 1) + (size_t)w * 16)
        return fail(ctx, MM_ERR_ARG, "mm_bind_output_external_buffer_fd: bad arguments (16-byte aligned offset and pitch >= 16*w inside the allocation)");
    CU(cudaSetDevice(ctx->device));
    release_output(ctx);
    cudaExternalMemoryHandleDesc hd = {};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = alloc_bytes;
    CU(cudaImportExternalMemory(&ctx->ext_mem, &hd));          // on success CUDA owns the fd
    cudaExternalMemoryBufferDesc bd = {};
    bd.offset = offset_bytes;
    bd.size = alloc_bytes - offset_bytes;
    cudaError_t e = cudaExternalMemoryGetMappedBuffer(&ctx->ext_linear, ctx->ext_mem, &bd);
    if (e != cudaSuccess) {
        cudaDestroyExternalMemory(ctx->ext_mem); ctx->ext_mem = nullptr;
        return fail(ctx, MM_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e));
    }
    ctx->out = static_cast<float *>(ctx->ext_linear); ctx->pitch = pitch_bytes; ctx->W = w; ctx->H = h; ctx->out_is_local = false;
    return size_counters(ctx);
}

// ---- synchronisation with the engine's queues.  The reference submits its compute command buffer with no fence or semaphore at all
// (VulkanApplication.cpp:168-177, vkQueueSubmit(computeQueue, 1, &submitInfo, VK_NULL_HANDLE)) and relies on vkQueueWaitIdle per frame;
// a CUDA producer needs the ordering made explicit.  The engine exports a VkSemaphore (VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT,
// binary, or a timeline semaphore) per direction; mm_wait_semaphore orders a stream behind the engine's last reader of the image,
// mm_signal_semaphore lets the graphics submits wait for the march.  Up to 8 imported semaphores per context.
int mm_import_semaphore_fd(mm_ctx *ctx, int fd, int timeline, int *slot_out) {
    if (!ctx || !slot_out) return MM_ERR_ARG;
    if (fd < 0) return fail(ctx, MM_ERR_ARG, "mm_import_semaphore_fd: bad fd");
    CU(cudaSetDevice(ctx->device));
    int slot = -1;
    for (int i = 0; i < 8; i++) if (!ctx->sems[i]) { slot = i; break; }
    if (slot < 0) return fail(ctx, MM_ERR_UNSUPPORTED, "mm_import_semaphore_fd: all 8 semaphore slots are in use");
    cudaExternalSemaphoreHandleDesc sd = {};
    sd.type = timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    sd.handle.fd = fd;
    CU(cudaImportExternalSemaphore(&ctx->sems[slot], &sd));
    *slot_out = slot;
    return MM_OK;
}

static int semaphore_op(mm_ctx *ctx, int slot, uint64_t value, void *stream_v, bool signal) {
    if (!ctx) return MM_ERR_ARG;
    if (slot < 0 || slot >= 8 || !ctx->sems[slot]) return fail(ctx, MM_ERR_STATE, "semaphore slot %d holds no imported semaphore", slot);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
    if (signal) {
        cudaExternalSemaphoreSignalParams sp = {};
        sp.params.fence.value = value;                          // ignored by binary semaphores
        CU(cudaSignalExternalSemaphoresAsync(&ctx->sems[slot], &sp, 1, stream));
    } else {
        cudaExternalSemaphoreWaitParams wp = {};
        wp.params.fence.value = value;
        CU(cudaWaitExternalSemaphoresAsync(&ctx->sems[slot], &wp, 1, stream));
    }
    return MM_OK;
}
int mm_signal_semaphore(mm_ctx *ctx, int slot, uint64_t value, void *stream) { return semaphore_op(ctx, slot, value, stream, true); }
int mm_wait_semaphore(mm_ctx *ctx, int slot, uint64_t value, void *stream) { return semaphore_op(ctx, slot, value, stream, false); }
int mm_release_semaphore(mm_ctx *ctx, int slot) {
    if (!ctx) return MM_ERR_ARG;
    if (slot < 0 || slot >= 8 || !ctx->sems[slot]) return fail(ctx, MM_ERR_STATE, "semaphore slot %d holds no imported semaphore", slot);
    CU(cudaSetDevice(ctx->device));
    CU(cudaDestroyExternalSemaphore(ctx->sems[slot]));
    ctx->sems[slot] = nullptr;
    return MM_OK;
}

// Peer access for several contexts of ONE process (mm_dispatch_multi): lets `ctx`'s kernels store into memory that lives on
// `peer`'s device.  (Across processes the CUDA-IPC open enables it; within a process nothing does it implicitly.)
int mm_enable_peer(mm_ctx *ctx, mm_ctx *peer) {
    if (!ctx || !peer) return MM_ERR_ARG;
    if (ctx->device == peer->device) return MM_OK;
    CU(cudaSetDevice(ctx->device));
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
    if (!can) return fail(ctx, MM_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, peer->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer->device, cudaGetErrorString(e));
    return MM_OK;
}

// One frame sharded over n contexts of ONE process (one per GPU, or several on one GPU): context i marches partition i of n.  The
// contexts' output bindings decide where the pixels land (the same image, mapped into every device, assembles the frame in place).
int mm_dispatch_multi(mm_ctx **ctxs, int n, int mode, int row_block, void **streams) {
    if (!ctxs || n <= 0) return MM_ERR_ARG;
    for (int i = 0; i < n; i++)
        if (!ctxs[i]) return MM_ERR_ARG;
    for (int i = 0; i < n; i++) {
        int rc = mm_dispatch(ctxs[i], mode, i, n, row_block, streams ? streams[i] : nullptr);
        if (rc != MM_OK) return rc;                              // mm_last_error(ctxs[i]) has the reason
    }
    return MM_OK;
}

int mm_set_lanes_per_ray(mm_ctx *ctx, int lanes) {
    if (!ctx) return MM_ERR_ARG;
    if (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8)
        return fail(ctx, MM_ERR_ARG, "mm_set_lanes_per_ray: %d (0 = per dispatch, 1, 2, 4, 8)", lanes);
    ctx->lanes_per_ray = lanes;
    return MM_OK;
}

int mm_set_scheduler(mm_ctx *ctx, int scheduler, int refill_lanes) {
    if (!ctx) return MM_ERR_ARG;
    if (scheduler != MM_SCHED_AUTO && scheduler != MM_SCHED_STATIC && scheduler != MM_SCHED_PERSISTENT && scheduler != MM_SCHED_PACKED)
        return fail(ctx, MM_ERR_ARG, "mm_set_scheduler: unknown scheduler %d", scheduler);
    if (refill_lanes == 0) refill_lanes = 32;
    if (refill_lanes != 32 && refill_lanes != 16 && refill_lanes != 8)
        return fail(ctx, MM_ERR_ARG, "mm_set_scheduler: refill_lanes %d (0 or 32 = a tile at a time, 16, 8)", refill_lanes);
    ctx->scheduler = scheduler;
    ctx->refill = refill_lanes;
    return MM_OK;
}

int mm_set_arithmetic(mm_ctx *ctx, int arith) {
    if (!ctx) return MM_ERR_ARG;
    if (arith != MM_ARITH_IEEE && arith != MM_ARITH_FMA) return fail(ctx, MM_ERR_ARG, "mm_set_arithmetic: unknown definition %d", arith);
    ctx->arith = arith;
    return MM_OK;
}

int mm_set_filter_mode(mm_ctx *ctx, int filter) {
    if (!ctx) return MM_ERR_ARG;
    if (filter < MM_FILTER_EXACT || filter > MM_FILTER_HYBRID) return fail(ctx, MM_ERR_ARG, "mm_set_filter_mode: unknown mode %d", filter);
    ctx->filter = filter;
    return MM_OK;
}

// CC:392-401: samples[i] = mat3(sun.directionBasis) * s_i, column-major, ((c0*x)+(c1*y))+(c2*z) per
// component, binary32, no contraction (this file's host code is built with -ffp-contract=off); under the contracted
// definition fma(c2, z, fma(c0, x, RN(c1*y))) (oracle/glsl_env_fma.h).
static void light_cone_samples(const float *sun, float out[18], int arith) {
    static const float sv[6][3] = {{0.f, 0.6f, 0.f}, {0.f, 0.5f, 0.05f}, {0.1f, 0.75f, 0.f}, {0.2f, 2.5f, 0.3f}, {0.f, 6.f, 0.f}, {-0.1f, 1.f, -0.2f}};
    const float *c0 = sun + 12, *c1 = sun + 16, *c2 = sun + 20;
    for (int i = 0; i < 6; i++)
        for (int r = 0; r < 3; r++) {
            if (arith == MM_ARITH_FMA) {
                volatile float b = c1[r] * sv[i][1];
                out[3 * i + r] = fmaf(c2[r], sv[i][2], fmaf(c0[r], sv[i][0], b));
                continue;
            }
            volatile float a = c0[r] * sv[i][0], b = c1[r] * sv[i][1], c = c2[r] * sv[i][2];
            volatile float ab = a + b;
            out[3 * i + r] = ab + c;
        }
}

// Order the block rows of one dispatch by expected cost, descending.  Cost proxy: the elevation of the ray through the middle
// column; below the horizon (CC:351) a ray is free, above it the path through the shell grows as the ray approaches the
// horizon, i.e. as rd.y falls.  A block row is judged by its extreme rows: free only if its TOP row is below the horizon, and as
// expensive as its LOWEST ray above it -- the block row that straddles the horizon holds the longest rays of the frame (250 loop
// trips, ~0.45 ms as a dependent chain) and must start first, not with the free rows at the end (measured: that mistake cost the
// rank owning it 15 % of its frame share).  Scheduling hint only.
static void order_block_rows(const MarchParams &p, uint16_t *order, int nblockrows, int block_h, bool free_rows_first = false, int *nfree_out = nullptr) {
    const float *cam = p.cam;
    struct Key { float k; uint16_t i; };
    static thread_local Key keys[4096];
    const int last_row = p.owned_rows - 1;
    int nfree = 0;
    auto elevation = [&](int j) {                                  // rd.y of the middle-column ray of owned row j
        if (j > last_row) j = last_row;
        int py;
        if (p.mode == DISPATCH_PHASE16) py = j * 4;
        else { int k = j / p.row_block; py = owned_block(k, p.row_begin, p.row_stride, p.row_snake) * p.row_block + (j - k * p.row_block); }
        if (py >= p.H) py = p.H - 1;
        double spy = 2.0 * py / p.H - 1.0;
        double tanH = cam[37];
        // rd ~ -look - spy*tanH*up (middle column: spx = 0); only the sign and size of its y matter
        double dx = -cam[2] - spy * tanH * cam[1], dy = -cam[6] - spy * tanH * cam[5], dz = -cam[10] - spy * tanH * cam[9];
        return dy / std::sqrt(dx * dx + dy * dy + dz * dz);
    };
    for (int b = 0; b < nblockrows; b++) {
        double y0 = elevation(b * block_h), y1 = elevation(b * block_h + block_h - 1);
        double hi = y0 > y1 ? y0 : y1, lo = y0 > y1 ? y1 : y0;
        // ascending key = descending cost; all-below-horizon rows (free: CC:351-354 stores the sky colour and returns) last -- unless every
        // pixel is also stored to HOST memory over PCIe (mm_render_to_host / mm_bind_host_mirror): then they go FIRST, so that their
        // burst of stores (28 % of a C3 frame, ~0.7 ms of PCIe time at 4K) drains behind the march instead of after it
        keys[b].k = hi < 0.0 ? (free_rows_first ? -1.0f : 2.0f) : (float)(lo > 0.0 ? lo : 0.0);
        keys[b].i = (uint16_t)b;
        if (hi < 0.0) nfree++;
    }
    if (nfree_out) *nfree_out = nfree;
    static const char *dbg = getenv("MM_DEBUG_ROW_ORDER");         // diagnostics: "identity" / "reverse" switch the cost order off
    if (dbg && dbg[0] == 'i') { for (int b = 0; b < nblockrows; b++) order[b] = (uint16_t)b; if (nfree_out) *nfree_out = 0; return; }
    std::stable_sort(keys, keys + nblockrows, [](const Key &a, const Key &b) { return a.k < b.k; });
    for (int b = 0; b < nblockrows; b++) order[b] = keys[(dbg && dbg[0] == 'r') ? nblockrows - 1 - b : b].i;
}

// Host-only: the execution order mm_dispatch would give the block rows of a dispatch (most expensive first).  No device needed.
int mm_plan_block_rows(const void *camera160, int h, int mode, int row_begin, int row_stride, int row_block, int block_h,
                       uint16_t *order_out, int *count_out) {
    if (!camera160 || !order_out || !count_out || h <= 0 || row_stride <= 0 || row_block <= 0 || block_h <= 0 || row_begin < 0 || row_begin >= row_stride) return MM_ERR_ARG;
    const int snake = (mode & MM_ROWS_SNAKE) ? 1 : 0;
    mode &= ~MM_ROWS_SNAKE;
    if (mode != MM_FULL && mode != MM_PHASE16) return MM_ERR_ARG;
    static thread_local MarchParams p;
    memcpy(p.cam, camera160, sizeof p.cam);
    p.H = h; p.mode = mode; p.row_begin = row_begin; p.row_stride = row_stride; p.row_block = row_block; p.row_snake = snake;
    if (mode == MM_FULL) {
        int nblocks = (h + row_block - 1) / row_block, owned = 0;
        while (owned_block(owned, row_begin, row_stride, snake) < nblocks) owned++;
        p.owned_rows = owned * row_block;
    } else {
        p.owned_rows = (h + 3) / 4;
    }
    int n = (p.owned_rows + block_h - 1) / block_h;
    if (n > 4096) return MM_ERR_UNSUPPORTED;
    order_block_rows(p, p.block_row_order, n, block_h);
    memcpy(order_out, p.block_row_order, sizeof(uint16_t) * (size_t)n);
    *count_out = n;
    return MM_OK;
}

int mm_dispatch(mm_ctx *ctx, int mode, int row_begin, int row_stride, int row_block, void *stream_v) {
    if (!ctx) return MM_ERR_ARG;
    const int snake = (mode & MM_ROWS_SNAKE) ? 1 : 0;
    mode &= ~MM_ROWS_SNAKE;
    if (mode != MM_FULL && mode != MM_PHASE16) return fail(ctx, MM_ERR_ARG, "mm_dispatch: unknown mode %d", mode);
    if (row_begin < 0 || row_stride <= 0 || row_block <= 0 || row_begin >= row_stride)
        return fail(ctx, MM_ERR_ARG, "mm_dispatch: bad row partition (%d,%d,%d): need 0 <= row_begin < row_stride, row_block > 0", row_begin, row_stride, row_block);
    if (!ctx->have_uniforms) return fail(ctx, MM_ERR_STATE, "mm_dispatch: uniforms were never set");
    if (!ctx->out && !ctx->surf) return fail(ctx, MM_ERR_STATE, "mm_dispatch: no output image bound");
    // the 1-of-16 pixel phase rides in sun.color.a (CC:292-298, VulkanApplication.cpp:384 keeps it in 0..15); the shader's uint
    // pixel coordinates reject anything else by wrapping far outside the image -- here it is an argument error
    if (mode == MM_PHASE16 && !(ctx->sun[11] >= 0.0f && ctx->sun[11] < 16.0f))
        return fail(ctx, MM_ERR_ARG, "mm_dispatch: MM_PHASE16 needs sun.color.a in [0,16), got %g", (double)ctx->sun[11]);
    static const int need[4] = {MM_TEX_PLACEMENT, MM_TEX_CURL, MM_TEX_LOWRES, MM_TEX_HIRES};
    for (int i = 0; i < 4; i++)
        if (!ctx->tex[need[i]].obj) return fail(ctx, MM_ERR_STATE, "mm_dispatch: texture slot %d not bound", need[i]);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
    if (ctx->filter != FILTER_HW)                            // the FP32-sampler modes read the pair-major copies (built on first use)
        for (int i = 0; i < TEX_COUNT; i++) {
            int rc = ensure_pairs(ctx, i);
            if (rc) return rc;
        }

    MarchParams p;
    memcpy(p.cam, ctx->cam, sizeof p.cam);
    memcpy(p.sun, ctx->sun, sizeof p.sun);
    memcpy(p.sky, ctx->sky, sizeof p.sky);
    light_cone_samples(ctx->sun, p.light, ctx->arith);
    for (int i = 0; i < TEX_COUNT; i++) {
        const TexSlot &s = ctx->tex[i];
        p.tex[i].pairs = s.pairs; p.tex[i].obj = s.obj;
        p.tex[i].w = s.w; p.tex[i].h = s.h; p.tex[i].d = s.d;
        p.tex[i].wf = (float)s.w; p.tex[i].hf = (float)s.h; p.tex[i].df = (float)s.d;
        p.tex[i].pow2 = is_pow2(s.w) && is_pow2(s.h) && is_pow2(s.d);
    }
    p.out = ctx->out; p.pitch = ctx->pitch; p.surf = ctx->surf;
    p.mirror = ctx->mirror ? ctx->mirror : ctx->host_mirror; p.mirror_pitch = (size_t)ctx->W * 16;
    p.counters = ctx->counters_on ? ctx->counters : nullptr;
    p.W = ctx->W; p.H = ctx->H; p.mode = mode;
    p.row_begin = row_begin; p.row_stride = row_stride; p.row_block = row_block; p.row_snake = snake;
    if (mode == MM_FULL) {
        int nblocks = (ctx->H + row_block - 1) / row_block;
        int owned = 0;                                      // blocks k = 0, 1, ... of this partition that exist in the image
        while (owned_block(owned, row_begin, row_stride, snake) < nblocks) owned++;
        p.owned_rows = owned * row_block;
        p.grid_w = ctx->W;
    } else {
        p.owned_rows = (ctx->H + 3) / 4;
        p.grid_w = (ctx->W + 3) / 4;
    }
    // Lanes per ray (cloud_march.cu, K1 / K1s): one thread per ray unless the dispatch is too small to hide the latency of a
    // single ray behind other blocks.  A B200 keeps 148 x 1024 one-lane rays resident; below MM_SPLIT_WAVES times that, rays are
    // split over 2, 4 or 8 lanes so that the launch has about that many waves again (MM_PHASE16 dispatches, row-sharded frames
    // on several GPUs).  Scheduling only: every variant produces the same bits.
    int lanes = ctx->lanes_per_ray;
    if (lanes == 0) {
        // measured (tools/ab_bench.py --phase16 / --shard, profiles/r02_scheduler_ab.txt): one lane per ray wins from ~3 waves up (4
        // for the sparser, less coherent rays of a phase dispatch); two lanes down to ~1.2 waves (1/8 of a 1080p frame: 0.38 -> 0.30 ms);
        // four down to ~0.45 (1080p phase dispatch, 0.86 wave: 0.45 -> 0.22 ms); eight below that (720p phase dispatch: 0.47 -> 0.15 ms)
        double waves = (double)p.grid_w * (double)p.owned_rows / ((double)ctx->sm_count * 1024.0);
        double full = (mode == MM_PHASE16) ? 4.0 : MM_SPLIT_WAVES;
        lanes = waves >= full ? 1 : waves >= 1.2 ? 2 : waves >= 0.45 ? 4 : 8;
    }
    // the ray-split kernels (K1s) exist for power-of-two march textures only (wrap by mask); the texture unit wraps any extent
    bool pow2 = true;
    for (int i = 0; i < 4; i++) pow2 = pow2 && p.tex[need[i]].pow2;
    if (ctx->filter != FILTER_HW && !pow2) lanes = 1;
    // K1p (persistent warps + dynamic queue) instead of the static grid: opt-in.  Measured on B200 (profiles/r02_scheduler_ab.txt):
    // equal to the static grid on whole frames (5.85 vs 5.82 ms at 4K), 2-11 % slower on the row shares of an 8-GPU frame
    const bool persistent = lanes == 1 && ctx->scheduler == MM_SCHED_PERSISTENT;
    // K1x2 (two rays per thread on packed FP32): the texture-unit mode without diagnostic counters; anything else runs K1
    const bool packed = lanes == 1 && ctx->scheduler == MM_SCHED_PACKED && ctx->filter == FILTER_HW && !p.counters;
    int block_w, block_h;
    if (persistent) { block_w = TILE_W; block_h = TILE_H; } else march_block_shape(lanes, &block_w, &block_h);
    int nblockrows = (p.owned_rows + block_h - 1) / block_h;
    if (nblockrows > 4096) return fail(ctx, MM_ERR_UNSUPPORTED, "mm_dispatch: too many rows per dispatch");
    int nfree = 0;
    order_block_rows(p, p.block_row_order, nblockrows, block_h, p.mirror != nullptr, &nfree);
    if (getenv("MM_DEBUG_ROW_ORDER")) nfree = 0;
    int persistent_blocks = 0;
    p.queue = nullptr; p.n_slots = 0; p.tiles_x = 0;
    if (persistent) {                                             // queue entries are pixel slots, 32 per 8x4 tile
        p.tiles_x = (unsigned)((p.grid_w + TILE_W - 1) / TILE_W);
        unsigned tiles = p.tiles_x * (unsigned)nblockrows;
        p.n_slots = tiles * 32u;
        p.queue = ctx->queue;
        unsigned resident = (unsigned)(ctx->sm_count * persistent_blocks_per_sm(ctx->filter));
        persistent_blocks = (int)std::min(resident, (tiles + 3u) / 4u);
        CU(cudaMemsetAsync(ctx->queue, 0, 2 * sizeof(unsigned), stream));
    }
    if (packed) persistent_blocks = -1;
    p.launch_block_rows = 0;
    // End-to-end split.  A kernel's stores into mapped host memory top out near 21 GB/s on this platform, the copy engine reaches 55 GB/s
    // (tools/e2e_probe.py): a 4K frame mirrored pixel by pixel is PCIe-bound (6.37 ms against 5.66 ms of march).  The rows below the
    // horizon cost no march time and are 28 % of the C3 frame's bytes: they are marched FIRST in a launch of their own that does not
    // mirror, then copied device -> host by the copy engine on a second stream while the rest of the frame marches and mirrors its own
    // pixels.  When the bound image is a PEER's (ranks != 0 of a sharded frame) the copy engine must not read it -- that would go through the
    // peer's engine and PCIe link -- so those rows are stored to the peer image and to a local staging image, and copied from there.
    static const char *split_env = getenv("MM_E2E_SPLIT");         // diagnostics: "0" keeps every pixel on the kernel-store path
    const char *rest_env = getenv("MM_E2E_REST");                  // diagnostics: "fused" | "copy" | "copy2" overrides the choice below
    // The rest of the frame: stores fused into the kernel (one launch, nothing after it; the default).  The alternatives -- the copy
    // engine again, behind one launch ("copy") or behind two launches of which the second overlaps the first one's copies ("copy2") --
    // are kept for platforms whose mapped stores are slower; on this one they lose at every shard size measured
    // (4K on 4 GPUs: 1.60 ms fused, 2.30 copy, 2.01 copy2; profiles/r02_e2e_rest_modes.txt).
    int rest_mode = 0;
    if (rest_env) rest_mode = !strcmp(rest_env, "copy2") ? 2 : !strcmp(rest_env, "copy") ? 1 : 0;
    if (p.mirror && p.out && mode == MM_FULL && !persistent && (nfree > 0 || rest_mode != 0) && !(split_env && split_env[0] == '0')) {
        auto launch = ctx->arith == MM_ARITH_FMA ? launch_cloud_march_fma : launch_cloud_march;
        float *const host_frame = p.mirror;
        const size_t host_pitch = p.mirror_pitch;
        const char *copy_src = reinterpret_cast<const char *>(p.out);
        size_t copy_pitch = p.pitch;
        float *stage = nullptr;                                    // what a launch whose rows go through the copy engine mirrors into
        if (!ctx->out_is_local) {
            const size_t need = (size_t)p.W * p.H * 16;
            if (ctx->e2e_scratch_bytes < need) {
                if (ctx->e2e_scratch) { CU(cudaFree(ctx->e2e_scratch)); ctx->e2e_scratch = nullptr; ctx->e2e_scratch_bytes = 0; }
                CU(cudaMalloc(&ctx->e2e_scratch, need));
                ctx->e2e_scratch_bytes = need;
            }
            stage = ctx->e2e_scratch;
            copy_src = reinterpret_cast<const char *>(ctx->e2e_scratch); copy_pitch = (size_t)p.W * 16;
        }
        static thread_local uint16_t order[sizeof(p.block_row_order) / sizeof(uint16_t)];
        memcpy(order, p.block_row_order, sizeof(uint16_t) * (size_t)nblockrows);
        static thread_local int rows[4096 * 8];
        // image rows of block rows order[first .. first+count) as runs of consecutive rows -> one 2D copy per run, on the side stream
        auto copy_rows = [&](int first, int count) -> cudaError_t {
            int nrows = 0;
            for (int i = first; i < first + count; i++)
                for (int r = 0; r < block_h; r++) {
                    int j = (int)order[i] * block_h + r;
                    if (j >= p.owned_rows) continue;
                    int k = j / row_block;
                    int py = owned_block(k, row_begin, row_stride, snake) * row_block + (j - k * row_block);
                    if (py < p.H && nrows < 4096 * 8) rows[nrows++] = py;
                }
            std::sort(rows, rows + nrows);
            int i = 0;
            while (i < nrows) {
                int n = 1;
                while (i + n < nrows && rows[i + n] == rows[i] + n) n++;
                cudaError_t e = cudaMemcpy2DAsync(reinterpret_cast<char *>(host_frame) + (size_t)rows[i] * host_pitch, host_pitch,
                                                  copy_src + (size_t)rows[i] * copy_pitch, copy_pitch, (size_t)p.W * 16, (size_t)n,
                                                  cudaMemcpyDefault, ctx->side);
                if (e != cudaSuccess) return e;
                i += n;
            }
            return cudaSuccess;
        };
        // one launch of block rows order[first .. first+count); through_copy_engine: the launch does not touch host memory and its rows
        // are queued on the side stream behind it
        auto chunk = [&](int first, int count, bool through_copy_engine) -> cudaError_t {
            if (count <= 0) return cudaSuccess;
            memcpy(p.block_row_order, order + first, sizeof(uint16_t) * (size_t)count);
            p.launch_block_rows = count;
            p.mirror = through_copy_engine ? stage : host_frame;
            p.mirror_pitch = through_copy_engine ? (size_t)p.W * 16 : host_pitch;
            cudaError_t e = launch(p, ctx->filter, lanes, persistent_blocks, ctx->refill, stream);
            if (e != cudaSuccess || !through_copy_engine) return e;
            if ((e = cudaEventRecord(ctx->ev_fork, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0)) != cudaSuccess) return e;
            return copy_rows(first, count);
        };
        CU(cudaEventRecord(ctx->ev0, stream));
        CU(chunk(0, nfree, true));                                 // the free block rows lead the order when a mirror is bound
        const int nrest = nblockrows - nfree;
        if (rest_mode == 0) {
            CU(chunk(nfree, nrest, false));
        } else if (rest_mode == 1 || nrest < 8) {
            CU(chunk(nfree, nrest, true));
        } else {
            const int first_half = (nrest * 5) / 8;
            CU(chunk(nfree, first_half, true));
            CU(chunk(nfree + first_half, nrest - first_half, true));
        }
        CU(cudaEventRecord(ctx->ev_join, ctx->side));
        CU(cudaStreamWaitEvent(stream, ctx->ev_join, 0));
        CU(cudaEventRecord(ctx->ev1, stream));
        ctx->timed = true;
        return MM_OK;
    }
    CU(cudaEventRecord(ctx->ev0, stream));
    CU((ctx->arith == MM_ARITH_FMA ? launch_cloud_march_fma : launch_cloud_march)(p, ctx->filter, lanes, persistent_blocks, ctx->refill, stream));
    CU(cudaEventRecord(ctx->ev1, stream));
    ctx->timed = true;
    return MM_OK;
}

int mm_bind_previous_linear(mm_ctx *ctx, const float *dptr, size_t pitch) {
    if (!ctx) return MM_ERR_ARG;
    if (!dptr || (pitch & 15) || ((uintptr_t)dptr & 15)) return fail(ctx, MM_ERR_ARG, "mm_bind_previous_linear: need a 16-byte aligned pointer and pitch");
    ctx->prev_image = dptr; ctx->prev_pitch = pitch;
    return MM_OK;
}

int mm_dispatch_reproject(mm_ctx *ctx, void *stream_v) {
    if (!ctx) return MM_ERR_ARG;
    if (!ctx->have_uniforms || !ctx->have_prev_camera) return fail(ctx, MM_ERR_STATE, "mm_dispatch_reproject: camera and previous camera blocks are required");
    if (!ctx->out) return fail(ctx, MM_ERR_STATE, "mm_dispatch_reproject: no linear output image bound");
    if (!ctx->prev_image) return fail(ctx, MM_ERR_STATE, "mm_dispatch_reproject: no previous image bound (mm_bind_previous_linear)");
    if (ctx->prev_image == ctx->out) return fail(ctx, MM_ERR_ARG, "mm_dispatch_reproject: previous and target image must differ (ping-pong)");
    if (ctx->prev_pitch < (size_t)ctx->W * 16) return fail(ctx, MM_ERR_ARG, "mm_dispatch_reproject: previous image pitch < 16*w");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
    ReprojectParams p;
    memcpy(p.cam, ctx->cam, 160);
    memcpy(p.cam_prev, ctx->cam_prev, 160);
    p.src = ctx->prev_image; p.src_pitch = ctx->prev_pitch;
    p.dst = ctx->out; p.dst_pitch = ctx->pitch;
    p.W = ctx->W; p.H = ctx->H;
    CU(cudaEventRecord(ctx->ev0, stream));
    CU(launch_reproject(p, stream));
    CU(cudaEventRecord(ctx->ev1, stream));
    ctx->timed = true;
    return MM_OK;
}

int mm_synchronize(mm_ctx *ctx) {
    if (!ctx) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaDeviceSynchronize());
    return MM_OK;
}

int mm_last_kernel_ms(mm_ctx *ctx, float *ms) {
    if (!ctx || !ms) return MM_ERR_ARG;
    if (!ctx->timed) return fail(ctx, MM_ERR_STATE, "mm_last_kernel_ms: nothing dispatched yet");
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return MM_OK;
}

int mm_read_output(mm_ctx *ctx, float *host_out) {
    if (!ctx || !host_out) return MM_ERR_ARG;
    if (!ctx->out) return fail(ctx, MM_ERR_STATE, "mm_read_output: no linear output bound");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy2D(host_out, (size_t)ctx->W * 16, ctx->out, ctx->pitch, (size_t)ctx->W * 16, ctx->H, cudaMemcpyDeviceToHost));
    return MM_OK;
}

int mm_render_to_host(mm_ctx *ctx, const void *camera160, const void *sun116, const void *sky52, int mode, float *out_host) {
    if (!ctx || !out_host) return MM_ERR_ARG;
    int rc = mm_set_uniforms(ctx, camera160, nullptr, sun116, sky52);
    if (rc) return rc;
    if (!ctx->out) return fail(ctx, MM_ERR_STATE, "mm_render_to_host: no linear output bound (mm_alloc_output)");
    // Pinned (page-locked) destination + full frame: the kernel stores every finished pixel to the device image AND
    // straight into the caller's buffer over PCIe, so the device->host transfer overlaps the march instead of following
    // it.  Pageable destination or a 1/16 phase dispatch: march, then copy the image.
    float *mapped = nullptr;
    if (mode == MM_FULL) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            mapped = static_cast<float *>(attr.devicePointer);
        else
            cudaGetLastError();
    }
    ctx->mirror = mapped;
    rc = mm_dispatch(ctx, mode, 0, 1, 1, nullptr);
    ctx->mirror = nullptr;
    if (rc) return rc;
    if (!mapped)
        CU(cudaMemcpy2DAsync(out_host, (size_t)ctx->W * 16, ctx->out, ctx->pitch, (size_t)ctx->W * 16, ctx->H, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MM_OK;
}

// ---- host frame shared by several processes (multi-GPU end-to-end path) -----------------------------------------
int mm_host_register(mm_ctx *ctx, void *host, size_t bytes) {
    if (!ctx || !host || bytes == 0) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return MM_OK;
}

int mm_host_unregister(mm_ctx *ctx, void *host) {
    if (!ctx || !host) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    if (ctx->host_mirror) { CU(cudaStreamSynchronize(ctx->stream)); ctx->host_mirror = nullptr; }
    CU(cudaHostUnregister(host));
    return MM_OK;
}

int mm_bind_host_mirror(mm_ctx *ctx, float *host_rgba32f) {
    if (!ctx) return MM_ERR_ARG;
    if (!host_rgba32f) { ctx->host_mirror = nullptr; return MM_OK; }
    if (!ctx->out) return fail(ctx, MM_ERR_STATE, "mm_bind_host_mirror: bind the device output image first");
    CU(cudaSetDevice(ctx->device));
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host_rgba32f) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
        cudaGetLastError();
        return fail(ctx, MM_ERR_ARG, "mm_bind_host_mirror: the frame must be page-locked and mapped (cudaHostAlloc / mm_host_register)");
    }
    ctx->host_mirror = static_cast<float *>(attr.devicePointer);
    return MM_OK;
}

int mm_tonemap_rgba8(mm_ctx *ctx, uint8_t *dst, int dst_is_device, void *stream_v) {
    if (!ctx || !dst) return MM_ERR_ARG;
    if (!ctx->out) return fail(ctx, MM_ERR_STATE, "mm_tonemap_rgba8: no linear output bound");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
    size_t bytes = (size_t)ctx->W * ctx->H * 4;
    if (dst_is_device) {
        CU(launch_tonemap(ctx->out, ctx->pitch, ctx->W, ctx->H, reinterpret_cast<uchar4 *>(dst), stream));
        return MM_OK;
    }
    uchar4 *tmp = nullptr;
    CU(cudaMalloc(&tmp, bytes));
    cudaError_t e = launch_tonemap(ctx->out, ctx->pitch, ctx->W, ctx->H, tmp, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "mm_tonemap_rgba8: %s", cudaGetErrorString(e));
    return MM_OK;
}

// ---- post chain (god-ray.frag, radialBlur.frag, tonemap.frag) ------------------------------------------------------
// (proj * view) * sun.location and the perspective divide (god-ray.frag:44-46, radialBlur.frag:48-49): uniform per
// frame, evaluated here in the oracle's order -- mat4*mat4 then mat4*vec4, ((m0*v0 + m1*v1) + m2*v2) + m3*v3 per
// component, binary32, no contraction (volatile keeps the host compiler from fusing or reassociating).
static float dot4_ordered(float a0, float b0, float a1, float b1, float a2, float b2, float a3, float b3) {
    volatile float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2, p3 = a3 * b3;
    volatile float s = p0 + p1;
    s = s + p2;
    s = s + p3;
    return s;
}
static void post_params(const float *cam, const float *sun, PostParams &p) {
    const float *V = cam, *P = cam + 16;
    float PV[16], r[4];
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++)
            PV[j * 4 + i] = dot4_ordered(P[0 * 4 + i], V[j * 4 + 0], P[1 * 4 + i], V[j * 4 + 1], P[2 * 4 + i], V[j * 4 + 2], P[3 * 4 + i], V[j * 4 + 3]);
    for (int i = 0; i < 4; i++) r[i] = dot4_ordered(PV[0 * 4 + i], sun[0], PV[1 * 4 + i], sun[1], PV[2 * 4 + i], sun[2], PV[3 * 4 + i], sun[3]);
    volatile float sx = r[0] / r[3], sy = r[1] / r[3];
    p.sun_x = sx; p.sun_y = sy;
    p.sun_dir_y = sun[5];
    for (int k = 0; k < 3; k++) { volatile float c = sun[8 + k] * sun[28]; p.sun_rgb[k] = c; }
}

static int post_check(mm_ctx *ctx, const char *who, const void *cam, const void *sun, const void *src, size_t src_pitch, const void *dst, size_t dst_pitch,
                      size_t dst_texel, int w, int h) {
    if (!ctx) return MM_ERR_ARG;
    if ((!cam || !sun) && who[3] != 'p') return fail(ctx, MM_ERR_ARG, "%s: camera and sun blocks are required", who);
    if (!src || !dst || w <= 0 || h <= 0) return fail(ctx, MM_ERR_ARG, "%s: null image or empty extent", who);
    if (src == dst) return fail(ctx, MM_ERR_ARG, "%s: the pass reads neighbours of the pixel it writes; source and destination must differ", who);
    if (((uintptr_t)src & 15) || (src_pitch & 15) || src_pitch < (size_t)w * 16) return fail(ctx, MM_ERR_ARG, "%s: source needs 16-byte alignment and pitch >= 16*w", who);
    if (((uintptr_t)dst & (dst_texel - 1)) || (dst_pitch & (dst_texel - 1)) || dst_pitch < (size_t)w * dst_texel)
        return fail(ctx, MM_ERR_ARG, "%s: destination needs %zu-byte alignment and pitch >= %zu*w", who, dst_texel, dst_texel);
    return MM_OK;
}

int mm_god_ray(mm_ctx *ctx, const void *camera160, const void *sun116, const float *src, size_t src_pitch, float *dst, size_t dst_pitch, int w, int h, void *stream_v) {
    int rc = post_check(ctx, "mm_god_ray", camera160, sun116, src, src_pitch, dst, dst_pitch, 16, w, h);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    PostParams p = {};
    post_params((const float *)camera160, (const float *)sun116, p);
    p.src = src; p.src_pitch = src_pitch; p.dst = dst; p.dst_pitch = dst_pitch; p.W = w; p.H = h;
    CU(launch_post(POST_GOD_RAY, p, stream_v ? (cudaStream_t)stream_v : ctx->stream));
    return MM_OK;
}

int mm_radial_blur(mm_ctx *ctx, const void *camera160, const void *sun116, const float *src, size_t src_pitch, float *dst, size_t dst_pitch, int w, int h, void *stream_v) {
    int rc = post_check(ctx, "mm_radial_blur", camera160, sun116, src, src_pitch, dst, dst_pitch, 16, w, h);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    PostParams p = {};
    post_params((const float *)camera160, (const float *)sun116, p);
    p.src = src; p.src_pitch = src_pitch; p.dst = dst; p.dst_pitch = dst_pitch; p.W = w; p.H = h;
    CU(launch_post(POST_RADIAL_BLUR, p, stream_v ? (cudaStream_t)stream_v : ctx->stream));
    return MM_OK;
}

int mm_tonemap_present(mm_ctx *ctx, const float *src, size_t src_pitch, uint8_t *dst8, size_t dst_pitch, int w, int h, int bgra, void *stream_v) {
    int rc = post_check(ctx, "mm_present", nullptr, nullptr, src, src_pitch, dst8, dst_pitch, 4, w, h);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    PostParams p = {};
    p.src = src; p.src_pitch = src_pitch; p.dst8 = dst8; p.dst8_pitch = dst_pitch; p.W = w; p.H = h; p.bgra = bgra != 0;
    CU(launch_post(POST_PRESENT, p, stream_v ? (cudaStream_t)stream_v : ctx->stream));
    return MM_OK;
}

int mm_post_chain(mm_ctx *ctx, const void *camera160, const void *sun116, const float *src, size_t src_pitch, uint8_t *dst8, size_t dst_pitch, int w, int h,
                  int bgra, void *stream_v) {
    int rc = post_check(ctx, "mm_post_chain", camera160, sun116, src, src_pitch, dst8, dst_pitch, 4, w, h);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    size_t need = (size_t)w * h * 4;
    if (ctx->post_plane_bytes < need) {                      // god-ray alpha plane, kept for the next frame
        if (ctx->post_plane) { CU(cudaFree(ctx->post_plane)); ctx->post_plane = nullptr; ctx->post_plane_bytes = 0; }
        CU(cudaMalloc(&ctx->post_plane, need));
        ctx->post_plane_bytes = need;
    }
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
    PostParams p = {};
    post_params((const float *)camera160, (const float *)sun116, p);
    p.src = src; p.src_pitch = src_pitch; p.plane = ctx->post_plane; p.plane_pitch = (size_t)w * 4;
    p.dst8 = dst8; p.dst8_pitch = dst_pitch; p.W = w; p.H = h; p.bgra = bgra != 0;
    CU(cudaEventRecord(ctx->ev0, stream));
    CU(launch_post(POST_GOD_RAY_ALPHA, p, stream));
    CU(launch_post(POST_BLUR_PRESENT, p, stream));
    CU(cudaEventRecord(ctx->ev1, stream));
    ctx->timed = true;
    return MM_OK;
}

// ---- cloud-shadow march of the mesh shader (model.frag:240-283) ------------------------------------------------------
int mm_cloud_shadow(mm_ctx *ctx, const float *positions_xyz, int n, int on_device, float *out_density, uint32_t *out_fetches, void *stream_v) {
    if (!ctx) return MM_ERR_ARG;
    if (!positions_xyz || !out_density || n < 0) return fail(ctx, MM_ERR_ARG, "mm_cloud_shadow: null array or negative count");
    if (!ctx->have_uniforms) return fail(ctx, MM_ERR_STATE, "mm_cloud_shadow: no uniforms set (mm_set_uniforms)");
    if (!ctx->tex[TEX_PLACEMENT].obj || !ctx->tex[TEX_LOWRES].obj) return fail(ctx, MM_ERR_STATE, "mm_cloud_shadow: cloudPlacement and lowResCloudShape must be bound");
    if (n == 0) return MM_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : ctx->stream;
    if (ctx->filter != MM_FILTER_HW) {
        int rc = ensure_pairs(ctx, TEX_PLACEMENT);
        if (rc == MM_OK) rc = ensure_pairs(ctx, TEX_LOWRES);
        if (rc) return rc;
    }
    ShadowParams p = {};
    memcpy(p.cam, ctx->cam, sizeof p.cam); memcpy(p.sun, ctx->sun, sizeof p.sun); memcpy(p.sky, ctx->sky, sizeof p.sky);
    {   // L = normalize((camera.view * vec4(sun.directionBasis[1].xyz, 0)).xyz); if (L.y < -0.05) L *= -1  (model.frag:216-217)
        const float *c = ctx->cam, *d = ctx->sun + 16;
        float l[3];
        for (int i = 0; i < 3; i++) l[i] = dot4_ordered(c[0 + i], d[0], c[4 + i], d[1], c[8 + i], d[2], c[12 + i], 0.0f);
        volatile float xx = l[0] * l[0], yy = l[1] * l[1], zz = l[2] * l[2];
        volatile float dd = xx + yy;
        dd = dd + zz;
        volatile float inv = 1.0f / sqrtf(dd);
        for (int i = 0; i < 3; i++) { volatile float v = l[i] * inv; p.L[i] = v; }
        if (p.L[1] < -0.05f) for (int i = 0; i < 3; i++) { volatile float v = -1.0f * p.L[i]; p.L[i] = v; }
    }
    const int slots[2] = {TEX_PLACEMENT, TEX_LOWRES};
    TexDev *td[2] = {&p.placement, &p.lowres};
    for (int k = 0; k < 2; k++) {
        const TexSlot &s = ctx->tex[slots[k]];
        *td[k] = TexDev{s.pairs, s.obj, s.w, s.h, s.d, (float)s.w, (float)s.h, (float)s.d, is_pow2(s.w) && is_pow2(s.h) && is_pow2(s.d)};
    }
    p.n = n;
    if (on_device) {
        p.pos = positions_xyz; p.out = out_density; p.fetches = out_fetches;
        CU(cudaEventRecord(ctx->ev0, stream));
        CU(launch_cloud_shadow(p, ctx->filter == MM_FILTER_HW ? FILTER_HW : FILTER_EXACT, stream));
        CU(cudaEventRecord(ctx->ev1, stream));
        ctx->timed = true;
        return MM_OK;
    }
    float *d = nullptr;
    CU(cudaMalloc(&d, (size_t)n * 20 + 16));                 // positions (12 B) + density (4 B) + fetches (4 B) per point
    p.pos = d; p.out = d + 3 * (size_t)n; p.fetches = out_fetches ? reinterpret_cast<uint32_t *>(d + 4 * (size_t)n) : nullptr;
    cudaError_t e = cudaMemcpyAsync(d, positions_xyz, (size_t)n * 12, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = launch_cloud_shadow(p, ctx->filter == MM_FILTER_HW ? FILTER_HW : FILTER_EXACT, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_density, p.out, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && out_fetches) e = cudaMemcpyAsync(out_fetches, p.fetches, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "mm_cloud_shadow: %s", cudaGetErrorString(e));
    return MM_OK;
}

int mm_enable_counters(mm_ctx *ctx, int enable) {
    if (!ctx) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->counters_on = enable != 0;
    return size_counters(ctx);
}

int mm_read_counters(mm_ctx *ctx, uint32_t *host_out) {
    if (!ctx || !host_out) return MM_ERR_ARG;
    if (!ctx->counters) return fail(ctx, MM_ERR_STATE, "mm_read_counters: counters are not enabled");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(host_out, ctx->counters, (size_t)ctx->W * ctx->H * 16, cudaMemcpyDeviceToHost));
    return MM_OK;
}

int mm_sample(mm_ctx *ctx, int slot, int filter, const float *uvw_host, int n, float *out_host) {
    if (!ctx || !uvw_host || !out_host || n < 0) return MM_ERR_ARG;
    if (slot < 0 || slot >= TEX_COUNT || !ctx->tex[slot].obj) return fail(ctx, MM_ERR_STATE, "mm_sample: slot %d not bound", slot);
    if (filter != MM_FILTER_EXACT && filter != MM_FILTER_HW) return fail(ctx, MM_ERR_ARG, "mm_sample: filter must be EXACT or HW");
    CU(cudaSetDevice(ctx->device));
    if (filter == MM_FILTER_EXACT) { int rc = ensure_pairs(ctx, slot); if (rc) return rc; }
    float *duvw = nullptr; float4 *dout = nullptr;
    CU(cudaMalloc(&duvw, (size_t)n * 12 + 16));
    cudaError_t e = cudaMalloc(&dout, (size_t)n * 16 + 16);
    const TexSlot &s = ctx->tex[slot];
    TexDev t = {s.pairs, s.obj, s.w, s.h, s.d, (float)s.w, (float)s.h, (float)s.d, is_pow2(s.w) && is_pow2(s.h) && is_pow2(s.d)};
    if (e == cudaSuccess) e = cudaMemcpyAsync(duvw, uvw_host, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_sample_probe(t, s.is3d, slot == MM_TEX_PLACEMENT, filter, duvw, n, dout, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(duvw); cudaFree(dout);
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "mm_sample: %s", cudaGetErrorString(e));
    return MM_OK;
}

int mm_measure_tex_peak(mm_ctx *ctx, int slot, int iters, float *ms_out, double *quads_per_second_out) {
    if (!ctx || !ms_out || !quads_per_second_out) return MM_ERR_ARG;
    if (slot < 0 || slot >= TEX_COUNT || !ctx->tex[slot].obj) return fail(ctx, MM_ERR_STATE, "mm_measure_tex_peak: slot %d not bound", slot);
    if (iters < 2 || iters > (1 << 20)) return fail(ctx, MM_ERR_ARG, "mm_measure_tex_peak: iters %d out of range", iters);
    CU(cudaSetDevice(ctx->device));
    const TexSlot &t = ctx->tex[slot];
    const int blocks = 148 * 8;                                   // one wave of 256-thread blocks on every SM
    iters += iters & 1;
    float4 *sink = nullptr;
    CU(cudaMalloc(&sink, (size_t)blocks * 256 * sizeof(float4)));
    cudaError_t e = launch_tex_peak(t.obj, t.is3d ? 1 : 0, t.w, 64, blocks, sink, ctx->stream);        // warm-up: fills L1, loads the module
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev0, ctx->stream);
    if (e == cudaSuccess) e = launch_tex_peak(t.obj, t.is3d ? 1 : 0, t.w, iters, blocks, sink, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev1, ctx->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev1);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    cudaFree(sink);
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "mm_measure_tex_peak: %s", cudaGetErrorString(e));
    *ms_out = ms;
    double fetches = (double)blocks * 256.0 * (double)iters;
    *quads_per_second_out = ms > 0.0f ? fetches * (t.is3d ? 2.0 : 1.0) / ((double)ms * 1e-3) : 0.0;
    return MM_OK;
}

int mm_selftest_div(mm_ctx *ctx, int which, float *constant_out, unsigned long long *mismatches_out) {
    if (!ctx || !constant_out || !mismatches_out) return MM_ERR_ARG;
    if (which < 0 || which >= selftest_div_count()) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CU(cudaMalloc(&d, 8));
    cudaError_t e = cudaMemsetAsync(d, 0, 8, ctx->stream);
    if (e == cudaSuccess) e = launch_selftest_div(which, constant_out, d, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mismatches_out, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "mm_selftest_div: %s", cudaGetErrorString(e));
    return MM_OK;
}

int mm_alloc_device(mm_ctx *ctx, size_t bytes, void **dptr_out) {
    if (!ctx || !dptr_out || bytes == 0) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(dptr_out, bytes));
    return MM_OK;
}

int mm_free_device(mm_ctx *ctx, void *dptr) {
    if (!ctx || !dptr) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(dptr));
    return MM_OK;
}

int mm_ipc_get_handle(mm_ctx *ctx, void *dptr, uint8_t handle_out[64]) {
    if (!ctx || !dptr || !handle_out) return MM_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle_out, &h, 64);
    return MM_OK;
}

int mm_ipc_open_handle(mm_ctx *ctx, const uint8_t handle[64], void **dptr_out) {
    if (!ctx || !handle || !dptr_out) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return MM_OK;
}

int mm_ipc_close_handle(mm_ctx *ctx, void *dptr) {
    if (!ctx || !dptr) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaIpcCloseMemHandle(dptr));
    return MM_OK;
}

int mm_det_pow(mm_ctx *ctx, const float *x, const float *y, int n, float *out) {
    if (!ctx || !x || !y || !out || n < 0) return MM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    float *d = nullptr;
    CU(cudaMalloc(&d, (size_t)n * 12 + 16));
    cudaError_t e = cudaMemcpyAsync(d, x, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, y, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = (ctx->arith == MM_ARITH_FMA ? launch_det_pow_fma : launch_det_pow)(d, d + n, n, d + 2 * (size_t)n, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, MM_ERR_CUDA, "mm_det_pow: %s", cudaGetErrorString(e));
    return MM_OK;
}

}  // extern "C"
