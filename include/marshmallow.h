/*
 * marshmallow.h -- C-ABI of the B200-native cloud ray-march pass and noise-texture build.
 *
 * This is the drop-in boundary for ONE path of mccannd/Project-Marshmallow: the compute dispatch of
 * SkyEngine/SkyEngine/Shaders/compute-clouds.comp and the textures it samples.  Each entry point
 * names the reference interface it replaces (paths relative to /root/reference/SkyEngine/SkyEngine).
 * Plain C types only: opaque handle, pointers, sizes, int status (0 = MM_OK, negative = error;
 * the reference throws std::runtime_error instead, main.cpp:8-14).  No exceptions cross this
 * boundary.  One context per GPU; a context is not thread-safe (the reference is single-threaded).
 * There is NO CPU fallback: every compute entry fails with MM_ERR_CUDA when no sm_100 device is
 * usable.
 */
#ifndef MARSHMALLOW_H
#define MARSHMALLOW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_API __attribute__((visibility("default")))

typedef struct mm_ctx mm_ctx;

enum mm_status {
    MM_OK = 0,
    MM_ERR_ARG = -1,        /* null / out-of-range argument */
    MM_ERR_CUDA = -2,       /* CUDA runtime error or no usable device; see mm_last_error */
    MM_ERR_STATE = -3,      /* a required texture / uniform / output was never bound */
    MM_ERR_UNSUPPORTED = -4
};

/* texture slots = the sampler bindings of compute-clouds.comp:43-47 (set 2, bindings 4..8) */
enum mm_tex_slot {
    MM_TEX_PLACEMENT = 0,   /* binding 4 cloudPlacement   (VulkanApplication.cpp:253-254) */
    MM_TEX_NIGHTSKY = 1,    /* binding 5 nightSkyMap      (VulkanApplication.cpp:255-256) */
    MM_TEX_CURL = 2,        /* binding 6 curlNoise        (VulkanApplication.cpp:257-258) */
    MM_TEX_LOWRES = 3,      /* binding 7 lowResCloudShape (VulkanApplication.cpp:259-260) */
    MM_TEX_HIRES = 4        /* binding 8 hiResCloudShape  (VulkanApplication.cpp:261-262) */
};

/* which pixels one dispatch marches */
enum mm_dispatch_mode {
    MM_FULL = 0,            /* every pixel: the union of the reference's 16 phase dispatches */
    MM_PHASE16 = 1,         /* one reference dispatch: pixels (4*gx + o%4, 4*gy + o/4), o = int(sun.color.a)
                               (compute-clouds.comp:291-301, VulkanApplication.cpp:1067-1071) */
    MM_ROWS_SNAKE = 0x100   /* flag, OR-ed to a mode: the row-cyclic partition runs boustrophedon -- round k of the assignment
                               hands its row_stride blocks out in rank order for even k and in reverse rank order for odd k,
                               which cancels the systematic cost gradient between ranks (rows nearer the horizon cost more) */
};

/* sampler arithmetic (see DESIGN.md "sampler modes").  Vulkan leaves linear-filter precision to the implementation;
 * each mode is one definition of it, and each has a bit-exact CPU statement in oracle/ that its decisions are tested
 * against (zero branch flips, bit-identical alpha). */
enum mm_filter_mode {
    MM_FILTER_EXACT = 0,    /* binary32 software filtering of every fetch (oracle sampler OM_FILTER_FP32) */
    MM_FILTER_HW = 1,       /* DEFAULT: the texture unit filters every fetch (cudaTextureObject; 8-bit corner weights,
                               UNORM16 result) -- what the reference shader gets on this GPU; oracle sampler
                               OM_FILTER_TEXUNIT is a bit-exact integer model of the unit */
    MM_FILTER_HYBRID = 2    /* march decisions: EXACT; the 6 light-cone samples: texture unit */
};

/* arithmetic definition of the march (see DESIGN.md "arithmetic definitions").  GLSL fixes neither the rounding of `a*b + c` nor
 * that of pow(); each value below is one complete definition with a bit-exact CPU statement in oracle/ that is pinned to the
 * reference's own shader text and against which the kernel's decisions are tested (zero branch flips, bit-identical alpha). */
enum mm_arith_mode {
    MM_ARITH_IEEE = 0,      /* DEFAULT: every operator rounded once, no contraction (oracle/cloud_march_oracle.c, glsl_env.h) */
    MM_ARITH_FMA = 1        /* the contraction GLSL permits and GPUs perform, as one lexical rule: a product that is directly an
                               operand of + or - is fused into it (oracle/cloud_march_oracle_fma.c, glsl_env_fma.h) */
};

/* ---- context: replaces ComputeShader's constructor/destructor (Shader.h:338-353, Shader.cpp:633-645) */
MM_API int mm_create(int device, mm_ctx **out);
MM_API int mm_destroy(mm_ctx *ctx);
MM_API const char *mm_last_error(const mm_ctx *ctx);        /* never NULL; ctx may be NULL for create errors */
MM_API const char *mm_version(void);

/* ---- textures: replace Texture::initFromFile / Texture3D::initFromFile + sampler creation
 *      (Texture.cpp:212-246, 502-538; samplers Texture.cpp:29-52, 315-338: LINEAR, REPEAT, 1 mip).
 *      rgba8 is HOST memory, RGBA8_UNORM, x fastest, then y, then z ([z][y][x][4]). */
MM_API int mm_upload_tex2d(mm_ctx *ctx, int slot, const uint8_t *rgba8, int w, int h);
MM_API int mm_upload_tex3d(mm_ctx *ctx, int slot, const uint8_t *rgba8, int w, int h, int d);

/* ---- noise-texture build on the GPU.
 * mm_build_curl_noise replaces the offline GenerateCurlNoise (ImageUtils.cpp:176-223): builds the
 * 128x128 RGBA8 curl-noise FBM, binds it to MM_TEX_CURL, and optionally copies it to host memory.
 * mm_build_noise_volumes builds the 128^3 Perlin-Worley/Worley-FBM and 32^3 Worley-FBM volumes the
 * reference ships as Houdini-baked TGA slices (no reference generator exists), binds them to
 * MM_TEX_LOWRES / MM_TEX_HIRES, and optionally copies them out ([z][y][x][4] bytes). */
MM_API int mm_build_curl_noise(mm_ctx *ctx, uint8_t *out_rgba8_128x128_or_null);
MM_API int mm_build_noise_volumes(mm_ctx *ctx, uint64_t seed, uint8_t *out_low128_or_null, uint8_t *out_hi32_or_null);

/* ---- uniforms: replaces ComputeShader::updateUniformBuffers (Shader.cpp:967-992).  The four blocks
 * are byte-identical to the reference structs: UniformCameraObject 160 B (Shader.h:24-29; view@0
 * proj@64 cameraPosition@128 cameraParams@144), UniformSunObject 116 B (SkyManager.h:8-14; location@0
 * direction@16 color@32 directionBasis@48 intensity@112), UniformSkyObject 52 B (SkyManager.h:28-36;
 * betaR@0 betaV@16 wind@32 mie_directional@48).  camera_prev may be NULL (unused by the march). */
MM_API int mm_set_uniforms(mm_ctx *ctx, const void *camera160, const void *camera_prev160_or_null,
                           const void *sun116, const void *sky52);

/* ---- output image: replaces descriptor set 0 (resultImage, rgba32f; VulkanApplication.cpp:247-250).
 * mm_bind_output_linear: a DEVICE pointer to pitch-linear float4 pixels (pitch in bytes, >= 16*w).
 *   The pointer may be peer memory of another GPU (multi-GPU gather is fused into the stores).
 * mm_bind_output_external_fd: a VkImage's exported opaque-fd memory (R32G32B32A32_SFLOAT, optimal
 *   tiling) imported with cudaImportExternalMemory -> mipmapped array -> surface object.
 *   Compile-checked only in this environment (no Vulkan loader exists here).
 * mm_alloc_output: convenience, allocates the image inside the context. */
MM_API int mm_bind_output_linear(mm_ctx *ctx, float *dptr_rgba32f, size_t pitch_bytes, int w, int h);
MM_API int mm_bind_output_external_fd(mm_ctx *ctx, int opaque_fd, size_t alloc_bytes, int w, int h);
/* the same import for LINEARLY laid out external memory (VkBuffer, or a VK_IMAGE_TILING_LINEAR image: pitch from
 * vkGetImageSubresourceLayout): mapped as a device buffer (cudaExternalMemoryGetMappedBuffer) and written with plain stores.
 * On success the library owns the fd.  Exercised in tests with an fd exported by CUDA's own virtual-memory allocator. */
MM_API int mm_bind_output_external_buffer_fd(mm_ctx *ctx, int opaque_fd, size_t alloc_bytes, size_t offset_bytes, size_t pitch_bytes,
                                             int w, int h);
/* ---- ordering against the engine's queues.  The reference submits the compute work with no fence or semaphore
 * (VulkanApplication.cpp:168-177) and idles the present queue every frame (:237); a CUDA producer imports the engine's exported
 * VkSemaphores (opaque fd; timeline != 0 for a timeline semaphore, `value` is then the payload) and waits / signals them on the
 * dispatch stream:  mm_wait_semaphore(image_free); mm_dispatch(...); mm_signal_semaphore(image_ready).  Slots 0..7 per context. */
MM_API int mm_import_semaphore_fd(mm_ctx *ctx, int opaque_fd, int timeline, int *slot_out);
MM_API int mm_wait_semaphore(mm_ctx *ctx, int slot, uint64_t value, void *stream);
MM_API int mm_signal_semaphore(mm_ctx *ctx, int slot, uint64_t value, void *stream);
MM_API int mm_release_semaphore(mm_ctx *ctx, int slot);
MM_API int mm_alloc_output(mm_ctx *ctx, int w, int h, float **dptr_out, size_t *pitch_out);

/* ---- dispatch: replaces bindShader + vkCmdDispatch + vkQueueSubmit (Shader.h:358-376,
 * VulkanApplication.cpp:1062-1071, 168-177).  Asynchronous on `stream` (a cudaStream_t, or NULL for
 * the context's own stream).  Rows are partitioned in blocks of row_block rows; this call marches
 * block b when b % row_stride == row_begin; 0 <= row_begin < row_stride is required (single GPU: 0,1,1).
 * MM_PHASE16 requires sun.color.a in [0,16) (the engine keeps it there, VulkanApplication.cpp:384). */
MM_API int mm_set_filter_mode(mm_ctx *ctx, int filter_mode);
MM_API int mm_set_arithmetic(mm_ctx *ctx, int arith_mode);   /* applies to mm_dispatch / mm_render_to_host and to mm_det_pow */
/* scheduling knob, never changes results: lanes that share one ray.  1 = one thread per ray; 2, 4, 8 = that many
 * consecutive loop trips of a ray are evaluated side by side and replayed through the loop's state machine in order
 * (shortens the dependent chain of a ray 1.9x / 3.7x / 6.7x for 2 / 8 / 17 % more density evaluations); 0 (default) =
 * chosen per dispatch from its size: small dispatches (MM_PHASE16, row-sharded frames on several GPUs) get more lanes. */
MM_API int mm_set_lanes_per_ray(mm_ctx *ctx, int lanes);
/* scheduling knob, never changes results: how the pixels of a dispatch reach the warps (one lane per ray only; the ray-split
 * kernels keep a static grid).  Replaces the fixed workgroup grid of VulkanApplication.cpp:1062-1071.
 *   MM_SCHED_STATIC      one thread block per 16x8 pixel tile, block rows launched in cost order
 *   MM_SCHED_PERSISTENT  one resident wave of warps pulling 8x4 pixel tiles, most expensive first, from an atomic queue in device
 *                        memory; refill_lanes = 32 (or 0): a warp takes a new tile when all its rays are done; 16 / 8: as soon as
 *                        that many lanes hold finished rays, those lanes store their pixels and are refilled from the queue
 *   MM_SCHED_PACKED      K1x2: static grid, TWO neighbouring rays per thread on packed FP32 instructions (FADD2 / FFMA2): every arithmetic
 *                        instruction of the decision path serves both rays; texture-unit mode without counters (anything else runs K1)
 *   MM_SCHED_AUTO        (default) MM_SCHED_STATIC: measured at least as fast on every workload */
enum mm_scheduler { MM_SCHED_AUTO = 0, MM_SCHED_STATIC = 1, MM_SCHED_PERSISTENT = 2, MM_SCHED_PACKED = 3 };
MM_API int mm_set_scheduler(mm_ctx *ctx, int scheduler, int refill_lanes);
MM_API int mm_dispatch(mm_ctx *ctx, int mode, int row_begin, int row_stride, int row_block, void *stream);
/* one frame over n contexts of one process (one per GPU): context i marches partition i of n (row blocks of row_block rows) on
 * streams[i] (NULL: each context's own stream).  Returns the first failing context's status. */
MM_API int mm_dispatch_multi(mm_ctx **ctxs, int n, int mode, int row_block, void **streams);
/* contexts on DIFFERENT devices of one process that store into one image: enable peer access from ctx's device to peer's
 * (no-op for the same device; across processes mm_ipc_open_handle does it) */
MM_API int mm_enable_peer(mm_ctx *ctx, mm_ctx *peer);
MM_API int mm_synchronize(mm_ctx *ctx);
/* host-only diagnostics (no device needed): the order in which mm_dispatch would execute the block rows (block_h rows of the
 * partition's compact row index each) of such a dispatch, most expensive first; order_out needs ceil(owned_rows / block_h)
 * entries (at most 4096).  The cost order is a scheduling hint and never changes results. */
MM_API int mm_plan_block_rows(const void *camera160, int h, int mode, int row_begin, int row_stride, int row_block, int block_h,
                              uint16_t *order_out, int *count_out);

/* ---- reprojection pass: replaces the ReprojectShader dispatch that precedes the cloud dispatch in the engine's
 * frame (Shaders/reproject.comp; Shader.h:380-452; VulkanApplication.cpp:1053-1059).  Reads the PREVIOUS frame's
 * image (descriptor set 1, backgroundTexturePrev) and writes the bound output image; needs camera and previous
 * camera blocks from mm_set_uniforms.  A frame of the engine's cadence is
 *     mm_set_uniforms; mm_dispatch_reproject; mm_dispatch(MM_PHASE16); swap the two images.
 * (On one stream the two launches are ordered -- the reference records no barrier between them.) */
MM_API int mm_bind_previous_linear(mm_ctx *ctx, const float *dptr_prev_rgba32f, size_t pitch_bytes);
MM_API int mm_dispatch_reproject(mm_ctx *ctx, void *stream);

/* host-buffer convenience (the end-to-end call): uniforms in, march, image out to HOST memory.
 * out_host: w*h*4 floats (packed).  Includes H2D of the uniforms and D2H of the image.  When out_host is
 * page-locked (cudaHostAlloc / cudaHostRegister) and mode is MM_FULL the D2H transfer is fused into the kernel
 * (each pixel is stored to the device image and to out_host); otherwise the image is copied after the march. */
MM_API int mm_render_to_host(mm_ctx *ctx, const void *camera160, const void *sun116, const void *sky52,
                             int mode, float *out_host_rgba32f);

/* ---- host frame for several GPUs (new; the end-to-end path of a sharded frame).  A packed w*h RGBA32F frame in HOST memory that
 * is page-locked and mapped into this context's device (mm_host_register on memory the caller owns -- e.g. a POSIX shared-memory
 * mapping that every rank's process opens -- or cudaHostAlloc).  Once bound, every mm_dispatch stores each finished pixel to
 * the device image AND to this frame over PCIe, so the device->host transfer of an N-GPU frame runs on N links in parallel and
 * overlaps the march; the frame is complete when every rank's stream has drained.  NULL unbinds. */
MM_API int mm_host_register(mm_ctx *ctx, void *host, size_t bytes);
MM_API int mm_host_unregister(mm_ctx *ctx, void *host);
MM_API int mm_bind_host_mirror(mm_ctx *ctx, float *host_rgba32f_packed);

/* ---- HDR -> RGBA8 (tonemap.frag:11-28, vignette omitted) of the bound output; device or host dst */
MM_API int mm_tonemap_rgba8(mm_ctx *ctx, uint8_t *dst, int dst_is_device, void *stream);

/* ---- post chain: replaces the three PostProcessShader passes that consume the cloud image
 * (Shader.h:460-520; VulkanApplication.cpp:309-317 construction, :943-968 and :1016 recording):
 *     god-ray.frag:41-76 -> radialBlur.frag:36-63 -> tonemap.frag:11-33 -> swapchain (B8G8R8A8_UNORM, :1436-1446).
 * camera160 carries view AND proj here (the passes project sun.location to the screen).  Images are DEVICE memory,
 * pitch-linear: RGBA32F (16-byte aligned) or 8-bit RGBA/BGRA.  Source and destination must differ.
 *   mm_god_ray / mm_radial_blur / mm_tonemap_present  one reference pass each (framebuffer in, framebuffer out)
 *   mm_post_chain                                     the whole chain in two kernels, byte-identical to running the
 *                                                     three passes in sequence: cloud image in, swapchain bytes out */
MM_API int mm_god_ray(mm_ctx *ctx, const void *camera160, const void *sun116, const float *src_rgba32f, size_t src_pitch,
                      float *dst_rgba32f, size_t dst_pitch, int w, int h, void *stream);
MM_API int mm_radial_blur(mm_ctx *ctx, const void *camera160, const void *sun116, const float *src_rgba32f, size_t src_pitch,
                          float *dst_rgba32f, size_t dst_pitch, int w, int h, void *stream);
MM_API int mm_tonemap_present(mm_ctx *ctx, const float *src_rgba32f, size_t src_pitch, uint8_t *dst_8888, size_t dst_pitch,
                              int w, int h, int bgra, void *stream);
MM_API int mm_post_chain(mm_ctx *ctx, const void *camera160, const void *sun116, const float *src_rgba32f, size_t src_pitch,
                         uint8_t *dst_8888, size_t dst_pitch, int w, int h, int bgra, void *stream);

/* ---- cloud shadows: the 6-step march of the low-res cloud field that the mesh shader runs per fragment
 * (model.frag:240-283, helpers :58-140) as a standalone pass over n world positions (fragPositionWC): a G-buffer
 * position image or a shadow-map grid.  Uses the uniforms of mm_set_uniforms (camera view + position, sun basis,
 * wind) and the bound cloudPlacement / lowResCloudShape textures; the filter mode selects the sampler.
 * out_density[i] = accumDensity of model.frag:266-272; the shader applies it as color *= 1 - 2*accumDensity (:281).
 * on_device != 0: all pointers are device memory and the launch is asynchronous on `stream`; otherwise host arrays. */
MM_API int mm_cloud_shadow(mm_ctx *ctx, const float *positions_xyz, int n, int on_device, float *out_density,
                           uint32_t *out_fetches /* optional */, void *stream);

/* ---- diagnostics: per-pixel work counters {loop trips, 2D fetches, 3D fetches, lit steps}
 * (uint32 x4 per pixel, device memory owned by the context; NULL disables).  Algorithmic counts:
 * what compute-clouds.comp would execute, whether or not the kernel skipped the work. */
MM_API int mm_enable_counters(mm_ctx *ctx, int enable);
MM_API int mm_read_counters(mm_ctx *ctx, uint32_t *host_out /* w*h*4 */);
MM_API int mm_read_output(mm_ctx *ctx, float *host_out /* w*h*4 packed */);
MM_API int mm_last_kernel_ms(mm_ctx *ctx, float *ms);       /* CUDA-event time of the last mm_dispatch */

/* sampler probe: filters `n` coordinates (u,v,w triples) of a slot with the given mode on the GPU */
MM_API int mm_sample(mm_ctx *ctx, int slot, int filter_mode, const float *uvw_host, int n, float *out_rgba_host);
/* the deterministic pow of the decision path, evaluated on the GPU (bit-exactness probe) */
MM_API int mm_det_pow(mm_ctx *ctx, const float *x_host, const float *y_host, int n, float *out_host);

/* texture-pipe ceiling: filtered fetches per second from the texture bound to `slot` when it is L1-resident (use MM_TEX_CURL, 64 KB,
 * for the bilinear figure and MM_TEX_HIRES, 128 KB, for the trilinear one); `iters` fetches per thread of one full wave of blocks.
 * Reported as bilinear-quad operations per second (a trilinear fetch counts as two), the unit of bench.py's texture roofline. */
MM_API int mm_measure_tex_peak(mm_ctx *ctx, int slot, int iters, float *ms_out, double *quads_per_second_out);

/* exhaustive self-test of the exact divide-by-constant sequence the march uses (csrc/cloud_march.cu,
 * div_const): for constant number `which` (0 .. count-1; MM_ERR_ARG beyond) compares it with the IEEE
 * divide over every binary32 dividend and returns the number of mismatching bit patterns (must be 0). */
MM_API int mm_selftest_div(mm_ctx *ctx, int which, float *constant_out, unsigned long long *mismatches_out);

/* ---- multi-GPU plumbing (new work: the reference is single-device, VulkanApplication.cpp:639-687).
 * One process per GPU.  Rank 0 exports the allocation behind its output image as a 64-byte CUDA-IPC
 * handle; the other ranks open it and bind the mapped pointer with mm_bind_output_linear, so their
 * march kernels store finished pixels straight into rank 0's image over NVLink (no gather step).
 * dptr must be the base of an allocation made by mm_alloc_output or mm_alloc_device.  For frame-parallel animation
 * (BASELINE config 5) rank 0 allocates one slot per rank with mm_alloc_device and every rank binds its own slot. */
MM_API int mm_alloc_device(mm_ctx *ctx, size_t bytes, void **dptr_out);    /* plain device allocation (IPC-exportable base) */
MM_API int mm_free_device(mm_ctx *ctx, void *dptr);
MM_API int mm_ipc_get_handle(mm_ctx *ctx, void *dptr, uint8_t handle_out[64]);
MM_API int mm_ipc_open_handle(mm_ctx *ctx, const uint8_t handle[64], void **dptr_out);
MM_API int mm_ipc_close_handle(mm_ctx *ctx, void *dptr);

/* ---- host-side value producers (CPU only; no device needed).  Restate the reference's
 * SkyManager (SkyManager.cpp:15-70) and Camera (camera.cpp:27-39,179-195; camera.h:72) so that a
 * caller without the engine can fill the uniform blocks exactly as VulkanApplication.cpp:351-386. */
MM_API int mm_host_sky(float elevation, float azimuth, float turbidity, float rayleigh, float mie,
                       float mie_directional, const float wind_xyz[3], float time, int pixel_phase,
                       void *sun116_out, void *sky52_out);
MM_API int mm_host_camera(const float position[3], float yaw, float pitch, float fov_deg, float aspect,
                          void *camera160_out);

#ifdef __cplusplus
}
#endif
#endif /* MARSHMALLOW_H */
