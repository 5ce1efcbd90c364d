// common.h -- device-side parameter blocks and kernel launchers shared by the C-ABI (capi.cu)
// and the kernels.  Internal; the public boundary is include/marshmallow.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

// One bound texture.  `pairs` is the EXACT-mode copy: per texel two float4 {A(x),B(x),A(x+1),B(x+1)} for
// the channel pairs (A,B), texel values 0..255 as binary32, wrap baked in, x fastest.  `obj` is the hardware-filtered view of
// the same bytes: uchar4 cudaArray, normalised coordinates, wrap addressing, linear filter, UNORM -> float
// (the reference sampler: Texture.cpp:29-52, 315-338).
struct TexDev {
    const float4 *pairs;         // EXACT-mode copy, pair-major: 2 float4 per texel (cloud_march.cu, "Sampler")
    cudaTextureObject_t obj;
    int w, h, d;
    float wf, hf, df;   // the extents as binary32 (exact)
    int pow2;   // all extents are powers of two -> wrap by mask
};

enum { TEX_PLACEMENT = 0, TEX_NIGHTSKY = 1, TEX_CURL = 2, TEX_LOWRES = 3, TEX_HIRES = 4, TEX_COUNT = 5 };
enum { FILTER_EXACT = 0, FILTER_HW = 1, FILTER_HYBRID = 2 };
enum { DISPATCH_FULL = 0, DISPATCH_PHASE16 = 1 };
// row block index of the k-th block owned by partition `begin` of `stride` (cyclic, or boustrophedon when snake)
__host__ __device__ inline int owned_block(int k, int begin, int stride, int snake) {
    return k * stride + ((snake && (k & 1)) ? (stride - 1 - begin) : begin);
}
__host__ __device__ inline bool owns_block(int blk, int begin, int stride, int snake) {
    int k = blk / stride, m = blk - k * stride;
    return ((snake && (k & 1)) ? (stride - 1 - m) : m) == begin;
}

// Pixel tile of one warp: MM_TILE_W x (32 / MM_TILE_W); a block is MM_WARPS_X x MM_WARPS_Y warps.
#ifndef MM_TILE_W
#define MM_TILE_W 8
#endif
#ifndef MM_WARPS_X
#define MM_WARPS_X 2
#endif
#ifndef MM_WARPS_Y
#define MM_WARPS_Y 2
#endif
enum { TILE_W = MM_TILE_W, TILE_H = 32 / MM_TILE_W, WARPS_X = MM_WARPS_X, WARPS_Y = MM_WARPS_Y, BLOCK_W = WARPS_X * TILE_W, BLOCK_H = WARPS_Y * TILE_H };

struct MarchParams {
    float cam[40];   // UniformCameraObject (Shader.h:24-29)
    float sun[29];   // UniformSunObject    (SkyManager.h:8-14)
    float sky[13];   // UniformSkyObject    (SkyManager.h:28-36)
    float light[18]; // the six cone-sample offsets mat3(sun.directionBasis) * s_i (CC:392-401): uniform per launch,
                     // evaluated once on the host in the same binary32 order (capi.cu, light_cone_samples)
    TexDev tex[TEX_COUNT];
    float *out;                  // pitch-linear float4 image (may be peer memory), or nullptr when surf is used
    size_t pitch;                // bytes
    float *mirror;               // optional second destination of every pixel store (pinned host memory mapped into the
    size_t mirror_pitch;         // device address space: the device->host transfer fused into the kernel), or nullptr
    cudaSurfaceObject_t surf;    // external (Vulkan) image, when out == nullptr
    uint32_t *counters;          // 4 x uint32 per pixel or nullptr
    int W, H;
    int mode;                    // DISPATCH_*
    int row_begin, row_stride, row_block;
    int row_snake;               // odd rounds of the row-cyclic assignment run in reverse rank order (MM_ROWS_SNAKE)
    int owned_rows;              // rows this dispatch enumerates (FULL), virtual rows (PHASE16)
    int grid_w;                  // pixel columns enumerated (W, or ceil(W/4) in PHASE16)
    // Execution order of the BLOCK_H-row block rows, most expensive first (rays nearest the horizon cross the longest
    // stretch of the cloud shell; rays below it are free): blockIdx.y -> block row.  Keeps the tail of the launch
    // cheap, which is what limits strong scaling when a GPU owns only a few waves of tiles.  Built on the host
    // per dispatch (capi.cu, order_block_rows); affects scheduling only, never results.
    uint16_t block_row_order[4096];
    // K1p (persistent warps, dynamic queue): one counter in device memory, zeroed on the stream before the launch; slots are
    // numbered tile-major (slot = tile * 32 + lane-in-tile, tiles_x tiles per tile row, tile rows in block_row_order)
    unsigned *queue;
    unsigned n_slots, tiles_x;
    int launch_block_rows;       // host side: entries of block_row_order this launch runs (0 = all of them); sizes grid.y of K1 / K1s / K1x2
};

struct ReprojectParams {
    float cam[40], cam_prev[40];     // UniformCameraObject, UniformCameraObjectPrev (reproject.comp:12-23)
    const float *src; size_t src_pitch;   // sourceImage (previous frame), pitch-linear float4
    float *dst; size_t dst_pitch;         // targetImage
    int W, H;
};
// K6 post chain (post_chain.cu): god-ray.frag / radialBlur.frag / tonemap.frag.  The sun's screen position
// ((proj*view)*sun.location, divided by w) and sun.color.xyz*sun.intensity are uniform per frame and evaluated on the
// host in the oracle's binary32 order (capi.cu, post_params).
struct PostParams {
    const float *src; size_t src_pitch;       // RGBA32F input of the pass (pitch-linear float4)
    float *dst; size_t dst_pitch;             // RGBA32F output (pass-by-pass kernels)
    float *plane; size_t plane_pitch;         // god-ray alpha plane (fused chain): written by pass 1, tapped by pass 2
    unsigned char *dst8; size_t dst8_pitch;   // UNORM8 output (present / fused chain)
    float sun_x, sun_y, sun_dir_y;
    float sun_rgb[3];
    int W, H, bgra;
};
enum { POST_GOD_RAY = 0, POST_GOD_RAY_ALPHA = 1, POST_RADIAL_BLUR = 2, POST_BLUR_PRESENT = 3, POST_PRESENT = 4 };
cudaError_t launch_post(int which, const PostParams &p, cudaStream_t stream);

// K7 cloud-shadow march of the mesh shader (model.frag:240-283) for n world positions (cloud_march.cu)
struct ShadowParams {
    float cam[40], sun[29], sky[13];
    float L[3];                               // view-space sun direction, flipped when L.y < -0.05 (model.frag:216-217); host
    TexDev placement, lowres;
    const float *pos;                         // n x (x, y, z) world positions (fragPositionWC)
    float *out;                               // n x accumDensity
    uint32_t *fetches;                        // optional: texture() calls per point
    int n;
};
cudaError_t launch_cloud_shadow(const ShadowParams &p, int filter, cudaStream_t stream);

cudaError_t launch_reproject(const ReprojectParams &p, cudaStream_t stream);
// persistent_blocks > 0: K1p with that many 128-thread blocks (lanes_per_ray must be 1, p.queue zeroed on `stream`); refill 32/16/8
// persistent_blocks < 0: K1x2, two rays per thread on packed FP32 (FILTER_HW, no counters, lanes_per_ray 1; block rows of 8)
cudaError_t launch_cloud_march(const MarchParams &p, int filter, int lanes_per_ray, int persistent_blocks, int refill, cudaStream_t stream);
int persistent_blocks_per_sm(int filter);
// the same kernels built under the contracted arithmetic definition (cloud_march_fma.cu)
cudaError_t launch_cloud_march_fma(const MarchParams &p, int filter, int lanes_per_ray, int persistent_blocks, int refill, cudaStream_t stream);
cudaError_t launch_det_pow_fma(const float *x, const float *y, int n, float *out, cudaStream_t stream);
void march_block_shape(int lanes_per_ray, int *block_w, int *block_h);   // pixels per block of the variant that will run
cudaError_t launch_sample_probe(const TexDev &t, int is3d, int placement_layout, int filter, const float *uvw, int n, float4 *out, cudaStream_t stream);
cudaError_t launch_tex_peak(cudaTextureObject_t obj, int is3d, int width, int iters, int blocks, float4 *sink, cudaStream_t stream);
cudaError_t launch_det_pow(const float *x, const float *y, int n, float *out, cudaStream_t stream);
cudaError_t launch_pack_pairs(const uchar4 *src, float4 *dst, int w, int h, int d, int placement_layout, cudaStream_t stream);
int selftest_div_count();
cudaError_t launch_selftest_div(int which, float *c_out, unsigned long long *mismatches_dev, cudaStream_t stream);
cudaError_t launch_tonemap(const float *src, size_t pitch, int W, int H, uchar4 *dst, cudaStream_t stream);
cudaError_t launch_curl_noise(uchar4 *dst128x128, float *scratch /* 15*128*128 + 8 floats */, const unsigned char *gradient_table /* 26^3, device */, cudaStream_t stream);
cudaError_t launch_noise_volumes(uint32_t seed, uchar4 *low128, uchar4 *hi32, cudaStream_t stream);

}  // namespace mm
