// aux_host.cu -- TEST INFRASTRUCTURE: the per-pixel source of K2 (curl_noise_pixel.h), K3 (noise_volume_pixel.h), K4 (tonemap_pixel.h), K5 (reproject_pixel.h) and K6 (post_chain_pixel.h) compiled for the HOST and run pixel by pixel on
// the CPU, so that the CPU test-suite (tests/test_host_build.py) can compare the product's kernel arithmetic with the oracle without a GPU.  The loops below do
// what the kernels of reproject.cu / post_chain.cu do with the same functions; nothing here is linked into the product library.
#define MM_HOST_BUILD 1
#include <cstdint>
#include <vector>

#include "../../project-marshmallow_b200/csrc/curl_noise_pixel.h"
#include "../../project-marshmallow_b200/csrc/curl_table.h"
#include "../../project-marshmallow_b200/csrc/noise_volume_pixel.h"
#include "../../project-marshmallow_b200/csrc/post_chain_pixel.h"
#include "../../project-marshmallow_b200/csrc/reproject_pixel.h"
#include "../../project-marshmallow_b200/csrc/tonemap_pixel.h"

using namespace mm;

static PostParams post(const float *src, int W, int H, float sun_x, float sun_y, float sun_dir_y, const float *sun_rgb) {
    PostParams p = {};
    p.src = src; p.src_pitch = (size_t)W * 16; p.W = W; p.H = H;
    p.sun_x = sun_x; p.sun_y = sun_y; p.sun_dir_y = sun_dir_y;
    for (int k = 0; k < 3; k++) p.sun_rgb[k] = sun_rgb ? sun_rgb[k] : 0.0f;
    return p;
}

extern "C" {

// god_ray_kernel<false>: RGBA32F in, RGBA32F out (rgb passes through)
int hb_god_ray(const float *src, int W, int H, float sun_x, float sun_y, float sun_dir_y, float *dst) {
    PostParams p = post(src, W, H, sun_x, sun_y, sun_dir_y, nullptr);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float *c = src + ((size_t)y * W + x) * 4;
            float *o = dst + ((size_t)y * W + x) * 4;
            o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
            o[3] = post_pixel::god_ray_alpha(p, x, y, c[3]);
        }
    return 0;
}

// radial_blur_kernel<false>
int hb_radial_blur(const float *src, int W, int H, float sun_x, float sun_y, float sun_dir_y, const float *sun_rgb, float *dst) {
    PostParams p = post(src, W, H, sun_x, sun_y, sun_dir_y, sun_rgb);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float *c = src + ((size_t)y * W + x) * 4;
            float3 rgb = post_pixel::radial_blur_rgb<false>(p, x, y, make_float3(c[0], c[1], c[2]));
            float *o = dst + ((size_t)y * W + x) * 4;
            o[0] = rgb.x; o[1] = rgb.y; o[2] = rgb.z; o[3] = 1.0f;
        }
    return 0;
}

// present_kernel
int hb_present(const float *src, int W, int H, int bgra, uint8_t *dst8) {
    PostParams p = post(src, W, H, 0.f, 0.f, 0.f, nullptr);
    p.bgra = bgra != 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float *c = src + ((size_t)y * W + x) * 4;
            uchar4 q = post_pixel::present_pixel(p, x, y, make_float3(c[0], c[1], c[2]));
            uint8_t *o = dst8 + ((size_t)y * W + x) * 4;
            o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
        }
    return 0;
}

// mm_post_chain: god_ray_kernel<true> into an alpha plane, then radial_blur_kernel<true> (taps from the plane, tone map in registers)
int hb_post_chain(const float *src, int W, int H, float sun_x, float sun_y, float sun_dir_y, const float *sun_rgb, int bgra, uint8_t *dst8) {
    PostParams p = post(src, W, H, sun_x, sun_y, sun_dir_y, sun_rgb);
    std::vector<float> plane((size_t)W * H);
    p.plane = plane.data(); p.plane_pitch = (size_t)W * 4; p.bgra = bgra != 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            plane[(size_t)y * W + x] = post_pixel::god_ray_alpha(p, x, y, src[((size_t)y * W + x) * 4 + 3]);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float *c = src + ((size_t)y * W + x) * 4;
            uchar4 q = post_pixel::present_pixel(p, x, y, post_pixel::radial_blur_rgb<true>(p, x, y, make_float3(c[0], c[1], c[2])));
            uint8_t *o = dst8 + ((size_t)y * W + x) * 4;
            o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
        }
    return 0;
}

// reproject_kernel
int hb_reproject(const float *camera160, const float *camera_prev160, const float *src, int W, int H, float *dst) {
    ReprojectParams p;
    for (int i = 0; i < 40; i++) { p.cam[i] = camera160[i]; p.cam_prev[i] = camera_prev160[i]; }
    p.src = src; p.src_pitch = (size_t)W * 16; p.dst = dst; p.dst_pitch = (size_t)W * 16; p.W = W; p.H = H;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float4 t = reproject_pixel::reproject_texel(p, x, y);
            float *o = dst + ((size_t)y * W + x) * 4;
            o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
        }
    return 0;
}

// launch_curl_noise (curl_noise.cu): 12 FBM probes per texel, curl, global per-channel bounds, normalise + quantise -> 128 x 128 RGBA8
int hb_curl_noise(uint8_t *dst_rgba8) {
    const int N = 128 * 128;
    std::vector<unsigned char> table(26 * 26 * 26);
    build_curl_gradient_table(table.data());
    std::vector<float> fbm12((size_t)12 * N), curls((size_t)3 * N);
    float bounds[6] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (int gid = 0; gid < 12 * N; gid++) fbm12[gid] = curl_pixel::curl_probe(table.data(), gid);
    for (int pix = 0; pix < N; pix++) curl_pixel::curl_combine(fbm12.data(), curls.data(), pix);
    for (int pix = 0; pix < N; pix++)                         // curl_bounds_kernel: min / max are order-independent
        for (int c = 0; c < 3; c++) { float v = curls[3 * pix + c]; bounds[c] = fminf(bounds[c], v); bounds[3 + c] = fmaxf(bounds[3 + c], v); }
    for (int pix = 0; pix < N; pix++) {
        uchar4 q = curl_pixel::curl_quantise(curls.data(), bounds, pix);
        dst_rgba8[4 * pix] = q.x; dst_rgba8[4 * pix + 1] = q.y; dst_rgba8[4 * pix + 2] = q.z; dst_rgba8[4 * pix + 3] = q.w;
    }
    return 0;
}

// lowres_kernel / hires_kernel (noise_volumes.cu): the 32^3 volume whole, and the slices z = z0, z0 + zstep, ... of the 128^3 volume (the others stay untouched)
int hb_noise_volumes(uint32_t seed, int z0, int zstep, uint8_t *low128_rgba8, uint8_t *hi32_rgba8) {
    if (zstep <= 0) return -1;
    for (int z = 0; z < 32; z++)
        for (int y = 0; y < 32; y++)
            for (int x = 0; x < 32; x++) {
                uchar4 q = volume_pixel::hires_voxel(seed, x, y, z);
                uint8_t *o = hi32_rgba8 + 4 * (((size_t)z * 32 + y) * 32 + x);
                o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
            }
    for (int z = z0; z < 128; z += zstep)
        for (int y = 0; y < 128; y++)
            for (int x = 0; x < 128; x++) {
                uchar4 q = volume_pixel::lowres_voxel(seed, x, y, z);
                uint8_t *o = low128_rgba8 + 4 * (((size_t)z * 128 + y) * 128 + x);
                o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
            }
    return 0;
}

// tonemap_kernel: the map the parity gate is defined on (RGBA32F -> RGBA8)
int hb_tonemap(const float *src, size_t npix, uint8_t *dst_rgba8) {
    for (size_t i = 0; i < npix; i++) {
        uchar4 q = tonemap_pixel::tonemap_texel(make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]));
        dst_rgba8[4 * i] = q.x; dst_rgba8[4 * i + 1] = q.y; dst_rgba8[4 * i + 2] = q.z; dst_rgba8[4 * i + 3] = q.w;
    }
    return 0;
}

}  // extern "C"
