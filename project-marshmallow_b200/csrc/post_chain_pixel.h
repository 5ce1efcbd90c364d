// post_chain_pixel.h -- the per-pixel arithmetic of K6 (god-ray.frag, radialBlur.frag, tonemap.frag), shared by the kernels of post_chain.cu and by the
// HOST build of the same source that the CPU test-suite checks against the oracle (tests/host_build/aux_host.cu, tests/test_host_build.py): every function is
// __host__ __device__, the two device-only spellings go through MM_FMAF / gamma_pow's host branch.  On the device nothing changes (the SASS of the kernels is
// byte-identical to the build that had these functions inside post_chain.cu).
#pragma once
#include <math.h>

#include "common.h"

#ifdef __CUDA_ARCH__
#define MM_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define MM_FMAF(a, b, c) fmaf((a), (b), (c))          // the host build: one rounding, like the device intrinsic
#endif
#define MM_HD __host__ __device__ __forceinline__

namespace mm {
namespace post_pixel {

MM_HD float lerpf(float p, float q, float a) { return MM_FMAF(a, q - p, p); }
MM_HD float clamp01n(float x) { float r = (x > 0.0f) ? x : 0.0f; return (r < 1.0f) ? r : 1.0f; }

// one axis of the LINEAR / CLAMP_TO_EDGE footprint (VulkanApplication.cpp:1290-1303)
MM_HD void tap_axis(float u, int n, int &i0, int &i1, float &a) {
    float U = (u * (float)n) - 0.5f;
    float fl = floorf(U);
    a = U - fl;
    if (!(fl >= -1.0f)) fl = -1.0f;
    if (fl > (float)n) fl = (float)n;
    int i = (int)fl;
    i0 = min(max(i, 0), n - 1);
    i1 = min(max(i + 1, 0), n - 1);
}

// alpha of an RGBA32F image (stride 16 B) or of a float plane (stride 4 B)
template <bool PLANE>
MM_HD float texel_alpha(const char *img, size_t pitch, int x, int y) {
    return PLANE ? *reinterpret_cast<const float *>(img + (size_t)y * pitch + (size_t)x * 4)
                 : *reinterpret_cast<const float *>(img + (size_t)y * pitch + (size_t)x * 16 + 12);
}
template <bool PLANE>
MM_HD float tap_alpha(const char *img, size_t pitch, int W, int H, float u, float v) {
    int x0, x1, y0, y1; float a, b;
    tap_axis(u, W, x0, x1, a);
    tap_axis(v, H, y0, y1, b);
    float t00 = texel_alpha<PLANE>(img, pitch, x0, y0), t10 = texel_alpha<PLANE>(img, pitch, x1, y0);
    float t01 = texel_alpha<PLANE>(img, pitch, x0, y1), t11 = texel_alpha<PLANE>(img, pitch, x1, y1);
    return lerpf(lerpf(t00, t10, a), lerpf(t01, t11, a), b);
}

// god-ray.frag:41-76 -> the pass's output alpha
MM_HD float god_ray_alpha(const PostParams &P, int x, int y, float alpha0) {
    if (P.sun_dir_y < 0.0f) return 1.0f;                                       // :51-53
    const char *src = reinterpret_cast<const char *>(P.src);
    float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
    float cx = (u * 2.0f) - 1.0f, cy = (v * 2.0f) - 1.0f;                      // :42-43
    const float k = (1.0f / 8.0f) * 0.75f;                                     // SAMPLE_WEIGHT * DENSITY
    float dx = (cx - P.sun_x) * k, dy = (cy - P.sun_y) * k;                    // :46-47
    float accum = alpha0 * 0.5f;                                               // :55
    float decay = 1.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {                                              // :58-73
        cx = cx - dx; cy = cy - dy;
        float s = tap_alpha<false>(src, P.src_pitch, P.W, P.H, (cx * 0.5f) + 0.5f, (cy * 0.5f) + 0.5f) * 0.5f;
        s = s * ((1.0f / 8.0f) * decay);
        accum = accum + s;
        decay = decay * 0.99f;
    }
    return accum * 0.9f;                                                       // :75
}

// radialBlur.frag:36-63 -> rgb of the pass's output (alpha is 1)
template <bool PLANE>
MM_HD float3 radial_blur_rgb(const PostParams &P, int x, int y, float3 cf) {
    if (P.sun_dir_y < 0.0f) return cf;                                         // :41-43
    const char *taps = PLANE ? reinterpret_cast<const char *>(P.plane) : reinterpret_cast<const char *>(P.src);
    size_t pitch = PLANE ? P.plane_pitch : P.src_pitch;
    const float samples[10] = {-0.08f, -0.05f, -0.03f, -0.02f, -0.01f, 0.01f, 0.02f, 0.03f, 0.05f, 0.08f};
    float u = ((float)x + 0.5f) / (float)P.W, v = ((float)y + 0.5f) / (float)P.H;
    float sx = (u * 2.0f) - 1.0f, sy = (v * 2.0f) - 1.0f;
    float lx = P.sun_x - sx, ly = P.sun_y - sy;                                // :51
    float dist = sqrtf((lx * lx) + (ly * ly));
    lx = lx / dist; ly = ly / dist;
    float accum = 0.0f;
#pragma unroll
    for (int i = 0; i < 10; i++) {                                             // :57-59
        float px = sx + (((samples[i] * lx) * 1.5f) * dist), py = sy + (((samples[i] * ly) * 1.5f) * dist);
        accum = accum + (tap_alpha<PLANE>(taps, pitch, P.W, P.H, (px * 0.5f) + 0.5f, (py * 0.5f) + 0.5f) * 1.1f);
    }
    accum = accum / 10.0f;                                                     // :60
    return make_float3((P.sun_rgb[0] * accum) + (0.5f * cf.x), (P.sun_rgb[1] * accum) + (0.5f * cf.y), (P.sun_rgb[2] * accum) + (0.5f * cf.z));   // :62
}

// tonemap.frag:11-33 -> UNORM8, alpha 255
MM_HD float uc2(float x) {
    return (((x * ((0.15f * x) + (0.1f * 0.5f))) + (0.2f * 0.02f)) / ((x * ((0.15f * x) + 0.5f)) + (0.2f * 0.3f))) - (0.02f / 0.3f);
}
// pow(t, 1/2.2) the way a GPU evaluates GLSL pow: exp2(y * log2(t)) on the special-function unit (the Vulkan precision of pow is the one
// inherited from that expression).  t = 0 -> 0, t < 0 -> NaN -> clamps to 0 like powf's NaN.  ~1e-6 relative: the UNORM8 result differs from
// the correctly rounded one by one step on ~0.05 % of channels, inside the pass's stated tolerance (tests/test_post_chain.py), and it takes
// the three powf calls (a third of the fused kernel's instructions) down to a dozen instructions.
MM_HD float gamma_pow(float t) {
#ifdef __CUDA_ARCH__
    float l, r;
    asm("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(t));
    asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(l * (1.0f / 2.2f)));
    return r;
#else
    return exp2f(log2f(t) * (1.0f / 2.2f));            // host build: the same expression through libm (the pass's tolerance is one UNORM8 step)
#endif
}
MM_HD uchar4 present_pixel(const PostParams &P, int x, int y, float3 c) {
    float whitemap = 1.0f / uc2(50.2f);
    float u = (((float)x + 0.5f) / (float)P.W) - 0.5f, v = (((float)y + 0.5f) / (float)P.H) - 0.5f;
    float vig = (u * u) + (v * v);                                             // :30
    float col[3] = {c.x, c.y, c.z};
    const float vc[3] = {0.1f, 0.05f, 0.13f};
    unsigned char q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float t = gamma_pow(uc2(0.7f * col[k]) * whitemap);                     // :14-20, 27-28
        t = (t * (1.0f - vig)) + (vc[k] * vig);                                // :32
        q[k] = (unsigned char)floorf((255.0f * clamp01n(t)) + 0.5f);
    }
    return P.bgra ? make_uchar4(q[2], q[1], q[0], 255) : make_uchar4(q[0], q[1], q[2], 255);
}

}  // namespace post_pixel
}  // namespace mm
