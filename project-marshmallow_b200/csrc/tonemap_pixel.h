// tonemap_pixel.h -- the per-pixel arithmetic of K4 (tonemap.frag:11-28), shared by tonemap_kernel and by the HOST build of the same source in the CPU
// test-suite (tests/host_build/aux_host.cu).  The product build never defines MM_HOST_BUILD; the kernel's SASS is byte-identical to the build that had this code
// inside tonemap.cu.
#pragma once
#include <math.h>

#include "common.h"

#if defined(MM_HOST_BUILD)
#define MM_HD __host__ __device__ __forceinline__
#else
#define MM_HD __device__ __forceinline__
#endif

namespace mm {
namespace tonemap_pixel {

MM_HD float uc2(float x) {
    return (((x * ((0.15f * x) + (0.1f * 0.5f))) + (0.2f * 0.02f)) / ((x * ((0.15f * x) + 0.5f)) + (0.2f * 0.3f))) - (0.02f / 0.3f);
}
MM_HD float clamp01n(float x) { float r = (x > 0.0f) ? x : 0.0f; return (r < 1.0f) ? r : 1.0f; }

// RGBA32F texel -> RGBA8: Uncharted-2 curve, exposure 0.7, gamma 1/2.2, white point 50.2; alpha = clamp(a, 0, 1)
MM_HD uchar4 tonemap_texel(float4 c) {
    float whitemap = 1.0f / uc2(50.2f);
    float r = powf(uc2(0.7f * c.x) * whitemap, 1.0f / 2.2f);
    float g = powf(uc2(0.7f * c.y) * whitemap, 1.0f / 2.2f);
    float b = powf(uc2(0.7f * c.z) * whitemap, 1.0f / 2.2f);
    return make_uchar4((unsigned char)floorf((255.0f * clamp01n(r)) + 0.5f), (unsigned char)floorf((255.0f * clamp01n(g)) + 0.5f),
                                         (unsigned char)floorf((255.0f * clamp01n(b)) + 0.5f), (unsigned char)floorf((255.0f * clamp01n(c.w)) + 0.5f));
}

}  // namespace tonemap_pixel
}  // namespace mm
