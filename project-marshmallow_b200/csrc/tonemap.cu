// tonemap.cu -- K4: HDR RGBA32F -> RGBA8, the map the parity gate is defined on.
// Restates tonemap.frag:11-28 (Uncharted-2 curve, exposure 0.7, gamma 1/2.2, white point 50.2);
// the vignette (tonemap.frag:30-32) is position-only and omitted.  alpha = clamp(a,0,1).
// HBM-bound: 16 B read + 4 B written per pixel, fully coalesced.
#include "common.h"

namespace mm {
namespace {

__device__ __forceinline__ float uc2(float x) {
    return (((x * ((0.15f * x) + (0.1f * 0.5f))) + (0.2f * 0.02f)) / ((x * ((0.15f * x) + 0.5f)) + (0.2f * 0.3f))) - (0.02f / 0.3f);
}
__device__ __forceinline__ float clamp01n(float x) { float r = (x > 0.0f) ? x : 0.0f; return (r < 1.0f) ? r : 1.0f; }

__global__ void tonemap_kernel(const float *src, size_t pitch, int W, int H, uchar4 *dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    float4 c = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(src) + (size_t)y * pitch + (size_t)x * 16);
    float whitemap = 1.0f / uc2(50.2f);
    float r = powf(uc2(0.7f * c.x) * whitemap, 1.0f / 2.2f);
    float g = powf(uc2(0.7f * c.y) * whitemap, 1.0f / 2.2f);
    float b = powf(uc2(0.7f * c.z) * whitemap, 1.0f / 2.2f);
    dst[(size_t)y * W + x] = make_uchar4((unsigned char)floorf((255.0f * clamp01n(r)) + 0.5f), (unsigned char)floorf((255.0f * clamp01n(g)) + 0.5f),
                                         (unsigned char)floorf((255.0f * clamp01n(b)) + 0.5f), (unsigned char)floorf((255.0f * clamp01n(c.w)) + 0.5f));
}

}  // namespace

cudaError_t launch_tonemap(const float *src, size_t pitch, int W, int H, uchar4 *dst, cudaStream_t stream) {
    if (W <= 0 || H <= 0) return cudaSuccess;
    tonemap_kernel<<<dim3((W + 127) / 128, H), 128, 0, stream>>>(src, pitch, W, H, dst);
    return cudaGetLastError();
}

}  // namespace mm
