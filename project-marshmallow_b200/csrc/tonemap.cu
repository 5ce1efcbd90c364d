// tonemap.cu -- K4: HDR RGBA32F -> RGBA8, the map the parity gate is defined on.
// Restates tonemap.frag:11-28 (Uncharted-2 curve, exposure 0.7, gamma 1/2.2, white point 50.2);
// the vignette (tonemap.frag:30-32) is position-only and omitted.  alpha = clamp(a,0,1).
// HBM-bound: 16 B read + 4 B written per pixel, fully coalesced.
#include "common.h"
#include "tonemap_pixel.h"

namespace mm {
namespace {

using namespace tonemap_pixel;

__global__ void tonemap_kernel(const float *src, size_t pitch, int W, int H, uchar4 *dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    float4 c = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(src) + (size_t)y * pitch + (size_t)x * 16);
    dst[(size_t)y * W + x] = tonemap_texel(c);
}

}  // namespace

cudaError_t launch_tonemap(const float *src, size_t pitch, int W, int H, uchar4 *dst, cudaStream_t stream) {
    if (W <= 0 || H <= 0) return cudaSuccess;
    tonemap_kernel<<<dim3((W + 127) / 128, H), 128, 0, stream>>>(src, pitch, W, H, dst);
    return cudaGetLastError();
}

}  // namespace mm
