// post_chain.cu -- K6: the three full-screen passes that consume the cloud image
//     Shaders/god-ray.frag:41-76      8 alpha taps toward the sun -> light-shaft alpha
//     Shaders/radialBlur.frag:36-63   10 alpha taps along the sun direction -> additive sun-coloured light
//     Shaders/tonemap.frag:11-33      Uncharted-2 curve, gamma, vignette -> swapchain UNORM8
// The reference records them back to back over three RGBA32F framebuffers (VulkanApplication.cpp:936-968, 1016):
// 16 B read + 16 B written per pixel and pass, plus the taps.  Here:
//   * pass-by-pass kernels (god_ray_kernel<false>, radial_blur_kernel<false>, present_kernel) reproduce each
//     framebuffer bit for bit against the oracle (oracle/post_chain_oracle.c) -- the parity surface;
//   * the production path mm_post_chain is TWO kernels: god_ray_kernel<true> writes only the new alpha (4 B/pixel
//     plane; rgb passes through god-ray.frag unchanged), radial_blur_kernel<true> reads rgb from the cloud image and
//     its 10 taps from that plane, tone-maps in registers and stores 4 B/pixel.  Same device functions, same
//     operation order, so the bytes equal the three-pass result; algorithmic HBM traffic 16+4 and 16+4+4 = 44 B/pixel
//     against 3 x 32 (+ the background copy) in the reference's chain.  HBM-bound; the taps hit L1/L2.
// One thread per pixel, 32x8 blocks: a warp covers 32 consecutive pixels of a row (512 B loads, 128 B / 512 B stores).
// Arithmetic: IEEE binary32 in GLSL order (-fmad=false); the only transcendental is the tone map's pow (gamma_pow below).
#include "common.h"
#include "post_chain_pixel.h"

namespace mm {
namespace {

using namespace post_pixel;

// ALPHA_ONLY = false: the pass as the reference runs it (RGBA32F in, RGBA32F out); true: only the new alpha, to a plane
template <bool ALPHA_ONLY>
__global__ void __launch_bounds__(256) god_ray_kernel(const __grid_constant__ PostParams P) {
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.H) return;
    const char *src = reinterpret_cast<const char *>(P.src);
    if (ALPHA_ONLY) {
        float a0 = texel_alpha<false>(src, P.src_pitch, x, y);
        *reinterpret_cast<float *>(reinterpret_cast<char *>(P.plane) + (size_t)y * P.plane_pitch + (size_t)x * 4) = god_ray_alpha(P, x, y, a0);
    } else {
        float4 cf = *reinterpret_cast<const float4 *>(src + (size_t)y * P.src_pitch + (size_t)x * 16);
        cf.w = god_ray_alpha(P, x, y, cf.w);
        *reinterpret_cast<float4 *>(reinterpret_cast<char *>(P.dst) + (size_t)y * P.dst_pitch + (size_t)x * 16) = cf;
    }
}

// FUSED = false: radialBlur.frag alone (RGBA32F in, RGBA32F out).  FUSED = true: rgb from the cloud image, taps from
// the god-ray alpha plane, tone map + vignette in registers, UNORM8 out.
template <bool FUSED>
__global__ void __launch_bounds__(256) radial_blur_kernel(const __grid_constant__ PostParams P) {
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.H) return;
    float4 cf = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(P.src) + (size_t)y * P.src_pitch + (size_t)x * 16);
    float3 rgb = radial_blur_rgb<FUSED>(P, x, y, make_float3(cf.x, cf.y, cf.z));
    if (FUSED) {
        *reinterpret_cast<uchar4 *>(reinterpret_cast<char *>(P.dst8) + (size_t)y * P.dst8_pitch + (size_t)x * 4) = present_pixel(P, x, y, rgb);
    } else {
        *reinterpret_cast<float4 *>(reinterpret_cast<char *>(P.dst) + (size_t)y * P.dst_pitch + (size_t)x * 16) = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
    }
}

__global__ void __launch_bounds__(256) present_kernel(const __grid_constant__ PostParams P) {
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.H) return;
    float4 c = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(P.src) + (size_t)y * P.src_pitch + (size_t)x * 16);
    *reinterpret_cast<uchar4 *>(reinterpret_cast<char *>(P.dst8) + (size_t)y * P.dst8_pitch + (size_t)x * 4) = present_pixel(P, x, y, make_float3(c.x, c.y, c.z));
}

}  // namespace

cudaError_t launch_post(int which, const PostParams &p, cudaStream_t stream) {
    if (p.W <= 0 || p.H <= 0) return cudaSuccess;
    dim3 grid((p.W + 31) / 32, (p.H + 7) / 8), block(32, 8);
    switch (which) {
        case POST_GOD_RAY: god_ray_kernel<false><<<grid, block, 0, stream>>>(p); break;
        case POST_GOD_RAY_ALPHA: god_ray_kernel<true><<<grid, block, 0, stream>>>(p); break;
        case POST_RADIAL_BLUR: radial_blur_kernel<false><<<grid, block, 0, stream>>>(p); break;
        case POST_BLUR_PRESENT: radial_blur_kernel<true><<<grid, block, 0, stream>>>(p); break;
        case POST_PRESENT: present_kernel<<<grid, block, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mm
