// reproject.cu -- K5: the reprojection pass that precedes the cloud dispatch in the engine's frame.
//
// Replaces the vkCmdDispatch of Shaders/reproject.comp recorded at VulkanApplication.cpp:1053-1059 (ReprojectShader,
// Shader.h:380-452): each pixel re-aims its ray at the inner atmosphere shell (reproject.comp:96-120), moves the hit
// into the PREVIOUS camera's view space (:125), converts it back to a screen position (:128-137) and averages 10 taps
// of the previous image along the motion vector (:142-148); alpha comes from the centre tap (:150-151).  With
// MM_PHASE16 cloud dispatches this is the engine's actual cadence: reproject everything, re-march 1/16 of the pixels.
//
// Every operation that selects a source pixel is IEEE binary32 in the order the GLSL writes it (-fmad=false), so
// source coordinates -- and therefore the whole image, including the 10-tap sums -- are bit-identical to the oracle.
// HBM-bound: 16 B written + 16 B of unique previous-image data read per pixel (the 10 taps of neighbouring pixels
// overlap and hit L1/L2); one thread per pixel, a warp covers 32 consecutive pixels of a row (512 B stores).
#include "common.h"
#include "reproject_pixel.h"

namespace mm {
namespace {

using namespace reproject_pixel;

__global__ void __launch_bounds__(256) reproject_kernel(const __grid_constant__ ReprojectParams P) {
    int gx = blockIdx.x * 32 + (threadIdx.x & 31), gy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (gx >= P.W || gy >= P.H) return;
    float4 acc = reproject_texel(P, gx, gy);
    *(reinterpret_cast<float4 *>(reinterpret_cast<char *>(P.dst) + (size_t)gy * P.dst_pitch) + gx) = acc;
}

}  // namespace

cudaError_t launch_reproject(const ReprojectParams &p, cudaStream_t stream) {
    if (p.W <= 0 || p.H <= 0) return cudaSuccess;
    reproject_kernel<<<dim3((p.W + 31) / 32, (p.H + 7) / 8), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace mm
