// noise_volumes.cu -- K3: hashed-cell Perlin-Worley / Worley-FBM volume build on the GPU.
//
// The reference ships its 128^3 low-res and 32^3 hi-res cloud-shape volumes as TGA slices baked by
// an external Houdini asset (README.md:68; loaded by Texture.cpp:502-538) and contains no generator,
// so this is our own design with the channel semantics compute-clouds.comp consumes: low-res .r =
// Perlin-Worley base shape (CC:240), .gba = Worley FBM at rising frequency (CC:247); hi-res .rgb =
// Worley FBM (CC:224), .a = 0.  Cells are hashed with a 32-bit integer mix (no tables, no state), all
// lattices wrap so the volumes tile, and the floating-point part uses only binary32 +,-,*,/,sqrt in a
// fixed order (-fmad=false) so the CPU statement of the same generator produces identical bytes.
//
// Layout: one thread per voxel, x fastest; a warp writes 32 consecutive uchar4 = one 128-byte line.
#include "common.h"
#include "noise_volume_pixel.h"

namespace mm {
namespace {

using namespace volume_pixel;

__global__ void __launch_bounds__(128) lowres_kernel(uint32_t seed, uchar4 *out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    out[((size_t)z * 128 + y) * 128 + x] = lowres_voxel(seed, x, y, z);
}

__global__ void __launch_bounds__(32) hires_kernel(uint32_t seed, uchar4 *out) {
    int x = threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    out[((size_t)z * 32 + y) * 32 + x] = hires_voxel(seed, x, y, z);
}

}  // namespace

cudaError_t launch_noise_volumes(uint32_t seed, uchar4 *low128, uchar4 *hi32, cudaStream_t stream) {
    if (low128) lowres_kernel<<<dim3(1, 128, 128), 128, 0, stream>>>(seed, low128);
    if (hi32) hires_kernel<<<dim3(1, 32, 32), 32, 0, stream>>>(seed, hi32);
    return cudaGetLastError();
}

}  // namespace mm
