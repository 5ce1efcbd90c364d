// curl_noise.cu -- K2: the 128x128 curl-noise FBM texture, built on the GPU.
//
// Replaces the reference's offline generator GenerateCurlNoise (ImageUtils.cpp:176-223) and its
// callees: sin-hash (IU:25-29) -> 12-gradient tiling Perlin (IU:41-73) -> 4-octave FBM, persistence
// 0.4 (IU:75-87) -> central-difference curl (IU:130-169) -> per-channel global min/max normalise ->
// roundf(x*255) (IU:200-217).  The result must equal the shipped Textures/CurlNoiseFBM.tga byte for
// byte, so every float/double promotion of the C++ source is followed: lerp promotes to binary64
// through its 1.0 literal, EPS is a binary64 literal, the rest is binary32.  Compiled with -fmad=false.
//
// The lattice hash is fract(sinf(dot(p, k)) * 43758.5453f): its value depends on the LAST BIT of the
// host C library's sinf (one ulp moves the fraction by up to 0.004 and flips the chosen gradient), and
// CUDA's sinf is not bit-identical to glibc's / MSVC's.  The lattice is tiny -- integer coordinates in
// [-1, 24]^3 for base frequency 3 and 4 octaves -- so the gradient index of every lattice point
// (26^3 = 17,576 bytes) is evaluated once on the host with the C library's sinf (capi.cu,
// build_curl_gradient_table, following IU:25-34) and the kernels look it up.  Everything downstream of
// the hash (Perlin interpolation, FBM, curl, min/max normalisation, quantisation) runs on the GPU.
//
// Work split: 12 FBM evaluations per pixel (six central differences) -> one thread per
// (pixel, evaluation): 196,608 threads; then curl + min/max; then normalise + quantise.
#include "common.h"
#include "curl_noise_pixel.h"

namespace mm {
namespace {

using namespace curl_pixel;

__global__ void curl_fbm_kernel(const unsigned char *__restrict__ table, float *fbm12) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= CURL_DIM * CURL_DIM * 12) return;
    fbm12[gid] = curl_probe(table, gid);
}

__global__ void curl_combine_kernel(const float *fbm12, float *curls) {
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= CURL_DIM * CURL_DIM) return;
    curl_combine(fbm12, curls, pix);
}

// one block: per-channel min / max over all pixels -> bounds[0..2] = min, bounds[3..5] = max
__global__ void curl_bounds_kernel(const float *curls, float *bounds) {
    __shared__ float smin[3][32], smax[3][32];
    float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (int i = threadIdx.x; i < CURL_DIM * CURL_DIM; i += blockDim.x)
        for (int c = 0; c < 3; c++) { float v = curls[3 * i + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { smin[c][threadIdx.x >> 5] = lo[c]; smax[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int c = threadIdx.x;
        float a = smin[c][0], b = smax[c][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { a = fminf(a, smin[c][w]); b = fmaxf(b, smax[c][w]); }
        bounds[c] = a; bounds[3 + c] = b;
    }
}

__global__ void curl_quantise_kernel(const float *curls, const float *bounds, uchar4 *dst) {
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= CURL_DIM * CURL_DIM) return;
    dst[pix] = curl_quantise(curls, bounds, pix);
}

}  // namespace

// scratch: 12*N (fbm) + 3*N (curl) + 6 floats, N = 128*128; table: 26^3 gradient indices (device)
cudaError_t launch_curl_noise(uchar4 *dst, float *scratch, const unsigned char *table, cudaStream_t stream) {
    const int N = CURL_DIM * CURL_DIM;
    float *fbm12 = scratch, *curls = scratch + 12 * N, *bounds = curls + 3 * N;
    curl_fbm_kernel<<<(N * 12 + 127) / 128, 128, 0, stream>>>(table, fbm12);
    curl_combine_kernel<<<(N + 127) / 128, 128, 0, stream>>>(fbm12, curls);
    curl_bounds_kernel<<<1, 1024, 0, stream>>>(curls, bounds);
    curl_quantise_kernel<<<(N + 127) / 128, 128, 0, stream>>>(curls, bounds, dst);
    return cudaGetLastError();
}

}  // namespace mm
