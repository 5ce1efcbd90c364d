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

namespace mm {
namespace {

#define CURL_DIM 128
#define CURL_EPS 0.0005

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ((ax * bx) + (ay * by)) + (az * bz);
}

#define CURL_LATTICE 26      // lattice coordinates -1..24, stored at +1

__device__ __forceinline__ int hash_index(const unsigned char *__restrict__ table, float x, float y, float z) {
    int ix = (int)x + 1, iy = (int)y + 1, iz = (int)z + 1;
    return __ldg(table + (iz * CURL_LATTICE + iy) * CURL_LATTICE + ix);
}

__device__ __forceinline__ void gradient(int idx, float &gx, float &gy, float &gz) {
    // the 12 edge directions scaled by 0.7071 (IU:10-23): axis pair = idx/4, signs = idx%4
    const float k = (float)0.7071;
    float s0 = (idx & 2) ? -k : k, s1 = (idx & 1) ? -k : k;
    int pair = idx >> 2;
    gx = pair == 2 ? 0.0f : s0;
    gy = pair == 0 ? s1 : (pair == 1 ? 0.0f : s0);
    gz = pair == 0 ? 0.0f : s1;
}

__device__ __forceinline__ float corner(const unsigned char *__restrict__ table, float cx, float cy, float cz, float dx, float dy, float dz) {
    float gx, gy, gz;
    gradient(hash_index(table, cx, cy, cz), gx, gy, gz);
    return dot3(gx, gy, gz, dx, dy, dz);
}

__device__ __forceinline__ float lerp_d(float a, float b, float t) {
    return (float)(((1.0 - (double)t) * (double)a) + (double)(t * b));
}

__device__ __forceinline__ float fade5(float r) { return ((r * r) * r) * ((r * ((r * 6.0f) - 15.0f)) + 10.0f); }

__device__ float perlin(const unsigned char *__restrict__ table, float x, float y, float z, float freq) {
    x *= freq; y *= freq; z *= freq;
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    float rx = x - fx, ry = y - fy, rz = z - fz;
    float ux = fade5(rx), uy = fade5(ry), uz = fade5(rz);
    float gx = fx + 1.0f, gy = fy + 1.0f, gz = fz + 1.0f;
    if (fabsf(gx - freq) < 0.001f) gx = freq;     // IU:49-51 as written
    if (fabsf(gy - freq) < 0.001f) gy = freq;
    if (fabsf(gz - freq) < 0.001f) gz = freq;
    float nnn = corner(table, fx, fy, fz, rx, ry, rz);
    float nnp = corner(table, fx, fy, gz, rx, ry, rz - 1.0f);
    float npn = corner(table, fx, gy, fz, rx, ry - 1.0f, rz);
    float npp = corner(table, fx, gy, gz, rx, ry - 1.0f, rz - 1.0f);
    float pnn = corner(table, gx, fy, fz, rx - 1.0f, ry, rz);
    float pnp = corner(table, gx, fy, gz, rx - 1.0f, ry, rz - 1.0f);
    float ppn = corner(table, gx, gy, fz, rx - 1.0f, ry - 1.0f, rz);
    float ppp = corner(table, gx, gy, gz, rx - 1.0f, ry - 1.0f, rz - 1.0f);
    float nn = lerp_d(nnn, pnn, ux), np = lerp_d(nnp, pnp, ux), pn = lerp_d(npn, ppn, ux), pp = lerp_d(npp, ppp, ux);
    float n = lerp_d(nn, pn, uy), p = lerp_d(np, pp, uy);
    return lerp_d(n, p, uz);
}

__device__ float fbm(const unsigned char *__restrict__ table, float x, float y, float z, float freq, int octaves) {
    float noise = 0.0f, weight = 1.0f, total = 0.0f;
    const float persistence = 0.4f;
    for (int i = 0; i < octaves; i++) {
        total += weight;
        noise += weight * perlin(table, x, y, z, freq);
        freq *= 2.0f;
        weight *= persistence;
    }
    return noise / total;
}

// evaluation k of pixel (col,row): the 12 FBM probes of curlNoiseFBM in source order (IU:133-166)
__global__ void curl_fbm_kernel(const unsigned char *__restrict__ table, float *fbm12) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= CURL_DIM * CURL_DIM * 12) return;
    int k = gid % 12, pix = gid / 12;
    float px = (float)(pix % CURL_DIM) / CURL_DIM, py = (float)(pix / CURL_DIM) / CURL_DIM;
    float xm = (float)((double)px - CURL_EPS), xp = (float)((double)px + CURL_EPS);
    float ym = (float)((double)py - CURL_EPS), yp = (float)((double)py + CURL_EPS);
    bool plus = k & 1;
    float x, y, z;
    switch (k >> 1) {
        case 0: x = plus ? xp : xm; y = py; z = 0.5f; break;              // dydx
        case 1: x = px; y = plus ? yp : ym; z = 0.5f; break;              // dxdy
        case 2: x = px; y = 0.5f; z = plus ? yp : ym; break;              // dxdz
        case 3: x = plus ? xp : xm; y = 0.5f; z = py; break;              // dzdx
        case 4: x = (float)0.5; y = plus ? yp : ym; z = px; break;        // dzdy
        default: x = 0.5f; y = py; z = plus ? xp : xm; break;             // dydz
    }
    fbm12[gid] = fbm(table, x, y, z, 3.f, 4);
}

__device__ __forceinline__ float cdiff(float a, float b) { return (float)((double)(b - a) / ((double)2.f * CURL_EPS)); }

__global__ void curl_combine_kernel(const float *fbm12, float *curls) {
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= CURL_DIM * CURL_DIM) return;
    const float *f = fbm12 + 12 * pix;
    float dydx = cdiff(f[0], f[1]), dxdy = cdiff(f[2], f[3]), dxdz = cdiff(f[4], f[5]);
    float dzdx = cdiff(f[6], f[7]), dzdy = cdiff(f[8], f[9]), dydz = cdiff(f[10], f[11]);
    curls[3 * pix + 0] = dzdy - dydz;
    curls[3 * pix + 1] = dxdz - dzdx;
    curls[3 * pix + 2] = dydx - dxdy;
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
    unsigned char q[3];
    for (int c = 0; c < 3; c++) {
        float lo = bounds[c], hi = bounds[3 + c];
        float m = 0.f + ((curls[3 * pix + c] - lo) / (hi - lo) * (1.f - 0.f));       // IU:171-174
        q[c] = (unsigned char)((int)roundf(m * 255.f));
    }
    dst[pix] = make_uchar4(q[0], q[1], q[2], 255);
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
