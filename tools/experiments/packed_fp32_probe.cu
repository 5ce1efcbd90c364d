// packed_fp32_probe.cu -- does halving the FP32 issue slots with FADD2/FMUL2/FFMA2 buy time in an issue-bound kernel whose other
// ~45 % of instructions (integer, compare, select) stay scalar?  Models the march's mix: per "trip" 24 FP32 ops + 20 other ops on a
// dependent-enough chain; variant A = one chain per thread, 32 warps/SM; variant B = two chains per thread on packed FP32, 16 warps/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o packed_probe packed_fp32_probe.cu && ./packed_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b, float2 nz) { return __ffma2_rn(a, b, nz); }   // exact product: addend is -0 at run time
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int ITERS>
__global__ void __launch_bounds__(128, 8) scalar_kernel(float *out, float seed, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x = seed + i * 1e-6f, y = 0.5f, z = 0.25f;
    unsigned k = i;
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int u = 0; u < ITERS; u++) {
            float a = x * 1.0001f, b = y * 0.9999f, c = z * 1.00003f;          // 3 FMUL
            float d = a + b, e = b + c, f = c + a;                               // 3 FADD
            float g = __fmaf_rn(d, e, f), h = __fmaf_rn(e, f, d);                // 2 FFMA
            x = g * 0.5f + 0.1f; y = h * 0.5f + 0.2f; z = (d + e) * 0.25f;       // 3 FMUL 3 FADD (no contraction: -fmad=false)
            // ~8 "other" ops: integer hash, compare, select
            k = k * 1664525u + 1013904223u; k ^= k >> 13;
            if ((k & 255u) == 0u) x = 0.5f;
            k += (x > 0.7f) ? 3u : 5u;
            z = (k & 1u) ? z : -z;
        }
    }
    out[i] = x + y + z + (float)(k & 7u);
}
template <int ITERS>
__global__ void __launch_bounds__(128, 4) packed_kernel(float *out, float seed, int n, float nzr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float2 nz = make_float2(nzr, nzr);
    float2 x = make_float2(seed + (2 * i) * 1e-6f, seed + (2 * i + 1) * 1e-6f), y = make_float2(0.5f, 0.5f), z = make_float2(0.25f, 0.25f);
    unsigned k0 = 2 * i, k1 = 2 * i + 1;
    const float2 c1 = make_float2(1.0001f, 1.0001f), c2 = make_float2(0.9999f, 0.9999f), c3 = make_float2(1.00003f, 1.00003f);
    const float2 hf = make_float2(0.5f, 0.5f), q = make_float2(0.25f, 0.25f), p1 = make_float2(0.1f, 0.1f), p2 = make_float2(0.2f, 0.2f);
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int u = 0; u < ITERS; u++) {
            float2 a = f2mul(x, c1, nz), b = f2mul(y, c2, nz), c = f2mul(z, c3, nz);
            float2 d = f2add(a, b), e = f2add(b, c), f = f2add(c, a);
            float2 g = f2fma(d, e, f), h = f2fma(e, f, d);
            x = f2add(f2mul(g, hf, nz), p1); y = f2add(f2mul(h, hf, nz), p2); z = f2mul(f2add(d, e), q, nz);
            k0 = k0 * 1664525u + 1013904223u; k0 ^= k0 >> 13;
            k1 = k1 * 1664525u + 1013904223u; k1 ^= k1 >> 13;
            if ((k0 & 255u) == 0u) x.x = 0.5f;
            if ((k1 & 255u) == 0u) x.y = 0.5f;
            k0 += (x.x > 0.7f) ? 3u : 5u; k1 += (x.y > 0.7f) ? 3u : 5u;
            z.x = (k0 & 1u) ? z.x : -z.x; z.y = (k1 & 1u) ? z.y : -z.y;
        }
    }
    out[2 * i] = x.x + y.x + z.x + (float)(k0 & 7u);
    out[2 * i + 1] = x.y + y.y + z.y + (float)(k1 & 7u);
}

int main() {
    const int chains = 148 * 8 * 128 * 8;          // 8 waves of one-chain threads
    float *out; cudaMalloc(&out, chains * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); scalar_kernel<8><<<chains / 128, 128>>>(out, 0.3f, 200); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("scalar: %.3f ms\n", ms);
        cudaEventRecord(e0); packed_kernel<8><<<chains / 256, 128>>>(out, 0.3f, 200, -0.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("packed: %.3f ms\n", ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
