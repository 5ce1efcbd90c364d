// noise_volume_pixel.h -- the per-voxel arithmetic of K3 (our own hashed-cell generator, see noise_volumes.cu), shared by the kernels and by the HOST build of
// the same source in the CPU test-suite (tests/host_build/aux_host.cu).  The product build never defines MM_HOST_BUILD; its kernels' SASS is byte-identical to
// the build that had this code inside noise_volumes.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#include "common.h"

#if defined(MM_HOST_BUILD)
#define MM_HD __host__ __device__ __forceinline__
#define MM_HD_PLAIN __host__ __device__
#else
#define MM_HD __device__ __forceinline__
#define MM_HD_PLAIN __device__
#endif

namespace mm {
namespace volume_pixel {

MM_HD uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
MM_HD uint32_t cell_hash(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
    return fmix32((x * 73856093u) ^ (y * 19349663u) ^ (z * 83492791u) ^ (seed * 0x9E3779B9u));
}

MM_HD_PLAIN float worley(float px, float py, float pz, int cells, uint32_t seed) {
    float fx = px * (float)cells, fy = py * (float)cells, fz = pz * (float)cells;
    int cx = (int)fx, cy = (int)fy, cz = (int)fz;
    float best = 1.0e9f;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                uint32_t h = cell_hash((uint32_t)((nx + cells) % cells), (uint32_t)((ny + cells) % cells), (uint32_t)((nz + cells) % cells), seed);
                float jx = (float)(h & 1023u) * (1.0f / 1024.0f);
                float jy = (float)((h >> 10) & 1023u) * (1.0f / 1024.0f);
                float jz = (float)((h >> 20) & 1023u) * (1.0f / 1024.0f);
                float ex = ((float)nx + jx) - fx, ey = ((float)ny + jy) - fy, ez = ((float)nz + jz) - fz;
                float d2 = ((ex * ex) + (ey * ey)) + (ez * ez);
                if (d2 < best) best = d2;
            }
    float d = sqrtf(best);
    if (d > 1.0f) d = 1.0f;
    return 1.0f - d;
}

MM_HD_PLAIN float worley_fbm(float x, float y, float z, int cells, uint32_t seed) {
    return ((0.625f * worley(x, y, z, cells, seed)) + (0.25f * worley(x, y, z, cells * 2, seed + 1u))) +
           (0.125f * worley(x, y, z, cells * 4, seed + 2u));
}

MM_HD float fade(float t) { return ((t * t) * t) * ((t * ((t * 6.0f) - 15.0f)) + 10.0f); }
MM_HD float lerpn(float a, float b, float t) { return a + (t * (b - a)); }

MM_HD_PLAIN float perlin(float px, float py, float pz, int cells, uint32_t seed) {
    float fx = px * (float)cells, fy = py * (float)cells, fz = pz * (float)cells;
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float rx = fx - (float)ix, ry = fy - (float)iy, rz = fz - (float)iz;
    float u = fade(rx), v = fade(ry), w = fade(rz);
    float c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int ox = k & 1, oy = (k >> 1) & 1, oz = (k >> 2) & 1;
        uint32_t h = cell_hash((uint32_t)((ix + ox) % cells), (uint32_t)((iy + oy) % cells), (uint32_t)((iz + oz) % cells), seed);
        uint32_t g = h % 12u;
        // 12 edge gradients: pair = g/4 picks the zero axis (2,1,0), low bits pick signs
        float s0 = (g & 1u) ? -1.0f : 1.0f, s1 = (g & 2u) ? -1.0f : 1.0f;
        uint32_t pair = g >> 2;
        float gx = pair == 2u ? 0.0f : s0;
        float gy = pair == 0u ? s1 : (pair == 1u ? 0.0f : s0);
        float gz = pair == 0u ? 0.0f : s1;
        c[k] = ((gx * (rx - (float)ox)) + (gy * (ry - (float)oy))) + (gz * (rz - (float)oz));
    }
    float x00 = lerpn(c[0], c[1], u), x10 = lerpn(c[2], c[3], u), x01 = lerpn(c[4], c[5], u), x11 = lerpn(c[6], c[7], u);
    return lerpn(lerpn(x00, x10, v), lerpn(x01, x11, v), w);
}

MM_HD_PLAIN float perlin_fbm(float x, float y, float z, int cells, int octaves, uint32_t seed) {
    float sum = 0.0f, amp = 1.0f, tot = 0.0f;
    for (int o = 0; o < octaves; o++) {
        sum = sum + (amp * perlin(x, y, z, cells, seed + (uint32_t)o));
        tot = tot + amp;
        amp = amp * 0.5f;
        cells = cells * 2;
    }
    return sum / tot;
}

MM_HD float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
MM_HD unsigned char quant(float v) { return (unsigned char)(int)((clamp01(v) * 255.0f) + 0.5f); }

// per-channel affine maps fitted once (seed 0) to the shipped volumes' channel means / standard deviations
#define NV_L0_GAIN 0.8718f
#define NV_L0_BIAS -0.1396f
#define NV_L1_GAIN 0.8817f
#define NV_L1_BIAS 0.2618f
#define NV_L2_GAIN 0.8671f
#define NV_L2_BIAS 0.2710f
#define NV_L3_GAIN 0.8721f
#define NV_L3_BIAS 0.2644f
#define NV_H0_GAIN 0.9660f
#define NV_H0_BIAS 0.2127f
#define NV_H1_GAIN 0.8610f
#define NV_H1_BIAS 0.2786f
#define NV_H2_GAIN 0.8877f
#define NV_H2_BIAS 0.2553f

// voxel (x, y, z) of the 128^3 low-res volume: .r Perlin-Worley base shape, .gba Worley FBM at rising frequency
MM_HD uchar4 lowres_voxel(uint32_t seed, int x, int y, int z) {
    float px = ((float)x + 0.5f) * (1.0f / 128.0f), py = ((float)y + 0.5f) * (1.0f / 128.0f), pz = ((float)z + 0.5f) * (1.0f / 128.0f);
    float pf = (perlin_fbm(px, py, pz, 4, 5, seed) * 1.2f) + 0.5f;
    float w0 = worley_fbm(px, py, pz, 4, seed + 100u);
    float w1 = worley_fbm(px, py, pz, 8, seed + 200u);
    float w2 = worley_fbm(px, py, pz, 16, seed + 300u);
    float w3 = worley_fbm(px, py, pz, 32, seed + 400u);
    float pw = w0 + (clamp01(pf) * (1.0f - w0));
    return make_uchar4(quant((pw * NV_L0_GAIN) + NV_L0_BIAS), quant((w1 * NV_L1_GAIN) + NV_L1_BIAS),
                       quant((w2 * NV_L2_GAIN) + NV_L2_BIAS), quant((w3 * NV_L3_GAIN) + NV_L3_BIAS));
}

// voxel (x, y, z) of the 32^3 hi-res volume: .rgb Worley FBM, .a = 0
MM_HD uchar4 hires_voxel(uint32_t seed, int x, int y, int z) {
    float px = ((float)x + 0.5f) * (1.0f / 32.0f), py = ((float)y + 0.5f) * (1.0f / 32.0f), pz = ((float)z + 0.5f) * (1.0f / 32.0f);
    float w0 = worley_fbm(px, py, pz, 2, seed + 500u);
    float w1 = worley_fbm(px, py, pz, 4, seed + 600u);
    float w2 = worley_fbm(px, py, pz, 8, seed + 700u);
    return make_uchar4(quant((w0 * NV_H0_GAIN) + NV_H0_BIAS), quant((w1 * NV_H1_GAIN) + NV_H1_BIAS), quant((w2 * NV_H2_GAIN) + NV_H2_BIAS), 0);
}

}  // namespace volume_pixel
}  // namespace mm
