/*
 * noise_volume_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * CPU statement of OUR hashed-cell 3D noise-volume generator (K3).  The reference holds NO code
 * for these volumes: its 128^3 "lowResCloudShape" and 32^3 "hiResCloudShape" textures were baked
 * by an external Houdini asset (reference README.md:68) and only loaded (Texture.cpp:502-538).
 * PARITY STATUS: "parity unpinned" -- there is no reference arithmetic to follow.  What is pinned:
 *   (i)  CPU (this file) vs GPU (csrc/noise_volumes.cu) byte-identical volumes and integer hashes;
 *   (ii) the channel statistics of the shipped volumes within a tolerance;
 *   (iii) seamless tiling on all three axes.
 * Channel semantics follow what compute-clouds.comp consumes: low-res .r = Perlin-Worley base
 * shape (CC:240), .gba = three Worley-FBM octaves of increasing frequency (CC:247); hi-res .rgb =
 * three Worley-FBM octaves (CC:224), .a = 0 (the shipped hi-res alpha is identically 0).
 *
 * Arithmetic contract: uint32 hash (wrapping multiply/xor/shift); binary32 +,-,*,/,sqrt in the
 * order written, no FMA contraction (-ffp-contract=off / -fmad=false); floor via exact int cast.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include "oracle.h"

static inline uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
uint32_t om_noise_hash(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
    return fmix32((x * 73856093u) ^ (y * 19349663u) ^ (z * 83492791u) ^ (seed * 0x9E3779B9u));
}

/* inverted Worley F1 with `cells` cells per axis, tiling; p in [0,1)^3 */
static float worley(float px, float py, float pz, int cells, uint32_t seed) {
    float fx = px * (float)cells, fy = py * (float)cells, fz = pz * (float)cells;
    int cx = (int)fx, cy = (int)fy, cz = (int)fz;          /* p >= 0: trunc == floor */
    float best = 1.0e9f;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                uint32_t wx = (uint32_t)((nx + cells) % cells), wy = (uint32_t)((ny + cells) % cells), wz = (uint32_t)((nz + cells) % cells);
                uint32_t h = om_noise_hash(wx, wy, wz, seed);
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

static float worley_fbm(float x, float y, float z, int cells, uint32_t seed) {
    return ((0.625f * worley(x, y, z, cells, seed)) + (0.25f * worley(x, y, z, cells * 2, seed + 1u))) + (0.125f * worley(x, y, z, cells * 4, seed + 2u));
}

static const float grad12[12][3] = {
    {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}};

static inline float fade(float t) { return ((t * t) * t) * ((t * ((t * 6.0f) - 15.0f)) + 10.0f); }
static inline float lerp_(float a, float b, float t) { return a + (t * (b - a)); }

/* tiling gradient noise, `cells` lattice cells per axis, result roughly in [-1,1] */
static float perlin(float px, float py, float pz, int cells, uint32_t seed) {
    float fx = px * (float)cells, fy = py * (float)cells, fz = pz * (float)cells;
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float rx = fx - (float)ix, ry = fy - (float)iy, rz = fz - (float)iz;
    float u = fade(rx), v = fade(ry), w = fade(rz);
    float c[8];
    for (int k = 0; k < 8; k++) {
        int ox = k & 1, oy = (k >> 1) & 1, oz = (k >> 2) & 1;
        uint32_t h = om_noise_hash((uint32_t)((ix + ox) % cells), (uint32_t)((iy + oy) % cells), (uint32_t)((iz + oz) % cells), seed);
        const float *g = grad12[h % 12u];
        c[k] = ((g[0] * (rx - (float)ox)) + (g[1] * (ry - (float)oy))) + (g[2] * (rz - (float)oz));
    }
    float x00 = lerp_(c[0], c[1], u), x10 = lerp_(c[2], c[3], u), x01 = lerp_(c[4], c[5], u), x11 = lerp_(c[6], c[7], u);
    return lerp_(lerp_(x00, x10, v), lerp_(x01, x11, v), w);
}

static float perlin_fbm(float x, float y, float z, int cells, int octaves, uint32_t seed) {
    float sum = 0.0f, amp = 1.0f, tot = 0.0f;
    for (int o = 0; o < octaves; o++) {
        sum = sum + (amp * perlin(x, y, z, cells, seed + (uint32_t)o));
        tot = tot + amp;
        amp = amp * 0.5f;
        cells = cells * 2;
    }
    return sum / tot;
}

static inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
static inline uint8_t quant(float v) { return (uint8_t)(int)((clamp01(v) * 255.0f) + 0.5f); }

/* per-channel affine maps fitted once (seed 0) so that the channel means / standard deviations equal the shipped
   volumes' (low-res mean .504 .686 .687 .687, std .105 .104 .102 .102; hi-res mean .687 .690 .685, std ~.10) */
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

void om_noise_lowres_voxel(uint32_t seed, int x, int y, int z, uint8_t out[4]) {
    float px = ((float)x + 0.5f) * (1.0f / 128.0f), py = ((float)y + 0.5f) * (1.0f / 128.0f), pz = ((float)z + 0.5f) * (1.0f / 128.0f);
    float pf = (perlin_fbm(px, py, pz, 4, 5, seed) * 1.2f) + 0.5f;         /* ~[0,1] */
    float w0 = worley_fbm(px, py, pz, 4, seed + 100u);
    float w1 = worley_fbm(px, py, pz, 8, seed + 200u);
    float w2 = worley_fbm(px, py, pz, 16, seed + 300u);
    float w3 = worley_fbm(px, py, pz, 32, seed + 400u);
    float pw = w0 + (clamp01(pf) * (1.0f - w0));                          /* remap(perlin, 0, 1, worley, 1) */
    out[0] = quant((pw * NV_L0_GAIN) + NV_L0_BIAS);
    out[1] = quant((w1 * NV_L1_GAIN) + NV_L1_BIAS);
    out[2] = quant((w2 * NV_L2_GAIN) + NV_L2_BIAS);
    out[3] = quant((w3 * NV_L3_GAIN) + NV_L3_BIAS);
}

void om_noise_hires_voxel(uint32_t seed, int x, int y, int z, uint8_t out[4]) {
    float px = ((float)x + 0.5f) * (1.0f / 32.0f), py = ((float)y + 0.5f) * (1.0f / 32.0f), pz = ((float)z + 0.5f) * (1.0f / 32.0f);
    float w0 = worley_fbm(px, py, pz, 2, seed + 500u);
    float w1 = worley_fbm(px, py, pz, 4, seed + 600u);
    float w2 = worley_fbm(px, py, pz, 8, seed + 700u);
    out[0] = quant((w0 * NV_H0_GAIN) + NV_H0_BIAS);
    out[1] = quant((w1 * NV_H1_GAIN) + NV_H1_BIAS);
    out[2] = quant((w2 * NV_H2_GAIN) + NV_H2_BIAS);
    out[3] = 0;
}

void om_build_noise_volumes(uint64_t seed64, uint8_t *low, uint8_t *hi) {
    uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
    if (low) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int z = 0; z < 128; z++)
            for (int y = 0; y < 128; y++)
                for (int x = 0; x < 128; x++)
                    om_noise_lowres_voxel(seed, x, y, z, low + 4 * (((size_t)z * 128 + y) * 128 + x));
    }
    if (hi) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int z = 0; z < 32; z++)
            for (int y = 0; y < 32; y++)
                for (int x = 0; x < 32; x++)
                    om_noise_hires_voxel(seed, x, y, z, hi + 4 * (((size_t)z * 32 + y) * 32 + x));
    }
}
