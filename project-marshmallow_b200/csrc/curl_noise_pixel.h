// curl_noise_pixel.h -- the per-texel arithmetic of K2 (ImageUtils.cpp:25-223 downstream of the lattice hash), shared by the kernels of curl_noise.cu and by the
// HOST build of the same source that the CPU test-suite compares with the reference's shipped CurlNoiseFBM texture (tests/host_build/aux_host.cu).  The product
// build never defines MM_HOST_BUILD; its kernels' SASS is byte-identical to the build that had this code inside curl_noise.cu.
#pragma once
#include <math.h>

#include "common.h"

#if defined(MM_HOST_BUILD)
#define MM_HD __host__ __device__ __forceinline__
#define MM_HD_PLAIN __host__ __device__
#else
#define MM_HD __device__ __forceinline__
#define MM_HD_PLAIN __device__
#endif
#if defined(__CUDA_ARCH__) || !defined(MM_HOST_BUILD)
#define MM_CURL_LDG(p) __ldg(p)
#else
#define MM_CURL_LDG(p) (*(p))
#endif

namespace mm {
namespace curl_pixel {

#define CURL_DIM 128
#define CURL_EPS 0.0005

MM_HD float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ((ax * bx) + (ay * by)) + (az * bz);
}

#define CURL_LATTICE 26      // lattice coordinates -1..24, stored at +1

MM_HD int hash_index(const unsigned char *__restrict__ table, float x, float y, float z) {
    int ix = (int)x + 1, iy = (int)y + 1, iz = (int)z + 1;
    return MM_CURL_LDG(table + (iz * CURL_LATTICE + iy) * CURL_LATTICE + ix);
}

MM_HD void gradient(int idx, float &gx, float &gy, float &gz) {
    // the 12 edge directions scaled by 0.7071 (IU:10-23): axis pair = idx/4, signs = idx%4
    const float k = (float)0.7071;
    float s0 = (idx & 2) ? -k : k, s1 = (idx & 1) ? -k : k;
    int pair = idx >> 2;
    gx = pair == 2 ? 0.0f : s0;
    gy = pair == 0 ? s1 : (pair == 1 ? 0.0f : s0);
    gz = pair == 0 ? 0.0f : s1;
}

MM_HD float corner(const unsigned char *__restrict__ table, float cx, float cy, float cz, float dx, float dy, float dz) {
    float gx, gy, gz;
    gradient(hash_index(table, cx, cy, cz), gx, gy, gz);
    return dot3(gx, gy, gz, dx, dy, dz);
}

MM_HD float lerp_d(float a, float b, float t) {
    return (float)(((1.0 - (double)t) * (double)a) + (double)(t * b));
}

MM_HD float fade5(float r) { return ((r * r) * r) * ((r * ((r * 6.0f) - 15.0f)) + 10.0f); }

MM_HD_PLAIN float perlin(const unsigned char *__restrict__ table, float x, float y, float z, float freq) {
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

MM_HD_PLAIN float fbm(const unsigned char *__restrict__ table, float x, float y, float z, float freq, int octaves) {
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

// evaluation k = gid % 12 of pixel gid / 12: the 12 FBM probes of curlNoiseFBM in source order (IU:133-166)
MM_HD float curl_probe(const unsigned char *__restrict__ table, int gid) {
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
    return fbm(table, x, y, z, 3.f, 4);
}

MM_HD float cdiff(float a, float b) { return (float)((double)(b - a) / ((double)2.f * CURL_EPS)); }

// curl of pixel `pix` from its 12 probes (IU:130-169)
MM_HD void curl_combine(const float *fbm12, float *curls, int pix) {
    const float *f = fbm12 + 12 * pix;
    float dydx = cdiff(f[0], f[1]), dxdy = cdiff(f[2], f[3]), dxdz = cdiff(f[4], f[5]);
    float dzdx = cdiff(f[6], f[7]), dzdy = cdiff(f[8], f[9]), dydz = cdiff(f[10], f[11]);
    curls[3 * pix + 0] = dzdy - dydz;
    curls[3 * pix + 1] = dxdz - dzdx;
    curls[3 * pix + 2] = dydx - dxdy;
}

// per-channel normalisation to the global bounds and quantisation (IU:171-174, 200-217)
MM_HD uchar4 curl_quantise(const float *curls, const float *bounds, int pix) {
    unsigned char q[3];
    for (int c = 0; c < 3; c++) {
        float lo = bounds[c], hi = bounds[3 + c];
        float m = 0.f + ((curls[3 * pix + c] - lo) / (hi - lo) * (1.f - 0.f));       // IU:171-174
        q[c] = (unsigned char)((int)roundf(m * 255.f));
    }
    return make_uchar4(q[0], q[1], q[2], 255);
}

}  // namespace curl_pixel
}  // namespace mm
