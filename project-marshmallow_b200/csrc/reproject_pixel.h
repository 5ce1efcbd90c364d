// reproject_pixel.h -- the per-pixel arithmetic of K5 (reproject.comp:91-152), shared by the kernel of reproject.cu and by the HOST build of the same source
// that the CPU test-suite checks against the oracle (tests/host_build/aux_host.cu).  On the device nothing changes (the kernel's SASS is byte-identical to the
// build that had this code inside reproject_kernel).
#pragma once
#include <math.h>

#include "common.h"

#define MM_HD __host__ __device__ __forceinline__

namespace mm {
namespace reproject_pixel {

struct v3 { float x, y, z; };
MM_HD v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
MM_HD float dot(v3 a, v3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
MM_HD v3 normalize(v3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return V3(a.x * inv, a.y * inv, a.z * inv); }
MM_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// ivec2(float): truncate, saturate, NaN -> 0 (cvt.rzi.s32.f32)
MM_HD int to_int(float f) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(f);
#else
    if (!(f == f)) return 0;                           // host build: the same saturating conversion spelled out
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return -2147483647 - 1;
    return (int)f;
#endif
}
// read-only load of one previous-image texel
MM_HD float4 load_texel(const float *img, size_t pitch, int x, int y) {
    const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(img) + (size_t)y * pitch) + x;
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}


// reproject.comp:56-86; only .point is consumed (vec3(0) on a miss, as initialised at :58-62)
MM_HD v3 shellHitPoint(v3 ro, v3 rd, v3 c, float w) {
    v3 o = V3((ro.x - c.x) / w, (ro.y - c.y) / w, (ro.z - c.z) / w);
    float A = dot(rd, rd);
    float B = 2.0f * dot(rd, o);
    float C = dot(o, o) - 0.25f;
    float disc = (B * B) - ((4.0f * A) * C);
    if (disc < 0.0f) return V3(0.f, 0.f, 0.f);
    float t = (((-sqrtf(disc)) - B) / A) * 0.5f;
    if (t < 0.0f) t = ((sqrtf(disc) - B) / A) * 0.5f;
    if (t >= 0.0f) {
        v3 p = V3(o.x + (rd.x * t), o.y + (rd.y * t), o.z + (rd.z * t));
        return V3((p.x * w) + c.x, (p.y * w) + c.y, (p.z * w) + c.z);
    }
    return V3(0.f, 0.f, 0.f);
}

// targetImage texel (gx, gy): reproject.comp:94-151
MM_HD float4 reproject_texel(const ReprojectParams &P, int gx, int gy) {
    const float *cam = P.cam, *prev = P.cam_prev;
    const float dimx = (float)P.W, dimy = (float)P.H;
    float uvx = (float)gx / dimx, uvy = (float)gy / dimy;                                   // :94
    float spx = (uvx * 2.0f) - 1.0f, spy = (uvy * 2.0f) - 1.0f;                             // :99
    v3 camLook = V3(cam[2], cam[6], cam[10]), camRight = V3(cam[0], cam[4], cam[8]), camUp = V3(cam[1], cam[5], cam[9]);
    v3 cameraPos = V3(cam[32], cam[33], cam[34]);
    float aspect = cam[36], tanH = cam[37];
    v3 ref = V3(cameraPos.x - camLook.x, cameraPos.y - camLook.y, cameraPos.z - camLook.z);
    float sr = (aspect * spx) * tanH, su = spy * tanH;                                       // :111
    v3 p = V3((ref.x + (sr * camRight.x)) - (su * camUp.x), (ref.y + (sr * camRight.y)) - (su * camUp.y), (ref.z + (sr * camRight.z)) - (su * camUp.z));
    v3 rd = normalize(V3(p.x - cameraPos.x, p.y - cameraPos.y, p.z - cameraPos.z));
    v3 earthCenter = V3(cameraPos.x, (-2000000.0f * 0.5f) * 0.995f, cameraPos.z);             // :115-117
    v3 hit = shellHitPoint(cameraPos, rd, earthCenter, 2000000.0f);                           // :120
    v3 q = V3((((prev[0] * hit.x) + (prev[4] * hit.y)) + (prev[8] * hit.z)) + (prev[12] * 1.0f),      // :125
              (((prev[1] * hit.x) + (prev[5] * hit.y)) + (prev[9] * hit.z)) + (prev[13] * 1.0f),
              (((prev[2] * hit.x) + (prev[6] * hit.y)) + (prev[10] * hit.z)) + (prev[14] * 1.0f));
    v3 od = normalize(q);                                                                    // :128
    float nz = -od.z;                                                                        // :132
    od = V3(od.x / nz, od.y / nz, od.z / nz);
    float oldU = (((od.x / tanH) / aspect) * 0.5f) + 0.5f;                                   // :133-134
    float oldV = (((-od.y) / tanH) * 0.5f) + 0.5f;                                           // :135-136
    float bvx = oldU - uvx, bvy = oldV - uvy;                                                // :138

    // :142-147, ten taps along the motion vector.  The tap index is a monotonic function of the tap number (every operation of
    // round((old - bv*k) * dim) is monotonic, k = s/9 - 0.5 increases with s), so when the FIRST and the LAST tap select the same
    // source pixel all ten do: one load, and the ten ordered additions of the reference on that one value.  That is every pixel of
    // a frame whose camera moved by less than a pixel over the blur span (a static camera in particular); otherwise the full loop.
    auto tap_index = [&](int s, int &sx, int &sy) {
        float k = ((float)s / 9.0f) - 0.5f;
        float ix = roundf((oldU - (bvx * k)) * dimx), iy = roundf((oldV - (bvy * k)) * dimy);
        sx = clampi(to_int(ix), 0, P.W - 1); sy = clampi(to_int(iy), 0, P.H - 1);
    };
    auto tap_load = [&](int sx, int sy) { return load_texel(P.src, P.src_pitch, sx, sy); };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int x0, y0, x9, y9;
    tap_index(0, x0, y0);
    tap_index(9, x9, y9);
    if (x0 == x9 && y0 == y9) {
        float4 t = tap_load(x0, y0);
#pragma unroll
        for (int s = 0; s < 10; ++s) { acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
    } else {
#pragma unroll
        for (int s = 0; s < 10; ++s) {
            int sx, sy;
            tap_index(s, sx, sy);
            float4 t = tap_load(sx, sy);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
    }
    acc.x = acc.x / 10.0f; acc.y = acc.y / 10.0f; acc.z = acc.z / 10.0f;                      // :148
    int cx = clampi(to_int(roundf(oldU * dimx)), 0, P.W - 1), cy = clampi(to_int(roundf(oldV * dimy)), 0, P.H - 1);   // :150-151
    acc.w = load_texel(P.src, P.src_pitch, cx, cy).w;
    return acc;
}

}  // namespace reproject_pixel
}  // namespace mm
