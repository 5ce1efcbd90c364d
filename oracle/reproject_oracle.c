/*
 * reproject_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restatement of the reference's reprojection pass
 *     /root/reference/SkyEngine/SkyEngine/Shaders/reproject.comp:91-152
 * (the step before the cloud dispatch in the engine's frame,
 * VulkanApplication.cpp:1053-1071): every pixel re-aims its ray at the inner atmosphere shell, expresses the hit
 * in the PREVIOUS camera's view space, and averages 10 taps of the previous image along the motion vector.
 * Same arithmetic contract as cloud_march_oracle.c (binary32, GLSL order, no contraction).  Additional definitions
 * where GLSL leaves freedom: round() = round half away from zero (roundf); ivec2(float) truncates, saturates at the
 * int32 range and sends NaN to 0 (what CUDA's cvt.rzi.s32.f32 does); a ray that misses the shell keeps
 * isect.point = vec3(0) exactly as reproject.comp:58-62 initialises it.
 * PARITY STATUS: pinned against the reference's own shader text -- reproject.comp rewritten lexically into C++ and run on the CPU
 * (oracle/_ref/libref_passes.so, oracle/glsl_env.h): images bit-identical (tests/test_reference_shader.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "oracle.h"

typedef struct { float x, y, z; } v3;
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline float dot3(v3 a, v3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
static inline v3 normalize3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return V3(a.x * inv, a.y * inv, a.z * inv); }

static inline int sat_int(float f) {
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* reproject.comp:56-86: only .point is used by the caller */
static v3 shellHitPoint(v3 ro, v3 rd, v3 c, float w) {
    v3 o = V3((ro.x - c.x) / w, (ro.y - c.y) / w, (ro.z - c.z) / w);
    float A = dot3(rd, rd);
    float B = 2.0f * dot3(rd, o);
    float C = dot3(o, o) - 0.25f;
    float disc = (B * B) - ((4.0f * A) * C);
    if (disc < 0.0f) return V3(0, 0, 0);
    float t = (((-sqrtf(disc)) - B) / A) * 0.5f;
    if (t < 0.0f) t = ((sqrtf(disc) - B) / A) * 0.5f;
    if (t >= 0.0f) {
        v3 p = V3(o.x + (rd.x * t), o.y + (rd.y * t), o.z + (rd.z * t));
        return V3((p.x * w) + c.x, (p.y * w) + c.y, (p.z * w) + c.z);
    }
    return V3(0, 0, 0);
}

/* camera160 / cameraPrev160: UniformCameraObject blocks (Shader.h:24-29).  src and dst: W*H float4, packed. */
int om_reproject(const void *camera160, const void *cameraPrev160, const float *src, int W, int H, float *dst) {
    if (!camera160 || !cameraPrev160 || !src || !dst || W <= 0 || H <= 0) return -1;
    float cam[40], prev[40];
    memcpy(cam, camera160, 160);
    memcpy(prev, cameraPrev160, 160);
    const float dimx = (float)W, dimy = (float)H;
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < H; gy++)
        for (int gx = 0; gx < W; gx++) {
            float uvx = (float)gx / dimx, uvy = (float)gy / dimy;                        /* :94 */
            float spx = (uvx * 2.0f) - 1.0f, spy = (uvy * 2.0f) - 1.0f;                  /* :99 */
            v3 camLook = V3(cam[2], cam[6], cam[10]), camRight = V3(cam[0], cam[4], cam[8]), camUp = V3(cam[1], cam[5], cam[9]);
            v3 cameraPos = V3(cam[32], cam[33], cam[34]);
            float aspect = cam[36], tanH = cam[37];
            v3 ref = V3(cameraPos.x - camLook.x, cameraPos.y - camLook.y, cameraPos.z - camLook.z);
            float sr = (aspect * spx) * tanH, su = spy * tanH;                            /* :111 */
            v3 p = V3((ref.x + (sr * camRight.x)) - (su * camUp.x), (ref.y + (sr * camRight.y)) - (su * camUp.y), (ref.z + (sr * camRight.z)) - (su * camUp.z));
            v3 rd = normalize3(V3(p.x - cameraPos.x, p.y - cameraPos.y, p.z - cameraPos.z));
            v3 earthCenter = V3(cameraPos.x, (-2000000.0f * 0.5f) * 0.995f, cameraPos.z);  /* :115-117 */
            v3 hit = shellHitPoint(cameraPos, rd, earthCenter, 2000000.0f);                /* :120 */
            /* :125  cameraPrev.view * vec4(hit, 1): column-major, ((c0*x + c1*y) + c2*z) + c3*1 */
            v3 q = V3((((prev[0] * hit.x) + (prev[4] * hit.y)) + (prev[8] * hit.z)) + (prev[12] * 1.0f),
                      (((prev[1] * hit.x) + (prev[5] * hit.y)) + (prev[9] * hit.z)) + (prev[13] * 1.0f),
                      (((prev[2] * hit.x) + (prev[6] * hit.y)) + (prev[10] * hit.z)) + (prev[14] * 1.0f));
            v3 od = normalize3(q);                                                        /* :128 */
            float nz = -od.z;                                                             /* :132 */
            od = V3(od.x / nz, od.y / nz, od.z / nz);
            float oldU = (((od.x / tanH) / aspect) * 0.5f) + 0.5f;                        /* :133-134 */
            float oldV = (((-od.y) / tanH) * 0.5f) + 0.5f;                                /* :135-136 */
            float bvx = oldU - uvx, bvy = oldV - uvy;                                     /* :138 */
            float acc[4] = {0, 0, 0, 0};
            for (int s = 0; s < 10; ++s) {                                                /* :142-147 */
                float k = ((float)s / 9.0f) - 0.5f;
                float ix = roundf((oldU - (bvx * k)) * dimx), iy = roundf((oldV - (bvy * k)) * dimy);
                int sx = clampi(sat_int(ix), 0, W - 1), sy = clampi(sat_int(iy), 0, H - 1);
                const float *t = src + 4 * ((size_t)sy * W + sx);
                for (int c = 0; c < 4; c++) acc[c] += t[c];
            }
            for (int c = 0; c < 4; c++) acc[c] = acc[c] / 10.0f;                          /* :148 */
            int cx = clampi(sat_int(roundf(oldU * dimx)), 0, W - 1), cy = clampi(sat_int(roundf(oldV * dimy)), 0, H - 1);   /* :150-151 */
            acc[3] = src[4 * ((size_t)cy * W + cx) + 3];
            memcpy(dst + 4 * ((size_t)gy * W + gx), acc, 16);
        }
    return 0;
}
