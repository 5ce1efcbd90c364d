/*
 * curl_noise_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restatement of the reference's offline curl-noise generator
 *     /root/reference/SkyEngine/SkyEngine/ImageUtils.cpp:25-223   ("IU")
 * following its float/double promotions exactly as g++ resolves them IN THAT TRANSLATION UNIT:
 * stb_image_write.h pulls in libstdc++'s <math.h> wrapper, which injects the float overloads of
 * sin/floor/fabs into the global namespace, so the unqualified calls on float arguments are sinf /
 * floorf / fabsf (MSVC, the reference's real toolchain, resolves them the same way; SURVEY.md section 7
 * hard part 5 guessed "double" -- measured here: the double path changes 214 of 900 probed hashes and
 * does NOT reproduce the shipped texture, the float path does).  EPS and the 1.0 in lerp are double
 * literals and do promote.  Everything else is binary32.  Build with -ffp-contract=off.
 *
 * PINNED: output equals Textures/CurlNoiseFBM.tga (the reference's shipped golden vector) byte for
 * byte, and equals the reference's own ImageUtils.cpp compiled verbatim (oracle/_ref) -- checked
 * by tests/test_curl_noise_oracle.py.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "oracle.h"

#define CURL_DIM 128                 /* IU:7 */
#define EPS 0.0005                   /* IU:8 (double) */

typedef struct { float x, y, z; } f3;

static const f3 basis[12] = {        /* IU:10-23: literals are double, narrowed to float by glm::vec3 */
    {(float)0.7071, (float)0.7071, 0}, {(float)0.7071, (float)-0.7071, 0}, {(float)-0.7071, (float)0.7071, 0}, {(float)-0.7071, (float)-0.7071, 0},
    {(float)0.7071, 0, (float)0.7071}, {(float)0.7071, 0, (float)-0.7071}, {(float)-0.7071, 0, (float)0.7071}, {(float)-0.7071, 0, (float)-0.7071},
    {0, (float)0.7071, (float)0.7071}, {0, (float)0.7071, (float)-0.7071}, {0, (float)-0.7071, (float)0.7071}, {0, (float)-0.7071, (float)-0.7071},
};

static inline float dotf(f3 a, f3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }   /* glm compute_dot<vec3> */

/* IU:25-29 */
static float hashNoise(float x, float y, float z) {
    f3 k = {12.9898f, 78.233f, (float)47.387};
    f3 p = {x, y, z};
    float n = sinf(dotf(p, k)) * 43758.5453f;
    n = n - floorf(n);
    return n;
}
float om_curl_hash(float x, float y, float z) { return hashNoise(x, y, z); }

/* IU:31-34 */
static int hashIndex(float x, float y, float z) {
    float a = hashNoise(x, y, z);
    return (int)floorf(12.f * a);
}
int om_curl_hash_index(float x, float y, float z) { return hashIndex(x, y, z); }
static f3 hashVec(float x, float y, float z) { return basis[hashIndex(x, y, z)]; }

/* IU:36-38 */
static float lerp_(float a, float b, float t) { return (float)(((1.0 - (double)t) * (double)a) + (double)(t * b)); }

/* IU:41-73 */
static float perlinNoise(f3 pt, float freq) {
    pt.x *= freq; pt.y *= freq; pt.z *= freq;
    f3 f = {floorf(pt.x), floorf(pt.y), floorf(pt.z)};
    f3 r = {pt.x - f.x, pt.y - f.y, pt.z - f.z};
    f3 u;
    u.x = ((r.x * r.x) * r.x) * ((r.x * ((r.x * 6.0f) - 15.0f)) + 10.0f);
    u.y = ((r.y * r.y) * r.y) * ((r.y * ((r.y * 6.0f) - 15.0f)) + 10.0f);
    u.z = ((r.z * r.z) * r.z) * ((r.z * ((r.z * 6.0f) - 15.0f)) + 10.0f);

    f3 f1 = {f.x + 1.0f, f.y + 1.0f, f.z + 1.0f};
    /* IU:49-51 "force tiling": as written these assign freq (NOT 0); kept */
    if (fabsf(f1.x - freq) < 0.001f) f1.x = freq;
    if (fabsf(f1.y - freq) < 0.001f) f1.y = freq;
    if (fabsf(f1.z - freq) < 0.001f) f1.z = freq;

    f3 d;
    d = (f3){r.x, r.y, r.z};                      float nnn = dotf(hashVec(f.x,  f.y,  f.z ), d);
    d = (f3){r.x, r.y, r.z - 1.0f};               float nnp = dotf(hashVec(f.x,  f.y,  f1.z), d);
    d = (f3){r.x, r.y - 1.0f, r.z};               float npn = dotf(hashVec(f.x,  f1.y, f.z ), d);
    d = (f3){r.x, r.y - 1.0f, r.z - 1.0f};        float npp = dotf(hashVec(f.x,  f1.y, f1.z), d);
    d = (f3){r.x - 1.0f, r.y, r.z};               float pnn = dotf(hashVec(f1.x, f.y,  f.z ), d);
    d = (f3){r.x - 1.0f, r.y, r.z - 1.0f};        float pnp = dotf(hashVec(f1.x, f.y,  f1.z), d);
    d = (f3){r.x - 1.0f, r.y - 1.0f, r.z};        float ppn = dotf(hashVec(f1.x, f1.y, f.z ), d);
    d = (f3){r.x - 1.0f, r.y - 1.0f, r.z - 1.0f}; float ppp = dotf(hashVec(f1.x, f1.y, f1.z), d);

    float nn = lerp_(nnn, pnn, u.x);
    float np = lerp_(nnp, pnp, u.x);
    float pn = lerp_(npn, ppn, u.x);
    float pp = lerp_(npp, ppp, u.x);
    float n = lerp_(nn, pn, u.y);
    float p = lerp_(np, pp, u.y);
    return lerp_(n, p, u.z);
}

/* IU:75-87 */
static float FBM(f3 pt, float freq, int octaves) {
    float noise = 0.0f, weight = 1.0f, persistence = 0.4f, totalWeight = 0.0f;
    for (int i = 0; i < octaves; i++) {
        totalWeight += weight;
        noise += weight * perlinNoise(pt, freq);
        freq *= 2.0f;
        weight *= persistence;
    }
    return noise / totalWeight;
}

static inline float minusEps(float v) { return (float)((double)v - EPS); }
static inline float plusEps(float v) { return (float)((double)v + EPS); }
static inline float fdiff(float a, float b) { return (float)((double)(b - a) / ((double)2.f * EPS)); }

/* IU:130-169 */
static f3 curlNoiseFBM(float px, float py, float freq, int octaves) {
    float a, b;
    a = FBM((f3){minusEps(px), py, 0.5f}, freq, octaves);
    b = FBM((f3){plusEps(px), py, 0.5f}, freq, octaves);
    float dydx = fdiff(a, b);
    a = FBM((f3){px, minusEps(py), 0.5f}, freq, octaves);
    b = FBM((f3){px, plusEps(py), 0.5f}, freq, octaves);
    float dxdy = fdiff(a, b);
    a = FBM((f3){px, 0.5f, minusEps(py)}, freq, octaves);
    b = FBM((f3){px, 0.5f, plusEps(py)}, freq, octaves);
    float dxdz = fdiff(a, b);
    a = FBM((f3){minusEps(px), 0.5f, py}, freq, octaves);
    b = FBM((f3){plusEps(px), 0.5f, py}, freq, octaves);
    float dzdx = fdiff(a, b);
    a = FBM((f3){(float)0.5, minusEps(py), px}, freq, octaves);
    b = FBM((f3){(float)0.5, plusEps(py), px}, freq, octaves);
    float dzdy = fdiff(a, b);
    a = FBM((f3){0.5f, py, minusEps(px)}, freq, octaves);
    b = FBM((f3){0.5f, py, plusEps(px)}, freq, octaves);
    float dydz = fdiff(a, b);
    return (f3){dzdy - dydz, dxdz - dzdx, dydx - dxdy};
}

/* IU:171-174 */
static float remap_(float x, float oldMin, float oldMax, float newMin, float newMax) {
    return newMin + ((x - oldMin) / (oldMax - oldMin) * (newMax - newMin));
}

/* IU:176-223, writing the RGBA8 pixels instead of a TGA file */
void om_generate_curl_noise(uint8_t *pixels) {
    f3 *curls = (f3 *)malloc(sizeof(f3) * CURL_DIM * CURL_DIM);
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma omp parallel for schedule(dynamic, 4)
    for (int row = 0; row < CURL_DIM; row++)
        for (int col = 0; col < CURL_DIM; col++)
            curls[row * CURL_DIM + col] = curlNoiseFBM((float)col / CURL_DIM, (float)row / CURL_DIM, 3.f, 4);
    for (int i = 0; i < CURL_DIM * CURL_DIM; i++) {
        float v[3] = {curls[i].x, curls[i].y, curls[i].z};
        for (int c = 0; c < 3; c++) { if (v[c] < lo[c]) lo[c] = v[c]; if (v[c] > hi[c]) hi[c] = v[c]; }
    }
    for (int i = 0; i < CURL_DIM * CURL_DIM; i++) {
        float v[3] = {curls[i].x, curls[i].y, curls[i].z};
        for (int c = 0; c < 3; c++) {
            float m = remap_(v[c], lo[c], hi[c], 0.f, 1.f);
            pixels[4 * i + c] = (uint8_t)((int)roundf(m * 255.f));
        }
        pixels[4 * i + 3] = 255;
    }
    free(curls);
}
