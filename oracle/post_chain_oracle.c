/*
 * post_chain_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) of the three full-screen passes that
 * consume the cloud image in the reference's frame:
 *     Shaders/god-ray.frag:41-76      radial light shafts from the alpha channel (8 taps toward the sun)
 *     Shaders/radialBlur.frag:36-63   10-tap radial blur of that alpha, added as sun-coloured light
 *     Shaders/tonemap.frag:11-33      Uncharted-2 tone map, gamma, vignette -> swapchain UNORM8
 * recorded back to back over three RGBA32F framebuffers in VulkanApplication.cpp:936-968 and :1016 (the first
 * pass, background.frag, is a texel-for-texel copy of the cloud image and is not restated).
 *
 * PARITY STATUS: pinned against the reference's own shader texts (god-ray.frag, radialBlur.frag, tonemap.frag rewritten lexically
 * into C++ and run on the CPU, oracle/_ref/libref_passes.so): both RGBA32F framebuffers bit-identical, the UNORM8 present
 * byte-identical (tests/test_reference_shader.py).  The reference has no tests or golden images of its own.  Restated under the
 * arithmetic contract of cloud_march_oracle.c: IEEE binary32, GLSL order, one rounding per operator, no contraction.
 * Fixed interpretations:
 *   - fragUV of the pixel (x, y) is ((x + 0.5)/W, (y + 0.5)/H): the quad's UVs (Geometry.cpp:94-99) interpolated at the
 *     pixel centre; texture(texColor, fragUV) at that centre is the texel itself (weights 0 at any filter precision);
 *   - the off-centre taps use the offscreen sampler (VulkanApplication.cpp:1290-1303: LINEAR, CLAMP_TO_EDGE) on an
 *     RGBA32F image: U = u*W - 0.5, i0 = floor(U), a = U - i0, indices clamped to [0, W-1], fused lerps in x then y
 *     (only the alpha channel is ever read at a tap);
 *   - mat4 * mat4 and mat4 * vec4 accumulate left to right: ((m0*v0 + m1*v1) + m2*v2) + m3*v3;
 *   - `camera.proj * camera.view * sun.location` is (proj * view) * location, as GLSL parses it;
 *   - the macro SAMPLE_WEIGHT expands textually: `x *= 1.0 / float(8) * d` is x = x * ((1.0/8.0) * d).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"

static inline float clampf(float x, float lo, float hi) { float r = (x > lo) ? x : lo; return (r < hi) ? r : hi; }
static inline float lerpf(float p, float q, float a) { return __builtin_fmaf(a, q - p, p); }

/* (proj * view) * sun.location, then the perspective divide (god-ray.frag:44-46, radialBlur.frag:48-49) */
void om_sun_screen_position(const void *camera160, const void *sun116, float out_xy[2]) {
    const float *cam = (const float *)camera160, *sun = (const float *)sun116;
    const float *V = cam, *P = cam + 16;                                     /* column-major: m[col*4 + row] */
    float PV[16];
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++)
            PV[j * 4 + i] = (((P[0 * 4 + i] * V[j * 4 + 0]) + (P[1 * 4 + i] * V[j * 4 + 1])) + (P[2 * 4 + i] * V[j * 4 + 2])) + (P[3 * 4 + i] * V[j * 4 + 3]);
    float r[4];
    for (int i = 0; i < 4; i++)
        r[i] = (((PV[0 * 4 + i] * sun[0]) + (PV[1 * 4 + i] * sun[1])) + (PV[2 * 4 + i] * sun[2])) + (PV[3 * 4 + i] * sun[3]);
    out_xy[0] = r[0] / r[3];
    out_xy[1] = r[1] / r[3];
}

/* one axis of the LINEAR / CLAMP_TO_EDGE footprint */
static inline void tap_axis(float u, int n, int *i0, int *i1, float *a) {
    float U = (u * (float)n) - 0.5f;
    float fl = floorf(U);
    *a = U - fl;
    if (!(fl >= -1.0f)) fl = -1.0f;                                          /* far off-screen (or NaN): clamped below anyway */
    if (fl > (float)n) fl = (float)n;
    int i = (int)fl;
    *i0 = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    *i1 = i + 1 < 0 ? 0 : (i + 1 > n - 1 ? n - 1 : i + 1);
}
static float tap_alpha(const float *img, int W, int H, float u, float v) {
    int x0, x1, y0, y1; float a, b;
    tap_axis(u, W, &x0, &x1, &a);
    tap_axis(v, H, &y0, &y1, &b);
    float t00 = img[4 * ((size_t)y0 * W + x0) + 3], t10 = img[4 * ((size_t)y0 * W + x1) + 3];
    float t01 = img[4 * ((size_t)y1 * W + x0) + 3], t11 = img[4 * ((size_t)y1 * W + x1) + 3];
    return lerpf(lerpf(t00, t10, a), lerpf(t01, t11, a), b);
}

/* god-ray.frag:41-76 */
int om_god_ray(const void *camera160, const void *sun116, const float *src, int W, int H, float *dst) {
    if (!camera160 || !sun116 || !src || !dst || W <= 0 || H <= 0) return -1;
    const float *sun = (const float *)sun116;
    float sp[2];
    om_sun_screen_position(camera160, sun116, sp);
    const float k = (1.0f / 8.0f) * 0.75f;                                   /* SAMPLE_WEIGHT * DENSITY, :47 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const float *cf = src + 4 * ((size_t)y * W + x);
            float *o = dst + 4 * ((size_t)y * W + x);
            o[0] = cf[0]; o[1] = cf[1]; o[2] = cf[2];
            if (sun[5] < 0.0f) { o[3] = 1.0f; continue; }                    /* :51-53 */
            float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
            float cx = (u * 2.0f) - 1.0f, cy = (v * 2.0f) - 1.0f;            /* :42-43 */
            float dx = (cx - sp[0]) * k, dy = (cy - sp[1]) * k;              /* :46-47 */
            float accum = cf[3] * 0.5f;                                      /* :55 */
            float decay = 1.0f;
            for (int i = 0; i < 8; i++) {                                    /* :58-73 */
                cx = cx - dx; cy = cy - dy;
                float s = tap_alpha(src, W, H, (cx * 0.5f) + 0.5f, (cy * 0.5f) + 0.5f) * 0.5f;
                s = s * ((1.0f / 8.0f) * decay);
                accum = accum + s;
                decay = decay * 0.99f;
            }
            o[3] = accum * 0.9f;                                             /* :75 */
        }
    }
    return 0;
}

/* radialBlur.frag:36-63 */
int om_radial_blur(const void *camera160, const void *sun116, const float *src, int W, int H, float *dst) {
    if (!camera160 || !sun116 || !src || !dst || W <= 0 || H <= 0) return -1;
    const float *sun = (const float *)sun116;
    static const float samples[10] = {-0.08f, -0.05f, -0.03f, -0.02f, -0.01f, 0.01f, 0.02f, 0.03f, 0.05f, 0.08f};
    float sp[2];
    om_sun_screen_position(camera160, sun116, sp);
    float cr = sun[8] * sun[28], cg = sun[9] * sun[28], cb = sun[10] * sun[28];   /* sun.color.xyz * sun.intensity, :62 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const float *cf = src + 4 * ((size_t)y * W + x);
            float *o = dst + 4 * ((size_t)y * W + x);
            o[3] = 1.0f;
            if (sun[5] < 0.0f) { o[0] = cf[0]; o[1] = cf[1]; o[2] = cf[2]; continue; }    /* :41-43 */
            float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
            float sx = (u * 2.0f) - 1.0f, sy = (v * 2.0f) - 1.0f;
            float lx = sp[0] - sx, ly = sp[1] - sy;                          /* :51 */
            float dist = sqrtf((lx * lx) + (ly * ly));
            lx = lx / dist; ly = ly / dist;
            float accum = 0.0f;
            for (int i = 0; i < 10; i++) {                                   /* :57-59 */
                float px = sx + (((samples[i] * lx) * 1.5f) * dist), py = sy + (((samples[i] * ly) * 1.5f) * dist);
                accum = accum + (tap_alpha(src, W, H, (px * 0.5f) + 0.5f, (py * 0.5f) + 0.5f) * 1.1f);
            }
            accum = accum / 10.0f;                                           /* :60 */
            o[0] = (cr * accum) + (0.5f * cf[0]);                            /* :62 */
            o[1] = (cg * accum) + (0.5f * cf[1]);
            o[2] = (cb * accum) + (0.5f * cf[2]);
        }
    }
    return 0;
}

/* tonemap.frag:11-33 -> UNORM8 (round half up), alpha = 1.  bgra != 0 writes the swapchain's B8G8R8A8 byte order
 * (VulkanApplication.cpp:1436-1446). */
static inline float uc2(float x) {
    return (((x * ((0.15f * x) + (0.1f * 0.5f))) + (0.2f * 0.02f)) / ((x * ((0.15f * x) + 0.5f)) + (0.2f * 0.3f))) - (0.02f / 0.3f);
}
int om_tonemap_present(const float *src, int W, int H, int bgra, uint8_t *dst) {
    if (!src || !dst || W <= 0 || H <= 0) return -1;
    float whitemap = 1.0f / uc2(50.2f);
    static const float vc[3] = {0.1f, 0.05f, 0.13f};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const float *c = src + 4 * ((size_t)y * W + x);
            float u = (((float)x + 0.5f) / (float)W) - 0.5f, v = (((float)y + 0.5f) / (float)H) - 0.5f;
            float vig = (u * u) + (v * v);                                   /* :30 */
            uint8_t q[3];
            for (int k = 0; k < 3; k++) {
                float col = powf(uc2(0.7f * c[k]) * whitemap, 1.0f / 2.2f);  /* :14-20, 27-28 */
                col = (col * (1.0f - vig)) + (vc[k] * vig);                  /* :32 */
                q[k] = (uint8_t)floorf((255.0f * clampf(col, 0.0f, 1.0f)) + 0.5f);
            }
            uint8_t *o = dst + 4 * ((size_t)y * W + x);
            o[0] = bgra ? q[2] : q[0]; o[1] = q[1]; o[2] = bgra ? q[0] : q[2]; o[3] = 255;
        }
    }
    return 0;
}
