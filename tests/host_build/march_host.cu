// march_host.cu -- TEST INFRASTRUCTURE: the per-ray source of K1 (csrc/cloud_march_ray.inl: ray_setup, cloudTest with its exact early-outs and the pow filter,
// cloudHiRes, the samplers, the relaxed light sample, litTerm, ray_finish) compiled for the HOST (MM_HOST_BUILD) and driven ray by ray on the CPU, so that the CPU
// test-suite (tests/test_host_build_march.py) can hold the product's own arithmetic to the oracle without a GPU.  What is NOT the product's source here is the
// ~40-line scalar loop below, which does for one ray what warp_trip + warpSharedLightSamples (cloud_march.cu) do for a warp -- same calls, same order -- and the
// host-side preparation capi.cu / pack_pairs_kernel do on the device side (cone samples, pair-major texel copies).  Texture-unit fetches go to a sampler callback
// (the oracle's bit-exact model of the unit).  Build with -DMM_FMA=1 for the contracted arithmetic definition.  Nothing here is linked into the product library.
#define MM_HOST_BUILD 1
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../project-marshmallow_b200/csrc/common.h"

namespace mm_host {
typedef void (*sampler_fn)(void *user, int slot, const float *uvw, float *out_rgba);
static sampler_fn g_sampler = nullptr;
static void *g_user = nullptr;
static inline float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline float2 fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float satf(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }          // __saturatef: NaN -> 0
static inline long long d2ll(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double ll2d(long long x) { double r; memcpy(&r, &x, 8); return r; }
// the texture unit: TexDev::obj carries slot + 1 in this build
static inline float4 tex(const mm::TexDev &t, float u, float v, float w, int) {
    float uvw[3] = {u, v, w}, o[4];
    g_sampler(g_user, (int)t.obj - 1, uvw, o);
    return make_float4(o[0], o[1], o[2], o[3]);
}
}  // namespace mm_host

namespace mm {
namespace {

#include "../../project-marshmallow_b200/csrc/cloud_march_ray.inl"

// one ray of CC:288-500: warp_trip (cloud_march.cu) for a single lane, its six light-cone samples evaluated in place and summed in the reference order
template <bool MARCH_HW, bool LIGHT_HW, bool CNT, bool P2>
static float4 march_ray(const MarchParams &P, int px, int py, Counters &cn) {
    Ray r;
    r.alive = false;
    cn.trips = cn.n2d = cn.n3d = cn.lit = 0u;
    ray_setup<MARCH_HW, CNT>(P, px, py, r, cn);
    const float timeOffset = P.sky[11];
    const v3 windXYZ = V3(P.sky[8], P.sky[9], P.sky[10]);
    const v3 cameraPos = V3(P.cam[32], P.cam[33], P.cam[34]);
    const v3 earthCenter = V3(cameraPos.x, (-ATMOSPHERE_RADIUS * 0.5f) * 0.995f, cameraPos.z);
    while (r.alive) {
        bool lit = false, skipTail = false;
        float density = 0.0f, loDensity = 0.0f, h = 0.0f;
        if (CNT) cn.trips++;
        v3 pos = mad3(r.t, r.rd, cameraPos);
        v3 proj = projectedShellPoint(pos, earthCenter);
        h = relativeHeight(pos, proj);
        v3 wo = windOffsetAt(windXYZ, timeOffset, h);
        density = cloudTest<MARCH_HW, CNT, P2>(P, pos + wo, h, earthCenter, cameraPos, cn);
        loDensity = density;
        if (density > 0.0f) {
            r.misses = 0;
            if (r.noHits) {
                r.t -= r.stepSize;
                r.stepSize *= 0.3f;
                r.noHits = false;
                skipTail = true;
            } else {
                density = cloudHiRes<MARCH_HW, CNT, P2>(P, pos + wo, r.stepSize, density, h, cn);
                if (density < 0.0001f) skipTail = true;
                else lit = true;
            }
        } else if (!r.noHits) {
            r.misses++;
            if (r.misses >= 10) {
                r.noHits = true;
                r.stepSize = DIVC(r.stepSize, 0.3f);
            }
        }
        if (lit) {
            if (CNT) cn.lit++;
            float dal = 0.0f;
            unsigned nh = 0u;
            for (int i = 0; i < 6; i++) {
                v3 smp = V3(P.light[3 * i], P.light[3 * i + 1], P.light[3 * i + 2]);
                v3 lsPos = mad3(3.0f * r.stepSize, smp, pos);
                float contrib = 0.0f;
                if (LIGHT_HW && !CNT) {
                    contrib = lightSampleFast(P, lsPos, r.stepSize, earthCenter, cameraPos, windXYZ, timeOffset);
                } else {
                    v3 lsProj = projectedShellPoint(lsPos, earthCenter);
                    float lsH = relativeHeight(lsPos, lsProj);
                    v3 lwo = windOffsetAt(windXYZ, timeOffset, lsH);
                    float lsD = cloudTest<LIGHT_HW, false, P2>(P, lsPos + lwo, lsH, earthCenter, cameraPos, cn);
                    if (lsD > 0.0f) { contrib = cloudHiRes<LIGHT_HW, false, P2>(P, lsPos + lwo, r.stepSize, lsD, lsH, cn); nh++; }
                }
                dal += contrib;
            }
            if (CNT) { cn.n2d += 6 + nh; cn.n3d += 6 + nh; }
            r.transmittance = mixg(r.transmittance, litTerm(dal, loDensity, h, r.cosTheta, r.hg), (1.0f - r.accum));
            r.accum += density;
        }
        if (!skipTail) {
            if (r.accum > 0.99f) {
                r.accum = 1.0f;
                r.alive = false;
            } else if (++r.steps > MAX_STEPS) {
                r.alive = false;
            }
        }
        if (r.alive) {
            r.t += r.stepSize;
            r.alive = r.t < r.tOuter;
        }
    }
    return ray_finish(P, r);
}

// capi.cu, light_cone_samples: mat3(sun.directionBasis) * s_i in the definition's order
static void cone_samples(const float *sun, float out[18]) {
    static const float sv[6][3] = {{0.f, 0.6f, 0.f}, {0.f, 0.5f, 0.05f}, {0.1f, 0.75f, 0.f}, {0.2f, 2.5f, 0.3f}, {0.f, 6.f, 0.f}, {-0.1f, 1.f, -0.2f}};
    const float *c0 = sun + 12, *c1 = sun + 16, *c2 = sun + 20;
    for (int i = 0; i < 6; i++)
        for (int r = 0; r < 3; r++) {
#if MM_FMA
            volatile float b = c1[r] * sv[i][1];
            out[3 * i + r] = fmaf(c2[r], sv[i][2], fmaf(c0[r], sv[i][0], b));
#else
            volatile float a = c0[r] * sv[i][0], b = c1[r] * sv[i][1], c = c2[r] * sv[i][2];
            volatile float ab = a + b;
            out[3 * i + r] = ab + c;
#endif
        }
}

// pack_pairs_kernel (cloud_march.cu) on the host: [z][y][x] RGBA8 -> two float4 per texel, {A(x), B(x), A(x+1), B(x+1)} per channel pair
static std::vector<float4> pack_pairs(const uint8_t *src, int w, int h, int d, bool placement_layout) {
    size_t n = (size_t)w * h * d;
    std::vector<float4> dst(2 * n);
    for (size_t i = 0; i < n; i++) {
        int x = (int)(i % w);
        const uint8_t *p = src + 4 * i, *q = src + 4 * (i - x + (size_t)((x + 1) % w));
        if (placement_layout) {
            dst[2 * i] = make_float4((float)p[2], (float)p[0], (float)q[2], (float)q[0]);
            dst[2 * i + 1] = make_float4((float)p[1], (float)p[3], (float)q[1], (float)q[3]);
        } else {
            dst[2 * i] = make_float4((float)p[0], (float)p[1], (float)q[0], (float)q[1]);
            dst[2 * i + 1] = make_float4((float)p[2], (float)p[3], (float)q[2], (float)q[3]);
        }
    }
    return dst;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <bool MH, bool LH, bool P2>
static void run(const MarchParams &P, bool cnt, float *out, uint32_t *counters) {
    for (int y = 0; y < P.H; y++)
        for (int x = 0; x < P.W; x++) {
            Counters cn;
            float4 c = cnt ? march_ray<MH, LH, true, P2>(P, x, y, cn) : march_ray<MH, LH, false, P2>(P, x, y, cn);
            size_t i = (size_t)y * P.W + x;
            out[4 * i] = c.x; out[4 * i + 1] = c.y; out[4 * i + 2] = c.z; out[4 * i + 3] = c.w;
            if (counters && cnt) { counters[4 * i] = cn.trips; counters[4 * i + 1] = cn.n2d; counters[4 * i + 2] = cn.n3d; counters[4 * i + 3] = cn.lit; }
        }
}

}  // namespace
}  // namespace mm

extern "C" {

int hm_arith(void) { return MM_FMA; }

// One MM_FULL frame.  tex[5]: RGBA8 texels per slot (placement, night sky or null, curl, low-res, hi-res), dims[5][3] = w, h, d.
// filter: 0 FILTER_EXACT, 1 FILTER_HW, 2 FILTER_HYBRID (csrc/common.h).  sampler / user: the texture unit of this build (needed by 1 and 2).
int hm_march(const float *camera160, const float *sun116, const float *sky52, const uint8_t *const *tex, const int *dims, int filter, int with_counters,
             int W, int H, void *sampler, void *user, float *out_rgba32f, uint32_t *out_counters) {
    using namespace mm;
    if (!camera160 || !sun116 || !sky52 || !tex || !dims || !out_rgba32f || W <= 0 || H <= 0) return -1;
    if (filter != FILTER_EXACT && !sampler) return -2;
    mm_host::g_sampler = (mm_host::sampler_fn)sampler;
    mm_host::g_user = user;
    static MarchParams P;
    memset(&P, 0, sizeof P);
    memcpy(P.cam, camera160, 160); memcpy(P.sun, sun116, 116); memcpy(P.sky, sky52, 52);
    cone_samples(P.sun, P.light);
    std::vector<float4> pairs[TEX_COUNT];
    bool p2 = true;
    for (int s = 0; s < TEX_COUNT; s++) {
        if (!tex[s]) continue;
        int w = dims[3 * s], h = dims[3 * s + 1], d = dims[3 * s + 2];
        pairs[s] = pack_pairs(tex[s], w, h, d, s == TEX_PLACEMENT);
        P.tex[s].pairs = pairs[s].data();
        P.tex[s].obj = (cudaTextureObject_t)(s + 1);
        P.tex[s].w = w; P.tex[s].h = h; P.tex[s].d = d;
        P.tex[s].wf = (float)w; P.tex[s].hf = (float)h; P.tex[s].df = (float)d;
        P.tex[s].pow2 = is_pow2(w) && is_pow2(h) && is_pow2(d);
        if (s != TEX_NIGHTSKY) p2 = p2 && P.tex[s].pow2;
    }
    P.W = W; P.H = H; P.mode = DISPATCH_FULL;
    const bool cnt = with_counters != 0;
    if (filter == FILTER_HW) run<true, true, true>(P, cnt, out_rgba32f, out_counters);
    else if (filter == FILTER_HYBRID) { if (p2) run<false, true, true>(P, cnt, out_rgba32f, out_counters); else run<false, true, false>(P, cnt, out_rgba32f, out_counters); }
    else { if (p2) run<false, false, true>(P, cnt, out_rgba32f, out_counters); else run<false, false, false>(P, cnt, out_rgba32f, out_counters); }
    return 0;
}

float hm_det_powf(float x, float y) { return mm::det_powf(x, y); }

// K7 (cloud_shadow_kernel): accumDensity and texture() counts for n world positions.  filter: 0 FILTER_EXACT, 1 FILTER_HW.  Not built under MM_FMA (the product
// has the pass once, in the uncontracted build).
int hm_cloud_shadow(const float *camera160, const float *sun116, const float *sky52, const uint8_t *placement, int pw, int ph, const uint8_t *lowres, int ln,
                    int filter, const float *positions_xyz, int n, void *sampler, void *user, float *out_density, uint32_t *out_fetches) {
    using namespace mm;
#if MM_FMA
    return -3;
#else
    if (!camera160 || !sun116 || !sky52 || !placement || !lowres || !positions_xyz || !out_density || n < 0) return -1;
    if (filter == FILTER_HW && !sampler) return -2;
    mm_host::g_sampler = (mm_host::sampler_fn)sampler;
    mm_host::g_user = user;
    ShadowParams P;
    memset(&P, 0, sizeof P);
    memcpy(P.cam, camera160, 160); memcpy(P.sun, sun116, 116); memcpy(P.sky, sky52, 52);
    {   // capi.cu, mm_cloud_shadow: L = normalize((camera.view * vec4(sun.directionBasis[1].xyz, 0)).xyz); if (L.y < -0.05) L *= -1  (model.frag:216-217)
        const float *c = P.cam, *d = P.sun + 16;
        float l[3];
        for (int i = 0; i < 3; i++) {
            volatile float p0 = c[0 + i] * d[0], p1 = c[4 + i] * d[1], p2 = c[8 + i] * d[2], p3 = c[12 + i] * 0.0f;
            volatile float sum = p0 + p1;
            sum = sum + p2;
            sum = sum + p3;
            l[i] = sum;
        }
        volatile float xx = l[0] * l[0], yy = l[1] * l[1], zz = l[2] * l[2];
        volatile float dd = xx + yy;
        dd = dd + zz;
        volatile float inv = 1.0f / sqrtf(dd);
        for (int i = 0; i < 3; i++) { volatile float v = l[i] * inv; P.L[i] = v; }
        if (P.L[1] < -0.05f) for (int i = 0; i < 3; i++) { volatile float v = -1.0f * P.L[i]; P.L[i] = v; }
    }
    std::vector<float4> pp = pack_pairs(placement, pw, ph, 1, true), lp = pack_pairs(lowres, ln, ln, ln, false);
    P.placement = TexDev{pp.data(), (cudaTextureObject_t)(TEX_PLACEMENT + 1), pw, ph, 1, (float)pw, (float)ph, 1.0f, is_pow2(pw) && is_pow2(ph)};
    P.lowres = TexDev{lp.data(), (cudaTextureObject_t)(TEX_LOWRES + 1), ln, ln, ln, (float)ln, (float)ln, (float)ln, is_pow2(ln)};
    P.pos = positions_xyz; P.n = n;
    const bool p2 = P.placement.pow2 && P.lowres.pow2;
    for (int i = 0; i < n; i++) {
        uint32_t nf;
        out_density[i] = filter == FILTER_HW ? shadowPoint<true, true>(P, i, nf) : p2 ? shadowPoint<false, true>(P, i, nf) : shadowPoint<false, false>(P, i, nf);
        if (out_fetches) out_fetches[i] = nf;
    }
    return 0;
#endif
}

}  // extern "C"
