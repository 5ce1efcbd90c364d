/*
 * cloud_march_oracle_fma.c -- CPU ORACLE (test infrastructure, NOT product code): compute-clouds.comp under the CONTRACTED
 * arithmetic definition (OM_ARITH_FMA).
 *
 * A second restatement of /root/reference/SkyEngine/SkyEngine/Shaders/compute-clouds.comp ("CC"), line by line like
 * cloud_march_oracle.c, but with the fused multiply-add contraction that GLSL permits and every GPU performs, pinned down as ONE
 * lexical rule (oracle/glsl_env_fma.h states it in full):
 *     a product that is DIRECTLY an operand of a + or - is not rounded:  a*b + c -> fma(a,b,c),  c - a*b -> fma(-a,b,c),
 *     a*b + c*d -> fma(a,b,RN(c*d))  (left product fuses);  `x += a*b` counts;  any other use of a product rounds it;
 *     built-ins are their formulas under the same rule: dot = fma(az,bz,fma(ax,bx,RN(ay*by))), mix(x,y,a) = fma(x,1-a,RN(y*a)),
 *     mat3*v = fma(c2,v.z,fma(c0,v.x,RN(c1*v.y))), smoothstep = RN(t*t)*fma(-2,t,3), remap = fma(q, newMax-newMin, newMin).
 * PARITY STATUS: PINNED AGAINST THE REFERENCE'S OWN SHADER TEXT -- oracle/_ref/libref_cc_fma.so is that text compiled in an
 * environment whose operators apply the rule mechanically (C++ overload resolution, no expression touched);
 * tests/test_reference_shader.py compares this file with it bit for bit, fetch counts included.  Sampler, deterministic pow
 * (its Horner steps fused: om_det_powf_fma) and quirks Q1..Q14 are those of cloud_march_oracle.c.  Each explicit fmaf below
 * carries the GLSL expression it restates.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"
#include "oracle_internal.h"

#define F(a, b, c) __builtin_fmaf((a), (b), (c))

typedef struct { float x, y, z; } v3;
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 scale3(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }                       /* a ROUNDED product */
static inline v3 mad3s(float s, v3 a, v3 c) { return V3(F(s, a.x, c.x), F(s, a.y, c.y), F(s, a.z, c.z)); }   /* s*a + c */
static inline float dot3(v3 a, v3 b) { return F(a.z, b.z, F(a.x, b.x, a.y * b.y)); }                    /* ((ax*bx)+(ay*by))+(az*bz) */
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 normalize3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return V3(a.x * inv, a.y * inv, a.z * inv); }
static inline float omaxf(float x, float y) { return (x < y) ? y : x; }
static inline float ominf(float x, float y) { return (y < x) ? y : x; }
static inline float clampf(float x, float lo, float hi) { float r = (x > lo) ? x : lo; return (r < hi) ? r : hi; }
static inline float mixf(float x, float y, float a) { return F(x, 1.0f - a, y * a); }                   /* x*(1-a) + y*a */
static inline float smoothstepf(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return (t * t) * F(-2.0f, t, 3.0f);                                                                 /* (t*t)*(3 - 2*t) */
}
/* CC:65-71: newMin + (((value - oldMin) / (oldMax - oldMin)) * (newMax - newMin)) */
static inline float remapf(float value, float oldMin, float oldMax, float newMin, float newMax) {
    return F((value - oldMin) / (oldMax - oldMin), newMax - newMin, newMin);
}
static inline float remapClampedf(float value, float oldMin, float oldMax, float newMin, float newMax) {
    return clampf(remapf(value, oldMin, oldMax, newMin, newMax), newMin, newMax);
}
/* column-major mat3 * vec3: ((c0*v.x) + (c1*v.y)) + (c2*v.z) */
static inline v3 mat3mul(const float m[9], v3 v) {
    return V3(F(m[6], v.z, F(m[0], v.x, m[3] * v.y)), F(m[7], v.z, F(m[1], v.x, m[4] * v.y)), F(m[8], v.z, F(m[2], v.x, m[5] * v.y)));
}

/* the deterministic pow of the decision path (cloud_march_oracle.c: om_det_powf) with every Horner step p*x + c fused: binary64 fma,
 * +, -, *, / only, so that the GPU reproduces it bit for bit.  Same domain and special cases. */
float om_det_powf_fma(float x, float y) {
    if (y == 1.0f) return x;
    if (!(x > 0.0f)) return 0.0f;
    if (x == 1.0f) return 1.0f;
    double dx = (double)x;
    uint64_t bits; memcpy(&bits, &dx, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m; memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e = e + 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 21.0;
    p = __builtin_fma(p, s2, 1.0 / 19.0);
    p = __builtin_fma(p, s2, 1.0 / 17.0);
    p = __builtin_fma(p, s2, 1.0 / 15.0);
    p = __builtin_fma(p, s2, 1.0 / 13.0);
    p = __builtin_fma(p, s2, 1.0 / 11.0);
    p = __builtin_fma(p, s2, 1.0 / 9.0);
    p = __builtin_fma(p, s2, 1.0 / 7.0);
    p = __builtin_fma(p, s2, 1.0 / 5.0);
    p = __builtin_fma(p, s2, 1.0 / 3.0);
    p = __builtin_fma(p, s2, 1.0);
    double l = __builtin_fma(s * p, 2.8853900817779268, (double)e);
    double t = (double)y * l;
    double n = floor(t + 0.5);
    double f = (t - n) * 0.6931471805599453;
    double q = 1.0 / 6227020800.0;
    q = __builtin_fma(q, f, 1.0 / 479001600.0);
    q = __builtin_fma(q, f, 1.0 / 39916800.0);
    q = __builtin_fma(q, f, 1.0 / 3628800.0);
    q = __builtin_fma(q, f, 1.0 / 362880.0);
    q = __builtin_fma(q, f, 1.0 / 40320.0);
    q = __builtin_fma(q, f, 1.0 / 5040.0);
    q = __builtin_fma(q, f, 1.0 / 720.0);
    q = __builtin_fma(q, f, 1.0 / 120.0);
    q = __builtin_fma(q, f, 1.0 / 24.0);
    q = __builtin_fma(q, f, 1.0 / 6.0);
    q = __builtin_fma(q, f, 0.5);
    q = __builtin_fma(q, f, 1.0);
    q = __builtin_fma(q, f, 1.0);
    int ni = (int)n;
    if (ni < -1000) return 0.0f;
    uint64_t sb = (uint64_t)(ni + 1023) << 52;
    double sc; memcpy(&sc, &sb, 8);
    return (float)(q * sc);
}

#define ATMOSPHERE_RADIUS 2000000.0f                       /* CC:56 */
#define ONE_OVER_FOURPI 0.07957747154594767f               /* CC:63 */
#define THREE_OVER_SIXTEENPI 0.05968310365946075f          /* CC:62 */
#define SUN_ANGULAR_COS 0.999956676946448443553574619906976478926848692873900859324f  /* CC:82 */
#define PI_F 3.14159265f                                   /* CC:59 */
#define WIND_STRENGTH 20.0f                                /* CC:279 */
#define MAX_STEPS 100                                      /* CC:286 */

typedef struct {
    const struct om_scene *s;
    v3 cameraPos, earthCenter, windXYZ;
    float timeOffset;
    px_counters *cnt;
} ctx_t;

/* CC:73-77: 1.0 / pow(1.0 - 2.0 * g * cosTheta + g2, 1.5) */
static float hgPhase(float cosTheta, float g) {
    float g2 = g * g;
    float inv = 1.0f / powf(F(-(2.0f * g), cosTheta, 1.0f) + g2, 1.5f);
    return ONE_OVER_FOURPI * ((1.0f - g2) * inv);
}
/* CC:84-86: THREE_OVER_SIXTEENPI * (1.0 + cosTheta * cosTheta) */
static float rayleighPhase(float cosTheta) { return THREE_OVER_SIXTEENPI * F(cosTheta, cosTheta, 1.0f); }

/* CC:88-127 (Q11) */
static v3 getAtmosphereColorPhysical(const struct om_scene *s, v3 dir, v3 sunDir) {
    float sunE = s->sun[28];
    v3 BetaR = V3(s->sky[0], s->sky[1], s->sky[2]);
    v3 BetaM = V3(s->sky[4], s->sky[5], s->sky[6]);
    float zenith = acosf(omaxf(0.0f, dir.y));
    float inverse = 1.0f / F(0.15f, powf(93.885f - ((zenith * 180.0f) / PI_F), -1.253f), cosf(zenith));   /* cos(zenith) + 0.15 * pow(..) */
    float sR = 8.4E3f * inverse;
    float sM = 1.25E3f * inverse;
    /* exp(-BetaR * sR + BetaM * sM) */
    v3 fex = V3(expf(F(-BetaR.x, sR, BetaM.x * sM)), expf(F(-BetaR.y, sR, BetaM.y * sM)), expf(F(-BetaR.z, sR, BetaM.z * sM)));
    float cosTheta = dot3(sunDir, dir);
    float rPhase = rayleighPhase(F(cosTheta, 0.5f, 0.5f));                                               /* cosTheta * 0.5 + 0.5 */
    v3 betaRTheta = scale3(rPhase, BetaR);
    float mPhase = hgPhase(cosTheta, s->sky[12]);
    v3 betaMTheta = scale3(mPhase, BetaM);
    float yDot = 1.0f - sunDir.y;
    yDot *= (((yDot * yDot) * yDot) * yDot);
    v3 sum = add3(BetaR, BetaM), num = add3(betaRTheta, betaMTheta);
    v3 betas = V3(num.x / sum.x, num.y / sum.y, num.z / sum.z);
    v3 sb = scale3(sunE, betas);
    v3 Lin = V3(powf(sb.x * (1.0f - fex.x), 1.5f), powf(sb.y * (1.0f - fex.y), 1.5f), powf(sb.z * (1.0f - fex.z), 1.5f));
    float yc = clampf(yDot, 0.0f, 1.0f);
    Lin = V3(Lin.x * mixf(1.0f, powf(sb.x * fex.x, 0.5f), yc), Lin.y * mixf(1.0f, powf(sb.y * fex.y, 0.5f), yc), Lin.z * mixf(1.0f, powf(sb.z * fex.z, 0.5f), yc));
    v3 L0 = scale3(0.1f, fex);
    float sunDisk = 0.0f;                                                                                /* CC:119-120 */
    v3 big = V3((sunE * 15000.0f) * fex.x, (sunE * 15000.0f) * fex.y, (sunE * 15000.0f) * fex.z);
    L0 = V3(F(big.x, sunDisk, L0.x), F(big.y, sunDisk, L0.y), F(big.z, sunDisk, L0.z));                  /* L0 += (sunE*15000*fex) * sunDisk */
    /* (Lin + L0) * 0.04 + vec3(0.0, 0.0003, 0.00075) */
    return V3(F(Lin.x + L0.x, 0.04f, 0.0f), F(Lin.y + L0.y, 0.04f, 0.0003f), F(Lin.z + L0.z, 0.04f, 0.00075f));
}

/* CC:147-177 (Q1) */
static void raySphereIntersection(v3 ro, v3 rd, v3 c, float w, float *t_out) {
    ro = sub3(ro, c);
    ro = V3(ro.x / w, ro.y / w, ro.z / w);
    float A = dot3(rd, rd);
    float B = 2.0f * dot3(rd, ro);
    float C = dot3(ro, ro) - 0.25f;
    float discriminant = F(B, B, -((4.0f * A) * C));                                                     /* B * B - 4.0 * A * C */
    *t_out = 0.0f;
    if (discriminant < 0.0f) return;
    float t = (((-sqrtf(discriminant)) - B) / A) * 0.5f;
    if (t < 0.0f) t = ((sqrtf(discriminant) - B) / A) * 0.5f;
    if (t >= 0.0f) {
        v3 p = V3(F(rd.x, t, ro.x), F(rd.y, t, ro.y), F(rd.z, t, ro.z));                                 /* ro + rd * t */
        p = scale3(w, p);
        p = add3(p, c);
        *t_out = length3(sub3(p, ro));
    }
}

/* CC:180-188: 0.5 * ATMOSPHERE_RADIUS * normalize(pt - center) + center */
static inline v3 getProjectedShellPoint(v3 pt, v3 center) { return mad3s(0.5f * ATMOSPHERE_RADIUS, normalize3(sub3(pt, center)), center); }
static inline float getRelativeHeight(v3 pt, v3 projectedPt, float thickness) { return clampf(length3(sub3(pt, projectedPt)) / thickness, 0.0f, 1.0f); }

/* CC:193-204 */
static float cloudLayerDensity(float relativeHeight, float cloudType) {
    relativeHeight = clampf(relativeHeight, 0.0f, 1.0f);
    float cumulus = omaxf(0.0f, remapf(relativeHeight, 0.0f, 0.2f, 0.0f, 1.0f) * remapf(relativeHeight, 0.7f, 0.9f, 1.0f, 0.0f));
    float stratocumulus = omaxf(0.0f, remapf(relativeHeight, 0.0f, 0.2f, 0.0f, 1.0f) * remapf(relativeHeight, 0.2f, 0.7f, 1.0f, 0.0f));
    float stratus = omaxf(0.0f, remapf(relativeHeight, 0.0f, 0.1f, 0.0f, 1.0f) * remapf(relativeHeight, 0.2f, 0.3f, 1.0f, 0.0f));
    float d1 = mixf(stratus, stratocumulus, clampf(cloudType * 2.0f, 0.0f, 1.0f));
    float d2 = mixf(stratocumulus, cumulus, clampf((cloudType - 0.5f) * 2.0f, 0.0f, 1.0f));
    return mixf(d1, d2, cloudType);
}
/* CC:206-208 */
static float heightBiasCoverage(const struct om_scene *s, float coverage, float height) {
    float k = clampf(remapf(height, 0.7f, 0.8f, 1.0f, 0.8f), 0.8f, 1.0f);
    return s->pow_mode == OM_POW_LIBM ? powf(coverage, k) : om_det_powf_fma(coverage, k);
}

/* CC:214-228 */
static float cloudHiRes(const ctx_t *cx, v3 pos, float curlStrength, float origDensity, float relativeHeight) {
    const struct om_scene *s = cx->s;
    float c = 0.0001f;
    float cu[4];
    om__sample2d(&s->curl, s->filter, c * pos.x, c * pos.z, cu);
    cx->cnt->n2d++;
    v3 curl = V3(F(2.0f, cu[0], -1.0f), F(2.0f, cu[1], -1.0f), F(2.0f, cu[2], -1.0f));                   /* 2.0 * curl - 1.0 */
    pos = mad3s(1.9f * curlStrength, curl, pos);                                                         /* pos += 1.9 * curlStrength * curl */
    float dn[4];
    om__sample3d(&s->hires, s->filter, 0.0004f * pos.x, 0.0004f * pos.y, 0.0004f * pos.z, dn);
    cx->cnt->n3d++;
    float erosion = F(0.125f, dn[2], F(0.625f, dn[0], 0.25f * dn[1]));                                   /* 0.625*r + 0.25*g + 0.125*b */
    erosion = mixf(erosion, 1.0f - erosion, clampf(relativeHeight * 10.0f, 0.0f, 1.0f));
    return remapClampedf(origDensity, 1.0f * erosion, 1.0f, 0.0f, 1.0f);
}

/* CC:231-253 (Q2) */
static float cloudTest(const ctx_t *cx, v3 pos, float relativeHeight) {
    const struct om_scene *s = cx->s;
    v3 currentProj = getProjectedShellPoint(pos, cx->earthCenter);
    float ci[4];
    om__sample2d(&s->placement, s->filter, 0.000009f * (currentProj.x - cx->cameraPos.x), 0.000009f * (currentProj.z - cx->cameraPos.z), ci);
    cx->cnt->n2d++;
    float layerDensity = cloudLayerDensity(relativeHeight, ci[2]);
    float dn[4];
    om__sample3d(&s->lowres, s->filter, 0.00002f * pos.x, 0.00002f * pos.y, 0.00002f * pos.z, dn);
    cx->cnt->n3d++;
    float density = layerDensity * remapClampedf(dn[0], 0.3f, 1.0f, 0.0f, 1.0f);
    if (density < 0.0001f) return 0.0f;
    float coverage = heightBiasCoverage(s, relativeHeight, ominf(0.85f, ci[0]));
    float erosion = F(0.125f, dn[3], F(0.625f, dn[1], 0.25f * dn[2]));                                   /* 0.625*y + 0.25*z + 0.125*w */
    erosion = remapClampedf(erosion, coverage, 1.0f, 0.0f, 1.0f);
    density = remapClampedf(density, erosion, 1.0f, 0.0f, 1.0f);
    return density;
}

/* CC:256-277 */
static void fromAngleAxis(v3 a, float angleRad, float rot[9]) {
    float cost = cosf(angleRad), sint = sinf(angleRad), omc = 1.f - cost;
    rot[0] = F(a.x * a.x, omc, cost);                 /* cost + angle.x * angle.x * (1.f - cost) */
    rot[1] = F(a.y * a.x, omc, a.z * sint);           /* angle.y * angle.x * (1.f - cost) + angle.z * sint */
    rot[2] = F(a.z * a.x, omc, -(a.y * sint));        /* ... - angle.y * sint */
    rot[3] = F(a.x * a.y, omc, -(a.z * sint));
    rot[4] = F(a.y * a.y, omc, cost);
    rot[5] = F(a.z * a.y, omc, a.x * sint);
    rot[6] = F(a.x * a.z, omc, a.y * sint);
    rot[7] = F(a.y * a.z, omc, -(a.x * sint));
    rot[8] = F(a.z * a.z, omc, cost);
}

/* CC:414 / CC:445: WIND_STRENGTH * (sky.wind.xyz + h * vec3(0.1, 0.05, 0)) * (timeOffset + h * 200.0) */
static inline v3 windOffsetAt(v3 windXYZ, float timeOffset, float h) {
    v3 w = V3(F(h, 0.1f, windXYZ.x), F(h, 0.05f, windXYZ.y), F(h, 0.0f, windXYZ.z));
    return scale3(F(h, 200.0f, timeOffset), scale3(WIND_STRENGTH, w));
}

/* CC:288-500 for one target pixel.  W,H replace the hard-coded 1920x1080 (Q7, CC:283-285). */
void om__march_pixel_fma(const struct om_scene *s, int px, int py, int W, int H, float out[4], px_counters *cnt) {
    ctx_t cx; cx.s = s; cx.cnt = cnt;
    const float *cam = s->cam, *sun = s->sun, *sky = s->sky;
    float timeOffset = sky[11];                                                   /* CC:289 */

    float uvx = (float)px / (float)W, uvy = (float)py / (float)H;                 /* CC:305 */
    float spx = F(uvx, 2.0f, -1.0f), spy = F(uvy, 2.0f, -1.0f);                   /* CC:309: uv * 2.0 - 1.0 */

    v3 camLook = V3(cam[2], cam[6], cam[10]);                                     /* CC:312-314 */
    v3 camRight = V3(cam[0], cam[4], cam[8]);
    v3 camUp = V3(cam[1], cam[5], cam[9]);
    v3 cameraPos = V3(cam[32], cam[33], cam[34]);                                 /* CC:317 */
    float aspect = cam[36], tanH = cam[37];
    v3 refPoint = sub3(cameraPos, camLook);
    /* CC:320: refPoint + params.x * screenPoint.x * params.y * camRight - screenPoint.y * params.y * camUp */
    v3 p = mad3s(-(spy * tanH), camUp, mad3s((aspect * spx) * tanH, camRight, refPoint));
    v3 rayDirection = normalize3(sub3(p, cameraPos));                             /* CC:322 */

    v3 sunDir = normalize3(V3(sun[16], sun[17], sun[18]));                        /* CC:324 */
    float sunDirectionY = sun[5];

    float dotToSun = omaxf(0.0f, dot3(sunDir, rayDirection));                     /* CC:326-340 */
    float skyAmbient = dotToSun * 0.18f;
    skyAmbient *= (skyAmbient * skyAmbient);
    float sunDisk = smoothstepf(SUN_ANGULAR_COS, SUN_ANGULAR_COS + 0.00003f, dotToSun);
    dotToSun *= ((dotToSun * dotToSun) * dotToSun);
    dotToSun *= ((dotToSun * dotToSun) * dotToSun);
    dotToSun *= ((dotToSun * dotToSun) * dotToSun);
    dotToSun *= ((dotToSun * dotToSun) * dotToSun);
    dotToSun *= (dotToSun * dotToSun);
    if (sunDirectionY < 0.0f)
        dotToSun *= (((((dotToSun * dotToSun) * dotToSun) * dotToSun) * dotToSun) * dotToSun);
    sunDisk = omaxf(sunDisk, dotToSun);
    sunDisk = omaxf(0.0f, sunDisk);

    float fr = 0, fg = 0, fb = 0, fa = 0;                                         /* CC:342-348 */
    v3 backgroundCol = V3(0, 0, 0);
    if (sunDirectionY >= 0.0f) {
        backgroundCol = getAtmosphereColorPhysical(s, rayDirection, sunDir);
        fa = omaxf(skyAmbient, sunDisk);
        fr = backgroundCol.x; fg = backgroundCol.y; fb = backgroundCol.z;
    }
    if (rayDirection.y < 0.0f) {                                                  /* CC:351-354: dot(rd, (0,1,0)) = rd.y */
        out[0] = fr; out[1] = fg; out[2] = fb; out[3] = fa;
        return;
    }

    v3 earthCenter = V3(cameraPos.x, (-ATMOSPHERE_RADIUS * 0.5f) * 0.995f, cameraPos.z);   /* CC:357-358 */
    float atmosphereThickness = (0.5f * ATMOSPHERE_RADIUS) * 0.02f;               /* CC:360 */
    float tInner, tOuter;
    raySphereIntersection(cameraPos, rayDirection, earthCenter, ATMOSPHERE_RADIUS, &tInner);          /* CC:362 */
    raySphereIntersection(cameraPos, rayDirection, earthCenter, ATMOSPHERE_RADIUS * 1.02f, &tOuter);  /* CC:363 */
    cx.cameraPos = cameraPos; cx.earthCenter = earthCenter;

    if (sunDirectionY < 0.0f) {                                                   /* CC:365-384 (night) */
        float rot[9];
        fromAngleAxis(normalize3(V3(1.0f, 0.0f, 1.0f)), sunDirectionY * 0.5f, rot);
        v3 rotatedRayDir = mat3mul(rot, rayDirection);
        v3 rotatedRayOrigin = mat3mul(rot, cameraPos);
        v3 point = mad3s(tOuter, rotatedRayDir, rotatedRayOrigin);                /* t * rotatedRayDir + rotatedRayOrigin */
        v3 projectedPoint = getProjectedShellPoint(point, earthCenter);
        float nu = F(0.00002f, projectedPoint.x - cameraPos.x, 0.35f);            /* 0.00002 * (pp.xz - cam.xz) + 0.35 */
        float nv = F(0.00002f, projectedPoint.z - cameraPos.z, 0.35f);
        float ns[4] = {0, 0, 0, 0};
        if (s->nightsky.texels) { om__sample2d(&s->nightsky, s->filter, nu, nv, ns); cnt->n2d++; }
        backgroundCol = V3(ns[0], ns[1], ns[2]);
        backgroundCol = V3(backgroundCol.x * (sqrtf(backgroundCol.x) * 0.75f), backgroundCol.y * (sqrtf(backgroundCol.y) * 0.75f), backgroundCol.z * (sqrtf(backgroundCol.z) * 0.75f));
        backgroundCol = V3(powf(backgroundCol.x, 2.2f), powf(backgroundCol.y, 2.2f), powf(backgroundCol.z, 2.2f));
        backgroundCol = scale3(10.0f, backgroundCol);
        float falloff = powf(rayDirection.y, 6.0f);
        backgroundCol = scale3(falloff, backgroundCol);
        float mt = powf(rayDirection.y, 0.03125f);
        backgroundCol = V3(mixf(0.3f * 0.05f, backgroundCol.x, mt), mixf(0.6f * 0.05f, backgroundCol.y, mt), mixf(4.0f * 0.05f, backgroundCol.z, mt));
        backgroundCol = add3(backgroundCol, V3(sunDisk, sunDisk, sunDisk));
        fa = sunDisk;
    }

    float cosTheta = dot3(rayDirection, sunDir);                                  /* CC:386-390 */
    float accumDensity = 0.0f;
    float transmittance = 1.0f;
    float stepSize = 0.05f * atmosphereThickness;

    float basis[9] = {sun[12], sun[13], sun[14], sun[16], sun[17], sun[18], sun[20], sun[21], sun[22]};  /* mat3(directionBasis) CC:392 */
    static const float sv[6][3] = {{0, 0.6f, 0}, {0, 0.5f, 0.05f}, {0.1f, 0.75f, 0}, {0.2f, 2.5f, 0.3f}, {0, 6, 0}, {-0.1f, 1, -0.2f}};
    v3 samples[6];
    for (int i = 0; i < 6; i++) samples[i] = mat3mul(basis, V3(sv[i][0], sv[i][1], sv[i][2]));        /* CC:393-401 */

    int noHits = 1, misses = 0, steps = 0;                                        /* CC:403-405 */
    v3 windXYZ = V3(sky[8], sky[9], sky[10]);
    float henyeyGreenstein = omaxf(hgPhase(cosTheta, 0.6f), 0.7f * hgPhase(cosTheta, 0.99f - 0.1f));   /* CC:407 */

    for (float t = tInner; t < tOuter; t += stepSize) {                           /* CC:408 */
        cnt->trips++;
        v3 currentPos = mad3s(t, rayDirection, cameraPos);                        /* cameraPos + t * rayDirection */
        v3 currentProj = getProjectedShellPoint(currentPos, earthCenter);
        float rHeight = getRelativeHeight(currentPos, currentProj, atmosphereThickness);
        v3 windOffset = windOffsetAt(windXYZ, timeOffset, rHeight);               /* CC:414 (Q8) */

        float density = cloudTest(&cx, add3(currentPos, windOffset), rHeight);    /* CC:421 */
        float loDensity = density;

        if (density > 0.0f) {                                                     /* CC:426 */
            misses = 0;
            if (noHits) {                                                         /* CC:428-434 (Q3, Q4) */
                t -= stepSize;
                stepSize *= 0.3f;
                noHits = 0;
                continue;
            }
            density = cloudHiRes(&cx, add3(currentPos, windOffset), stepSize, density, rHeight);       /* CC:436 (Q9) */
            if (density < 0.0001f) continue;                                      /* CC:437 (Q3) */
            cnt->lit++;
            float densityAlongLight = 0.0f;
            for (int i = 0; i < 6; i++) {                                         /* CC:441-453 */
                v3 lsPos = mad3s(3.0f * stepSize, samples[i], currentPos);        /* currentPos + 3.0 * stepSize * samples[i] */
                v3 lsProj = getProjectedShellPoint(lsPos, earthCenter);
                float lsHeight = getRelativeHeight(lsPos, lsProj, atmosphereThickness);
                windOffset = windOffsetAt(windXYZ, timeOffset, lsHeight);
                float lsDensity = cloudTest(&cx, add3(lsPos, windOffset), lsHeight);
                if (lsDensity > 0.0f) {
                    lsDensity = cloudHiRes(&cx, add3(lsPos, windOffset), stepSize, lsDensity, lsHeight);
                    densityAlongLight += lsDensity;
                }
            }
            float beersLaw = expf(-densityAlongLight);                            /* CC:456-466 (Q10) */
            float beersModulated = omaxf(beersLaw, 0.7f * expf(-0.25f * densityAlongLight));
            beersLaw = mixf(beersLaw, beersModulated, F(-cosTheta, 0.5f, 0.5f));  /* -cosTheta * 0.5 + 0.5 */
            float inScatter = 0.09f + powf(loDensity, remapClampedf(rHeight, 0.3f, 0.85f, 0.5f, 2.0f));
            inScatter *= powf(remapClampedf(rHeight, 0.07f, 0.34f, 0.1f, 1.0f), 0.8f);
            transmittance = mixf(transmittance, (inScatter * henyeyGreenstein) * beersLaw, (1.0f - accumDensity));
            accumDensity += density;
        } else if (!noHits) {                                                     /* CC:468-474 */
            misses++;
            if (misses >= 10) {
                noHits = 1;
                stepSize /= 0.3f;
            }
        }
        if (accumDensity > 0.99f) {                                               /* CC:476-479 */
            accumDensity = 1.0f;
            break;
        }
        if (++steps > MAX_STEPS) break;                                           /* CC:481 (Q5) */
    }

    accumDensity *= smoothstepf(0.0f, 1.0f, ominf(1.0f, remapf(rayDirection.y, 0.0f, 0.1f, 0.0f, 1.0f)));  /* CC:485 */
    accumDensity = ominf(accumDensity, 0.999f);                                   /* CC:486 */

    v3 sunColor = V3(sun[8], sun[9], sun[10]);
    float sunI = sun[28];
    float e = expf(-transmittance);
    float direct = omaxf(0.0f, transmittance);
    v3 amb;                                                                       /* CC:489-493 */
    if (sunDirectionY >= 0.0f) {
        amb = scale3(e, scale3(0.08f, backgroundCol));                            /* 0.08 * backgroundCol * exp(-T) */
    } else {
        float pw = powf(rayDirection.y, 0.03125f);
        v3 nightAmb = scale3(pw, scale3(0.05f, V3(0.3f, 0.6f, 4.0f)));
        amb = scale3(e, scale3(0.08f, nightAmb));
    }
    /* sun.color.xyz * (sun.intensity * vec3(max(0, T)) + amb) */
    v3 cloudColor = V3(sunColor.x * F(sunI, direct, amb.x), sunColor.y * F(sunI, direct, amb.y), sunColor.z * F(sunI, direct, amb.z));
    out[0] = mixf(backgroundCol.x, cloudColor.x, accumDensity);                   /* CC:495 */
    out[1] = mixf(backgroundCol.y, cloudColor.y, accumDensity);
    out[2] = mixf(backgroundCol.z, cloudColor.z, accumDensity);
    out[3] = fa * omaxf(1.0f - accumDensity, 0.0f);                               /* CC:496 */
}
