// cloud_march_x2.inl -- K1x2: the march with TWO rays per thread on Blackwell's packed binary32 instructions (FADD2 / FFMA2).
// Included by cloud_march.cu inside its anonymous namespace (and therefore built under both arithmetic definitions).
//
// Why.  K1 is bound by instruction issue (86.6 % of the issue slots busy, FMA pipe 53 %, texture pipe 26 %; profiles/r02_C3_hw.*), and
// 57 % of what it issues is FADD / FMUL / FFMA.  sm_100 executes those on register PAIRS in one issue slot (FADD2, FMUL2, FFMA2): with two
// neighbouring rays per thread -- component .x = the left pixel, .y = the right one -- every arithmetic instruction of the decision path
// serves both rays, and the per-thread overhead that does not depend on the ray (constants, loop control, reconvergence) is paid once
// per pair.  What stays per ray: comparisons and selects, MUFU seeds, texture fetches, the binary64 pow, the loop's state machine.
//
// Bits.  Each component of a packed operation is the same IEEE operation as its scalar counterpart, so the frame is bit-identical to K1's
// (tests/test_packed_gpu.py) -- with one trap: ptxas 12.9 contracts a packed multiply feeding a packed add into FFMA2 even under
// -fmad=false (mul.rn.f32x2 + add.rn.f32x2 -> FFMA2; the scalar pair correctly stays FMUL + FADD).  Every stand-alone product is therefore
// issued as FFMA2(a, b, -0) with the -0 read from constant memory, which ptxas cannot fold: a*b + (-0) is RN(a*b) for every a*b, zeros and
// denormals included, and an FFMA2 is not a candidate for further contraction.  Fused steps that the arithmetic definition asks for
// (sampler-free exact sequences: div_const, in-range sqrt / rcp / divide; and every PMADD under MM_FMA) are explicit FFMA2.
//
// MEASURED (profiles/r02m_packed.txt): bit-identical to K1, and SLOWER -- 7.14 ms against 5.75 ms for the 4K frame.  Packing halves the
// arithmetic (3.25 G scalar FADD/FMUL/FFMA warp-instructions per frame become 1.54 G FFMA2/FADD2 + 0.37 G scalar), but (a) the
// texture unit returns each ray's texel in its own register quad and the state machine is per ray, so 0.44 G extra MOVs, 0.16 G
// register-pair clears and ~0.3 G more mask / select / compare instructions come back: 5.30 G against 5.73 G in total; and (b) the pair's
// state needs 96-118 registers, i.e. 16-20 warps per SM instead of 32, and two rays in ONE instruction stream add no instruction-level
// parallelism: 3.7 warps per scheduler keep only 58 % of the issue slots busy (K1: 86.6 %).  Kept as an opt-in scheduler
// (MM_SCHED_PACKED) because it is the evidence for that conclusion; K1 remains the default.
//
// Scope: the production mode (texture-unit sampler for march and light samples, no diagnostic counters), MM_FULL and MM_PHASE16, any row
// partition; a warp covers a 16 x 4 pixel tile, a block 32 x 8.  Everything outside the loop (ray_setup, ray_finish) is K1's scalar code.

typedef float2 f2;
struct v3p { f2 x, y, z; };
__constant__ float2 g_negzero = {-0.0f, -0.0f};

#define PF __device__ __forceinline__
PF f2 F2(float a, float b) { return make_float2(a, b); }
PF f2 S2(float a) { return make_float2(a, a); }
PF f2 pneg(f2 a) { return make_float2(-a.x, -a.y); }
PF f2 padd(f2 a, f2 b) { return __fadd2_rn(a, b); }
PF f2 pfma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }                 // one rounding per component
PF f2 psub(f2 a, f2 b) { return __ffma2_rn(b, S2(-1.0f), a); }               // a - b: b * (-1) is exact
PF f2 pmul(f2 a, f2 b) { return __ffma2_rn(a, b, g_negzero); }               // RN(a * b), protected from contraction (see above)
#if MM_FMA
#define PMADD(a, b, c) pfma((a), (b), (c))                                   // a*b + c, fused by definition
#else
#define PMADD(a, b, c) padd(pmul((a), (b)), (c))                             // a*b + c, two roundings
#endif
PF v3p V3P(f2 x, f2 y, f2 z) { v3p r; r.x = x; r.y = y; r.z = z; return r; }
PF v3p pair3(v3 a, v3 b) { return V3P(F2(a.x, b.x), F2(a.y, b.y), F2(a.z, b.z)); }
PF v3p splat3(v3 a) { return V3P(S2(a.x), S2(a.y), S2(a.z)); }
PF v3p padd3(v3p a, v3p b) { return V3P(padd(a.x, b.x), padd(a.y, b.y), padd(a.z, b.z)); }
PF v3p psub3(v3p a, v3p b) { return V3P(psub(a.x, b.x), psub(a.y, b.y), psub(a.z, b.z)); }
PF v3p pmad3(f2 s, v3p a, v3p c) { return V3P(PMADD(s, a.x, c.x), PMADD(s, a.y, c.y), PMADD(s, a.z, c.z)); }   // s*a + c
PF f2 pdot(v3p a, v3p b) { return PMADD(a.z, b.z, PMADD(a.x, b.x, pmul(a.y, b.y))); }                            // ((ax*bx)+(ay*by))+(az*bz)
PF f2 pclamp(f2 x, float lo, float hi) { return F2(clampg(x.x, lo, hi), clampg(x.y, lo, hi)); }
PF f2 pmax0(f2 x) { return F2(gmax(0.0f, x.x), gmax(0.0f, x.y)); }
// exact sequences of cloud_march.cu, per component the same operations
PF f2 psqrt_inrange(f2 d) {
    f2 y = F2(frsqrt(d.x), frsqrt(d.y));
    f2 s = pmul(d, y), hy = pmul(S2(0.5f), y);
    f2 e = pfma(pneg(s), s, d);
    return pfma(e, hy, s);
}
PF f2 prcp_inrange(f2 x) {
    f2 y = F2(frcp(x.x), frcp(x.y));
    f2 e = pfma(pneg(x), y, S2(1.0f));
    return pfma(y, e, y);
}
PF f2 pdiv_inrange(f2 x, f2 y) {
    f2 r = F2(frcp(y.x), frcp(y.y));
    f2 e = pfma(pneg(y), r, S2(1.0f));
    r = pfma(r, e, r);
    f2 q = pfma(x, r, S2(0.0f));
    f2 rem = pfma(pneg(y), q, x);
    return pfma(r, rem, q);
}
PF f2 pdiv_const(f2 x, float c, float rc) {
    f2 q = pmul(x, S2(rc));
    f2 r = pfma(q, S2(-c), x);
    return pfma(r, S2(rc), q);
}
#define PDIVC(x, c) pdiv_const((x), (c), 1.0f / (c))
PF f2 plength(v3p a) { return psqrt_inrange(pdot(a, a)); }
PF v3p pnormalize(v3p a) { f2 inv = prcp_inrange(psqrt_inrange(pdot(a, a))); return V3P(pmul(a.x, inv), pmul(a.y, inv), pmul(a.z, inv)); }
PF v3p pshellPoint(v3p pt, v3p center) { return pmad3(S2(0.5f * ATMOSPHERE_RADIUS), pnormalize(psub3(pt, center)), center); }          // CC:180-182
PF f2 prelativeHeight(v3p pt, v3p proj) { return pclamp(PDIVC(plength(psub3(pt, proj)), SHELL_THICKNESS), 0.0f, 1.0f); }                // CC:186-188
PF f2 pmix(f2 x, f2 y, f2 a) { return PMADD(x, psub(S2(1.0f), a), pmul(y, a)); }                                                         // x*(1-a) + y*a
// remapClamped(v, m, 1, 0, 1) per component (remapClampedTo1): the divide is evaluated for both rays and the special cases selected
PF f2 premapTo1(f2 v, f2 m) {
    f2 num = psub(v, m), den = psub(S2(1.0f), m);
    f2 q = pdiv_inrange(num, den);
    float r0 = !(num.x > 0.0f) ? 0.0f : (den.x < 5.9604645e-08f ? 1.0f : ((q.x < 1.0f) ? q.x : 1.0f));
    float r1 = !(num.y > 0.0f) ? 0.0f : (den.y < 5.9604645e-08f ? 1.0f : ((q.y < 1.0f) ? q.y : 1.0f));
    return F2(r0, r1);
}
// CC:414 / 445; wz20 = WIND_STRENGTH * (wind.z + 0.0f), uniform
PF v3p pwindOffset(v3 windXYZ, float wz20, float timeOffset, f2 h) {
    f2 wx = PMADD(h, S2(0.1f), S2(windXYZ.x)), wy = PMADD(h, S2(0.05f), S2(windXYZ.y));
    f2 s = PMADD(h, S2(200.0f), S2(timeOffset));
    return V3P(pmul(s, pmul(S2(WIND_STRENGTH), wx)), pmul(s, pmul(S2(WIND_STRENGTH), wy)), pmul(s, S2(wz20)));
}

struct Grad2 { f2 cumulus, stratocumulus, stratus; };
// CC:193-198.  REMAP_C(h, 0, c, 0, 1) = h/c (+0, the identity on a non-negative quotient); REMAP_C(h, a, b, 1, 0) = q*(-1) + 1 = 1 - q
PF Grad2 pLayerGradients(f2 h) {                      // h comes clamped from prelativeHeight (layerGradients)
    f2 up02 = PDIVC(h, 0.2f - 0.0f), up01 = pmul(S2(2.0f), up02);                     // h / 0.1f = 2 * (h / 0.2f), exactly (layerGradients)
    Grad2 g;
    g.cumulus = pmax0(pmul(up02, psub(S2(1.0f), PDIVC(padd(h, S2(-0.7f)), 0.9f - 0.7f))));
    g.stratocumulus = pmax0(pmul(up02, psub(S2(1.0f), pmul(S2(2.0f), padd(h, S2(-0.2f))))));   // (h - 0.2f) / 0.5f = 2 * (h - 0.2f)
    g.stratus = pmax0(pmul(up01, psub(S2(1.0f), PDIVC(padd(h, S2(-0.2f)), 0.3f - 0.2f))));
    return g;
}
PF f2 pBlendLayers(const Grad2 &g, f2 cloudType) {                                     // CC:200-203
    f2 d1 = pmix(g.stratus, g.stratocumulus, pclamp(pmul(cloudType, S2(2.0f)), 0.0f, 1.0f));
    f2 d2 = pmix(g.stratocumulus, g.cumulus, pclamp(pmul(padd(cloudType, S2(-0.5f)), S2(2.0f)), 0.0f, 1.0f));
    return pmix(d1, d2, cloudType);
}

// CC:231-253 for two rays; `on` = which components are wanted (bit 0 = .x, bit 1 = .y).  Components that are not wanted, or that CC would
// return 0 for, yield exactly +0.  Mirrors cloudTest<true, false, *> operation for operation.
PF f2 pCloudTest(const MarchParams &P, v3p pos, f2 h, v3p earthCenter, v3 cameraPos, unsigned on) {
    Grad2 lg = pLayerGradients(h);
    if (lg.cumulus.x == 0.0f && lg.stratocumulus.x == 0.0f && lg.stratus.x == 0.0f) on &= ~1u;
    if (lg.cumulus.y == 0.0f && lg.stratocumulus.y == 0.0f && lg.stratus.y == 0.0f) on &= ~2u;
    if (!on) return S2(0.0f);
    f2 u = pmul(S2(0.00002f), pos.x), v = pmul(S2(0.00002f), pos.y), w = pmul(S2(0.00002f), pos.z);
    float4 dn0 = make_float4(0.f, 0.f, 0.f, 0.f), dn1 = dn0, ci0 = dn0, ci1 = dn0;
    if (on & 1u) dn0 = tex3D<float4>(P.tex[TEX_LOWRES].obj, u.x, v.x, w.x);
    if (on & 2u) dn1 = tex3D<float4>(P.tex[TEX_LOWRES].obj, u.y, v.y, w.y);
    v3p proj = pshellPoint(pos, earthCenter);
    f2 pu = pmul(S2(0.000009f), psub(proj.x, S2(cameraPos.x))), pv = pmul(S2(0.000009f), psub(proj.z, S2(cameraPos.z)));
    if (on & 1u) ci0 = tex2D<float4>(P.tex[TEX_PLACEMENT].obj, pu.x, pv.x);
    if (on & 2u) ci1 = tex2D<float4>(P.tex[TEX_PLACEMENT].obj, pu.y, pv.y);
    f2 layerDensity = pBlendLayers(lg, F2(ci0.z, ci1.z));                              // .b = cloud type
    if (layerDensity.x == 0.0f) on &= ~1u;                                             // 0 * remapClamped(finite) = 0 < 0.0001
    if (layerDensity.y == 0.0f) on &= ~2u;
    f2 density = pmul(layerDensity, pclamp(PDIVC(padd(F2(dn0.x, dn1.x), S2(-0.3f)), 1.0f - 0.3f), 0.0f, 1.0f));
    if (density.x < 0.0001f) on &= ~1u;
    if (density.y < 0.0001f) on &= ~2u;
    if (!on) return S2(0.0f);
    f2 cov = F2(gmin(0.85f, ci0.x), gmin(0.85f, ci1.x));                               // .r = coverage
    f2 k = pclamp(PMADD(PDIVC(padd(cov, S2(-0.7f)), 0.8f - 0.7f), S2(0.8f - 1.0f), S2(1.0f)), 0.8f, 1.0f);
    f2 coverage = h;
    if ((on & 1u) && k.x != 1.0f) coverage.x = det_powf(h.x, k.x);
    if ((on & 2u) && k.y != 1.0f) coverage.y = det_powf(h.y, k.y);
    f2 erosion = PMADD(S2(0.125f), F2(dn0.w, dn1.w), PMADD(S2(0.625f), F2(dn0.y, dn1.y), pmul(S2(0.25f), F2(dn0.z, dn1.z))));
    {   // exact early-out of cloudTest: density * (1 - coverage) - (erosion - coverage) <= 0  =>  +0
        f2 num = psub(erosion, coverage), den = psub(S2(1.0f), coverage);
        f2 t = pfma(density, den, pneg(num));
        if (num.x > 0.0f && !(t.x > 0.0f)) on &= ~1u;
        if (num.y > 0.0f && !(t.y > 0.0f)) on &= ~2u;
        if (!on) return S2(0.0f);
    }
    erosion = premapTo1(erosion, coverage);
    f2 res = premapTo1(density, erosion);
    return F2((on & 1u) ? res.x : 0.0f, (on & 2u) ? res.y : 0.0f);
}

// CC:214-228 for the components in `on`; others yield +0.  Mirrors cloudHiRes<true, false, *>.
PF f2 pCloudHiRes(const MarchParams &P, v3p pos, f2 curlStrength, f2 origDensity, f2 h, unsigned on) {
    f2 cu_u = pmul(S2(0.0001f), pos.x), cu_v = pmul(S2(0.0001f), pos.z);
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, d0 = c0, d1 = c0;
    if (on & 1u) c0 = tex2D<float4>(P.tex[TEX_CURL].obj, cu_u.x, cu_v.x);
    if (on & 2u) c1 = tex2D<float4>(P.tex[TEX_CURL].obj, cu_u.y, cu_v.y);
    v3p curl = V3P(PMADD(S2(2.0f), F2(c0.x, c1.x), S2(-1.0f)), PMADD(S2(2.0f), F2(c0.y, c1.y), S2(-1.0f)), PMADD(S2(2.0f), F2(c0.z, c1.z), S2(-1.0f)));
    pos = pmad3(pmul(S2(1.9f), curlStrength), curl, pos);
    f2 u = pmul(S2(0.0004f), pos.x), v = pmul(S2(0.0004f), pos.y), w = pmul(S2(0.0004f), pos.z);
    if (on & 1u) d0 = tex3D<float4>(P.tex[TEX_HIRES].obj, u.x, v.x, w.x);
    if (on & 2u) d1 = tex3D<float4>(P.tex[TEX_HIRES].obj, u.y, v.y, w.y);
    f2 erosion = PMADD(S2(0.125f), F2(d0.z, d1.z), PMADD(S2(0.625f), F2(d0.x, d1.x), pmul(S2(0.25f), F2(d0.y, d1.y))));
    erosion = pmix(erosion, psub(S2(1.0f), erosion), pclamp(pmul(h, S2(10.0f)), 0.0f, 1.0f));
    f2 res = premapTo1(origDensity, erosion);                                          // `1.0 * erosion` is erosion
    return F2((on & 1u) ? res.x : 0.0f, (on & 2u) ? res.y : 0.0f);
}

// lightSampleFast (the relaxed light-cone sample of the texture-unit modes) for two (lit ray, sample) pairs; same operations per component
PF f2 psat(f2 x) { return F2(__saturatef(x.x), __saturatef(x.y)); }
PF f2 premapSatFast(f2 v, f2 oMin) { f2 d = psub(S2(1.0f), oMin); return psat(pmul(psub(v, oMin), F2(frcp(d.x), frcp(d.y)))); }
PF f2 pLightSampleFast(const MarchParams &P, v3p lsPos, f2 stepSize, v3p earthCenter, v3 windXYZ, float wz20, float timeOffset, unsigned on) {
    v3p d = psub3(lsPos, earthCenter);
    f2 dd = pdot(d, d);
    v3p proj = pmad3(pmul(S2(0.5f * ATMOSPHERE_RADIUS), F2(frsqrt(dd.x), frsqrt(dd.y))), d, earthCenter);
    v3p e = psub3(lsPos, proj);
    f2 e2 = pdot(e, e);
    f2 h = psat(pmul(pmul(e2, F2(frsqrt(fmaxf(e2.x, 1e-30f)), frsqrt(fmaxf(e2.y, 1e-30f)))), S2(1.0f / SHELL_THICKNESS)));
    v3p pos = padd3(lsPos, pwindOffset(windXYZ, wz20, timeOffset, h));
    f2 up02 = pmul(h, S2(5.0f)), up01 = pmul(h, S2(10.0f));
    f2 cumulus, stratocumulus, stratus;
    {
        f2 a = pmul(up02, PMADD(padd(h, S2(-0.7f)), S2(-5.0f), S2(1.0f)));
        f2 b = pmul(up02, PMADD(padd(h, S2(-0.2f)), S2(-2.0f), S2(1.0f)));
        f2 c = pmul(up01, PMADD(padd(h, S2(-0.2f)), S2(-10.0f), S2(1.0f)));
        cumulus = F2(fmaxf(0.0f, a.x), fmaxf(0.0f, a.y));
        stratocumulus = F2(fmaxf(0.0f, b.x), fmaxf(0.0f, b.y));
        stratus = F2(fmaxf(0.0f, c.x), fmaxf(0.0f, c.y));
    }
    if (cumulus.x == 0.0f && stratocumulus.x == 0.0f && stratus.x == 0.0f) on &= ~1u;
    if (cumulus.y == 0.0f && stratocumulus.y == 0.0f && stratus.y == 0.0f) on &= ~2u;
    if (!on) return S2(0.0f);
    f2 u = pmul(S2(0.00002f), pos.x), v = pmul(S2(0.00002f), pos.y), w = pmul(S2(0.00002f), pos.z);
    float4 dn0 = make_float4(0.f, 0.f, 0.f, 0.f), dn1 = dn0, ci0 = dn0, ci1 = dn0;
    if (on & 1u) dn0 = tex3D<float4>(P.tex[TEX_LOWRES].obj, u.x, v.x, w.x);
    if (on & 2u) dn1 = tex3D<float4>(P.tex[TEX_LOWRES].obj, u.y, v.y, w.y);
    v3p d2 = psub3(pos, earthCenter);
    f2 dd2 = pdot(d2, d2);
    f2 inv2 = pmul(S2(0.5f * ATMOSPHERE_RADIUS), F2(frsqrt(dd2.x), frsqrt(dd2.y)));
    f2 pu = pmul(S2(0.000009f), pmul(d2.x, inv2)), pv = pmul(S2(0.000009f), pmul(d2.z, inv2));
    if (on & 1u) ci0 = tex2D<float4>(P.tex[TEX_PLACEMENT].obj, pu.x, pv.x);
    if (on & 2u) ci1 = tex2D<float4>(P.tex[TEX_PLACEMENT].obj, pu.y, pv.y);
    f2 t = F2(ci0.z, ci1.z);
    f2 d1 = pmix(stratus, stratocumulus, psat(pmul(t, S2(2.0f))));
    f2 dmix = pmix(stratocumulus, cumulus, psat(pmul(padd(t, S2(-0.5f)), S2(2.0f))));
    f2 layerDensity = pmix(d1, dmix, t);
    f2 density = pmul(layerDensity, psat(pmul(padd(F2(dn0.x, dn1.x), S2(-0.3f)), S2(1.0f / 0.7f))));
    if (density.x < 0.0001f) on &= ~1u;
    if (density.y < 0.0001f) on &= ~2u;
    if (!on) return S2(0.0f);
    f2 kk = PMADD(padd(F2(fminf(0.85f, ci0.x), fminf(0.85f, ci1.x)), S2(-0.7f)), S2(-2.0f), S2(1.0f));
    f2 k = F2(fminf(fmaxf(kk.x, 0.8f), 1.0f), fminf(fmaxf(kk.y, 0.8f), 1.0f));
    f2 coverage = F2(__powf(h.x, k.x), __powf(h.y, k.y));
    f2 erosion = PMADD(S2(0.125f), F2(dn0.w, dn1.w), PMADD(S2(0.625f), F2(dn0.y, dn1.y), pmul(S2(0.25f), F2(dn0.z, dn1.z))));
    erosion = premapSatFast(erosion, coverage);
    density = premapSatFast(density, erosion);
    if (!(density.x > 0.0f)) on &= ~1u;
    if (!(density.y > 0.0f)) on &= ~2u;
    if (!on) return S2(0.0f);
    f2 cu_u = pmul(S2(0.0001f), pos.x), cu_v = pmul(S2(0.0001f), pos.z);
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, h0 = c0, h1 = c0;
    if (on & 1u) c0 = tex2D<float4>(P.tex[TEX_CURL].obj, cu_u.x, cu_v.x);
    if (on & 2u) c1 = tex2D<float4>(P.tex[TEX_CURL].obj, cu_u.y, cu_v.y);
    f2 cs = pmul(S2(1.9f), stepSize);
    v3p hp = pmad3(cs, V3P(PMADD(S2(2.0f), F2(c0.x, c1.x), S2(-1.0f)), PMADD(S2(2.0f), F2(c0.y, c1.y), S2(-1.0f)), PMADD(S2(2.0f), F2(c0.z, c1.z), S2(-1.0f))), pos);
    f2 hu = pmul(S2(0.0004f), hp.x), hv = pmul(S2(0.0004f), hp.y), hw = pmul(S2(0.0004f), hp.z);
    if (on & 1u) h0 = tex3D<float4>(P.tex[TEX_HIRES].obj, hu.x, hv.x, hw.x);
    if (on & 2u) h1 = tex3D<float4>(P.tex[TEX_HIRES].obj, hu.y, hv.y, hw.y);
    f2 er = PMADD(S2(0.125f), F2(h0.z, h1.z), PMADD(S2(0.625f), F2(h0.x, h1.x), pmul(S2(0.25f), F2(h0.y, h1.y))));
    er = pmix(er, psub(S2(1.0f), er), psat(pmul(h, S2(10.0f))));
    f2 res = premapSatFast(density, er);
    return F2((on & 1u) ? res.x : 0.0f, (on & 2u) ? res.y : 0.0f);
}

// the loop's state machine for one ray of the pair (CC:426-437, 468-482), after its density evaluation; returns what warp_trip's locals hold
struct TripState { bool skipTail, wantHiRes; };
PF TripState trip_decide(float density, float &t, float &stepSize, int &misses, bool &noHits) {
    TripState s; s.skipTail = false; s.wantHiRes = false;
    if (density > 0.0f) {                                                              // CC:426
        misses = 0;
        if (noHits) { t -= stepSize; stepSize *= 0.3f; noHits = false; s.skipTail = true; }   // CC:428-434
        else s.wantHiRes = true;                                                       // CC:436
    } else if (!noHits) {                                                              // CC:468-474
        misses++;
        if (misses >= 10) { noHits = true; stepSize = DIVC(stepSize, 0.3f); }
    }
    return s;
}
PF void trip_advance(bool skipTail, float &accum, int &steps, float &t, float stepSize, float tOuter, bool &alive) {
    if (!skipTail) {
        if (accum > 0.99f) { accum = 1.0f; alive = false; }                            // CC:476-479
        else if (++steps > MAX_STEPS) alive = false;                                   // CC:481
    }
    if (alive) { t += stepSize; alive = t < tOuter; }                                  // CC:408
}

// A warp covers 16 x 4 pixels: lane l owns pixels (2*(l % 8), l / 8) and (2*(l % 8) + 1, l / 8) of its tile; a block is 2 x 2 warps.
enum { X2_TILE_W = 16, X2_TILE_H = 4, X2_BLOCK_W = 32, X2_BLOCK_H = 8 };
template <int MINBLOCKS>
__global__ void __launch_bounds__(128, MINBLOCKS) cloud_march_x2_kernel(const __grid_constant__ MarchParams P) {
    __shared__ float4 s_item[4][64];                     // lit rays of the warp: (pos.xyz, stepSize)
    __shared__ float s_res[4][384];                      // contribution of (item, sample)
    __shared__ float s_light[18];
    const unsigned FULL = 0xffffffffu;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 18) s_light[threadIdx.x] = P.light[threadIdx.x];
    __syncthreads();

    int gx = blockIdx.x * X2_BLOCK_W + (warp & 1) * X2_TILE_W + 2 * (lane & 7);
    int j = (int)P.block_row_order[blockIdx.y] * X2_BLOCK_H + (warp >> 1) * X2_TILE_H + (lane >> 3);
    int pxa = 0, pya = 0, pxb = 0, pyb = 0;
    bool va = dispatch_pixel(P, gx, j, pxa, pya), vb = dispatch_pixel(P, gx + 1, j, pxb, pyb);
    Counters cn = {0u, 0u, 0u, 0u};
    Ray ra, rb;
    ra.alive = rb.alive = false;
    ra.rd = rb.rd = V3(0.f, 1.f, 0.f); ra.t = rb.t = ra.tOuter = rb.tOuter = 0.0f;
    ra.accum = rb.accum = 0.0f; ra.transmittance = rb.transmittance = 1.0f; ra.stepSize = rb.stepSize = 0.05f * SHELL_THICKNESS;
    ra.cosTheta = rb.cosTheta = ra.hg = rb.hg = 0.0f;
    if (va) ray_setup<true, false>(P, pxa, pya, ra, cn);
    if (vb) ray_setup<true, false>(P, pxb, pyb, rb, cn);

    const float timeOffset = P.sky[11];                                                // CC:289
    const v3 windXYZ = V3(P.sky[8], P.sky[9], P.sky[10]);
    const float wz20 = WIND_STRENGTH * (windXYZ.z + 0.0f);
    const v3 cameraPos = V3(P.cam[32], P.cam[33], P.cam[34]);
    const v3 earthCenterS = V3(cameraPos.x, (-ATMOSPHERE_RADIUS * 0.5f) * 0.995f, cameraPos.z);   // CC:357-358
    const v3p earthCenter = splat3(earthCenterS), cam2 = splat3(cameraPos);

    // loop state of the pair
    v3p rd = pair3(ra.rd, rb.rd);
    float ta = ra.t, tb = rb.t, stepA = ra.stepSize, stepB = rb.stepSize, accA = 0.0f, accB = 0.0f, trA = 1.0f, trB = 1.0f;
    int missA = 0, missB = 0, stepsA = 0, stepsB = 0;
    bool noHitsA = true, noHitsB = true, aliveA = ra.alive, aliveB = rb.alive;

    while (__any_sync(FULL, aliveA || aliveB)) {                                       // CC:408
        unsigned on = (aliveA ? 1u : 0u) | (aliveB ? 2u : 0u);
        bool litA = false, litB = false, skipA = false, skipB = false;
        f2 density = S2(0.0f), lo = S2(0.0f), h = S2(0.0f), step2 = F2(stepA, stepB);
        v3p pos = V3P(S2(0.f), S2(0.f), S2(0.f));
        if (on) {
            pos = pmad3(F2(ta, tb), rd, cam2);
            v3p proj = pshellPoint(pos, earthCenter);
            h = prelativeHeight(pos, proj);
            v3p pw = padd3(pos, pwindOffset(windXYZ, wz20, timeOffset, h));
            density = pCloudTest(P, pw, h, earthCenter, cameraPos, on);               // CC:421
            lo = density;
            unsigned hi = 0u;
            if (aliveA) { TripState s = trip_decide(density.x, ta, stepA, missA, noHitsA); skipA = s.skipTail; if (s.wantHiRes) hi |= 1u; }
            if (aliveB) { TripState s = trip_decide(density.y, tb, stepB, missB, noHitsB); skipB = s.skipTail; if (s.wantHiRes) hi |= 2u; }
            if (hi) {                                                                  // CC:436-437 (stepSize is unchanged on this path)
                f2 d = pCloudHiRes(P, pw, step2, density, h, hi);
                if (hi & 1u) { density.x = d.x; if (d.x < 0.0001f) skipA = true; else litA = true; }
                if (hi & 2u) { density.y = d.y; if (d.y < 0.0001f) skipB = true; else litB = true; }
            }
        }

        unsigned maskA = __ballot_sync(FULL, litA), maskB = __ballot_sync(FULL, litB);
        if (maskA | maskB) {                                                           // CC:438-466: light-cone samples shared by the warp
            int nA = __popc(maskA), nItems = nA + __popc(maskB);
            int itemA = __popc(maskA & ((1u << lane) - 1u)), itemB = nA + __popc(maskB & ((1u << lane) - 1u));
            if (litA) s_item[warp][itemA] = make_float4(pos.x.x, pos.y.x, pos.z.x, stepA);
            if (litB) s_item[warp][itemB] = make_float4(pos.x.y, pos.y.y, pos.z.y, stepB);
            __syncwarp();
            for (int base = 0; base < 6 * nItems; base += 64) {                        // CC:441-453, two (item, sample) pairs per thread
                int q0 = base + lane, q1 = base + 32 + lane;
                unsigned won = (q0 < 6 * nItems ? 1u : 0u) | (q1 < 6 * nItems ? 2u : 0u);
                if (won) {
                    int i0 = (won & 1u) ? q0 / 6 : 0, i1 = (won & 2u) ? q1 / 6 : 0;
                    int m0 = (won & 1u) ? q0 - 6 * i0 : 0, m1 = (won & 2u) ? q1 - 6 * i1 : 0;
                    float4 it0 = s_item[warp][i0], it1 = s_item[warp][i1];
                    v3p smp = V3P(F2(s_light[3 * m0], s_light[3 * m1]), F2(s_light[3 * m0 + 1], s_light[3 * m1 + 1]), F2(s_light[3 * m0 + 2], s_light[3 * m1 + 2]));
                    f2 sw = F2(it0.w, it1.w);
                    v3p lsPos = pmad3(pmul(S2(3.0f), sw), smp, V3P(F2(it0.x, it1.x), F2(it0.y, it1.y), F2(it0.z, it1.z)));
                    f2 c = pLightSampleFast(P, lsPos, sw, earthCenter, windXYZ, wz20, timeOffset, won);
                    if (won & 1u) s_res[warp][q0] = c.x;
                    if (won & 2u) s_res[warp][q1] = c.y;
                }
            }
            __syncwarp();
            if (litA) {
                float dal = 0.0f;
#pragma unroll
                for (int i = 0; i < 6; i++) dal += s_res[warp][6 * itemA + i];
                trA = mixg(trA, litTerm(dal, lo.x, h.x, ra.cosTheta, ra.hg), (1.0f - accA));   // CC:464
                accA += density.x;
            }
            if (litB) {
                float dal = 0.0f;
#pragma unroll
                for (int i = 0; i < 6; i++) dal += s_res[warp][6 * itemB + i];
                trB = mixg(trB, litTerm(dal, lo.y, h.y, rb.cosTheta, rb.hg), (1.0f - accB));
                accB += density.y;
            }
            __syncwarp();
        }
        if (aliveA) trip_advance(skipA, accA, stepsA, ta, stepA, ra.tOuter, aliveA);
        if (aliveB) trip_advance(skipB, accB, stepsB, tb, stepB, rb.tOuter, aliveB);
    }

    ra.accum = accA; ra.transmittance = trA; rb.accum = accB; rb.transmittance = trB;
    if (va) store_pixel<false>(P, pxa, pya, ray_finish(P, ra), cn);
    if (vb) store_pixel<false>(P, pxb, pyb, ray_finish(P, rb), cn);
}
#undef PF
