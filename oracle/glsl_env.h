// glsl_env.h -- a GLSL 4.50 execution environment in C++ for the reference's OWN shader text (TEST INFRASTRUCTURE).
//
// oracle/glsl_to_cpp.py rewrites /root/reference/SkyEngine/SkyEngine/Shaders/compute-clouds.comp mechanically (float
// suffixes on literals, `.xyz` -> `.xyz()`, in/inout qualifiers -> C++ parameters, layout(...) uniform blocks -> structs)
// into oracle/_ref/compute_clouds_gen.inc -- never into the repository -- and ref_cc_shim.cpp compiles that text inside this
// environment into oracle/_ref/libref_cc.so.  The shader's control flow, constants, argument orders and operator order are
// then the reference's, executed on the CPU; what this header supplies is the language: vector types and the built-ins, defined
// exactly as the arithmetic contract of cloud_march_oracle.c defines them (binary32, one rounding per operator,
// dot = ((ax*bx)+(ay*by))+(az*bz), normalize(v) = v*(1/sqrt(dot)), mix = x*(1-a)+y*a, GLSL-spec min/max/clamp/smoothstep,
// libm transcendentals) and texture() routed to the oracle's sampler.  tests/test_reference_shader.py compares the two.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef unsigned int uint;
struct vec2; struct vec3; struct vec4; struct uvec2; struct ivec2;

struct vec2 {
    union { float x, r; }; union { float y, g; };
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit inline vec2(const uvec2 &u);
    explicit inline vec2(const ivec2 &u);
    vec2 &operator+=(const vec2 &o) { x = x + o.x; y = y + o.y; return *this; }
    vec2 &operator-=(const vec2 &o) { x = x - o.x; y = y - o.y; return *this; }
    vec2 &operator*=(float s) { x = x * s; y = y * s; return *this; }
    vec2 &operator/=(float s) { x = x / s; y = y / s; return *this; }
};
// float -> int as GLSL's ivec2(vec2): truncation; out-of-range saturates and NaN gives 0 (the contract of reproject_oracle.c)
inline int to_int(float f) { if (!(f == f)) return 0; if (f >= 2147483648.0f) return 2147483647; if (f <= -2147483648.0f) return (-2147483647 - 1); return (int)f; }
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(uint a, uint b) : x((int)a), y((int)b) {}
    explicit ivec2(const uvec2 &u) : x((int)u.x), y((int)u.y) {}
    explicit inline ivec2(const vec2 &v);
};
struct uvec3 { uint x, y, z; uvec2 xy() const { return uvec2(x, y); } };

// assignable swizzles: v.xyz() = ..., v.rgb() = ...
struct ref3 {
    float &a, &b, &c;
    ref3(float &a_, float &b_, float &c_) : a(a_), b(b_), c(c_) {}
    inline ref3 &operator=(const vec3 &v);
    inline operator vec3() const;
};
struct ref2 {
    float &a, &b;
    ref2(float &a_, float &b_) : a(a_), b(b_) {}
    inline operator vec2() const;
};
inline ivec2::ivec2(const vec2 &v) : x(to_int(v.x)), y(to_int(v.y)) {}
inline vec2::vec2(const uvec2 &u) : x((float)u.x), y((float)u.y) {}
inline vec2::vec2(const ivec2 &u) : x((float)u.x), y((float)u.y) {}

struct vec3 {
    union { float x, r; }; union { float y, g; }; union { float z, b; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float c, float d) : x(a), y(c), z(d) {}
    vec3(const vec3 &o) : x(o.x), y(o.y), z(o.z) {}
    vec3 &operator=(const vec3 &o) { x = o.x; y = o.y; z = o.z; return *this; }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec3 &operator+=(const vec3 &o) { x = x + o.x; y = y + o.y; z = z + o.z; return *this; }
    vec3 &operator-=(const vec3 &o) { x = x - o.x; y = y - o.y; z = z - o.z; return *this; }
    vec3 &operator*=(const vec3 &o) { x = x * o.x; y = y * o.y; z = z * o.z; return *this; }
    vec3 &operator+=(float s) { x = x + s; y = y + s; z = z + s; return *this; }
    vec3 &operator*=(float s) { x = x * s; y = y * s; z = z * s; return *this; }
    vec3 &operator/=(float s) { x = x / s; y = y / s; z = z / s; return *this; }
    ref2 xz() { return ref2(x, z); }
    ref2 xy() { return ref2(x, y); }
    inline vec2 xz() const;
    inline vec2 xy() const;
    ref3 xyz() { return ref3(x, y, z); }
    vec3 xyz() const { return *this; }
    ref3 rgb() { return ref3(x, y, z); }
    vec3 rgb() const { return *this; }
};
inline vec2 vec3::xz() const { return vec2(x, z); }
inline vec2 vec3::xy() const { return vec2(x, y); }

struct vec4 {
    union { float x, r; }; union { float y, g; }; union { float z, b; }; union { float w, a; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(const vec3 &v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(const vec4 &o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4 &operator=(const vec4 &o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    vec4 &operator+=(const vec4 &o) { x = x + o.x; y = y + o.y; z = z + o.z; w = w + o.w; return *this; }
    vec4 &operator/=(float s) { x = x / s; y = y / s; z = z / s; w = w / s; return *this; }
    ref3 xyz() { return ref3(x, y, z); }
    vec3 xyz() const { return vec3(x, y, z); }
    ref3 rgb() { return ref3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
    ref2 xz() { return ref2(x, z); }
    vec2 xz() const { return vec2(x, z); }
    ref2 xy() { return ref2(x, y); }
    vec2 xy() const { return vec2(x, y); }
};
inline ref3 &ref3::operator=(const vec3 &v) { a = v.x; b = v.y; c = v.z; return *this; }
inline ref3 &operator*=(ref3 &&r, const vec3 &v) { r.a = r.a * v.x; r.b = r.b * v.y; r.c = r.c * v.z; return r; }
inline ref3::operator vec3() const { return vec3(a, b, c); }
inline ref2::operator vec2() const { return vec2(a, b); }

// ---- operators (componentwise, one rounding each)
inline vec2 operator+(const vec2 &a, const vec2 &b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2 &a, const vec2 &b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator/(const vec2 &a, const vec2 &b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(const vec2 &a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2 &a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator+(const vec2 &a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(const vec2 &a, float s) { return vec2(a.x - s, a.y - s); }
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(const vec3 &a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(float s, const vec3 &a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator-(const vec3 &a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator/(const vec4 &a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec2 operator*(const vec2 &a, const vec2 &b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(const vec2 &a, const ivec2 &b) { return vec2(a.x / (float)b.x, a.y / (float)b.y); }
inline vec2 operator*(const vec2 &a, const ivec2 &b) { return vec2(a.x * (float)b.x, a.y * (float)b.y); }
inline vec2 operator/(const vec2 &a, float s) { return vec2(a.x / s, a.y / s); }
inline vec3 operator/(float s, const vec3 &a) { return vec3(s / a.x, s / a.y, s / a.z); }

struct mat4 {
    vec4 c[4];
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(const mat4 &m) { for (int i = 0; i < 3; i++) c[i] = vec3(m.c[i].x, m.c[i].y, m.c[i].z); }   // upper-left 3x3
    vec3 &operator[](int i) { return c[i]; }
    const vec3 &operator[](int i) const { return c[i]; }
};
// column-major products, accumulated left to right: ((c0*v.x) + (c1*v.y)) + (c2*v.z) [+ (c3*v.w)]
inline vec4 operator*(const vec4 &a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator+(const vec4 &a, const vec4 &b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const mat4 &m, const vec4 &v) { return (((m.c[0] * v.x) + (m.c[1] * v.y)) + (m.c[2] * v.z)) + (m.c[3] * v.w); }
inline mat4 operator*(const mat4 &a, const mat4 &b) { mat4 r; for (int j = 0; j < 4; j++) r.c[j] = a * b.c[j]; return r; }
inline vec3 operator*(const mat3 &m, const vec3 &v) { return ((m.c[0] * v.x) + (m.c[1] * v.y)) + (m.c[2] * v.z); }

// ---- built-ins (the contract of cloud_march_oracle.c)
inline float dot(const vec3 &a, const vec3 &b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline vec3 sqrt(const vec3 &v) { return vec3(::sqrtf(v.x), ::sqrtf(v.y), ::sqrtf(v.z)); }
inline float length(const vec3 &a) { return ::sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3 &a) { float inv = 1.0f / ::sqrtf(dot(a, a)); return vec3(a.x * inv, a.y * inv, a.z * inv); }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float clamp(float x, float lo, float hi) { float r = (x > lo) ? x : lo; return (r < hi) ? r : hi; }
inline float mix(float x, float y, float a) { return (x * (1.0f - a)) + (y * a); }
inline vec3 mix(const vec3 &x, const vec3 &y, float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
inline float smoothstep(float e0, float e1, float x) { float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return (t * t) * (3.0f - (2.0f * t)); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline vec3 pow(const vec3 &x, const vec3 &y) { return vec3(::powf(x.x, y.x), ::powf(x.y, y.y), ::powf(x.z, y.z)); }
inline float exp(float x) { return ::expf(x); }
inline vec3 exp(const vec3 &x) { return vec3(::expf(x.x), ::expf(x.y), ::expf(x.z)); }
inline float acos(float x) { return ::acosf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline float dot(const vec2 &a, const vec2 &b) { return (a.x * b.x) + (a.y * b.y); }
inline float length(const vec2 &a) { return ::sqrtf(dot(a, a)); }
inline vec3 max(const vec3 &a, const vec3 &b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 mix(const vec3 &x, const vec3 &y, const vec3 &a) { return vec3(mix(x.x, y.x, a.x), mix(x.y, y.y, a.y), mix(x.z, y.z, a.z)); }
inline float round(float x) { return ::roundf(x); }                       // half away from zero (reproject_oracle.c)
inline vec2 round(const vec2 &v) { return vec2(::roundf(v.x), ::roundf(v.y)); }
inline ivec2 clamp(const ivec2 &v, const ivec2 &lo, const ivec2 &hi) {
    return ivec2(v.x < lo.x ? lo.x : (v.x > hi.x ? hi.x : v.x), v.y < lo.y ? lo.y : (v.y > hi.y ? hi.y : v.y));
}
inline float cos(float x) { return ::cosf(x); }
inline float sin(float x) { return ::sinf(x); }

// ---- resources
struct sampler2D { int slot; };
struct sampler3D { int slot; };
struct image2D { int id; };
typedef void (*sample_fn)(void *user, int slot, const float uvw[3], float out[4]);
struct Env {
    sample_fn sample = nullptr; void *user = nullptr;
    float *out = nullptr; uint8_t *written = nullptr; int out_w = 0, out_h = 0;
    const float *src = nullptr;                       // RGBA32F source image of the same extent (imageLoad / float-image sampler)
    unsigned long long n2d = 0, n3d = 0;
};
inline Env &env() { static Env e; return e; }
inline vec4 texture(const sampler2D &s, const vec2 &uv) {
    float c[3] = {uv.x, uv.y, 0.0f}, o[4];
    env().sample(env().user, s.slot, c, o); env().n2d++;
    return vec4(o[0], o[1], o[2], o[3]);
}
inline vec4 texture(const sampler3D &s, const vec3 &p) {
    float c[3] = {p.x, p.y, p.z}, o[4];
    env().sample(env().user, s.slot, c, o); env().n3d++;
    return vec4(o[0], o[1], o[2], o[3]);
}
inline ivec2 imageSize(const image2D &) { return ivec2(env().out_w, env().out_h); }
inline vec4 imageLoad(const image2D &, const ivec2 &p) {
    const float *t = env().src + 4 * ((size_t)p.y * env().out_w + p.x);
    return vec4(t[0], t[1], t[2], t[3]);
}
// texture() of an RGBA32F framebuffer through the offscreen sampler (VulkanApplication.cpp:1290-1303: LINEAR, CLAMP_TO_EDGE), as
// post_chain_oracle.c fixes it: a fetch at a pixel centre returns the texel (weights vanish at any filter precision of 12 bits or
// fewer; the rasteriser's fragUV is a centre), every other tap is binary32 bilinear with fused lerps in x then y.
struct fsampler2D { int id; };
inline void ftap_axis(float u, int n, int &i0, int &i1, float &a, bool &centre) {
    float U = (u * (float)n) - 0.5f;
    float fl = ::floorf(U);
    a = U - fl;
    centre = (a < (1.0f / 4096.0f)) || (a > 1.0f - (1.0f / 4096.0f));
    if (!(fl >= -1.0f)) fl = -1.0f;
    if (fl > (float)n) fl = (float)n;
    int i = (int)fl;
    i0 = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    i1 = i + 1 < 0 ? 0 : (i + 1 > n - 1 ? n - 1 : i + 1);
}
inline vec4 texture(const fsampler2D &, const vec2 &uv) {
    int W = env().out_w, H = env().out_h, x0, x1, y0, y1; float a, b; bool cx, cy;
    ftap_axis(uv.x, W, x0, x1, a, cx);
    ftap_axis(uv.y, H, y0, y1, b, cy);
    const float *s = env().src;
    if (cx && cy) {                                    // a pixel centre: the texel nearest to it
        int x = a > 0.5f ? x1 : x0, y = b > 0.5f ? y1 : y0;
        const float *t = s + 4 * ((size_t)y * W + x);
        return vec4(t[0], t[1], t[2], t[3]);
    }
    float o[4];
    for (int c = 0; c < 4; c++) {
        float t00 = s[4 * ((size_t)y0 * W + x0) + c], t10 = s[4 * ((size_t)y0 * W + x1) + c];
        float t01 = s[4 * ((size_t)y1 * W + x0) + c], t11 = s[4 * ((size_t)y1 * W + x1) + c];
        float top = __builtin_fmaf(a, t10 - t00, t00), bot = __builtin_fmaf(a, t11 - t01, t01);
        o[c] = __builtin_fmaf(b, bot - top, top);
    }
    return vec4(o[0], o[1], o[2], o[3]);
}
inline void imageStore(const image2D &, const ivec2 &p, const vec4 &v) {
    if (p.x < 0 || p.y < 0 || p.x >= env().out_w || p.y >= env().out_h) return;
    size_t i = (size_t)p.y * env().out_w + p.x;
    env().out[4 * i] = v.x; env().out[4 * i + 1] = v.y; env().out[4 * i + 2] = v.z; env().out[4 * i + 3] = v.w;
    if (env().written) env().written[i] = 1;
}

}  // namespace glsl
