// glsl_env_fma.h -- the GLSL execution environment of glsl_env.h with FUSED MULTIPLY-ADD CONTRACTION (TEST INFRASTRUCTURE).
//
// GLSL lets an implementation contract `a * b + c` into one fused operation (the spec's "precise" qualifier exists to forbid it,
// and compute-clouds.comp does not use it); every GPU the reference ran on did.  Which products fuse is the implementation's choice,
// so this project pins ONE definition -- the lexical one -- and states it three times: here (applied mechanically to the reference's
// own shader text by C++ overload resolution), in oracle/cloud_march_oracle_fma.c (by hand) and in csrc/cloud_march.cu (explicit
// __fmaf_rn; arithmetic mode MM_ARITH_FMA).  tests/test_reference_shader.py compares the first two bit for bit.
//
// THE RULE.  A binary `*` whose value is DIRECTLY an operand of a binary `+` or `-` (parentheses do not matter; compound `+=` / `-=`
// count as `x = x + (...)`) is not rounded: the add consumes the exact product,
//        a*b + c  ->  fma(a, b, c)          c + a*b  ->  fma(a, b, c)
//        a*b - c  ->  fma(a, b, -c)         c - a*b  ->  fma(-a, b, c)
//        a*b + c*d  ->  fma(a, b, RN(c*d))  a*b - c*d  ->  fma(a, b, -RN(c*d))        (the LEFT product fuses; the right one is rounded)
// componentwise for vectors, a scalar factor or addend broadcast.  A product used in any other way (assigned, passed to a function,
// multiplied or divided further, compared, negated and then used elsewhere) is rounded to binary32 first.  Products of two literal
// constants are constants (folded, rounded once) and never fuse.  Built-ins are the formulas of glsl_env.h under the same rule:
//        dot(a,b)    = fma(a.z, b.z, fma(a.x, b.x, RN(a.y*b.y)))                       [((ax*bx)+(ay*by))+(az*bz)]
//        mix(x,y,a)  = fma(x, 1-a, RN(y*a))                                            [x*(1-a) + y*a]
//        mat3 * v    = fma(c2, v.z, fma(c0, v.x, RN(c1*v.y)))  per component           [((c0*v.x)+(c1*v.y))+(c2*v.z)]
//        smoothstep  = RN(t*t) * fma(-2, t, 3)                                         [(t*t)*(3-(2*t))]
//        length, normalize, min, max, clamp, pow, exp, ... as in glsl_env.h (normalize(v) = v*(1/sqrt(dot(v,v))), no fusion: a product
//        by the reciprocal is not added to anything).
// Mechanism: the shader's `float` is rewritten to the class Float (glsl_to_cpp.py --float-class); Float * Float yields a Prod that
// remembers both factors, `+`/`-` are overloaded on Prod, and every other use converts Prod to Float, i.e. rounds it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glslf {

typedef unsigned int uint;
inline float fma_(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline int to_int(float f) { if (!(f == f)) return 0; if (f >= 2147483648.0f) return 2147483647; if (f <= -2147483648.0f) return (-2147483647 - 1); return (int)f; }

struct ProdF;
struct Float {
    float v;
    Float() = default;
    Float(float f) : v(f) {}
    Float(int i) : v((float)i) {}
    Float(unsigned u) : v((float)u) {}
    explicit operator int() const { return to_int(v); }                 // GLSL int(x): truncation
    Float operator-() const { return Float(-v); }
    Float &operator+=(Float o) { v = v + o.v; return *this; }
    Float &operator-=(Float o) { v = v - o.v; return *this; }
    Float &operator*=(Float o) { v = v * o.v; return *this; }
    Float &operator/=(Float o) { v = v / o.v; return *this; }
    inline Float &operator+=(const ProdF &p);
    inline Float &operator-=(const ProdF &p);
};
struct ProdF {                                                          // an unrounded product a*b
    float a, b;
    operator Float() const { return Float(a * b); }
    ProdF operator-() const { return ProdF{-a, b}; }
};
inline Float &Float::operator+=(const ProdF &p) { v = fma_(p.a, p.b, v); return *this; }
inline Float &Float::operator-=(const ProdF &p) { v = fma_(-p.a, p.b, v); return *this; }
inline ProdF operator*(Float a, Float b) { return ProdF{a.v, b.v}; }
inline Float operator/(Float a, Float b) { return Float(a.v / b.v); }
inline Float operator+(Float a, Float b) { return Float(a.v + b.v); }
inline Float operator-(Float a, Float b) { return Float(a.v - b.v); }
inline Float operator+(ProdF p, Float c) { return Float(fma_(p.a, p.b, c.v)); }
inline Float operator+(Float c, ProdF p) { return Float(fma_(p.a, p.b, c.v)); }
inline Float operator+(ProdF p, ProdF q) { return Float(fma_(p.a, p.b, q.a * q.b)); }
inline Float operator-(ProdF p, Float c) { return Float(fma_(p.a, p.b, -c.v)); }
inline Float operator-(Float c, ProdF p) { return Float(fma_(-p.a, p.b, c.v)); }
inline Float operator-(ProdF p, ProdF q) { return Float(fma_(p.a, p.b, -(q.a * q.b))); }
inline bool operator<(Float a, Float b) { return a.v < b.v; }
inline bool operator>(Float a, Float b) { return a.v > b.v; }
inline bool operator<=(Float a, Float b) { return a.v <= b.v; }
inline bool operator>=(Float a, Float b) { return a.v >= b.v; }
inline bool operator==(Float a, Float b) { return a.v == b.v; }
inline bool operator!=(Float a, Float b) { return a.v != b.v; }

struct vec2; struct vec3; struct vec4; struct uvec2; struct ivec2;
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    ivec2(uint a, uint b) : x((int)a), y((int)b) {}
};
struct uvec3 { uint x, y, z; uvec2 xy() const { return uvec2(x, y); } };

struct vec2 {
    union { Float x, r; }; union { Float y, g; };
    vec2() : x(0.0f), y(0.0f) {}
    explicit vec2(Float s) : x(s), y(s) {}
    vec2(Float a, Float b) : x(a), y(b) {}
};
struct ref3 {
    Float &a, &b, &c;
    ref3(Float &a_, Float &b_, Float &c_) : a(a_), b(b_), c(c_) {}
    inline ref3 &operator=(const vec3 &v);
    inline operator vec3() const;
};
struct ref2 {
    Float &a, &b;
    ref2(Float &a_, Float &b_) : a(a_), b(b_) {}
    operator vec2() const { return vec2(a, b); }
};
struct Prod3;
struct vec3 {
    union { Float x, r; }; union { Float y, g; }; union { Float z, b; };
    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    explicit vec3(Float s) : x(s), y(s), z(s) {}
    vec3(Float a, Float c, Float d) : x(a), y(c), z(d) {}
    vec3(const vec3 &o) : x(o.x), y(o.y), z(o.z) {}
    vec3 &operator=(const vec3 &o) { x = o.x; y = o.y; z = o.z; return *this; }
    Float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3 &operator-=(const vec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3 &operator*=(const vec3 &o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    vec3 &operator+=(Float s) { x += s; y += s; z += s; return *this; }
    vec3 &operator*=(Float s) { x *= s; y *= s; z *= s; return *this; }
    vec3 &operator/=(Float s) { x /= s; y /= s; z /= s; return *this; }
    inline vec3 &operator+=(const Prod3 &p);
    ref2 xz() { return ref2(x, z); }
    ref2 xy() { return ref2(x, y); }
    vec2 xz() const { return vec2(x, z); }
    vec2 xy() const { return vec2(x, y); }
    ref3 xyz() { return ref3(x, y, z); }
    vec3 xyz() const { return *this; }
    ref3 rgb() { return ref3(x, y, z); }
    vec3 rgb() const { return *this; }
};
struct vec4 {
    union { Float x, r; }; union { Float y, g; }; union { Float z, b; }; union { Float w, a; };
    vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    vec4(Float a_, Float b_, Float c_, Float d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(const vec3 &v, Float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(const vec4 &o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4 &operator=(const vec4 &o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    Float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    ref3 xyz() { return ref3(x, y, z); }
    vec3 xyz() const { return vec3(x, y, z); }
    ref3 rgb() { return ref3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
    ref2 xz() { return ref2(x, z); }
    vec2 xz() const { return vec2(x, z); }
};
inline ref3 &ref3::operator=(const vec3 &v) { a = v.x; b = v.y; c = v.z; return *this; }
inline ref3::operator vec3() const { return vec3(a, b, c); }

// ---- unrounded componentwise products; a scalar factor is broadcast
struct Prod2 {
    vec2 a, b;
    operator vec2() const { return vec2(Float(a.x.v * b.x.v), Float(a.y.v * b.y.v)); }
};
struct Prod3 {
    vec3 a, b;
    operator vec3() const { return vec3(Float(a.x.v * b.x.v), Float(a.y.v * b.y.v), Float(a.z.v * b.z.v)); }
};
inline vec3 fma3(const vec3 &a, const vec3 &b, const vec3 &c) { return vec3(Float(fma_(a.x.v, b.x.v, c.x.v)), Float(fma_(a.y.v, b.y.v, c.y.v)), Float(fma_(a.z.v, b.z.v, c.z.v))); }
inline vec2 fma2(const vec2 &a, const vec2 &b, const vec2 &c) { return vec2(Float(fma_(a.x.v, b.x.v, c.x.v)), Float(fma_(a.y.v, b.y.v, c.y.v))); }
inline vec3 neg3(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline vec2 neg2(const vec2 &a) { return vec2(-a.x, -a.y); }
inline vec3 &vec3::operator+=(const Prod3 &p) { *this = fma3(p.a, p.b, *this); return *this; }

inline Prod3 operator*(const vec3 &a, const vec3 &b) { return Prod3{a, b}; }
inline Prod3 operator*(const vec3 &a, Float s) { return Prod3{a, vec3(s)}; }
inline Prod3 operator*(Float s, const vec3 &a) { return Prod3{vec3(s), a}; }
inline Prod2 operator*(const vec2 &a, const vec2 &b) { return Prod2{a, b}; }
inline Prod2 operator*(const vec2 &a, Float s) { return Prod2{a, vec2(s)}; }
inline Prod2 operator*(Float s, const vec2 &a) { return Prod2{vec2(s), a}; }
inline vec3 operator-(const vec3 &a) { return neg3(a); }

inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator/(const vec3 &a, Float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(const vec3 &a, Float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3 &a, Float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(Float s, const vec3 &a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator+(const Prod3 &p, const vec3 &c) { return fma3(p.a, p.b, c); }
inline vec3 operator+(const vec3 &c, const Prod3 &p) { return fma3(p.a, p.b, c); }
inline vec3 operator+(const Prod3 &p, const Prod3 &q) { return fma3(p.a, p.b, (vec3)q); }
inline vec3 operator-(const Prod3 &p, const vec3 &c) { return fma3(p.a, p.b, neg3(c)); }
inline vec3 operator-(const vec3 &c, const Prod3 &p) { return fma3(neg3(p.a), p.b, c); }
inline vec3 operator-(const Prod3 &p, const Prod3 &q) { return fma3(p.a, p.b, neg3((vec3)q)); }
inline vec3 operator+(const Prod3 &p, Float s) { return fma3(p.a, p.b, vec3(s)); }
inline vec3 operator-(const Prod3 &p, Float s) { return fma3(p.a, p.b, vec3(-s)); }

inline vec2 operator+(const vec2 &a, const vec2 &b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2 &a, const vec2 &b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator/(const vec2 &a, const vec2 &b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator+(const Prod2 &p, const vec2 &c) { return fma2(p.a, p.b, c); }
inline vec2 operator+(const Prod2 &p, Float s) { return fma2(p.a, p.b, vec2(s)); }
inline vec2 operator-(const Prod2 &p, Float s) { return fma2(p.a, p.b, vec2(-s)); }

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
// column-major product, ((c0*v.x) + (c1*v.y)) + (c2*v.z) under the rule
inline vec3 operator*(const mat3 &m, const vec3 &v) { return ((m.c[0] * v.x) + (m.c[1] * v.y)) + (m.c[2] * v.z); }

// ---- built-ins: the formulas of glsl_env.h, written with this header's operators so that the rule applies inside them
inline Float dot(const vec3 &a, const vec3 &b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
inline Float sqrt(Float x) { return Float(::sqrtf(x.v)); }
inline vec3 sqrt(const vec3 &v) { return vec3(sqrt(v.x), sqrt(v.y), sqrt(v.z)); }
inline Float length(const vec3 &a) { return sqrt(dot(a, a)); }
inline vec3 normalize(const vec3 &a) { Float inv = Float(1.0f) / sqrt(dot(a, a)); return vec3(a.x * inv, a.y * inv, a.z * inv); }
inline Float max(Float x, Float y) { return (x < y) ? y : x; }
inline Float min(Float x, Float y) { return (y < x) ? y : x; }
inline Float clamp(Float x, Float lo, Float hi) { Float r = (x > lo) ? x : lo; return (r < hi) ? r : hi; }
inline Float mix(Float x, Float y, Float a) { return (x * (Float(1.0f) - a)) + (y * a); }
inline vec3 mix(const vec3 &x, const vec3 &y, Float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
inline Float smoothstep(Float e0, Float e1, Float x) { Float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return (t * t) * (Float(3.0f) - (Float(2.0f) * t)); }
inline Float pow(Float x, Float y) { return Float(::powf(x.v, y.v)); }
inline vec3 pow(const vec3 &x, const vec3 &y) { return vec3(pow(x.x, y.x), pow(x.y, y.y), pow(x.z, y.z)); }
inline Float exp(Float x) { return Float(::expf(x.v)); }
inline vec3 exp(const vec3 &x) { return vec3(exp(x.x), exp(x.y), exp(x.z)); }
inline Float acos(Float x) { return Float(::acosf(x.v)); }
inline Float cos(Float x) { return Float(::cosf(x.v)); }
inline Float sin(Float x) { return Float(::sinf(x.v)); }

// ---- resources (as glsl_env.h)
struct sampler2D { int slot; };
struct sampler3D { int slot; };
struct image2D { int id; };
typedef void (*sample_fn)(void *user, int slot, const float uvw[3], float out[4]);
struct Env {
    sample_fn sample = nullptr; void *user = nullptr;
    float *out = nullptr; uint8_t *written = nullptr; int out_w = 0, out_h = 0;
    unsigned long long n2d = 0, n3d = 0;
};
inline Env &env() { static Env e; return e; }
inline vec4 texture(const sampler2D &s, const vec2 &uv) {
    float c[3] = {uv.x.v, uv.y.v, 0.0f}, o[4];
    env().sample(env().user, s.slot, c, o); env().n2d++;
    return vec4(o[0], o[1], o[2], o[3]);
}
inline vec4 texture(const sampler3D &s, const vec3 &p) {
    float c[3] = {p.x.v, p.y.v, p.z.v}, o[4];
    env().sample(env().user, s.slot, c, o); env().n3d++;
    return vec4(o[0], o[1], o[2], o[3]);
}
inline ivec2 imageSize(const image2D &) { return ivec2(env().out_w, env().out_h); }
inline void imageStore(const image2D &, const ivec2 &p, const vec4 &v) {
    if (p.x < 0 || p.y < 0 || p.x >= env().out_w || p.y >= env().out_h) return;
    size_t i = (size_t)p.y * env().out_w + p.x;
    env().out[4 * i] = v.x.v; env().out[4 * i + 1] = v.y.v; env().out[4 * i + 2] = v.z.v; env().out[4 * i + 3] = v.w.v;
    if (env().written) env().written[i] = 1;
}

}  // namespace glslf
