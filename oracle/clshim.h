// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product.
//
// OpenCL-C 1.1 compatibility layer that lets g++ compile the reference's *unmodified*
// kernel text (/root/reference/Assign*/code.cl) as C++ so that it can be run on host
// cores as the parity oracle.  Only the subset of OpenCL C the reference
// uses is provided (inventory: grep over Assign{01,02,03,07,08,09,10}/code.cl):
//   types      float2/3/4/16, int3, uint2, uchar4, uint, uchar
//   swizzles   .x .y .z .w .s0-.sF (scalars); .xyz .xy .s012 (read+write);
//              .s345 .s678 .s9AB (read)
//   builtins   dot cross normalize length distance fabs min max fmin fmax clamp mad
//              sqrt cos sin get_global_id
//   constants  HUGE_VALF INFINITY UINT_MAX M_PI_4_F M_PI_2_F
//
// Arithmetic policy (OpenCL leaves these implementation-defined within its ulp bounds;
// the reference pins no compiler, README.md:4-6, so ONE evaluation order is fixed here
// and the CUDA product is written to the same one):
//   * all vector ops are component-wise fp32, evaluated left to right, no contraction
//     (build with -ffp-contract=off);
//   * dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z           (left to right)
//   * length(v)     = sqrtf(dot(v,v));  distance(a,b) = length(a-b)
//   * normalize(v)  = v / length(v)                          (three IEEE divisions)
//   * mad(a,b,c)    = fmaf(a,b,c)                            (single rounding)
//   * sqrt          = sqrtf (correctly rounded)
//   * cos/sin       = (float)cos((double)x) / (float)sin((double)x)
//   * int*int wraps (build with -fwrapv) -- quirk Q6 of SURVEY.md
#pragma once
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

typedef unsigned int uint;
typedef unsigned char uchar;

#define __kernel
#define __global
#define __constant const
#define __local
#define __private
#define M_PI_4_F 0.78539816339744830962f
#define M_PI_2_F 1.57079632679489661923f

// ---- NDRange plumbing (set by the driver loop before every work-item call) ----------
extern thread_local size_t cl_gid[3];
static inline size_t get_global_id(uint d) { return cl_gid[d]; }

struct float2;
struct float3;
struct float4;

// A window of K consecutive floats starting at offset O inside a parent vector of N
// floats.  Lives in an anonymous union with the parent's storage; converts to/from the
// K-wide value type.  Writes touch only the K addressed lanes.
template <typename V, int K, int O, int N>
struct swz {
    float v[N];
    inline operator V() const;
    inline swz& operator=(const V& a);
    inline swz& operator=(const swz& a) { for (int i = 0; i < K; i++) v[O + i] = a.v[O + i]; return *this; }
    template <int O2, int N2>
    inline swz& operator=(const swz<V, K, O2, N2>& a) { for (int i = 0; i < K; i++) v[O + i] = a.v[O2 + i]; return *this; }
    inline swz& operator*=(const V& a);
};

struct alignas(8) float2 {
    union {
        struct { float x, y; };
        struct { float s0, s1; };
        float v[2];
    };
    float2() {}
    float2(float a, float b) : x(a), y(b) {}
};

struct alignas(16) float3 {
    union {
        struct { float x, y, z; };
        struct { float s0, s1, s2; };
        float v[4];
        swz<float3, 3, 0, 4> xyz, s012;
        swz<float2, 2, 0, 4> xy;
    };
    float3() {}
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    float3(const float3& o) { x = o.x; y = o.y; z = o.z; }
    float3& operator=(const float3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};

struct alignas(16) float4 {
    union {
        struct { float x, y, z, w; };
        struct { float s0, s1, s2, s3; };
        float v[4];
        swz<float3, 3, 0, 4> xyz, s012;
        swz<float2, 2, 0, 4> xy;
    };
    float4() {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(const float4& o) { memcpy(v, o.v, sizeof v); }
    float4& operator=(const float4& o) { memcpy(v, o.v, sizeof v); return *this; }
};

struct alignas(64) float16 {
    union {
        struct { float s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sA, sB, sC, sD, sE, sF; };
        float v[16];
        swz<float3, 3, 0, 16> s012;
        swz<float3, 3, 3, 16> s345;
        swz<float3, 3, 6, 16> s678;
        swz<float3, 3, 9, 16> s9AB;
    };
    float16() {}
    float16(const float16& o) { memcpy(v, o.v, sizeof v); }
    float16& operator=(const float16& o) { memcpy(v, o.v, sizeof v); return *this; }
};

struct alignas(16) int3 {
    int x, y, z, _pad;
    int3() {}
    int3(int a, int b, int c) : x(a), y(b), z(c), _pad(0) {}
};
struct alignas(8) uint2 {
    union {
        struct { uint x, y; };
        struct { uint s0, s1; };
    };
    uint2() {}
    uint2(uint a, uint b) : x(a), y(b) {}
};
struct alignas(4) uchar4 {
    union {
        struct { uchar x, y, z, w; };
        struct { uchar s0, s1, s2, s3; };
    };
    uchar4() {}
};

static_assert(sizeof(float2) == 8 && sizeof(float3) == 16 && sizeof(float4) == 16, "vector sizes");
static_assert(sizeof(float16) == 64 && sizeof(int3) == 16 && sizeof(uchar4) == 4, "vector sizes");

// ---- swizzle window <-> value --------------------------------------------------------
template <> inline swz<float3, 3, 0, 4>::operator float3() const { return float3(v[0], v[1], v[2]); }
template <> inline swz<float3, 3, 0, 16>::operator float3() const { return float3(v[0], v[1], v[2]); }
template <> inline swz<float3, 3, 3, 16>::operator float3() const { return float3(v[3], v[4], v[5]); }
template <> inline swz<float3, 3, 6, 16>::operator float3() const { return float3(v[6], v[7], v[8]); }
template <> inline swz<float3, 3, 9, 16>::operator float3() const { return float3(v[9], v[10], v[11]); }
template <> inline swz<float2, 2, 0, 4>::operator float2() const { return float2(v[0], v[1]); }
template <> inline swz<float3, 3, 0, 4>& swz<float3, 3, 0, 4>::operator=(const float3& a) { v[0] = a.x; v[1] = a.y; v[2] = a.z; return *this; }
template <> inline swz<float2, 2, 0, 4>& swz<float2, 2, 0, 4>::operator=(const float2& a) { v[0] = a.x; v[1] = a.y; return *this; }
template <> inline swz<float3, 3, 0, 4>& swz<float3, 3, 0, 4>::operator*=(const float3& a) { v[0] *= a.x; v[1] *= a.y; v[2] *= a.z; return *this; }

// ---- vector literals: "(float3)(a,b,c)" is rewritten to "mk_float3(a,b,c)" by cl2cpp.py
static inline float2 mk_float2(float a, float b) { return float2(a, b); }
static inline float2 mk_float2(float a) { return float2(a, a); }
static inline float3 mk_float3(float a, float b, float c) { return float3(a, b, c); }
static inline float3 mk_float3(float a) { return float3(a, a, a); }
static inline float4 mk_float4(float a, float b, float c, float d) { return float4(a, b, c, d); }
static inline float4 mk_float4(float a) { return float4(a, a, a, a); }
static inline float4 mk_float4(const float3& a, float d) { return float4(a.x, a.y, a.z, d); }
static inline int3 mk_int3(int a, int b, int c) { return int3(a, b, c); }
static inline uint2 mk_uint2(uint a, uint b) { return uint2(a, b); }
template <typename A, typename B, typename C, typename D>
static inline uchar4 mk_uchar4(A a, B b, C c, D d) {
    uchar4 r; r.x = (uchar)a; r.y = (uchar)b; r.z = (uchar)c; r.w = (uchar)d; return r;
}

// ---- component-wise arithmetic ---------------------------------------------------------
static inline float2 operator*(const float2& a, float s) { return float2(a.x * s, a.y * s); }
static inline float2 operator*(float s, const float2& a) { return float2(s * a.x, s * a.y); }
static inline float2 operator-(const float2& a, float s) { return float2(a.x - s, a.y - s); }
static inline float2 operator+(const float2& a, const float2& b) { return float2(a.x + b.x, a.y + b.y); }

static inline float3 operator+(const float3& a, const float3& b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(const float3& a, const float3& b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(const float3& a, const float3& b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator/(const float3& a, const float3& b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float3 operator*(const float3& a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(float s, const float3& a) { return float3(s * a.x, s * a.y, s * a.z); }
static inline float3 operator/(const float3& a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
static inline float3 operator-(const float3& a) { return float3(-a.x, -a.y, -a.z); }
static inline float3& operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
static inline float3& operator-=(float3& a, const float3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
static inline float3& operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
static inline float3& operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }

static inline float4 operator+(const float4& a, const float4& b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline float4 operator*(const float4& a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline float4 operator*(float s, const float4& a) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
static inline float4 operator*(const float4& a, const float4& b) { return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline float4& operator+=(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; return a; }
static inline float4& operator*=(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; return a; }

// ---- builtins ----------------------------------------------------------------------------
static inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 cross(const float3& a, const float3& b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float sqrt(float a) { return sqrtf(a); }
static inline float length(const float3& a) { return sqrtf(dot(a, a)); }
static inline float distance(const float3& a, const float3& b) { return length(a - b); }
static inline float3 normalize(const float3& a) { return a / length(a); }
static inline float fabs(float a) { return fabsf(a); }
static inline float3 fabs(const float3& a) { return float3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// OpenCL fmin/fmax/min/max on floats: operands are never NaN-vs-number sensitive in the
// reference except through comparisons, so the C ternary forms of the OpenCL spec
// (min(x,y) = y < x ? y : x ; max(x,y) = x < y ? y : x) are used for min/max and
// IEEE fminf/fmaxf for fmin/fmax.
static inline float min(float x, float y) { return y < x ? y : x; }
static inline float max(float x, float y) { return x < y ? y : x; }
static inline float fmin(float x, float y) { return fminf(x, y); }
static inline float fmax(float x, float y) { return fmaxf(x, y); }
static inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
static inline float4 clamp(const float4& a, float lo, float hi) {
    return float4(clamp(a.x, lo, hi), clamp(a.y, lo, hi), clamp(a.z, lo, hi), clamp(a.w, lo, hi));
}
static inline float mad(float a, float b, float c) { return fmaf(a, b, c); }
static inline float cos(float a) { return (float)::cos((double)a); }
static inline float sin(float a) { return (float)::sin((double)a); }
