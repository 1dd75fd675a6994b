// rt_device.cuh -- device-side building blocks of the B200 ray tracer.
//
// Every function here computes exactly what one helper of the reference's OpenCL kernels
// computes, in the same fp32 operation order (file:line cited per function, paths relative
// to /root/reference; A10 = Assign10-Path_Tracing).  The translation unit is compiled with
// -fmad=false (no contraction; the one explicit `mad` of the reference is an explicit
// fmaf), IEEE division and square root (nvcc defaults -prec-div=true -prec-sqrt=true,
// -ftz=false), so results are bit-comparable with the reference text evaluated under the
// evaluation order documented in DESIGN.md ("arithmetic contract").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

// ---------------------------------------------------------------------------------------
// Memory layouts (reference: A10/code.cl:22-62)
// ---------------------------------------------------------------------------------------
struct __align__(16) Ray {   // 48 B
    float ox, oy, oz, _p0;
    float dx, dy, dz, _p1;
    float mint, maxt, _p2, _p3;
};
struct __align__(16) Poi10 { // 64 B (A10)
    float px, py, pz, _p0;
    float nx, ny, nz, _p1;
    float ax, ay, az, _p2;   // atte
    int matId, _p3, _p4, _p5;
};
struct __align__(16) Poi8 {  // 48 B (A08/A09)
    float px, py, pz, _p0;
    float nx, ny, nz, _p1;
    int matId, _p2, _p3, _p4;
};
static_assert(sizeof(Ray) == 48 && sizeof(Poi10) == 64 && sizeof(Poi8) == 48, "reference layouts");

struct f3 { float x, y, z; };
struct f2 { float x, y; };
struct AABB { f3 pmin, pmax; };
struct Camera {           // A10/code.cl:33-38, floatToCamera :73-84
    f3 eye, U, V, W;
    float width, height;
    unsigned cols, rows;
};

#define RT_DEV __device__ __forceinline__
#define RT_INF __int_as_float(0x7f800000)

RT_DEV f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_DEV f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEV f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_DEV f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
RT_DEV f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
RT_DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_DEV float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_DEV f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT_DEV float length(f3 a) { return sqrtf(dot(a, a)); }
RT_DEV float distance(f3 a, f3 b) { return length(a - b); }
RT_DEV f3 normalize(f3 a) { return a / length(a); }
// OpenCL min/max on floats as the C ternaries of the spec (NaN behaviour differs from fminf).
RT_DEV float cl_min(float x, float y) { return y < x ? y : x; }
RT_DEV float cl_max(float x, float y) { return x < y ? y : x; }
RT_DEV float cl_clamp(float x, float lo, float hi) { return cl_min(cl_max(x, lo), hi); }
RT_DEV float cl_cos(float a) { return (float)cos((double)a); }
RT_DEV float cl_sin(float a) { return (float)sin((double)a); }

RT_DEV Camera floatToCamera(const float* in) {   // A10/code.cl:73-84
    Camera c;
    c.eye = mk3(in[0], in[1], in[2]);
    c.U = mk3(in[3], in[4], in[5]);
    c.V = mk3(in[6], in[7], in[8]);
    c.W = mk3(in[9], in[10], in[11]);
    c.width = in[12];
    c.height = in[13];
    c.cols = (unsigned)in[14];
    c.rows = (unsigned)in[15];
    return c;
}

struct CamArg { float v[16]; };    // by-value float16 kernel argument
struct AabbArg { float v[8]; };    // by-value AABB (pmin.xyzw, pmax.xyzw)
struct LightArg { float v[16]; };  // by-value float16 light packing (A10/code.js:323-352)

RT_DEV AABB toAABB(const AabbArg& a) {
    AABB b;
    b.pmin = mk3(a.v[0], a.v[1], a.v[2]);
    b.pmax = mk3(a.v[4], a.v[5], a.v[6]);
    return b;
}

// ---------------------------------------------------------------------------------------
// Ray load/store helpers (vectorised: 3 x 16 B)
// ---------------------------------------------------------------------------------------
struct RayR {   // register form
    f3 o, d;
    float mint, maxt;
};
RT_DEV RayR loadRay(const Ray* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = q[0], b = q[1], c = q[2];
    RayR r;
    r.o = mk3(a.x, a.y, a.z);
    r.d = mk3(b.x, b.y, b.z);
    r.mint = c.x;
    r.maxt = c.y;
    return r;
}
RT_DEV void storeRay(Ray* p, const RayR& r) {
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4(r.o.x, r.o.y, r.o.z, 0.0f);
    q[1] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
    q[2] = make_float4(r.mint, r.maxt, 0.0f, 0.0f);
}
RT_DEV void storeDeadRay(Ray* p) {   // "ray.mint = ray.maxt = HUGE_VALF" with the rest unspecified
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    q[2] = make_float4(RT_INF, RT_INF, 0.0f, 0.0f);
}

// ---------------------------------------------------------------------------------------
// Camera rays
// ---------------------------------------------------------------------------------------
RT_DEV f3 getPoint(f3 o, f3 d, float t) { return o + t * d; }   // A10/code.cl:86-88

// A10/code.cl:108-119.  col,row arrive as float (uint -> float at the call site).
RT_DEV void getRay(const Camera& cam, float col, float row, f3& o, f3& d) {
    f3 cop = (-0.5f + (col + 0.5f) / (float)cam.cols) * cam.width * cam.U +
             (0.5f - (row + 0.5f) / (float)cam.rows) * cam.height * cam.V +
             (-1.0f) * cam.W;
    d = normalize(cop);
    o = cam.eye;
}

#ifndef RT_SINCOS
#define RT_SINCOS 1
#endif
// A10/code.cl:143-172 (Shirley/Whittle concentric map).
RT_DEV f2 concentric_distort(f2 in) {
    if (in.x == 0.0f && in.y == 0.0f) return in;
    float phi = 0.0f;
    float radius = 1.0f;
    float a = (2.0f * in.x) - 1.0f;
    float b = (2.0f * in.y) - 1.0f;
    if ((a * a) > (b * b)) {
        radius *= a;
        phi = 0.78539816339744830962f * (b / a);
    } else {
        radius *= b;
        phi = 1.57079632679489661923f - (0.78539816339744830962f * (a / b));
    }
    f2 r;
#if RT_SINCOS
    double sn, cs;   // one range reduction for both; same values as cos() and sin() (tests/test_gpu_a10.py checks every float of the range)
    sincos((double)phi, &sn, &cs);
    r.x = (float)cs * radius;
    r.y = (float)sn * radius;
#else
    r.x = cl_cos(phi) * radius;
    r.y = cl_sin(phi) * radius;
#endif
    return r;
}

// A10/code.cl:174-181.
RT_DEV f3 getFocalPoint(const Camera& cam, float col, float row, float focal_length) {
    f3 o, d;
    getRay(cam, col, row, o, d);
    f3 pip = cam.eye + focal_length * cam.W * (-1.0f);
    f3 N = cam.W;
    float dd = -dot(pip, N);
    float t = -(dot(o, N) + dd) / dot(d, N);
    return getPoint(o, d, t);
}

// A10/code.cl:183-197.
RT_DEV void getThinLensRay(const Camera& cam, f3 focal_point, float lens_rad, f2 coord, f3& o, f3& d) {
    f2 dc = concentric_distort(coord);
    dc.x = dc.x * lens_rad;
    dc.y = dc.y * lens_rad;
    o = cam.eye + dc.x * cam.U + dc.y * cam.V;
    d = normalize(focal_point - o);
}

// A10/code.cl:121-129.
RT_DEV RayR makeRay(f3 ori, f3 dst) {
    RayR r;
    r.o = ori;
    r.d = normalize(dst - ori);
    r.mint = 0.0f;
    r.maxt = length(dst - ori);
    return r;
}

// Read-only loads of the scene data the queue walkers come back to all the time (cell table, occupancy bits, face
// vectors).  RT_KEEP_L2 = 1 gives them an L2 evict-last policy (createpolicy + ld.global.nc.L2::cache_hint) so that the
// gigabytes of per-slot state streaming through the same L2 do not push them out.
#ifndef RT_KEEP_L2
#define RT_KEEP_L2 0
#endif
RT_DEV float4 ldKeep(const float4* p) {
#if RT_KEEP_L2
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
RT_DEV unsigned ldKeep(const unsigned* p) {
#if RT_KEEP_L2
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    unsigned v;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}

// ---------------------------------------------------------------------------------------
// Intersections
// ---------------------------------------------------------------------------------------
struct AabbHit { float tmin, tmax; bool v; };

// A10/code.cl:335-389.  Starts from [0,+inf) and ignores ray.mint/maxt.
RT_DEV AabbHit interAABB(f3 o, f3 d, const AABB& box) {
    AabbHit h;
    h.tmin = 0.0f;
    h.tmax = RT_INF;
    h.v = false;
    float ttmin, ttmax, tmp;
    ttmin = (box.pmin.x - o.x) / d.x;
    ttmax = (box.pmax.x - o.x) / d.x;
    if (d.x < 0) { tmp = ttmin; ttmin = ttmax; ttmax = tmp; }
    h.tmin = cl_max(ttmin, h.tmin);
    h.tmax = cl_min(ttmax, h.tmax);
    if (h.tmin > h.tmax) return h;
    ttmin = (box.pmin.y - o.y) / d.y;
    ttmax = (box.pmax.y - o.y) / d.y;
    if (d.y < 0) { tmp = ttmin; ttmin = ttmax; ttmax = tmp; }
    h.tmin = cl_max(ttmin, h.tmin);
    h.tmax = cl_min(ttmax, h.tmax);
    if (h.tmin > h.tmax) return h;
    ttmin = (box.pmin.z - o.z) / d.z;
    ttmax = (box.pmax.z - o.z) / d.z;
    if (d.z < 0) { tmp = ttmin; ttmin = ttmax; ttmax = tmp; }
    h.tmin = cl_max(ttmin, h.tmin);
    h.tmax = cl_min(ttmax, h.tmax);
    if (h.tmin > h.tmax) return h;
    h.v = true;
    return h;
}

// interAABB that also hands out the three far-plane quotients (the per-axis ttmax after the swap) for the 1-cell walk.
struct AabbFar { float tmin, tmax, fx, fy, fz; bool v; };
RT_DEV AabbFar interAABBFar(f3 o, f3 d, const AABB& box) {
    AabbFar h;
    h.tmin = 0.0f;
    h.tmax = RT_INF;
    h.fx = h.fy = h.fz = 0.0f;
    h.v = false;
    float ttmin, ttmax, tmp;
    ttmin = (box.pmin.x - o.x) / d.x;
    ttmax = (box.pmax.x - o.x) / d.x;
    if (d.x < 0) { tmp = ttmin; ttmin = ttmax; ttmax = tmp; }
    h.fx = ttmax;
    h.tmin = cl_max(ttmin, h.tmin);
    h.tmax = cl_min(ttmax, h.tmax);
    if (h.tmin > h.tmax) return h;
    ttmin = (box.pmin.y - o.y) / d.y;
    ttmax = (box.pmax.y - o.y) / d.y;
    if (d.y < 0) { tmp = ttmin; ttmin = ttmax; ttmax = tmp; }
    h.fy = ttmax;
    h.tmin = cl_max(ttmin, h.tmin);
    h.tmax = cl_min(ttmax, h.tmax);
    if (h.tmin > h.tmax) return h;
    ttmin = (box.pmin.z - o.z) / d.z;
    ttmax = (box.pmax.z - o.z) / d.z;
    if (d.z < 0) { tmp = ttmin; ttmin = ttmax; ttmax = tmp; }
    h.fz = ttmax;
    h.tmin = cl_max(ttmin, h.tmin);
    h.tmax = cl_min(ttmax, h.tmax);
    if (h.tmin > h.tmax) return h;
    h.v = true;
    return h;
}

// A10/code.cl:199-242 (A07-A10 form: stored w = r^2, inclusive range test).
// `a` = dot(d,d) is loop-invariant and hoisted by the caller (same value every call).
RT_DEV bool interSphere(f3 o, f3 d, float a_dd, float mint, float maxt, float4 s, float& t_out) {
    f3 omc = o - mk3(s.x, s.y, s.z);
    float a = a_dd;
    float b = 2.0f * dot(omc, d);
    float c = dot(omc, omc) - s.w;
    float dis = fmaf(-4.0f * c, a, b * b);
    if (dis < 0.0f) return false;
    a = 1.0f / (2.0f * a);
    dis = sqrtf(dis);
    float t0 = (-b - dis) * a;
    float t1 = (-b + dis) * a;
    float tmin = fminf(t0, t1);
    float tmax = fmaxf(t0, t1);
    if (tmin >= mint && tmin <= maxt) { t_out = tmin; return true; }
    if (tmax >= mint && tmax <= maxt) { t_out = tmax; return true; }
    return false;
}

// A10/code.cl:250-288 (INCL = true: A08-A10 inclusive `>= / <=`; false: A07/code.cl:195
// exclusive `> / <`).  One-sided: rejects div <= 0.
template <bool INCL>
RT_DEV bool interTriangle(f3 o, f3 d, float mint, float maxt, f3 p0, f3 p1, f3 p2, float& beta_o, float& gamma_o, float& t_out) {
    f3 e1 = p1 - p0;
    f3 e2 = p2 - p0;
    float div = dot(cross(e2, e1), d);
    if (div <= 0) return false;
    float idiv = 1.0f / div;
    f3 s = o - p0;
    float beta = dot(cross(s, d), e2) * idiv;
    if (beta < 0.0f || beta > 1.0f) return false;
    float gamma = dot(cross(s, e1), d) * idiv;
    if (gamma < 0.0f || (gamma + beta) < 0.0f || (gamma + beta) > 1.0f) return false;
    float t = dot(cross(s, e2), e1) * -idiv;
    bool ok = INCL ? (t >= mint && t <= maxt) : (t > mint && t < maxt);
    if (!ok) return false;
    beta_o = beta;
    gamma_o = gamma;
    t_out = t;
    return true;
}

// Same test on the precomputed form (p0, e1, e2) with div = dot(ng, d) supplied by the caller,
// ng = cross(e2, e1) -- identical operations, hence identical bits, as interTriangle<true>.
RT_DEV bool interTrianglePre(f3 o, f3 d, float mint, float maxt, float div, f3 p0, f3 e1, f3 e2, float& beta_o, float& gamma_o, float& t_out) {
    if (div <= 0) return false;
    float idiv = 1.0f / div;
    f3 s = o - p0;
    float beta = dot(cross(s, d), e2) * idiv;
    if (beta < 0.0f || beta > 1.0f) return false;
    float gamma = dot(cross(s, e1), d) * idiv;
    if (gamma < 0.0f || (gamma + beta) < 0.0f || (gamma + beta) > 1.0f) return false;
    float t = dot(cross(s, e2), e1) * -idiv;
    if (!(t >= mint && t <= maxt)) return false;
    beta_o = beta;
    gamma_o = gamma;
    t_out = t;
    return true;
}

// The same test once more, arranged for throughput: the two cheap sign rejections come BEFORE the
// IEEE division.  With div > 0 the reciprocal idiv = RN(1 / div) is positive, so beta = nb * idiv is negative
// exactly when nb is -- UNLESS the product underflows to -0, which the reference's `beta < 0` does not
// reject.  The early exit is therefore taken only where underflow is impossible: |numerator| >= FLT_MIN
// (2^-126) and div <= 2^22 give |numerator * idiv| >= 2^-126 * 2^-22 = 2^-148, a non-zero subnormal of the
// numerator's sign.  Everything else (tiny numerators, huge or infinite div, NaNs) falls through to the
// reference's own order of operations below.  Rejections have no side effects, so their order is free.
// Checked against interTriangle<true> on subnormal numerators and huge divisors in tests/tri_fast_check.cu.
constexpr float kFastDivMax = 4194304.0f;          // 2^22
constexpr float kFastNumMin = 1.17549435e-38f;     // FLT_MIN = 2^-126
// INCL = true: the inclusive range test of A08-A10; false: the exclusive one of A04-A07 (quirk Q9).
template <bool INCL = true>
RT_DEV bool interTriangleFast(f3 o, f3 d, float mint, float maxt, float div, f3 p0, f3 e1, f3 e2, float& beta_o, float& gamma_o, float& t_out) {
    if (div <= 0) return false;
    f3 s = o - p0;
    float nb = dot(cross(s, d), e2);
    float ngm = dot(cross(s, e1), d);
    if ((nb <= -kFastNumMin || ngm <= -kFastNumMin) && div <= kFastDivMax) return false;
    float idiv = 1.0f / div;
    float beta = nb * idiv;
    if (beta < 0.0f || beta > 1.0f) return false;
    float gamma = ngm * idiv;
    if (gamma < 0.0f || (gamma + beta) < 0.0f || (gamma + beta) > 1.0f) return false;
    float t = dot(cross(s, e2), e1) * -idiv;
    if (!(INCL ? (t >= mint && t <= maxt) : (t > mint && t < maxt))) return false;
    beta_o = beta;
    gamma_o = gamma;
    t_out = t;
    return true;
}

// A10/code.cl:391-403 (no t > 0 test -- quirk Q5).
RT_DEV bool interLight(f3 o, f3 d, f3 light_pos, f3 light_normal, float radius, float& t_out) {
    float den = dot(d, light_normal);
    if (den == 0.0f) return false;
    float num = dot((light_pos - o), light_normal);
    if (num == 0.0f) return false;
    float t = num / den;
    f3 poi = getPoint(o, d, t);
    if (distance(poi, light_pos) > radius) return false;
    t_out = t;
    return true;
}

// A10/code.cl:409-411.
RT_DEV f3 interp(float beta, float gamma, f3 v1, f3 v2, f3 v3) {
    return (1.0f - beta - gamma) * v1 + beta * v2 + gamma * v3;
}

// ---------------------------------------------------------------------------------------
// RNG -- A10/code.cl:420-434 (quirk Q6: 32-bit wrapping product, signed %, fabs)
// ---------------------------------------------------------------------------------------
RT_DEV float nextRand(int& seed) {
    int s = (int)((unsigned)seed * 16807u);
    seed = (int)((long long)s % 2147483647LL);
    const float im = 1.0f / 2147483647.0f;
    return fabsf((float)seed * im);
}

// ---------------------------------------------------------------------------------------
// 3D-DDA walk over the uniform grid (A10/code.cl:694-786 and its four textual copies)
// ---------------------------------------------------------------------------------------
struct GridView {
    const float4* prim;      // spheres: 1 float4 per ref; triangles: 3 float4 per ref
    const unsigned* box;     // n^3 + 1 exclusive prefix sums
    const unsigned* occ;     // optional: 1 bit per cell, set when the cell is non-empty (ours, built with the grid)
    AABB bound;
    unsigned n;
};

struct Hit {
    float t;          // champ_t
    unsigned i;       // champ_i (reference index inside this set), 0xFFFFFFFF = none
    float beta, gamma;
    int cx, cy, cz;   // champ_slab (A07 debug colouring)
};

struct WalkStats { unsigned long long cells, tests, front; };   // front: triangle tests that pass the face cull (div > 0) and run the full test

enum PrimKind { PRIM_SPHERE = 0, PRIM_TRIANGLE = 1 };

// Optional statistics sinks of the grid-walk launchers (rt_set_walk_stats / rt_set_walk_totals): per work-item
// champ_i / cells / tests, and launch-spanning totals {alive rays, walks started, cells, tests, hits, front-facing tests}.
struct StatPtrs { unsigned* hit; unsigned* cells; unsigned* tests; unsigned long long* totals; };
RT_DEV void tallyWalk(unsigned long long* t, unsigned walked, const WalkStats& ws, bool hit) {
    if (!t) return;
    const unsigned m = __activemask();   // the lanes that reached this point together
    const bool lead = (int)(threadIdx.x & 31) == __ffs(m) - 1;
    const unsigned v[6] = {1u, walked, (unsigned)ws.cells, (unsigned)ws.tests, hit ? 1u : 0u, (unsigned)ws.front};
#pragma unroll
    for (int q = 0; q < 6; q++) {
        const unsigned r = __reduce_add_sync(m, v[q]);
        if (lead && r) atomicAdd(t + q, (unsigned long long)r);
    }
}

// One axis of the DDA preparation, A10/code.cl:696-707.
struct Axis {
    float t_next, delta_t;
    int slab, step, limit;
};
RT_DEV Axis ddaAxis(float o, float d, float tmin, float pmin, float pmax, unsigned n) {
    Axis a;
    float x = o + tmin * d;
    float delta = (pmax - pmin) / (float)n;
    int slab = (int)((x - pmin) / delta);
    if (slab < 0) slab = 0;
    if ((unsigned)slab >= n) slab = (int)n - 1;
    a.step = (d >= 0) ? 1 : -1;
    a.limit = (d >= 0) ? (int)n : -1;
    a.delta_t = delta / fabsf(d);
    float nxt = pmin + (float)(slab + ((d >= 0) ? 1 : 0)) * delta;
    a.t_next = (nxt - o) / d;
    a.slab = slab;
    return a;
}

// Resumable 3D-DDA walker: the state of the reference's `while(true)` loop (A10/code.cl:745-786)
// so that the loop can be run either to completion by one thread (gridWalk) or one cell at a
// time by a persistent warp that refills idle lanes from a queue (rt_wavefront.cu).  Both use
// walkCell, hence produce identical floats.
struct Walker {
    f3 o, d;
    float a_dd;          // dot(d,d), spheres only
    Axis ax, ay, az;
    float t;             // entry parameter of the current cell
    float tmax_b;        // binter.tmax
    Hit h;
};

RT_DEV void walkInit(Walker& w, int prim, f3 o, f3 d, float maxt_in, const GridView& g, const AabbHit& binter) {
    w.o = o; w.d = d;
    w.ax = ddaAxis(o.x, d.x, binter.tmin, g.bound.pmin.x, g.bound.pmax.x, g.n);
    w.ay = ddaAxis(o.y, d.y, binter.tmin, g.bound.pmin.y, g.bound.pmax.y, g.n);
    w.az = ddaAxis(o.z, d.z, binter.tmin, g.bound.pmin.z, g.bound.pmax.z, g.n);
    w.h.t = maxt_in;     // champ_t starts at the STORED ray.maxt
    w.h.i = 0xFFFFFFFFu;
    w.h.beta = 0.f; w.h.gamma = 0.f;
    w.h.cx = w.h.cy = w.h.cz = (int)g.n;
    w.t = binter.tmin;
    w.tmax_b = binter.tmax;
    w.a_dd = (prim == PRIM_SPHERE) ? dot(d, d) : 0.f;
}

// One iteration of the reference loop: test the current cell's references, then advance.
// Returns true when the walk is over (a hit was accepted in this cell, or the ray left the grid).
// OCC = true consults the occupancy bitmap first and touches the cell table only for
// non-empty cells -- an empty cell contributes nothing to the reference's loop either, so the
// walk (and every float it produces) is unchanged.
template <int PRIM, bool ANY, bool TRI_INCL, bool STATS, bool OCC>
RT_DEV bool walkCell(Walker& w, const GridView& g, WalkStats* st) {
    float mint = w.t;
    float maxt = cl_min(cl_min(w.ax.t_next, w.ay.t_next), w.az.t_next);
    unsigned cell = (unsigned)w.az.slab * (g.n * g.n) + (unsigned)w.ay.slab * g.n + (unsigned)w.ax.slab;
    unsigned begin = 0, end = 0;
    if (!OCC || ((__ldg(g.occ + (cell >> 5)) >> (cell & 31)) & 1u)) {
        begin = __ldg(g.box + cell);
        end = __ldg(g.box + cell + 1);
    }
    if (STATS) st->cells++;
    for (unsigned i = begin; i < end; i++) {
        float ti;
        bool v;
        float be = 0.f, ga = 0.f;
        if (PRIM == PRIM_SPHERE) {
            float4 s = __ldg(g.prim + i);
            v = interSphere(w.o, w.d, w.a_dd, mint, maxt, s, ti);
        } else {
            float4 q0 = __ldg(g.prim + 3 * i), q1 = __ldg(g.prim + 3 * i + 1), q2 = __ldg(g.prim + 3 * i + 2);
            v = interTriangle<TRI_INCL>(w.o, w.d, mint, maxt, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z),
                                        mk3(q2.x, q2.y, q2.z), be, ga, ti);
        }
        if (STATS) {
            st->tests++;
            if (PRIM == PRIM_TRIANGLE) {   // instrumentation only: the reference's first rejection (A10/code.cl:257-261)
                float4 a0 = __ldg(g.prim + 3 * i), a1 = __ldg(g.prim + 3 * i + 1), a2 = __ldg(g.prim + 3 * i + 2);
                f3 c0 = mk3(a0.x, a0.y, a0.z);
                if (dot(cross(mk3(a2.x, a2.y, a2.z) - c0, mk3(a1.x, a1.y, a1.z) - c0), w.d) > 0) st->front++;
            }
        }
        if (v && ti < w.h.t) {
            w.h.t = ti;
            w.h.i = i;
            w.h.beta = be;
            w.h.gamma = ga;
            w.h.cx = w.ax.slab; w.h.cy = w.ay.slab; w.h.cz = w.az.slab;
            if (ANY) break;
        }
    }
    if (w.h.i != 0xFFFFFFFFu) return true;
    float t = maxt;
    w.t = t;
    if (t == w.ax.t_next) {
        w.ax.t_next += w.ax.delta_t;
        if (t >= w.tmax_b) return true;
        w.ax.slab += w.ax.step;
        if (w.ax.slab == w.ax.limit) return true;
    } else if (t == w.ay.t_next) {
        w.ay.t_next += w.ay.delta_t;
        if (t >= w.tmax_b) return true;
        w.ay.slab += w.ay.step;
        if (w.ay.slab == w.ay.limit) return true;
    } else {
        w.az.t_next += w.az.delta_t;
        if (t >= w.tmax_b) return true;
        w.az.slab += w.az.step;
        if (w.az.slab == w.az.limit) return true;
    }
    return false;
}

// RT_SPEC_BOX = 1 fetches the occupancy word and the two cell-table entries of a cell together instead of one after
// the other.  Measured on B200 at the full config: 6085 vs 6076 Mrays/s (noise), so it stays off.
#ifndef RT_SPEC_BOX
#define RT_SPEC_BOX 0
#endif
// ---- flattened form of the same loop, for the queue walkers --------------------------------
// The unit of work is ONE primitive test or ONE cell change, so that lanes of a warp that sit in
// cells of very different population (0 .. 100+ references) all make progress every iteration.
// FlatWalker = Walker + the current cell's [mint,maxt] and reference cursor.
struct FlatWalker {
    Walker w;
    float mint, maxt;     // parameter interval of the current cell
    unsigned i, end;      // next reference to test, end of the cell's list
};

// Enter the cell the walker stands in: its parameter interval and reference range
// (A10/code.cl:747-752; occupancy bitmap first, see walkCell).
RT_DEV void flatEnter(FlatWalker& f, const GridView& g) {
    Walker& w = f.w;
    f.mint = w.t;
    f.maxt = cl_min(cl_min(w.ax.t_next, w.ay.t_next), w.az.t_next);
    unsigned cell = (unsigned)w.az.slab * (g.n * g.n) + (unsigned)w.ay.slab * g.n + (unsigned)w.ax.slab;
    unsigned begin = 0, end = 0;
    if ((__ldg(g.occ + (cell >> 5)) >> (cell & 31)) & 1u) {
        begin = __ldg(g.box + cell);
        end = __ldg(g.box + cell + 1);
    }
    f.i = begin;
    f.end = end;
}

// Same, consulting a coarse occupancy bitmap held in shared memory first (1 bit per
// (2^shift)^3 cells): most cells a ray crosses lie in empty coarse blocks and cost no global load.
RT_DEV void flatEnterMacro(FlatWalker& f, const GridView& g, const unsigned* s_macro, unsigned shift, unsigned nm) {
    Walker& w = f.w;
    f.mint = w.t;
    f.maxt = cl_min(cl_min(w.ax.t_next, w.ay.t_next), w.az.t_next);
    unsigned begin = 0, end = 0;
    unsigned mc = ((unsigned)w.az.slab >> shift) * (nm * nm) + ((unsigned)w.ay.slab >> shift) * nm + ((unsigned)w.ax.slab >> shift);
    if ((s_macro[mc >> 5] >> (mc & 31)) & 1u) {
        unsigned cell = (unsigned)w.az.slab * (g.n * g.n) + (unsigned)w.ay.slab * g.n + (unsigned)w.ax.slab;
#if RT_SPEC_BOX
        // the occupancy word and the two cell-table entries are fetched together instead of one after the other
        // (about 4 in 10 cells of an occupied coarse block hold references; the walkers wait on latency, not bandwidth)
        // (asm volatile: the compiler would otherwise sink the table loads back under the occupancy test)
        unsigned ow, b0, b1;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(ow) : "l"(g.occ + (cell >> 5)));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(b0) : "l"(g.box + cell));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(b1) : "l"(g.box + cell + 1));
        if ((ow >> (cell & 31)) & 1u) { begin = b0; end = b1; }
#else
        if ((ldKeep(g.occ + (cell >> 5)) >> (cell & 31)) & 1u) {
            begin = ldKeep(g.box + cell);
            end = ldKeep(g.box + cell + 1);
        }
#endif
    }
    f.i = begin;
    f.end = end;
}

// Leave the current cell (A10/code.cl:766-785).  Returns true when the walk is over.  Written without the
// reference's three-way branch (lanes of a warp step along different axes and would serialise it): the axis
// is chosen with the same float equalities in the same x, y, z priority; t_next is advanced before the
// `t >= tmax` test exactly as in the reference (the slab update of a finished walk is dead either way).
RT_DEV bool flatLeave(FlatWalker& f) {
    Walker& w = f.w;
    if (w.h.i != 0xFFFFFFFFu) return true;
    const float t = f.maxt;
    w.t = t;
    const bool bx = (t == w.ax.t_next);
    const bool by = !bx && (t == w.ay.t_next);
    const bool bz = !bx && !by;
    if (bx) { w.ax.t_next += w.ax.delta_t; w.ax.slab += w.ax.step; }
    if (by) { w.ay.t_next += w.ay.delta_t; w.ay.slab += w.ay.step; }
    if (bz) { w.az.t_next += w.az.delta_t; w.az.slab += w.az.step; }
    const bool out = (bx && w.ax.slab == w.ax.limit) || (by && w.ay.slab == w.ay.limit) || (bz && w.az.slab == w.az.limit);
    return (t >= w.tmax_b) || out;
}

// Leave the current cell, or -- when `skip` says the cell lies in an EMPTY aligned 2x2x2 block of cells -- the
// whole block, in one go.  Exact: it produces the state the reference's cell-by-cell loop (A10/code.cl:766-785)
// has after the same steps.  Each axis' t_next sequence depends only on that axis (t_next += delta_t per step),
// and the loop executes the pending steps of the three axes in lexicographic order of (t_next value, axis x<y<z):
// `t = min(...)`, then `t == tx_next` first, `t == ty_next` second.  So: the block is left by the axis whose
// boundary-crossing step (its 1st or 2nd step from here: r = 1 or 2) comes first in that order; every other axis
// takes its first step iff that step precedes the crossing step in the same order (it then has r = 2 and stays
// inside the block).  All cells passed on the way are empty, the `t >= tmax` test of a skipped step can only
// fire when it also fires for the crossing step (its t is larger), and with an even n a skipped step never
// reaches the grid limit.  With skip = false every r is 1 and this IS the reference step.
// Callers pass skip = true only for rays whose t_next / delta_t are all finite (no zero direction component).
RT_DEV bool flatAdvance(FlatWalker& f, bool skip) {
    Walker& w = f.w;
    if (w.h.i != 0xFFFFFFFFu) return true;
    const bool two_x = skip && (((w.ax.slab & 1) != 0) == (w.ax.step < 0));   // two steps to the block face
    const bool two_y = skip && (((w.ay.slab & 1) != 0) == (w.ay.step < 0));
    const bool two_z = skip && (((w.az.slab & 1) != 0) == (w.az.step < 0));
    const float Tx = two_x ? w.ax.t_next + w.ax.delta_t : w.ax.t_next;         // t of the crossing step
    const float Ty = two_y ? w.ay.t_next + w.ay.delta_t : w.ay.t_next;
    const float Tz = two_z ? w.az.t_next + w.az.delta_t : w.az.t_next;
    const float t = cl_min(cl_min(Tx, Ty), Tz);
    w.t = t;
    const bool ex = (t == Tx);
    const bool ey = !ex && (t == Ty);
    const bool ez = !ex && !ey;
    // first steps of the non-crossing axes that come before the crossing step (only possible when they have two)
    const bool kx = !ex && two_x && (w.ax.t_next <= t);                         // x precedes y and z on ties
    const bool ky = !ey && two_y && (ex ? (w.ay.t_next < t) : (w.ay.t_next <= t));
    const bool kz = !ez && two_z && (w.az.t_next < t);
    if (ex) { w.ax.t_next = Tx + w.ax.delta_t; w.ax.slab += two_x ? 2 * w.ax.step : w.ax.step; }
    if (ey) { w.ay.t_next = Ty + w.ay.delta_t; w.ay.slab += two_y ? 2 * w.ay.step : w.ay.step; }
    if (ez) { w.az.t_next = Tz + w.az.delta_t; w.az.slab += two_z ? 2 * w.az.step : w.az.step; }
    if (kx) { w.ax.t_next += w.ax.delta_t; w.ax.slab += w.ax.step; }
    if (ky) { w.ay.t_next += w.ay.delta_t; w.ay.slab += w.ay.step; }
    if (kz) { w.az.t_next += w.az.delta_t; w.az.slab += w.az.step; }
    const bool out = (ex && w.ax.slab == w.ax.limit) || (ey && w.ay.slab == w.ay.limit) || (ez && w.az.slab == w.az.limit);
    return (t >= w.tmax_b) || out;
}

// Enter the cell the walker stands in; returns true when its coarse block is empty (nothing loaded, the
// cell interval is not needed: flatAdvance works from t_next).  `coarse2` = the coarse bitmap has 2x2x2 blocks.
RT_DEV bool flatEnterSkip(FlatWalker& f, const GridView& g, const unsigned* s_macro, unsigned shift, unsigned nm) {
    Walker& w = f.w;
    unsigned mc = ((unsigned)w.az.slab >> shift) * (nm * nm) + ((unsigned)w.ay.slab >> shift) * nm + ((unsigned)w.ax.slab >> shift);
    f.i = 0;
    f.end = 0;
    if (!((s_macro[mc >> 5] >> (mc & 31)) & 1u)) return true;
    f.mint = w.t;
    f.maxt = cl_min(cl_min(w.ax.t_next, w.ay.t_next), w.az.t_next);
    unsigned cell = (unsigned)w.az.slab * (g.n * g.n) + (unsigned)w.ay.slab * g.n + (unsigned)w.ax.slab;
    if ((__ldg(g.occ + (cell >> 5)) >> (cell & 31)) & 1u) {
        f.i = __ldg(g.box + cell);
        f.end = __ldg(g.box + cell + 1);
    }
    return false;
}

// Test reference f.i of the current cell (A10/code.cl:882-897) and advance the cursor.
template <int PRIM, bool ANY>
RT_DEV void flatTest(FlatWalker& f, const GridView& g) {
    Walker& w = f.w;
    unsigned i = f.i;
    float ti, be = 0.f, ga = 0.f;
    bool v;
    if (PRIM == PRIM_SPHERE) {
        float4 s = __ldg(g.prim + i);
        v = interSphere(w.o, w.d, w.a_dd, f.mint, f.maxt, s, ti);
    } else {
        float4 q0 = __ldg(g.prim + 3 * i), q1 = __ldg(g.prim + 3 * i + 1), q2 = __ldg(g.prim + 3 * i + 2);
        v = interTriangle<true>(w.o, w.d, f.mint, f.maxt, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), be, ga, ti);
    }
    f.i = i + 1;
    if (v && ti < w.h.t) {
        w.h.t = ti;
        w.h.i = i;
        w.h.beta = be;
        w.h.gamma = ga;
        if (ANY) f.i = f.end;   // `break` of the any-hit kernels
    }
}

// Closest-hit (ANY = false) or any-hit (ANY = true) walk run to completion.  `maxt_in` is the
// STORED ray.maxt.  Returns hit.i != ~0u when a primitive was accepted.
template <int PRIM, bool ANY, bool TRI_INCL, bool STATS, bool OCC = false>
RT_DEV Hit gridWalk(f3 o, f3 d, float maxt_in, const GridView& g, const AabbHit& binter, WalkStats* st) {
    Walker w;
    walkInit(w, PRIM, o, d, maxt_in, g, binter);
    while (!walkCell<PRIM, ANY, TRI_INCL, STATS, OCC>(w, g, st)) {}
    return w.h;
}

// Walk over a ONE-cell grid (n_slabs == 1: every XML sphere/triangle set of Assignment 10, whose global
// n_slabs is 1, A10/code.js:399).  Same floats as gridWalk: with n = 1 the entry slab clamps to 0 whatever
// (int)((x - pmin) / delta) is, the cell interval is [binter.tmin, min(t_next)] with
// t_next = (pmin + (d >= 0 ? 1 : 0) * delta - o) / d, delta = (pmax - pmin) / 1.0f, and after the cell the
// first step always reaches slab == limit (A10/code.cl:770-785), so the loop body runs exactly once.
// Triangles come in the precomputed form (face vector, p0, e1, e2 -- f_precomputeTriangles).
RT_DEV void cellExit(const AabbHit&, bool, float&, float&, float&) {}
RT_DEV void cellExit(const AabbFar& b, bool far_ok, float& tx, float& ty, float& tz) {
    if (far_ok) { tx = b.fx; ty = b.fy; tz = b.fz; }
}
template <int PRIM, bool ANY, bool STATS, typename BI>
RT_DEV Hit singleCellWalk(f3 o, f3 d, float maxt_in, const GridView& g, const float4* __restrict__ pre_ng, const float4* __restrict__ pre_pe,
                          const BI& binter, WalkStats* st, bool far_ok = false) {
    Hit h;
    h.t = maxt_in;
    h.i = 0xFFFFFFFFu;
    h.beta = 0.f; h.gamma = 0.f;
    h.cx = h.cy = h.cz = 1;
    float tx, ty, tz;
    if (!far_ok) {
        const float dx = (g.bound.pmax.x - g.bound.pmin.x) / 1.0f, dy = (g.bound.pmax.y - g.bound.pmin.y) / 1.0f,
                    dz = (g.bound.pmax.z - g.bound.pmin.z) / 1.0f;
        tx = ((g.bound.pmin.x + (float)((d.x >= 0) ? 1 : 0) * dx) - o.x) / d.x;
        ty = ((g.bound.pmin.y + (float)((d.y >= 0) ? 1 : 0) * dy) - o.y) / d.y;
        tz = ((g.bound.pmin.z + (float)((d.z >= 0) ? 1 : 0) * dz) - o.z) / d.z;
    }
    cellExit(binter, far_ok, tx, ty, tz);   // the box's own far-plane quotients when they are the same numbers (farPlanesShared)
    const float mint = binter.tmin;
    const float maxt = cl_min(cl_min(tx, ty), tz);
    const unsigned begin = __ldg(g.box), end = __ldg(g.box + 1);
    const float a_dd = (PRIM == PRIM_SPHERE) ? dot(d, d) : 0.f;
    if (STATS) st->cells++;
    if (PRIM == PRIM_SPHERE) {
        for (unsigned i = begin; i < end; i++) {
            float ti;
            bool v = interSphere(o, d, a_dd, mint, maxt, __ldg(g.prim + i), ti);
            if (STATS) st->tests++;
            if (v && ti < h.t) {
                h.t = ti;
                h.i = i;
                h.cx = h.cy = h.cz = 0;
                if (ANY) break;
            }
        }
    } else {
    // Triangles, 32 references at a time, in three passes so that the lanes of a warp (which all loop over the
    // SAME references but reject them at different points) do not serialise each other's rejections:
    //   1. face-vector cull for every reference (uniform, cheap)            -> bit mask of front-facing ones
    //   2. the two barycentric numerators for the survivors (each lane walks ITS OWN set bits) -> mask of
    //      references that survive interTriangleFast's sign rejections (numerator <= -2^-126 with div <= 2^22)
    //   3. the division and the range tests for what is left, in ascending reference order, applying the
    //      reference's `inter.v && inter.t < champ_t` update (and the any-hit break) exactly as the plain loop does.
    // Rejections have no side effects, so the outcome (winner and floats) is that of the sequential loop.
    for (unsigned base = begin; base < end; base += 32) {
        const unsigned cnt = min(32u, end - base);
        unsigned m1 = 0, minf = 0;
        for (unsigned j = 0; j < cnt; j++) {
            float4 q = __ldg(pre_ng + base + j);
            float dv = dot(mk3(q.x, q.y, q.z), d);
            if (dv > 0) m1 |= 1u << j;
            if (dv > kFastDivMax) minf |= 1u << j;   // interTriangleFast does not apply the sign rejections then
        }
        unsigned m2 = minf;
        for (unsigned m = m1 & ~minf; m; m &= m - 1) {
            const unsigned j = __ffs(m) - 1, r = base + j;
            float4 q0 = __ldg(pre_pe + 3 * r), q1 = __ldg(pre_pe + 3 * r + 1), q2 = __ldg(pre_pe + 3 * r + 2);
            f3 s = o - mk3(q0.x, q0.y, q0.z);
            float nb = dot(cross(s, d), mk3(q2.x, q2.y, q2.z));
            float ngm = dot(cross(s, mk3(q1.x, q1.y, q1.z)), d);
            if (!(nb <= -kFastNumMin || ngm <= -kFastNumMin)) m2 |= 1u << j;
        }
        bool stop = false;
        for (unsigned m = m2; m; m &= m - 1) {
            const unsigned j = __ffs(m) - 1, r = base + j;
            float4 q = __ldg(pre_ng + r);
            float div = dot(mk3(q.x, q.y, q.z), d);
            float4 q0 = __ldg(pre_pe + 3 * r), q1 = __ldg(pre_pe + 3 * r + 1), q2 = __ldg(pre_pe + 3 * r + 2);
            float ti, be, ga;
            if (interTriangleFast(o, d, mint, maxt, div, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), be, ga, ti) && ti < h.t) {
                h.t = ti;
                h.i = r;
                h.beta = be;
                h.gamma = ga;
                h.cx = h.cy = h.cz = 0;
                if (ANY) {
                    if (STATS) { st->tests += j + 1; st->front += __popc(m1 & (0xFFFFFFFFu >> (31 - j))); }   // the reference's loop broke here
                    stop = true;
                    break;
                }
            }
        }
        if (stop) break;
        if (STATS) { st->tests += cnt; st->front += __popc(m1); }
    }
    }
    return h;
}

// ---------------------------------------------------------------------------------------
// A10 shading helpers
// ---------------------------------------------------------------------------------------
// getHemisphereRay, A10/code.cl:545-579.  Draws s.x then s.y.
RT_DEV void getHemisphereRay(f3 p, f3 normal, int& seed, f3& o, f3& d) {
    f3 N = mk3(fabsf(normal.x), fabsf(normal.y), fabsf(normal.z));
    f3 B = normal;
    float nmin = cl_min(cl_min(N.x, N.y), N.z);
    if (N.x == nmin) B.x = 1.0f;
    else if (N.y == nmin) B.y = 1.0f;
    else B.z = 1.0f;
    N = normal;
    B = normalize(B);
    f3 T = cross(B, N);
    B = cross(N, T);
    f2 sxy;
    sxy.x = nextRand(seed);
    sxy.y = nextRand(seed);
    sxy = concentric_distort(sxy);
    float sz = sqrtf(cl_max(0.0f, 1.0f - sxy.x * sxy.x - sxy.y * sxy.y));
    o = p;
    d = normalize(sxy.x * T + sxy.y * B + sz * N);
}

// initShadowTrace body for a live hit record, A10/code.cl:652-672.  Draws x then y.
RT_DEV RayR makeShadowRay(f3 p, f3 normal, const LightArg& L, int& seed) {
    f3 light_pos = mk3(L.v[0], L.v[1], L.v[2]);
    f3 T = mk3(L.v[3], L.v[4], L.v[5]);
    f3 B = mk3(L.v[6], L.v[7], L.v[8]);
    float light_radius = L.v[9];
    p = p + normal * 0.001f;
    f2 xy;
    xy.x = nextRand(seed);
    xy.y = nextRand(seed);
    xy = concentric_distort(xy);
    xy.x = xy.x * light_radius;
    xy.y = xy.y * light_radius;
    light_pos = light_pos + (xy.x * T + xy.y * B);
    return makeRay(p, light_pos);
}

// sceneRender shade term, A10/code.cl:1343-1357.  `lit` = shadow.maxt != shadow.mint.
RT_DEV f3 neeShade(f3 p, f3 normal, f3 shadow_d, bool lit, const LightArg& L) {
    f3 shade = mk3(0.0f, 0.0f, 0.0f);
    if (lit) {
        f3 lpos = mk3(L.v[0], L.v[1], L.v[2]);
        f3 lnor = mk3(L.v[3], L.v[4], L.v[5]);
        f3 es = mk3(L.v[6], L.v[7], L.v[8]);
        float area = L.v[9];
        float r = distance(p, lpos);
        float cosx = cl_clamp(dot(shadow_d, normal), 0.0f, 1.0f);
        float cosy = cl_clamp(dot(-shadow_d, lnor), 0.0f, 1.0f);
        shade = area * ((cosx * cosy) / (r * r)) * es;
    }
    return shade;
}

}  // namespace rt
