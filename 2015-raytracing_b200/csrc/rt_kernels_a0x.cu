// rt_kernels_a0x.cu -- the kernels of the earlier assignments that BASELINE.json's configs 1-4
// name (A01 = Assign01-Sphere_Ray_Tracing, A02, A03, A07, A08, A09; paths relative to
// /root/reference), one CUDA kernel + C-ABI launcher per OpenCL kernel, same names, argument
// order and buffer layouts.  The grid walks reuse rt_device.cuh (the DDA text is identical in
// A07-A10 apart from the inclusive/exclusive range test of interTriangle, quirk Q9).
//
// float -> uchar conversions of out-of-range values are undefined in OpenCL C (quirk Q11: A01's
// depth shade, A02's unclamped shade, A08/A09's unclamped copyToPixel).  We pick the behaviour
// of the oracle build (x86: truncate to int32, keep the low byte) so images compare exactly;
// the float buffers are the authoritative comparison.
#include "rt_device.cuh"
#include "rt_frame.h"
#include "rt_internal.h"

using namespace rt;

namespace {

constexpr unsigned kBlock = 256;

RT_DEV unsigned char f2uc(float f) { return (unsigned char)(int)f; }
RT_DEV uchar4 mkPixel(float r, float g, float b) { return make_uchar4(f2uc(r), f2uc(g), f2uc(b), 255); }

AabbArg mkAabb(const float* b) { AabbArg a; memcpy(a.v, b, sizeof a.v); return a; }
CamArg mkCam(const float* c) { CamArg a; memcpy(a.v, c, sizeof a.v); return a; }
GridView mkGrid(const void* prim, const void* box, const float* bound, unsigned n, const unsigned* occ = nullptr) {
    GridView g;
    g.prim = (const float4*)prim;
    g.box = (const unsigned*)box;
    g.occ = n > 1 ? occ : nullptr;   // occupancy bitmap of a grid built by rt_grid_build_* (rt_occupancy_of), else none
    g.bound.pmin = f3{bound[0], bound[1], bound[2]};
    g.bound.pmax = f3{bound[4], bound[5], bound[6]};
    g.n = n;
    return g;
}

// ---------------------------------------------------------------------------------------
// A01 -- one hard-coded sphere (A01/code.cl:50-61, 63-111, 116-147).  The camera keeps rows and
// cols as FLOATS and packs rows before cols (sE, sF).
// ---------------------------------------------------------------------------------------
__global__ void k_a01_raytrace(uchar4* pixels, CamArg fcam, unsigned cols, unsigned rows) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= cols * rows) return;
    unsigned col = id % cols, row = id / cols;
    f3 eye = mk3(fcam.v[0], fcam.v[1], fcam.v[2]), U = mk3(fcam.v[3], fcam.v[4], fcam.v[5]);
    f3 V = mk3(fcam.v[6], fcam.v[7], fcam.v[8]), W = mk3(fcam.v[9], fcam.v[10], fcam.v[11]);
    float width = fcam.v[12], height = fcam.v[13], frows = fcam.v[14], fcols = fcam.v[15];
    f3 cop = (-0.5f + ((float)col + 0.5f) / fcols) * width * U + (0.5f - ((float)row + 0.5f) / frows) * height * V + (-1.0f) * W;
    f3 o = eye;
    f3 d = normalize(cop - o);
    // interSphere, A01/code.cl:63-111: c = (0,0,1), r = 0.5; `/ 2*a` is (x/2)*a; exclusive range test, t0 before t1
    f3 sc = mk3(0.0f, 0.0f, 1.0f);
    float sr = 0.5f;
    float a = dot(d, d);
    float b = 2.0f * dot(o - sc, d);
    float c = dot(o - sc, o - sc) - sr * sr;
    float dis = b * b - 4.0f * a * c;
    uchar4 color = make_uchar4(0, 0, 0, 255);
    if (!(dis < 0.0f)) {
        float sq = sqrtf(dis);
        float t0 = (-b - sq) / 2 * a;
        float t1 = (-b + sq) / 2 * a;
        float t = 0.f;
        bool v = false;
        if (t0 > 0.0f && t0 < RT_INF) { t = t0; v = true; }
        else if (t1 > 0.0f && t1 < RT_INF) { t = t1; v = true; }
        if (v) {
            unsigned char base = f2uc((1.0f - t) * 255.0f);
            color = make_uchar4(base, base, base, 255);
        }
    }
    pixels[(unsigned)fcols * row + col] = color;
}

// ---------------------------------------------------------------------------------------
// A02 / A03 -- brute-force closest sphere (A02/code.cl:158-232, A03/code.cl:128-187).
// Sphere records are float4 (c.xyz, RADIUS); interSphere squares the radius per test, uses
// `/ 2*a`, fmin/fmax and an EXCLUSIVE range test (A02/code.cl:92-140).
//
// B200 mapping: the N-sphere loop is FP32-issue bound and every thread of a block reads the same
// record, so records are staged through shared memory in tiles (one coalesced 16 B load per
// thread per tile, then broadcast reads); dot(d,d) and -b's ray part are loop invariants.
// ---------------------------------------------------------------------------------------
constexpr unsigned kSphereTile = 1024;   // 16 KB of shared memory

struct BruteHit { float t; unsigned i; };

RT_DEV bool interSphereA02(f3 o, f3 d, float a, float mint, float maxt, float4 s, float& t_out) {
    f3 omc = o - mk3(s.x, s.y, s.z);
    float b = 2.0f * dot(omc, d);
    float c = dot(omc, omc) - s.w;   // s.w = r*r, squared once when the tile was staged (same rounding)
    float dis = b * b - 4.0f * a * c;
    if (dis < 0.0f) return false;
    float sq = sqrtf(dis);
    float t0 = (-b - sq) / 2 * a;
    float t1 = (-b + sq) / 2 * a;
    float tmin = fminf(t0, t1), tmax = fmaxf(t0, t1);
    if (tmin > mint && tmin < maxt) { t_out = tmin; return true; }
    if (tmax > mint && tmax < maxt) { t_out = tmax; return true; }
    return false;
}

// Sphere loop on Blackwell's packed fp32 pipe (RT_PACKED_F32X2, default on).  FADD2 / FMUL2 (PTX add.f32x2 / mul.f32x2,
// sm_100+) perform two independent round-to-nearest operations per lane and issue slot, so one iteration evaluates the
// discriminant of TWO consecutive spheres (multiplications and the sums of non-products packed): tiles are staged as pairs, component-wise and pre-negated -- (-cx_j, -cx_j+1),
// ..., (-r^2_j, -r^2_j+1) -- and the subtraction chain of interSphere becomes additions of exact negatives
// (a - b == a + (-b), 2*x == x + x, -(p*q) == (-p)*q in IEEE arithmetic), so every discriminant has the bits of the scalar
// form.  A sphere whose discriminant is not negative (rare) goes through the scalar test, in index order.
#ifndef RT_PACKED_F32X2
#define RT_PACKED_F32X2 1
#endif
RT_DEV BruteHit bruteForcePacked(bool live, f3 o, f3 d, float mint, float maxt, unsigned s_size, const float4* __restrict__ s_atoms, float4* tile) {
    BruteHit h;
    h.t = RT_INF;
    h.i = s_size;
    const float a = dot(d, d);
    const float a4n = -(4.0f * a);
    // pair q of the tile = two float4: (-cx_j, -cx_j+1, -cy_j, -cy_j+1) and (-cz_j, -cz_j+1, -r2_j, -r2_j+1), j = 2q: two LDS.128 per pair
    float* tf = reinterpret_cast<float*>(tile);
    const float2 ox = make_float2(o.x, o.x), oy = make_float2(o.y, o.y), oz = make_float2(o.z, o.z);
    const float2 dx = make_float2(d.x, d.x), dy = make_float2(d.y, d.y), dz = make_float2(d.z, d.z);
    const float2 a4 = make_float2(a4n, a4n);
    for (unsigned base = 0; base < s_size; base += kSphereTile) {
        const unsigned n = min(kSphereTile, s_size - base);
        const unsigned n2 = (n + 1u) & ~1u;
        __syncthreads();
        for (unsigned j = threadIdx.x; j < n2; j += blockDim.x) {
            float* slot = tf + 8 * (j >> 1) + (j & 1);
            if (j < n) {
                float4 s = __ldg(s_atoms + base + j);
                slot[0] = -s.x; slot[2] = -s.y; slot[4] = -s.z; slot[6] = -(s.w * s.w);
            } else {   // padding of an odd tile: c = +inf makes the discriminant -inf, never a hit
                slot[0] = 0.f; slot[2] = 0.f; slot[4] = 0.f; slot[6] = RT_INF;
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll 2
            for (unsigned q = 0; q < n2 / 2; q++) {
                const float4 A = tile[2 * q], B = tile[2 * q + 1];
                const float2 pxq = make_float2(A.x, A.y), pyq = make_float2(A.z, A.w), pzq = make_float2(B.x, B.y), pwq = make_float2(B.z, B.w);
                // Products are added with SCALAR adds: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
                // --fmad=false (it keeps scalar mul.rn / add.rn apart), and a fused product would change the rounding.  Sums
                // of non-products (o - c, dt + dt, mm - r*r) stay packed.
                const float2 mx = __fadd2_rn(ox, pxq), my = __fadd2_rn(oy, pyq), mz = __fadd2_rn(oz, pzq);   // o - c
                const float2 p0 = __fmul2_rn(mx, dx), p1 = __fmul2_rn(my, dy), p2 = __fmul2_rn(mz, dz);
                const float2 dt = make_float2(__fadd_rn(__fadd_rn(p0.x, p1.x), p2.x), __fadd_rn(__fadd_rn(p0.y, p1.y), p2.y));
                const float2 b = __fadd2_rn(dt, dt);                                                               // 2 * dot(omc, d)
                const float2 s0 = __fmul2_rn(mx, mx), s1 = __fmul2_rn(my, my), s2 = __fmul2_rn(mz, mz);
                const float2 mm = make_float2(__fadd_rn(__fadd_rn(s0.x, s1.x), s2.x), __fadd_rn(__fadd_rn(s0.y, s1.y), s2.y));
                const float2 c = __fadd2_rn(mm, pwq);                                                              // dot(omc, omc) - r*r
                const float2 bb = __fmul2_rn(b, b), ac = __fmul2_rn(a4, c);
                const float2 dis = make_float2(__fadd_rn(bb.x, ac.x), __fadd_rn(bb.y, ac.y));                      // b*b - 4*a*c
                if (!(dis.x < 0.0f)) {
                    float t;
                    const unsigned j = 2 * q;
                    if (interSphereA02(o, d, a, mint, maxt, make_float4(-A.x, -A.z, -B.x, -B.z), t) && t < h.t) { h.t = t; h.i = base + j; }
                }
                if (!(dis.y < 0.0f)) {
                    float t;
                    const unsigned j = 2 * q + 1;
                    if (j < n && interSphereA02(o, d, a, mint, maxt, make_float4(-A.y, -A.w, -B.y, -B.w), t) && t < h.t) { h.t = t; h.i = base + j; }
                }
            }
        }
    }
    return h;
}

// every thread of the block must call this (barriers inside); `live` lanes own a ray; (mint, maxt) is the ray's
// EXCLUSIVE parameter range (0, +inf for a fresh primary ray)
RT_DEV BruteHit bruteForce(bool live, f3 o, f3 d, float mint, float maxt, unsigned s_size, const float4* __restrict__ s_atoms, float4* tile) {
#if RT_PACKED_F32X2
    return bruteForcePacked(live, o, d, mint, maxt, s_size, s_atoms, tile);
#else
    BruteHit h;
    h.t = RT_INF;
    h.i = s_size;
    float a = dot(d, d);
    for (unsigned base = 0; base < s_size; base += kSphereTile) {
        unsigned n = min(kSphereTile, s_size - base);
        __syncthreads();
        for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
            float4 s = __ldg(s_atoms + base + j);
            s.w = s.w * s.w;
            tile[j] = s;
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (unsigned j = 0; j < n; j++) {
                float t;
                if (interSphereA02(o, d, a, mint, maxt, tile[j], t) && t < h.t) { h.t = t; h.i = base + j; }
            }
        }
    }
    return h;
#endif
}

// pinhole ray of A02-A10 (getRay, A02/code.cl:78-90 = A10/code.cl:108-119)
RT_DEV void pinhole(const CamArg& fcam, unsigned id, Camera& cam, unsigned& col, unsigned& row, f3& o, f3& d) {
    cam = floatToCamera(fcam.v);
    col = id % cam.cols;
    row = id / cam.cols;
    getRay(cam, (float)col, (float)row, o, d);
}

template <bool CLAMP_SHADE>
RT_DEV uchar4 shadeSphere(const Camera& cam, f3 o, f3 d, float t, float4 atom, float4 color) {
    f3 ipoint = getPoint(o, d, t);
    float shade = dot(cam.W, normalize(ipoint - mk3(atom.x, atom.y, atom.z)));
    if (CLAMP_SHADE) shade = cl_clamp(shade, 0.0f, 1.0f);
    return mkPixel(color.x * shade * 255.0f, color.y * shade * 255.0f, color.z * shade * 255.0f);
}

__global__ void __launch_bounds__(kBlock) k_a02_raytrace(uchar4* pixels, CamArg fcam, unsigned s_size, const float4* s_atoms,
                                                         const float4* s_colors) {
    __shared__ float4 tile[kSphereTile];
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam;
    unsigned col, row;
    f3 o, d;
    pinhole(fcam, id, cam, col, row, o, d);
    bool live = id < cam.cols * cam.rows;
    BruteHit h = bruteForce(live, o, d, 0.0f, RT_INF, s_size, s_atoms, tile);
    if (!live) return;
    uchar4 color = make_uchar4(0, 0, 0, 255);
    if (h.i < s_size) color = shadeSphere<false>(cam, o, d, h.t, __ldg(s_atoms + h.i), __ldg(s_colors + h.i));
    pixels[id] = color;
}

__global__ void k_a03_initTrace(uchar4* pixels, CamArg fcam, Ray* rays) {   // A03/code.cl:132-143
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam;
    unsigned col, row;
    RayR ray;
    pinhole(fcam, id, cam, col, row, ray.o, ray.d);
    if (id >= cam.cols * cam.rows) return;
    ray.mint = 0.0f;
    ray.maxt = RT_INF;
    storeRay(rays + id, ray);
    pixels[id] = make_uchar4(0, 0, 0, 255);
}

__global__ void __launch_bounds__(kBlock) k_a03_molTrace(uchar4* pixels, CamArg fcam, Ray* rays, unsigned s_size, const float4* s_atoms,
                                                         const float4* s_colors) {   // A03/code.cl:145-187
    __shared__ float4 tile[kSphereTile];
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam = floatToCamera(fcam.v);
    bool live = id < cam.cols * cam.rows;
    RayR ray;
    ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 1.f); ray.mint = 0.f; ray.maxt = RT_INF;
    if (live) ray = loadRay(rays + id);
    BruteHit h = bruteForce(live, ray.o, ray.d, ray.mint, ray.maxt, s_size, s_atoms, tile);
    if (!live || h.i >= s_size) return;
    rays[id].mint = h.t;
    pixels[id] = shadeSphere<true>(cam, ray.o, ray.d, h.t, __ldg(s_atoms + h.i), __ldg(s_colors + h.i));
}

// ---------------------------------------------------------------------------------------
// A04 / A05 -- brute force over spheres AND over a triangle soup, one ray buffer shared by both
// passes (A04/code.cl:204-315, A05/code.cl:304-452).  A05 adds the scene box in initTrace and a
// per-set box test in front of each loop (BOX = true).  Both loops start champ_t at INFINITY and
// test against the STORED (mint, maxt) exclusively, so the second pass only accepts what is
// nearer than the first pass's hit.
//
// B200 mapping of the triangle loop: like the spheres, every thread of a block reads the same
// triangle, so triangles are staged through shared memory in tiles -- and staged in the
// ray-independent form the test starts from (p0, e1 = p1 - p0, e2 = p2 - p0, cross(e2, e1)),
// computed ONCE per tile by the staging thread with the reference's own operations instead of
// once per (ray, triangle).  The per-ray part is interTriangleFast's arrangement (sign rejections
// before the IEEE division, see rt_device.cuh) with the exclusive range test.
// ---------------------------------------------------------------------------------------
constexpr unsigned kTriTile = 256;   // 4 float4 per triangle = 16 KB of shared memory

// interTriangle of A04-A07 (A04/code.cl:156-188): exclusive range test.  A04/A05 also test `gamma > 1`, which
// can only fire together with `gamma + beta > 1` (beta >= 0 there and float addition is monotonic), so the
// one text serves A04-A07.
RT_DEV bool interTriangleExcl(f3 o, f3 d, float mint, float maxt, float div, f3 p0, f3 e1, f3 e2, float& beta_o, float& gamma_o, float& t_out) {
    if (div <= 0) return false;
    f3 s = o - p0;
    float nb = dot(cross(s, d), e2);
    float ngm = dot(cross(s, e1), d);
    if ((nb <= -kFastNumMin || ngm <= -kFastNumMin) && div <= kFastDivMax) return false;   // only where nb * idiv cannot underflow to -0 (rt_device.cuh)
    float idiv = 1.0f / div;
    float beta = nb * idiv;
    if (beta < 0.0f || beta > 1.0f) return false;
    float gamma = ngm * idiv;
    if (gamma < 0.0f || (gamma + beta) < 0.0f || (gamma + beta) > 1.0f) return false;
    float t = dot(cross(s, e2), e1) * -idiv;
    if (!(t > mint && t < maxt)) return false;
    beta_o = beta;
    gamma_o = gamma;
    t_out = t;
    return true;
}

struct TriHit { float t, beta, gamma; unsigned i; };

// every thread of the block must call this (barriers inside)
RT_DEV TriHit bruteForceTriangles(bool live, f3 o, f3 d, float mint, float maxt, unsigned t_size, const float4* __restrict__ t_pos, float4* tile) {
    TriHit h;
    h.t = RT_INF;
    h.i = t_size;
    h.beta = h.gamma = 0.f;
    for (unsigned base = 0; base < t_size; base += kTriTile) {
        unsigned n = min(kTriTile, t_size - base);
        __syncthreads();
        for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
            const float4* q = t_pos + 3ull * (base + j);
            float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            f3 p0 = mk3(q0.x, q0.y, q0.z);
            f3 e1 = mk3(q1.x, q1.y, q1.z) - p0;
            f3 e2 = mk3(q2.x, q2.y, q2.z) - p0;
            f3 ng = cross(e2, e1);
            tile[4 * j] = make_float4(ng.x, ng.y, ng.z, 0.f);
            tile[4 * j + 1] = make_float4(p0.x, p0.y, p0.z, 0.f);
            tile[4 * j + 2] = make_float4(e1.x, e1.y, e1.z, 0.f);
            tile[4 * j + 3] = make_float4(e2.x, e2.y, e2.z, 0.f);
        }
        __syncthreads();
        if (live) {
#pragma unroll 2
            for (unsigned j = 0; j < n; j++) {
                float4 g = tile[4 * j];
                float div = dot(mk3(g.x, g.y, g.z), d);
                if (div <= 0) continue;
                float4 a = tile[4 * j + 1], b = tile[4 * j + 2], c = tile[4 * j + 3];
                float t, be, ga;
                if (interTriangleExcl(o, d, mint, maxt, div, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), be, ga, t) && t < h.t) {
                    h.t = t; h.i = base + j; h.beta = be; h.gamma = ga;
                }
            }
        }
    }
    return h;
}

template <bool BOX>
__global__ void __launch_bounds__(kBlock) k_a045_molTrace(uchar4* pixels, CamArg fcam, Ray* rays, unsigned s_size, const float4* s_atoms,
                                                          const float4* s_colors, AabbArg bound_a) {
    __shared__ float4 tile[kSphereTile];
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam = floatToCamera(fcam.v);
    bool live = id < cam.cols * cam.rows;
    RayR ray;
    ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 1.f); ray.mint = 0.f; ray.maxt = RT_INF;
    if (live) ray = loadRay(rays + id);
    if (BOX && live) live = ray.mint != ray.maxt && interAABB(ray.o, ray.d, toAABB(bound_a)).v;   // A05/code.cl:341-351
    BruteHit h = bruteForce(live, ray.o, ray.d, ray.mint, ray.maxt, s_size, s_atoms, tile);
    if (!live || h.i >= s_size) return;
    rays[id].maxt = h.t;
    pixels[id] = shadeSphere<true>(cam, ray.o, ray.d, h.t, __ldg(s_atoms + h.i), __ldg(s_colors + h.i));
}

template <bool BOX>
__global__ void __launch_bounds__(kBlock) k_a045_meshTrace(uchar4* pixels, CamArg fcam, Ray* rays, unsigned t_size, const float4* t_pos,
                                                           const float4* t_normal, const unsigned* t_mindex, const float4* m_color,
                                                           AabbArg bound_a) {
    __shared__ float4 tile[4 * kTriTile];
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam = floatToCamera(fcam.v);
    bool live = id < cam.cols * cam.rows;
    RayR ray;
    ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 1.f); ray.mint = 0.f; ray.maxt = RT_INF;
    if (live) ray = loadRay(rays + id);
    if (BOX && live) live = ray.mint != ray.maxt && interAABB(ray.o, ray.d, toAABB(bound_a)).v;   // A05/code.cl:401-411
    TriHit h = bruteForceTriangles(live, ray.o, ray.d, ray.mint, ray.maxt, t_size, t_pos, tile);
    if (!live || h.i >= t_size) return;
    rays[id].maxt = h.t;
    float4 n0 = __ldg(t_normal + 3ull * h.i), n1 = __ldg(t_normal + 3ull * h.i + 1), n2 = __ldg(t_normal + 3ull * h.i + 2);
    f3 nrm = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
    float shade = cl_clamp(dot(cam.W, nrm), 0.0f, 1.0f);
    float4 c = __ldg(m_color + __ldg(t_mindex + h.i));
    pixels[id] = mkPixel(c.x * 255.0f * shade, c.y * 255.0f * shade, c.z * 255.0f * shade);   // m_color * 255.0f * shade, A04/code.cl:311
}

// ---------------------------------------------------------------------------------------
// A06 -- 1-D uniform slabs along x (A06/code.cl:336-533): the x axis of the later 3-D DDA.  One
// thread per pixel; the slab lists are short and the walk has at most n_slabs steps.
// Spheres carry the RADIUS (c = mad(-r, r, |o-c|^2), A06/code.cl:117) and use the inclusive
// range test; triangles the exclusive one.  Debug colour = slab number mod 3.
// ---------------------------------------------------------------------------------------
RT_DEV bool interSphereA06(f3 o, f3 d, float a_dd, float mint, float maxt, float4 s, float& t_out) {
    f3 omc = o - mk3(s.x, s.y, s.z);
    float a = a_dd;
    float b = 2.0f * dot(omc, d);
    float c = fmaf(-s.w, s.w, dot(omc, omc));
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

template <int PRIM>
__global__ void k_a06_trace(uchar4* pixels, CamArg fcam, Ray* rays, const float4* prim, const float4* normals, AabbArg bound_a,
                            unsigned n_slabs, const unsigned* __restrict__ slab_size) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam = floatToCamera(fcam.v);
    if (id >= cam.cols * cam.rows) return;
    RayR ray = loadRay(rays + id);
    if (ray.mint == ray.maxt) return;
    AABB bound = toAABB(bound_a);
    AabbHit binter = interAABB(ray.o, ray.d, bound);
    if (!binter.v) return;
    Axis ax = ddaAxis(ray.o.x, ray.d.x, binter.tmin, bound.pmin.x, bound.pmax.x, n_slabs);   // A06/code.cl:357-370
    const float a_dd = (PRIM == PRIM_SPHERE) ? dot(ray.d, ray.d) : 0.f;
    float champ_t = ray.maxt, champ_b = 0.f, champ_g = 0.f;
    unsigned champ_i = 0xFFFFFFFFu;
    int champ_slab = (int)n_slabs;
    float t = binter.tmin;
    while (true) {
        const float mint = t, maxt = ax.t_next;
        const unsigned end = __ldg(slab_size + ax.slab + 1);
        for (unsigned i = __ldg(slab_size + ax.slab); i < end; i++) {
            float ti, be = 0.f, ga = 0.f;
            bool v;
            if (PRIM == PRIM_SPHERE) {
                v = interSphereA06(ray.o, ray.d, a_dd, mint, maxt, __ldg(prim + i), ti);
            } else {
                float4 q0 = __ldg(prim + 3ull * i), q1 = __ldg(prim + 3ull * i + 1), q2 = __ldg(prim + 3ull * i + 2);
                v = interTriangle<false>(ray.o, ray.d, mint, maxt, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), be, ga, ti);
            }
            if (v && ti < champ_t) { champ_t = ti; champ_i = i; champ_b = be; champ_g = ga; champ_slab = ax.slab; }
        }
        if (champ_slab < (int)n_slabs) break;
        t = ax.t_next;
        if (t >= binter.tmax) break;
        ax.t_next += ax.delta_t;
        ax.slab += ax.step;
        if (ax.slab == ax.limit) break;
    }
    if (champ_slab >= (int)n_slabs) return;
    rays[id].maxt = champ_t;
    float shade;
    if (PRIM == PRIM_SPHERE) {
        float4 s = __ldg(prim + champ_i);
        f3 ipoint = getPoint(ray.o, ray.d, champ_t);
        shade = cl_clamp(dot(cam.W, normalize(ipoint - mk3(s.x, s.y, s.z))), 0.0f, 1.0f);
    } else {
        float4 n0 = __ldg(normals + 3ull * champ_i), n1 = __ldg(normals + 3ull * champ_i + 1), n2 = __ldg(normals + 3ull * champ_i + 2);
        f3 nrm = normalize(interp(champ_b, champ_g, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
        shade = cl_clamp(dot(cam.W, nrm), 0.0f, 1.0f);
    }
    float s127 = shade * 127.0f;   // fcolor *= shade * 127.0f, A06/code.cl:420-421
    pixels[id] = mkPixel((float)(champ_slab % 3) * s127, (float)((champ_slab + 1) % 3) * s127, (float)((champ_slab + 2) % 3) * s127);
}

// ---------------------------------------------------------------------------------------
// A07 -- 3-D uniform grid, primary rays, debug colouring by cell parity
// (A07/code.cl:311-335, 337-473, 475-626)
// ---------------------------------------------------------------------------------------
__global__ void k_a07_initTrace(uchar4* pixels, CamArg fcam, Ray* rays, AabbArg bound_a) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam;
    unsigned col, row;
    RayR ray;
    pinhole(fcam, id, cam, col, row, ray.o, ray.d);
    if (id >= cam.cols * cam.rows) return;
    AabbHit inter = interAABB(ray.o, ray.d, toAABB(bound_a));
    if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
    else { ray.mint = RT_INF; ray.maxt = RT_INF; }   // ray.mint = ray.maxt (= HUGE_VALF from getRay)
    storeRay(rays + id, ray);
    pixels[id] = make_uchar4(0, 0, 0, 255);
}

RT_DEV uchar4 cellParityColor(const Hit& h, float shade) {
    float s = shade * 127.0f;
    return mkPixel((float)((h.cx % 2) + 1) * s, (float)((h.cy % 2) + 1) * s, (float)((h.cz % 2) + 1) * s);
}

template <int PRIM, bool OCC, bool STATS>
__global__ void k_a07_trace(uchar4* pixels, CamArg fcam, Ray* rays, GridView g, const float4* normals, StatPtrs sp) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam = floatToCamera(fcam.v);
    if (id >= cam.cols * cam.rows) return;
    if (STATS) {
        if (sp.hit) sp.hit[id] = 0xFFFFFFFFu;
        if (sp.cells) sp.cells[id] = 0;
        if (sp.tests) sp.tests[id] = 0;
    }
    RayR ray = loadRay(rays + id);
    if (ray.mint == ray.maxt) return;
    AabbHit binter = interAABB(ray.o, ray.d, g.bound);
    WalkStats ws = {0, 0, 0};
    if (!binter.v) { if (STATS) tallyWalk(sp.totals, 0, ws, false); return; }
    // spheres: inclusive test; triangles: EXCLUSIVE in A07 (A07/code.cl:195, quirk Q9)
    Hit h = gridWalk<PRIM, false, false, STATS, OCC>(ray.o, ray.d, ray.maxt, g, binter, &ws);
    if (STATS) {
        tallyWalk(sp.totals, 1, ws, h.i != 0xFFFFFFFFu);
        if (sp.hit) sp.hit[id] = h.i;
        if (sp.cells) sp.cells[id] = (unsigned)ws.cells;
        if (sp.tests) sp.tests[id] = (unsigned)ws.tests;
    }
    if (h.i == 0xFFFFFFFFu) return;
    rays[id].maxt = h.t;
    float shade;
    if (PRIM == PRIM_SPHERE) {
        float4 s = __ldg(g.prim + h.i);
        f3 ipoint = getPoint(ray.o, ray.d, h.t);
        shade = cl_clamp(dot(cam.W, normalize(ipoint - mk3(s.x, s.y, s.z))), 0.0f, 1.0f);
    } else {
        float4 n0 = __ldg(normals + 3 * h.i), n1 = __ldg(normals + 3 * h.i + 1), n2 = __ldg(normals + 3 * h.i + 2);
        f3 nrm = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
        shade = cl_clamp(dot(cam.W, nrm), 0.0f, 1.0f);
    }
    pixels[id] = cellParityColor(h, shade);
}

// ---------------------------------------------------------------------------------------
// A08 / A09 -- deterministic shading with point lights (A08/code.cl:331-951, A09/code.cl:400-1040).
// The trace kernels are textually A10's with a 48-byte Poi (no atte); A08 launches them 2-D over
// (cols, rows), A09 1-D over total_rays -- both index slot id = cols*row+col resp. id.
// ---------------------------------------------------------------------------------------
__global__ void k_a08_initTrace(float4* acu, Ray* rays, Poi8* pois, AabbArg bound_a, CamArg fcam) {   // A08/code.cl:331-363
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    Camera cam;
    unsigned col, row;
    RayR ray;
    pinhole(fcam, id, cam, col, row, ray.o, ray.d);
    if (id >= cam.cols * cam.rows) return;
    AabbHit inter = interAABB(ray.o, ray.d, toAABB(bound_a));
    if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
    else { ray.mint = RT_INF; ray.maxt = RT_INF; }
    storeRay(rays + id, ray);
    acu[id] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    pois[id].matId = -1;
}

// A09/code.cl:400-461: one thread per slot replays the fp32 `coord += delta` accumulation (see
// k_initTrace_strat in rt_kernels_a10.cu); rays_per_pixel == 1 is the same path (side 1, coord 0.5).
__global__ void k_a09_initTrace(float4* acu, Ray* rays, Poi8* pois, AabbArg bound_a, CamArg fcam, float focal_length, float lens_rad,
                                unsigned rays_per_pixel, unsigned long long total) {
    unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    Camera cam = floatToCamera(fcam.v);
    unsigned pix = (unsigned)(id / rays_per_pixel), k = (unsigned)(id % rays_per_pixel);
    unsigned col = pix % cam.cols, row = pix / cam.cols;
    acu[id] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    pois[id].matId = -1;
    unsigned side = (unsigned)sqrtf((float)rays_per_pixel);
    if (k >= side * side) return;
    f3 focal_point = getFocalPoint(cam, (float)col, (float)row, focal_length);
    unsigned i = k / side, j = k % side;
    float delta = 1.0f / (float)side;
    f2 coord;
    coord.y = delta / 2.0f;
    for (unsigned a = 0; a < i; a++) coord.y += delta;
    coord.x = delta / 2.0f;
    for (unsigned a = 0; a < j; a++) coord.x += delta;
    RayR ray;
    getThinLensRay(cam, focal_point, lens_rad, coord, ray.o, ray.d);
    AabbHit inter = interAABB(ray.o, ray.d, toAABB(bound_a));
    if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
    else { ray.mint = RT_INF; ray.maxt = RT_INF; }
    storeRay(rays + id, ray);
}

__global__ void k_a089_initShadowTrace(Ray* shadow_rays, const Poi8* pois, unsigned total, f3 light_pos) {   // A08/code.cl:365-390
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    if (pois[id].matId < 0) { storeDeadRay(shadow_rays + id); return; }
    const float4* pq = reinterpret_cast<const float4*>(pois + id);
    float4 p = pq[0], n = pq[1];
    f3 org = mk3(p.x, p.y, p.z) + mk3(n.x, n.y, n.z) * 0.001f;
    storeRay(shadow_rays + id, makeRay(org, light_pos));
}

// STATS = the sinks of rt_set_walk_stats / rt_set_walk_totals are on (instrumented runs only)
RT_DEV void a089Stats(const StatPtrs& sp, unsigned id, const Hit& h, const WalkStats& ws) {
    tallyWalk(sp.totals, 1, ws, h.i != 0xFFFFFFFFu);
    if (sp.hit) sp.hit[id] = h.i;
    if (sp.cells) sp.cells[id] = (unsigned)ws.cells;
    if (sp.tests) sp.tests[id] = (unsigned)ws.tests;
}
RT_DEV void a089StatsInit(const StatPtrs& sp, unsigned id) {
    if (sp.hit) sp.hit[id] = 0xFFFFFFFFu;
    if (sp.cells) sp.cells[id] = 0;
    if (sp.tests) sp.tests[id] = 0;
}

template <int PRIM, bool STATS>
__global__ void k_a089_closest(unsigned total, Poi8* pois, Ray* rays, GridView g, const float4* normals, const unsigned* matid, StatPtrs sp) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    if (STATS) a089StatsInit(sp, id);
    RayR ray = loadRay(rays + id);
    if (ray.mint == ray.maxt) return;
    AabbHit binter = interAABB(ray.o, ray.d, g.bound);
    WalkStats ws = {0, 0, 0};
    if (!binter.v) { if (STATS) tallyWalk(sp.totals, 0, ws, false); return; }
    Hit h = gridWalk<PRIM, false, true, STATS>(ray.o, ray.d, ray.maxt, g, binter, &ws);
    if (STATS) a089Stats(sp, id, h, ws);
    if (h.i == 0xFFFFFFFFu) return;
    rays[id].maxt = h.t;
    f3 p = getPoint(ray.o, ray.d, h.t);
    f3 nrm;
    if (PRIM == PRIM_SPHERE) {
        float4 s = __ldg(g.prim + h.i);
        nrm = normalize(p - mk3(s.x, s.y, s.z));
    } else {
        float4 n0 = __ldg(normals + 3 * h.i), n1 = __ldg(normals + 3 * h.i + 1), n2 = __ldg(normals + 3 * h.i + 2);
        nrm = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
    }
    float4* pq = reinterpret_cast<float4*>(pois + id);
    pq[0] = make_float4(p.x, p.y, p.z, 0.f);
    pq[1] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
    pois[id].matId = (int)__ldg(matid + h.i);
}

template <int PRIM, bool STATS>
__global__ void k_a089_any(unsigned total, Ray* shadow_rays, GridView g, StatPtrs sp) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    if (STATS) a089StatsInit(sp, id);
    RayR ray = loadRay(shadow_rays + id);
    if (ray.mint == ray.maxt) return;
    AabbHit binter = interAABB(ray.o, ray.d, g.bound);
    WalkStats ws = {0, 0, 0};
    if (!binter.v) { if (STATS) tallyWalk(sp.totals, 0, ws, false); return; }
    Hit h = gridWalk<PRIM, true, true, STATS>(ray.o, ray.d, ray.maxt, g, binter, &ws);
    if (STATS) a089Stats(sp, id, h, ws);
    if (h.i != 0xFFFFFFFFu) {
        shadow_rays[id].maxt = h.t;
        shadow_rays[id].mint = h.t;
    } else {
        shadow_rays[id].maxt = h.t;
    }
}

__global__ void k_a089_sceneRender(float4* acu, const Poi8* pois, const Ray* shadow_rays, const float4* material, unsigned total) {   // A08/code.cl:916-939
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int matId = pois[id].matId;
    if (matId < 0) return;
    float4 n = reinterpret_cast<const float4*>(pois + id)[1];
    float shade = 0.2f;
    RayR sr = loadRay(shadow_rays + id);
    if (sr.maxt != sr.mint) shade += cl_clamp(dot(sr.d, mk3(n.x, n.y, n.z)), 0.0f, 1.0f);
    float4 color = __ldg(material + matId);
    float s = cl_clamp(shade, 0.0f, 1.0f);
    float4 a = acu[id];
    acu[id] = make_float4(a.x + color.x * s, a.y + color.y * s, a.z + color.z * s, a.w + 1.0f);
}

__global__ void k_a08_copyToPixel(uchar4* pixel, const float4* acu, float m, unsigned pixels) {   // A08/code.cl:941-951
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pixels) return;
    float4 a = acu[id];
    pixel[id] = mkPixel(a.x * 255.0f * m, a.y * 255.0f * m, a.z * 255.0f * m);   // unclamped (Q11)
}

__global__ void k_a09_copyToPixel(uchar4* pixel, const float4* acu, float m, unsigned pixels, unsigned rays_per_pixel) {   // A09/code.cl:1023-1040
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pixels) return;
    const float4* a = acu + (size_t)id * rays_per_pixel;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (unsigned i = 0; i < rays_per_pixel; i++) {
        float4 v = a[i];
        cx += v.x; cy += v.y; cz += v.z;
    }
    float s = 255.0f * m;
    pixel[id] = mkPixel(cx * s, cy * s, cz * s);   // unclamped (Q11)
}

// ---------------------------------------------------------------------------------------
// A08 / A09 whole frame in one kernel (ours; the launchers above stay the 1:1 drop-ins).  render() of
// A08/code.js:1194-1232 / A09/code.js:1256-1294 enqueues initTrace, the two closest-hit traces and per light
// initShadowTrace + two any-hit traces + sceneRender, each streaming 48-byte Ray / Poi records of EVERY slot through HBM
// (at the default 100 rays per pixel of Assignment 9: 207 M slots, ~30 GB per kernel pair).  Every one of those kernels
// touches only its own slot, so one thread can run the whole sequence with the ray, the hit record and the shadow ray in
// registers and store the slot's accumulator once: same device functions in the same per-slot order, hence the same bits
// (tests compare it with the launcher-by-launcher frame).  copyToPixel stays a second kernel: its per-pixel sum is
// sequential in k (A09/code.cl:1030-1034).
// ---------------------------------------------------------------------------------------
constexpr int kMaxPointLights = 16;
struct A089Frame {
    GridView spheres, triangles;     // prim == nullptr: the set is absent
    const unsigned* s_matid;
    const unsigned* t_matid;
    const float4* t_normal;
    AABB t_shadow_bound;             // A08 hands triangleShadowTrace the SPHERE bounds (quirk Q10), A09 the triangle bounds
    const float4* material;
    AabbArg bound;
    CamArg cam;
    float focal_length, lens_rad;
    unsigned rays_per_pixel, thin_lens, n_lights;
    f3 light[kMaxPointLights];
};

template <int PRIM>
RT_DEV void frameClosest(const GridView& g, const float4* normals, const unsigned* matid, RayR& ray, f3& p, f3& nrm, int& matId) {
    if (ray.mint == ray.maxt) return;
    AabbHit binter = interAABB(ray.o, ray.d, g.bound);
    if (!binter.v) return;
    Hit h = gridWalk<PRIM, false, true, false>(ray.o, ray.d, ray.maxt, g, binter, nullptr);
    if (h.i == 0xFFFFFFFFu) return;
    ray.maxt = h.t;
    p = getPoint(ray.o, ray.d, h.t);
    if (PRIM == PRIM_SPHERE) {
        float4 s = __ldg(g.prim + h.i);
        nrm = normalize(p - mk3(s.x, s.y, s.z));
    } else {
        float4 n0 = __ldg(normals + 3 * h.i), n1 = __ldg(normals + 3 * h.i + 1), n2 = __ldg(normals + 3 * h.i + 2);
        nrm = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
    }
    matId = (int)__ldg(matid + h.i);
}

template <int PRIM>
RT_DEV void frameAny(const GridView& g, const AABB& bound, RayR& sr) {
    if (sr.mint == sr.maxt) return;
    AabbHit binter = interAABB(sr.o, sr.d, bound);
    if (!binter.v) return;
    GridView gb = g;
    gb.bound = bound;
    Hit h = gridWalk<PRIM, true, true, false>(sr.o, sr.d, sr.maxt, gb, binter, nullptr);
    if (h.i != 0xFFFFFFFFu) { sr.maxt = h.t; sr.mint = h.t; }
    else sr.maxt = h.t;
}

__global__ void __launch_bounds__(kBlock) k_a089_frame(const __grid_constant__ A089Frame a, float4* acu, int* out_matid, float* out_maxt,
                                                       unsigned long long total) {
    unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    Camera cam = floatToCamera(a.cam.v);
    unsigned pix = (unsigned)(id / a.rays_per_pixel), k = (unsigned)(id % a.rays_per_pixel);
    unsigned col = pix % cam.cols, row = pix / cam.cols;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    int matId = -1;
    RayR ray;
    ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 0.f); ray.mint = RT_INF; ray.maxt = RT_INF;
    bool have_ray = true;
    if (a.thin_lens) {   // A09/code.cl:400-461
        unsigned side = (unsigned)sqrtf((float)a.rays_per_pixel);
        if (k >= side * side) have_ray = false;
        else {
            f3 focal_point = getFocalPoint(cam, (float)col, (float)row, a.focal_length);
            unsigned i = k / side, j = k % side;
            float delta = 1.0f / (float)side;
            f2 coord;
            coord.y = delta / 2.0f;
            for (unsigned q = 0; q < i; q++) coord.y += delta;
            coord.x = delta / 2.0f;
            for (unsigned q = 0; q < j; q++) coord.x += delta;
            getThinLensRay(cam, focal_point, a.lens_rad, coord, ray.o, ray.d);
        }
    } else {             // A08/code.cl:331-363
        getRay(cam, (float)col, (float)row, ray.o, ray.d);
    }
    f3 p = mk3(0.f, 0.f, 0.f), nrm = p;
    if (have_ray) {
        AabbHit inter = interAABB(ray.o, ray.d, toAABB(a.bound));
        if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
        else { ray.mint = RT_INF; ray.maxt = RT_INF; }
        if (a.spheres.prim) frameClosest<PRIM_SPHERE>(a.spheres, nullptr, a.s_matid, ray, p, nrm, matId);
        if (a.triangles.prim) frameClosest<PRIM_TRIANGLE>(a.triangles, a.t_normal, a.t_matid, ray, p, nrm, matId);
        if (matId >= 0) {
            float4 color = __ldg(a.material + matId);
            f3 org = p + nrm * 0.001f;
            for (unsigned l = 0; l < a.n_lights; l++) {   // initShadowTrace + shadow traces + sceneRender, A08/code.cl:365-390, 916-939
                RayR sr = makeRay(org, a.light[l]);
                if (a.spheres.prim) frameAny<PRIM_SPHERE>(a.spheres, a.spheres.bound, sr);
                if (a.triangles.prim) frameAny<PRIM_TRIANGLE>(a.triangles, a.t_shadow_bound, sr);
                float shade = 0.2f;
                if (sr.maxt != sr.mint) shade += cl_clamp(dot(sr.d, nrm), 0.0f, 1.0f);
                float s = cl_clamp(shade, 0.0f, 1.0f);
                acc = make_float4(acc.x + color.x * s, acc.y + color.y * s, acc.z + color.z * s, acc.w + 1.0f);
            }
        }
    }
    acu[id] = acc;
    if (out_matid) out_matid[id] = matId;
    if (out_maxt) out_maxt[id] = ray.maxt;
}

#define RT_GRID1(n) rt_blocks((n), kBlock), kBlock, 0, ctx->stream

// The queue-walker route of the Assignment-7 traces (rt_walk_a07) applies to a grid this library built (its cell table is registered
// with the context, with occupancy bits) that is worth the three launches: >= 32 slabs per axis and >= 64 K references, a frame of
// >= 64 K pixels, and no statistics sinks (those live in the one-thread-per-ray kernel).  RT2015_A07_WALKER=0 switches it off.
rt_ctx::GridAux* a07WalkerGrid(rt_ctx* ctx, const void* slab_size, const void* prim, unsigned n_slabs, size_t npix) {
    static const bool on = !(getenv("RT2015_A07_WALKER") && atoi(getenv("RT2015_A07_WALKER")) == 0);
    if (!on || ctx->st_hit || ctx->st_cells || ctx->st_tests || ctx->st_totals) return nullptr;
    rt_ctx::GridAux* g = rt_grid_aux_of(ctx, slab_size);
    if (!g || g->prim != prim || g->n_slabs != n_slabs || !rt_grid_wants_walker(*g) || !g->aux_ready) return nullptr;
    if (npix < 65536 || npix > 0x7FFFFFFFull) return nullptr;
    return g;
}

template <int PRIM>
void launchA089Closest(rt_ctx* ctx, unsigned total, void* pois, void* rays, const GridView& g, const void* normals, const void* matid) {
    StatPtrs sp = {ctx->st_hit, ctx->st_cells, ctx->st_tests, ctx->st_totals};
    if (sp.hit || sp.cells || sp.tests || sp.totals)
        k_a089_closest<PRIM, true><<<RT_GRID1(total)>>>(total, (Poi8*)pois, (Ray*)rays, g, (const float4*)normals, (const unsigned*)matid, sp);
    else
        k_a089_closest<PRIM, false><<<RT_GRID1(total)>>>(total, (Poi8*)pois, (Ray*)rays, g, (const float4*)normals, (const unsigned*)matid, sp);
}
template <int PRIM>
void launchA089Any(rt_ctx* ctx, unsigned total, void* shadow_rays, const GridView& g) {
    StatPtrs sp = {ctx->st_hit, ctx->st_cells, ctx->st_tests, ctx->st_totals};
    if (sp.hit || sp.cells || sp.tests || sp.totals) k_a089_any<PRIM, true><<<RT_GRID1(total)>>>(total, (Ray*)shadow_rays, g, sp);
    else k_a089_any<PRIM, false><<<RT_GRID1(total)>>>(total, (Ray*)shadow_rays, g, sp);
}

template <int PRIM>
void launchA07(rt_ctx* ctx, size_t n, void* pixels, const float* fcam, void* rays, const GridView& g, const void* normals) {
    StatPtrs sp = {ctx->st_hit, ctx->st_cells, ctx->st_tests, ctx->st_totals};
    const bool stats = sp.hit || sp.cells || sp.tests || sp.totals;
    if (stats) {
        if (g.occ) k_a07_trace<PRIM, true, true><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, g, (const float4*)normals, sp);
        else k_a07_trace<PRIM, false, true><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, g, (const float4*)normals, sp);
    } else {
        if (g.occ) k_a07_trace<PRIM, true, false><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, g, (const float4*)normals, sp);
        else k_a07_trace<PRIM, false, false><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, g, (const float4*)normals, sp);
    }
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------- A01 - A03
int rt_a01_raytrace(rt_ctx* ctx, void* pixels, const float* fcam) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam) return RT_ERR_INVALID;
    unsigned rows = (unsigned)fcam[14], cols = (unsigned)fcam[15];   // A01 packs rows first
    if (!rows || !cols) return RT_OK;
    k_a01_raytrace<<<RT_GRID1((size_t)cols * rows)>>>((uchar4*)pixels, mkCam(fcam), cols, rows);
    RT_LAUNCH_CHECK(ctx, "A01 raytrace");
    return RT_OK;
}

int rt_a02_raytrace(rt_ctx* ctx, void* pixels, const float* fcam, unsigned s_size, const void* s_atoms, const void* s_colors) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam || (s_size && (!s_atoms || !s_colors))) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a02_raytrace<<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), s_size, (const float4*)s_atoms, (const float4*)s_colors);
    RT_LAUNCH_CHECK(ctx, "A02 raytrace");
    return RT_OK;
}

int rt_a03_initTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam || !rays) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a03_initTrace<<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays);
    RT_LAUNCH_CHECK(ctx, "A03 initTrace");
    return RT_OK;
}

int rt_a03_molTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms, const void* s_colors) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam || !rays || (s_size && (!s_atoms || !s_colors))) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a03_molTrace<<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, s_size, (const float4*)s_atoms, (const float4*)s_colors);
    RT_LAUNCH_CHECK(ctx, "A03 molTrace");
    return RT_OK;
}

// ------------------------------------------------------------------------------- A04 - A06
int rt_a04_initTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays) { return rt_a03_initTrace(ctx, pixels, fcam, rays); }
int rt_a04_raytrace(rt_ctx* ctx, void* pixels, const float* fcam, unsigned s_size, const void* s_atoms, const void* s_colors) {
    return rt_a02_raytrace(ctx, pixels, fcam, s_size, s_atoms, s_colors);
}

static int launchMol045(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms, const void* s_colors,
                        const float* bound) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam || !rays || (s_size && (!s_atoms || !s_colors))) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    AabbArg none;
    memset(&none, 0, sizeof none);
    if (bound) k_a045_molTrace<true><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, s_size, (const float4*)s_atoms, (const float4*)s_colors, mkAabb(bound));
    else k_a045_molTrace<false><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, s_size, (const float4*)s_atoms, (const float4*)s_colors, none);
    RT_LAUNCH_CHECK(ctx, "A04/A05 molTrace");
    return RT_OK;
}
static int launchMesh045(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos, const void* t_normal,
                         const void* t_mindex, const void* m_color, const float* bound) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam || !rays || (t_size && (!t_pos || !t_normal || !t_mindex || !m_color))) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    AabbArg none;
    memset(&none, 0, sizeof none);
    if (bound) k_a045_meshTrace<true><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, t_size, (const float4*)t_pos, (const float4*)t_normal,
                                                         (const unsigned*)t_mindex, (const float4*)m_color, mkAabb(bound));
    else k_a045_meshTrace<false><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, t_size, (const float4*)t_pos, (const float4*)t_normal,
                                                   (const unsigned*)t_mindex, (const float4*)m_color, none);
    RT_LAUNCH_CHECK(ctx, "A04/A05 meshTrace");
    return RT_OK;
}
int rt_a04_molTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms, const void* s_colors) {
    return launchMol045(ctx, pixels, fcam, rays, s_size, s_atoms, s_colors, nullptr);
}
int rt_a04_meshTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos, const void* t_normal,
                     const void* t_mindex, const void* m_color) {
    return launchMesh045(ctx, pixels, fcam, rays, t_size, t_pos, t_normal, t_mindex, m_color, nullptr);
}
int rt_a05_initTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, const float* bound) {
    return rt_a07_initTrace(ctx, pixels, fcam, rays, bound);   // same text, A05/code.cl:304-328 = A07/code.cl:311-335
}
int rt_a05_molTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms, const void* s_colors,
                    const float* bound) {
    if (!bound) return RT_ERR_INVALID;
    return launchMol045(ctx, pixels, fcam, rays, s_size, s_atoms, s_colors, bound);
}
int rt_a05_meshTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos, const void* t_normal,
                     const void* t_mindex, const void* m_color, const float* bound) {
    if (!bound) return RT_ERR_INVALID;
    return launchMesh045(ctx, pixels, fcam, rays, t_size, t_pos, t_normal, t_mindex, m_color, bound);
}
int rt_a06_initTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, const float* bound) {
    return rt_a07_initTrace(ctx, pixels, fcam, rays, bound);   // A06/code.cl:310-334
}
int rt_a06_molTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms, const void* s_colors,
                    const float* bound, unsigned n_slabs, const void* slab_size) {
    RT_CHECK_CTX(ctx);
    (void)s_size; (void)s_colors;   // the kernel colours by slab number; the material lookup is commented out in the reference
    if (!pixels || !fcam || !rays || !s_atoms || !bound || !n_slabs || !slab_size) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a06_trace<PRIM_SPHERE><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, (const float4*)s_atoms, nullptr, mkAabb(bound), n_slabs,
                                              (const unsigned*)slab_size);
    RT_LAUNCH_CHECK(ctx, "A06 molTrace");
    return RT_OK;
}
int rt_a06_meshTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos, const void* t_normal,
                     const void* t_mindex, const void* m_color, const float* bound, unsigned n_slabs, const void* slab_size) {
    RT_CHECK_CTX(ctx);
    (void)t_size; (void)t_mindex; (void)m_color;
    if (!pixels || !fcam || !rays || !t_pos || !t_normal || !bound || !n_slabs || !slab_size) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a06_trace<PRIM_TRIANGLE><<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, (const float4*)t_pos, (const float4*)t_normal, mkAabb(bound),
                                                n_slabs, (const unsigned*)slab_size);
    RT_LAUNCH_CHECK(ctx, "A06 meshTrace");
    return RT_OK;
}

// ------------------------------------------------------------------------------- A07
int rt_a07_initTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, const float* bound) {
    RT_CHECK_CTX(ctx);
    if (!pixels || !fcam || !rays || !bound) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a07_initTrace<<<RT_GRID1(n)>>>((uchar4*)pixels, mkCam(fcam), (Ray*)rays, mkAabb(bound));
    RT_LAUNCH_CHECK(ctx, "A07 initTrace");
    return RT_OK;
}

int rt_a07_molTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms, const void* s_mindex,
                    const void* m_color, const float* bound, unsigned n_slabs, const void* slab_size) {
    RT_CHECK_CTX(ctx);
    (void)s_size; (void)s_mindex; (void)m_color;   // the kernel colours by cell parity; the material lookup is commented out in the reference
    if (!pixels || !fcam || !rays || !s_atoms || !bound || !n_slabs || !slab_size) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    if (rt_ctx::GridAux* aux = a07WalkerGrid(ctx, slab_size, s_atoms, n_slabs, n)) {   // big grid of ours: through the queue walker
        int rc = rt_walk_a07(ctx, *aux, PRIM_SPHERE, pixels, fcam, rays, nullptr, bound, n);
        if (rc) return rc;
        return RT_OK;
    }
    GridView g = mkGrid(s_atoms, slab_size, bound, n_slabs, rt_occupancy_of(ctx, slab_size));
    launchA07<PRIM_SPHERE>(ctx, n, pixels, fcam, rays, g, nullptr);
    RT_LAUNCH_CHECK(ctx, "A07 molTrace");
    return RT_OK;
}

int rt_a07_meshTrace(rt_ctx* ctx, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos, const void* t_normal,
                     const void* t_mindex, const void* m_color, const float* bound, unsigned n_slabs, const void* slab_size) {
    RT_CHECK_CTX(ctx);
    (void)t_size; (void)t_mindex; (void)m_color;
    if (!pixels || !fcam || !rays || !t_pos || !t_normal || !bound || !n_slabs || !slab_size) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    if (rt_ctx::GridAux* aux = a07WalkerGrid(ctx, slab_size, t_pos, n_slabs, n)) {
        int rc = rt_walk_a07(ctx, *aux, PRIM_TRIANGLE, pixels, fcam, rays, t_normal, bound, n);
        if (rc) return rc;
        return RT_OK;
    }
    GridView g = mkGrid(t_pos, slab_size, bound, n_slabs, rt_occupancy_of(ctx, slab_size));
    launchA07<PRIM_TRIANGLE>(ctx, n, pixels, fcam, rays, g, t_normal);
    RT_LAUNCH_CHECK(ctx, "A07 meshTrace");
    return RT_OK;
}

// ------------------------------------------------------------------------------- A08
int rt_a08_initTrace(rt_ctx* ctx, void* acu, void* rays, void* pois, const float* bound, const float* fcam) {
    RT_CHECK_CTX(ctx);
    if (!acu || !rays || !pois || !bound || !fcam) return RT_ERR_INVALID;
    size_t n = (size_t)(unsigned)fcam[14] * (unsigned)fcam[15];
    if (!n) return RT_OK;
    k_a08_initTrace<<<RT_GRID1(n)>>>((float4*)acu, (Ray*)rays, (Poi8*)pois, mkAabb(bound), mkCam(fcam));
    RT_LAUNCH_CHECK(ctx, "A08 initTrace");
    return RT_OK;
}

int rt_a09_initShadowTrace(rt_ctx* ctx, void* shadow_rays, void* pois, unsigned total_rays, const float* light_pos) {
    RT_CHECK_CTX(ctx);
    if (!shadow_rays || !pois || !light_pos) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_a089_initShadowTrace<<<RT_GRID1(total_rays)>>>((Ray*)shadow_rays, (const Poi8*)pois, total_rays, f3{light_pos[0], light_pos[1], light_pos[2]});
    RT_LAUNCH_CHECK(ctx, "initShadowTrace");
    return RT_OK;
}
int rt_a08_initShadowTrace(rt_ctx* ctx, void* shadow_rays, void* pois, unsigned cols, unsigned rows, const float* light_pos) {
    return rt_a09_initShadowTrace(ctx, shadow_rays, pois, cols * rows, light_pos);
}

int rt_a09_sphereTrace(rt_ctx* ctx, unsigned total_rays, void* pois, void* rays, const void* spheres, const void* s_matid,
                       const void* s_box_size, const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !spheres || !s_matid || !s_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    launchA089Closest<PRIM_SPHERE>(ctx, total_rays, pois, rays, mkGrid(spheres, s_box_size, bound, n_slabs), nullptr, s_matid);
    RT_LAUNCH_CHECK(ctx, "sphereTrace");
    return RT_OK;
}
int rt_a08_sphereTrace(rt_ctx* ctx, unsigned cols, unsigned rows, void* pois, void* rays, const void* spheres, const void* s_matid,
                       const void* s_box_size, const float* bound, unsigned n_slabs) {
    return rt_a09_sphereTrace(ctx, cols * rows, pois, rays, spheres, s_matid, s_box_size, bound, n_slabs);
}

int rt_a09_triangleTrace(rt_ctx* ctx, unsigned total_rays, void* pois, void* rays, const void* t_pos, const void* t_normal,
                         const void* t_matid, const void* t_box_size, const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !t_pos || !t_normal || !t_matid || !t_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    launchA089Closest<PRIM_TRIANGLE>(ctx, total_rays, pois, rays, mkGrid(t_pos, t_box_size, bound, n_slabs), t_normal, t_matid);
    RT_LAUNCH_CHECK(ctx, "triangleTrace");
    return RT_OK;
}
int rt_a08_triangleTrace(rt_ctx* ctx, unsigned cols, unsigned rows, void* pois, void* rays, const void* t_pos, const void* t_normal,
                         const void* t_matid, const void* t_box_size, const float* bound, unsigned n_slabs) {
    return rt_a09_triangleTrace(ctx, cols * rows, pois, rays, t_pos, t_normal, t_matid, t_box_size, bound, n_slabs);
}

int rt_a09_sphereShadowTrace(rt_ctx* ctx, unsigned total_rays, void* shadow_rays, const void* spheres, const void* s_box_size,
                             const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!shadow_rays || !spheres || !s_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    launchA089Any<PRIM_SPHERE>(ctx, total_rays, shadow_rays, mkGrid(spheres, s_box_size, bound, n_slabs));
    RT_LAUNCH_CHECK(ctx, "sphereShadowTrace");
    return RT_OK;
}
int rt_a08_sphereShadowTrace(rt_ctx* ctx, unsigned cols, unsigned rows, void* shadow_rays, const void* spheres, const void* s_box_size,
                             const float* bound, unsigned n_slabs) {
    return rt_a09_sphereShadowTrace(ctx, cols * rows, shadow_rays, spheres, s_box_size, bound, n_slabs);
}

int rt_a09_triangleShadowTrace(rt_ctx* ctx, unsigned total_rays, void* shadow_rays, const void* t_pos, const void* t_box_size,
                               const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!shadow_rays || !t_pos || !t_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    launchA089Any<PRIM_TRIANGLE>(ctx, total_rays, shadow_rays, mkGrid(t_pos, t_box_size, bound, n_slabs));
    RT_LAUNCH_CHECK(ctx, "triangleShadowTrace");
    return RT_OK;
}
int rt_a08_triangleShadowTrace(rt_ctx* ctx, unsigned cols, unsigned rows, void* shadow_rays, const void* t_pos, const void* t_box_size,
                               const float* bound, unsigned n_slabs) {
    return rt_a09_triangleShadowTrace(ctx, cols * rows, shadow_rays, t_pos, t_box_size, bound, n_slabs);
}

int rt_a09_sceneRender(rt_ctx* ctx, void* acu, void* pois, const void* shadow_rays, const void* material, unsigned total_rays) {
    RT_CHECK_CTX(ctx);
    if (!acu || !pois || !shadow_rays || !material) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_a089_sceneRender<<<RT_GRID1(total_rays)>>>((float4*)acu, (const Poi8*)pois, (const Ray*)shadow_rays, (const float4*)material, total_rays);
    RT_LAUNCH_CHECK(ctx, "sceneRender");
    return RT_OK;
}
int rt_a08_sceneRender(rt_ctx* ctx, void* acu, void* pois, const void* shadow_rays, const void* material, unsigned pixels) {
    return rt_a09_sceneRender(ctx, acu, pois, shadow_rays, material, pixels);
}

int rt_a08_copyToPixel(rt_ctx* ctx, void* pixel, const void* acu, float m, unsigned pixels) {
    RT_CHECK_CTX(ctx);
    if (!pixel || !acu) return RT_ERR_INVALID;
    if (!pixels) return RT_OK;
    k_a08_copyToPixel<<<RT_GRID1(pixels)>>>((uchar4*)pixel, (const float4*)acu, m, pixels);
    RT_LAUNCH_CHECK(ctx, "A08 copyToPixel");
    return RT_OK;
}

// ------------------------------------------------------------------------------- A09
int rt_a09_initTrace(rt_ctx* ctx, void* acu, void* rays, void* pois, const float* bound, const float* fcam, float focal_length,
                     float lens_rad, unsigned rays_per_pixel) {
    RT_CHECK_CTX(ctx);
    if (!acu || !rays || !pois || !bound || !fcam || !rays_per_pixel) return RT_ERR_INVALID;
    unsigned long long total = (unsigned long long)(unsigned)fcam[14] * (unsigned)fcam[15] * rays_per_pixel;
    if (!total) return RT_OK;
    if (total > 0xFFFFFFFFull) return rt_fail(ctx, RT_ERR_INVALID, "A09 initTrace: total_rays exceeds the reference's uint range");
    k_a09_initTrace<<<RT_GRID1(total)>>>((float4*)acu, (Ray*)rays, (Poi8*)pois, mkAabb(bound), mkCam(fcam), focal_length, lens_rad,
                                         rays_per_pixel, total);
    RT_LAUNCH_CHECK(ctx, "A09 initTrace");
    return RT_OK;
}

int rt_a089_render_frame(rt_ctx* ctx, const rt_a089_frame* f, void* acu, void* out_matid, void* out_maxt) {
    RT_CHECK_CTX(ctx);
    if (!f || !acu || !f->material || !f->rays_per_pixel || f->n_lights > (unsigned)kMaxPointLights || (f->n_lights && !f->light_pos))
        return RT_ERR_INVALID;
    if (f->spheres && (!f->s_matid || !f->s_box_size || !f->s_n_slabs)) return RT_ERR_INVALID;
    if (f->t_pos && (!f->t_normal || !f->t_matid || !f->t_box_size || !f->t_n_slabs)) return RT_ERR_INVALID;
    unsigned long long total = (unsigned long long)(unsigned)f->fcam[14] * (unsigned)f->fcam[15] * f->rays_per_pixel;
    if (!total) return RT_OK;
    if (total > 0xFFFFFFFFull) return rt_fail(ctx, RT_ERR_INVALID, "a089_render_frame: total_rays exceeds the reference's uint range");
    A089Frame a;
    memset(&a, 0, sizeof a);
    if (f->spheres) a.spheres = mkGrid(f->spheres, f->s_box_size, f->s_bound, f->s_n_slabs);
    if (f->t_pos) a.triangles = mkGrid(f->t_pos, f->t_box_size, f->t_bound, f->t_n_slabs);
    a.s_matid = (const unsigned*)f->s_matid;
    a.t_matid = (const unsigned*)f->t_matid;
    a.t_normal = (const float4*)f->t_normal;
    a.t_shadow_bound.pmin = f3{f->t_shadow_bound[0], f->t_shadow_bound[1], f->t_shadow_bound[2]};
    a.t_shadow_bound.pmax = f3{f->t_shadow_bound[4], f->t_shadow_bound[5], f->t_shadow_bound[6]};
    a.material = (const float4*)f->material;
    a.bound = mkAabb(f->bound);
    a.cam = mkCam(f->fcam);
    a.focal_length = f->focal_length;
    a.lens_rad = f->lens_rad;
    a.rays_per_pixel = f->rays_per_pixel;
    a.thin_lens = f->thin_lens ? 1u : 0u;
    a.n_lights = f->n_lights;
    for (unsigned l = 0; l < f->n_lights; l++) a.light[l] = f3{f->light_pos[4 * l], f->light_pos[4 * l + 1], f->light_pos[4 * l + 2]};
    k_a089_frame<<<RT_GRID1(total)>>>(a, (float4*)acu, (int*)out_matid, (float*)out_maxt, total);
    RT_LAUNCH_CHECK(ctx, "A08/A09 frame");
    return RT_OK;
}

int rt_a09_copyToPixel(rt_ctx* ctx, void* pixel, const void* acu, float m, unsigned pixels, unsigned rays_per_pixel) {
    RT_CHECK_CTX(ctx);
    if (!pixel || !acu) return RT_ERR_INVALID;
    if (!pixels) return RT_OK;
    k_a09_copyToPixel<<<RT_GRID1(pixels)>>>((uchar4*)pixel, (const float4*)acu, m, pixels, rays_per_pixel);
    RT_LAUNCH_CHECK(ctx, "A09 copyToPixel");
    return RT_OK;
}

}  // extern "C"
