/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product.
 *
 * rt_oracle.c -- plain-C restatement of the Assignment-10 kernels of eaymerich/2015-RayTracing
 * (the north-star path: thin-lens primaries, 3D-DDA closest-hit / any-hit walks, disk-light next
 * event estimation, cosine bounces, accumulation).  Every function cites the lines of
 * /root/reference/Assign10-Path_Tracing/code.cl (= "A10") it follows and keeps their fp32 operation
 * order; build with -ffp-contract=off -fwrapv (oracle/Makefile).  It exports the same entry points
 * as oracle/ref_driver.cpp with the prefix port_ (kind "port" in oracle/refcl.py) for the a10_*
 * kernels, and is itself pinned: tests/test_oracle_golden.py::test_port_equals_reference_kernels
 * requires it to reproduce, bit for bit, the golden fixtures that the reference's own kernel text
 * (oracle/_ref) produced.  It exists so that the parity suite still has an oracle where /root/reference
 * and the prebuilt oracle/_ref are both absent, and as an independent reading of the kernel text.
 *
 * Conventions of the restatement: float3 is a 16-byte slot (x,y,z,pad) as in OpenCL; `Ray` is 48 B,
 * `Poi` 64 B (A10:27-31, 57-62); cos/sin are the double functions rounded to float (as clshim.h);
 * min/max/clamp are the OpenCL ternaries; the one `mad` (A10:209) is fmaf.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

typedef unsigned int uint;
typedef struct { float x, y, z, w; } f3;              /* float3 in memory: 16 B */
typedef struct { float x, y; } f2;
typedef struct { f3 o, d; float mint, maxt, pad0, pad1; } Ray;                       /* 48 B */
typedef struct { f3 p, normal, atte; int matId, pad0, pad1, pad2; } Poi;             /* 64 B */
typedef struct { f3 pmin, pmax; } AABB;
typedef struct { f3 eye, U, V, W; float width, height; uint cols, rows; } Camera;
typedef struct { float tmin, tmax; int v; } AabbInter;
typedef struct { float t; int v; } Inter;
typedef struct { float t, beta, gamma; int v; } TriInter;
typedef char ray_is_48[sizeof(Ray) == 48 ? 1 : -1];
typedef char poi_is_64[sizeof(Poi) == 64 ? 1 : -1];

static inline f3 v3(float x, float y, float z) { f3 r = {x, y, z, 0.0f}; return r; }
static inline f3 add(f3 a, f3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub(f3 a, f3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul(f3 a, f3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 scl(float s, f3 a) { return v3(s * a.x, s * a.y, s * a.z); }      /* float * float3 */
static inline f3 lcs(f3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }      /* float3 * float */
static inline f3 neg(f3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross(f3 a, f3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length3(f3 a) { return sqrtf(dot(a, a)); }
static inline f3 normalize(f3 a) { float l = length3(a); return v3(a.x / l, a.y / l, a.z / l); }
static inline float cl_min(float x, float y) { return y < x ? y : x; }
static inline float cl_max(float x, float y) { return x < y ? y : x; }
static inline float cl_clamp(float x, float lo, float hi) { return cl_min(cl_max(x, lo), hi); }
static inline float cl_cos(float a) { return (float)cos((double)a); }
static inline float cl_sin(float a) { return (float)sin((double)a); }
static inline f3 arg3(const float* c) { return v3(c[0], c[1], c[2]); }

static Camera floatToCamera(const float* in) {   /* A10:73-84 */
    Camera c;
    c.eye = arg3(in); c.U = arg3(in + 3); c.V = arg3(in + 6); c.W = arg3(in + 9);
    c.width = in[12]; c.height = in[13];
    c.cols = (uint)in[14]; c.rows = (uint)in[15];
    return c;
}
static AABB toAABB(const float* b) { AABB r; r.pmin = arg3(b); r.pmax = arg3(b + 4); return r; }

static inline f3 getPoint(Ray r, float t) { return add(r.o, scl(t, r.d)); }   /* A10:86-88 */

static Ray getRay(Camera cam, float col, float row) {   /* A10:108-119 */
    Ray ray;
    memset(&ray, 0, sizeof ray);
    f3 cop = add(add(scl((-0.5f + (col + 0.5f) / cam.cols) * cam.width, cam.U), scl((0.5f - (row + 0.5f) / cam.rows) * cam.height, cam.V)),
                 scl(-1.0f, cam.W));
    ray.d = normalize(cop);
    ray.o = cam.eye;
    ray.mint = 0.0f;
    ray.maxt = HUGE_VALF;
    return ray;
}

static Ray makeRay(f3 ori, f3 dst) {   /* A10:121-129 */
    Ray ray;
    memset(&ray, 0, sizeof ray);
    ray.o = ori;
    ray.d = normalize(sub(dst, ori));
    ray.mint = 0.0f;
    ray.maxt = length3(sub(dst, ori));
    return ray;
}

static f2 concentric_distort(f2 in) {   /* A10:143-172 */
    if (in.x == 0.0f && in.y == 0.0f) return in;
    float phi = 0.0f, radius = 1.0f;
    float a = (2.0f * in.x) - 1.0f;
    float b = (2.0f * in.y) - 1.0f;
    if ((a * a) > (b * b)) {
        radius *= a;
        phi = 0.78539816339744830962f * (b / a);
    } else {
        radius *= b;
        phi = 1.57079632679489661923f - (0.78539816339744830962f * (a / b));
    }
    f2 r = {cl_cos(phi) * radius, cl_sin(phi) * radius};
    return r;
}

static f3 getFocalPoint(Camera cam, float col, float row, float focal_length) {   /* A10:174-181 */
    Ray ray = getRay(cam, col, row);
    f3 pip = add(cam.eye, lcs(scl(focal_length, cam.W), -1.0f));
    f3 N = cam.W;
    float d = -dot(pip, N);
    float t = -(dot(ray.o, N) + d) / dot(ray.d, N);
    return getPoint(ray, t);
}

static Ray getThinLensRay(Camera cam, f3 focal_point, float lens_rad, f2 coord) {   /* A10:183-197 */
    Ray ray;
    memset(&ray, 0, sizeof ray);
    ray.mint = 0.0f;
    ray.maxt = HUGE_VALF;
    f2 dc = concentric_distort(coord);
    dc.x = dc.x * lens_rad;
    dc.y = dc.y * lens_rad;
    ray.o = add(add(cam.eye, scl(dc.x, cam.U)), scl(dc.y, cam.V));
    ray.d = normalize(sub(focal_point, ray.o));
    return ray;
}

static Inter interSphere(Ray r, f3 c, float r2) {   /* A10:199-242: stored radius is already squared; inclusive range */
    Inter inter;
    f3 omc = sub(r.o, c);
    float a = dot(r.d, r.d);
    float b = 2.0f * dot(omc, r.d);
    float cc = dot(omc, omc) - r2;
    float dis = fmaf(-4.0f * cc, a, b * b);
    inter.v = 0;
    inter.t = HUGE_VALF;
    if (dis < 0.0f) return inter;
    a = 1.0f / (2.0f * a);
    dis = sqrtf(dis);
    float t0 = (-b - dis) * a, t1 = (-b + dis) * a;
    float tmin = fminf(t0, t1), tmax = fmaxf(t0, t1);
    if (tmin >= r.mint && tmin <= r.maxt) { inter.t = tmin; inter.v = 1; return inter; }
    if (tmax >= r.mint && tmax <= r.maxt) { inter.t = tmax; inter.v = 1; return inter; }
    return inter;
}

static TriInter interTriangle(Ray ray, f3 p0, f3 p1, f3 p2) {   /* A10:250-288: one-sided, inclusive range */
    TriInter inter;
    inter.v = 0; inter.t = 0.f; inter.beta = 0.f; inter.gamma = 0.f;
    f3 e1 = sub(p1, p0), e2 = sub(p2, p0);
    float div = dot(cross(e2, e1), ray.d);
    if (div <= 0) return inter;
    float idiv = 1.0f / div;
    f3 s = sub(ray.o, p0);
    float beta = dot(cross(s, ray.d), e2) * idiv;
    if (beta < 0.0f || beta > 1.0f) return inter;
    float gamma = dot(cross(s, e1), ray.d) * idiv;
    if (gamma < 0.0f || (gamma + beta) < 0.0f || (gamma + beta) > 1.0f) return inter;
    inter.t = dot(cross(s, e2), e1) * -idiv;
    if (inter.t >= ray.mint && inter.t <= ray.maxt) { inter.beta = beta; inter.gamma = gamma; inter.v = 1; }
    return inter;
}

static AabbInter interAABB(Ray ray, AABB box) {   /* A10:335-389: starts from [0, +inf), early outs per axis */
    AabbInter inter;
    float ttmin, ttmax, temp;
    inter.tmin = 0.0f; inter.tmax = HUGE_VALF; inter.v = 0;
    ttmin = (box.pmin.x - ray.o.x) / ray.d.x; ttmax = (box.pmax.x - ray.o.x) / ray.d.x;
    if (ray.d.x < 0) { temp = ttmin; ttmin = ttmax; ttmax = temp; }
    inter.tmin = cl_max(ttmin, inter.tmin); inter.tmax = cl_min(ttmax, inter.tmax);
    if (inter.tmin > inter.tmax) return inter;
    ttmin = (box.pmin.y - ray.o.y) / ray.d.y; ttmax = (box.pmax.y - ray.o.y) / ray.d.y;
    if (ray.d.y < 0) { temp = ttmin; ttmin = ttmax; ttmax = temp; }
    inter.tmin = cl_max(ttmin, inter.tmin); inter.tmax = cl_min(ttmax, inter.tmax);
    if (inter.tmin > inter.tmax) return inter;
    ttmin = (box.pmin.z - ray.o.z) / ray.d.z; ttmax = (box.pmax.z - ray.o.z) / ray.d.z;
    if (ray.d.z < 0) { temp = ttmin; ttmin = ttmax; ttmax = temp; }
    inter.tmin = cl_max(ttmin, inter.tmin); inter.tmax = cl_min(ttmax, inter.tmax);
    if (inter.tmin > inter.tmax) return inter;
    inter.v = 1;
    return inter;
}

static Inter interLight(Ray ray, f3 light_pos, f3 light_normal, float radius) {   /* A10:391-403: no t > 0 test (Q5) */
    Inter inter;
    inter.v = 0; inter.t = 0.f;
    float den = dot(ray.d, light_normal);
    if (den == 0.0f) return inter;
    float num = dot(sub(light_pos, ray.o), light_normal);
    if (num == 0.0f) return inter;
    inter.t = num / den;
    f3 poi = getPoint(ray, inter.t);
    if (length3(sub(poi, light_pos)) > radius) return inter;
    inter.v = 1;
    return inter;
}

static inline f3 interp(float beta, float gamma, f3 v1, f3 v2, f3 v3_) {   /* A10:409-411 */
    return add(add(scl(1.0f - beta - gamma, v1), scl(beta, v2)), scl(gamma, v3_));
}

static float getRand(int* seeds, size_t gid0) {   /* A10:420-434: 32-bit wrapping product, signed %, fabs (Q6) */
    const float im = 1.0f / 2147483647.0f;
    int seed = seeds[gid0];
    seed = (int)((long long)(int)((uint)seed * 16807u) % 2147483647LL);
    seeds[gid0] = seed;
    return fabsf((float)seed * im);
}

/* ---- the 3D-DDA walk shared by the five trace kernels (A10:694-786 and its textual copies).
 * kind 0 = spheres (float4 per ref), 1 = triangles (3 x float3 slots per ref); any != 0 breaks on the first hit.
 * Returns champ_i (0xFFFFFFFF = none); champ_t starts at the STORED ray.maxt. */
typedef struct { float t; uint i; float beta, gamma; } Champ;

static Champ walk(Ray ray, AabbInter binter, const float* prim, const uint* box_size, AABB bound, uint n_slabs, int kind, int any) {
    float delta[3], delta_t[3], t_next[3];
    int slab[3], dslab[3], limit[3];
    const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
    const float pmin[3] = {bound.pmin.x, bound.pmin.y, bound.pmin.z}, pmax[3] = {bound.pmax.x, bound.pmax.y, bound.pmax.z};
    for (int a = 0; a < 3; a++) {   /* A10:696-731 */
        float x = o[a] + binter.tmin * d[a];
        delta[a] = (pmax[a] - pmin[a]) / n_slabs;
        slab[a] = (int)((x - pmin[a]) / delta[a]);
        if (slab[a] < 0) slab[a] = 0;
        if ((uint)slab[a] >= n_slabs) slab[a] = (int)n_slabs - 1;
        dslab[a] = (d[a] >= 0) ? 1 : -1;
        limit[a] = (d[a] >= 0) ? (int)n_slabs : -1;
        delta_t[a] = delta[a] / fabsf(d[a]);
        float nxt = pmin[a] + (slab[a] + ((d[a] >= 0) ? 1 : 0)) * delta[a];
        t_next[a] = (nxt - o[a]) / d[a];
    }
    Champ ch;
    ch.t = ray.maxt; ch.i = 0xFFFFFFFFu; ch.beta = 0.f; ch.gamma = 0.f;
    float t = binter.tmin;
    const uint z_stride = n_slabs * n_slabs, y_stride = n_slabs;
    for (;;) {   /* A10:745-786 */
        ray.mint = t;
        ray.maxt = cl_min(cl_min(t_next[0], t_next[1]), t_next[2]);
        uint cell = (uint)slab[2] * z_stride + (uint)slab[1] * y_stride + (uint)slab[0];
        uint begin = box_size[cell], end = box_size[cell + 1];
        for (uint i = begin; i < end; i++) {
            float ti;
            int v;
            float be = 0.f, ga = 0.f;
            if (kind == 0) {
                const float* s = prim + 4 * (size_t)i;
                Inter it = interSphere(ray, v3(s[0], s[1], s[2]), s[3]);
                v = it.v; ti = it.t;
            } else {
                const float* p = prim + 12 * (size_t)i;
                TriInter it = interTriangle(ray, v3(p[0], p[1], p[2]), v3(p[4], p[5], p[6]), v3(p[8], p[9], p[10]));
                v = it.v; ti = it.t; be = it.beta; ga = it.gamma;
            }
            if (v && ti < ch.t) {
                ch.t = ti; ch.i = i; ch.beta = be; ch.gamma = ga;
                if (any) break;
            }
        }
        if (ch.i < 0xFFFFFFFFu) break;
        t = ray.maxt;
        int a = (t == t_next[0]) ? 0 : (t == t_next[1]) ? 1 : 2;
        t_next[a] += delta_t[a];
        if (t >= binter.tmax) break;
        slab[a] += dslab[a];
        if (slab[a] == limit[a]) break;
    }
    return ch;
}

/* ================================================================ kernels (extern "C" surface) */
int port_num_threads(void) { return omp_get_max_threads(); }
void port_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
uint port_a10_sizeofRay(void) { return (uint)sizeof(Ray); }   /* A10:440-446 */
uint port_a10_sizeofPoi(void) { return (uint)sizeof(Poi); }

void port_a10_initAcu(float* acu, uint total_rays) {   /* A10:448-456 */
#pragma omp parallel for
    for (long long id = 0; id < (long long)total_rays; id++) memset(acu + 4 * id, 0, 16);
}

static void initTracePixel(int* seeds, Ray* rays, Poi* pois, AABB bound, Camera cam, float focal_length, float lens_rad, uint rpp, uint col, uint row) {
    /* A10:458-543; `rays`/`pois` already point at the pixel's first slot */
    f3 focal_point = getFocalPoint(cam, (float)col, (float)row, focal_length);
    if (rpp > 1) {
        uint side = (uint)sqrtf((float)rpp);
        float delta = 1.0f / side;
        f2 coord;
        coord.y = delta / 2.0f;
        for (uint i = 0; i < side; i++) {
            coord.x = delta / 2.0f;
            for (uint j = 0; j < side; j++) {
                Ray ray = getThinLensRay(cam, focal_point, lens_rad, coord);
                AabbInter inter = interAABB(ray, bound);
                if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; } else { ray.mint = ray.maxt; }
                rays[i * side + j] = ray;
                coord.x += delta;
            }
            coord.y += delta;
        }
    } else {
        f2 coord;
        coord.y = getRand(seeds, col);   /* seeds[get_global_id(0)] = seeds[col]: the race of quirk Q7 */
        coord.x = getRand(seeds, col);
        Ray ray = getThinLensRay(cam, focal_point, lens_rad, coord);
        AabbInter inter = interAABB(ray, bound);
        if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; } else { ray.mint = ray.maxt; }
        rays[0] = ray;
    }
    for (uint i = 0; i < rpp; i++) {
        pois[i].matId = -1;
        pois[i].atte = v3(1.0f, 1.0f, 1.0f);
    }
}

void port_a10_initTrace(int* seeds, void* rays, void* pois, const float* bound, const float* cam16, float focal_length, float lens_rad,
                        uint rpp, uint cols, uint rows, int serial) {
    Camera cam = floatToCamera(cam16);
    AABB b = toAABB(bound);
    if (serial || rpp == 1) {   /* the only defined outcome of the seeds[col] race: row-major order */
        for (uint row = 0; row < rows; row++)
            for (uint col = 0; col < cols; col++) {
                size_t base = ((size_t)cam.cols * row + col) * rpp;
                initTracePixel(seeds, (Ray*)rays + base, (Poi*)pois + base, b, cam, focal_length, lens_rad, rpp, col, row);
            }
        return;
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (long long row = 0; row < (long long)rows; row++)
        for (uint col = 0; col < cols; col++) {
            size_t base = ((size_t)cam.cols * (uint)row + col) * rpp;
            initTracePixel(seeds, (Ray*)rays + base, (Poi*)pois + base, b, cam, focal_length, lens_rad, rpp, col, (uint)row);
        }
}

/* row tile [row0, row0+nrows): rays/pois hold only the tile's slots (rpp > 1, no RNG in initTrace) */
void port_a10_initTrace_rows(int* seeds, void* rays, void* pois, const float* bound, const float* cam16, float focal_length, float lens_rad,
                             uint rpp, uint cols, uint rows, uint row0, uint nrows) {
    Camera cam = floatToCamera(cam16);
    AABB b = toAABB(bound);
    (void)rows;
#pragma omp parallel for schedule(dynamic, 1)
    for (long long row = row0; row < (long long)row0 + nrows; row++)
        for (uint col = 0; col < cols; col++) {
            size_t base = ((size_t)cam.cols * ((uint)row - row0) + col) * rpp;
            initTracePixel(seeds, (Ray*)rays + base, (Poi*)pois + base, b, cam, focal_length, lens_rad, rpp, col, (uint)row);
        }
}

void port_a10_bouncePaths(void* pois_, void* rays_, int* seeds, uint total_rays) {   /* A10:545-598 */
    Poi* pois = (Poi*)pois_;
    Ray* rays = (Ray*)rays_;
#pragma omp parallel for schedule(dynamic, 2048)
    for (long long id = 0; id < (long long)total_rays; id++) {
        Poi poi = pois[id];
        Ray ray;
        memset(&ray, 0, sizeof ray);
        if (poi.matId >= 0) {
            f3 N = v3(fabsf(poi.normal.x), fabsf(poi.normal.y), fabsf(poi.normal.z));
            f3 B = poi.normal;
            float nmin = cl_min(cl_min(N.x, N.y), N.z);
            if (N.x == nmin) B.x = 1.0f; else if (N.y == nmin) B.y = 1.0f; else B.z = 1.0f;
            N = poi.normal;
            B = normalize(B);
            f3 T = cross(B, N);
            B = cross(N, T);
            f2 sxy;
            sxy.x = getRand(seeds, (size_t)id);
            sxy.y = getRand(seeds, (size_t)id);
            sxy = concentric_distort(sxy);
            float sz = sqrtf(cl_max(0.0f, 1.0f - sxy.x * sxy.x - sxy.y * sxy.y));
            ray.o = poi.p;
            ray.d = normalize(add(add(scl(sxy.x, T), scl(sxy.y, B)), scl(sz, N)));
            ray.mint = 0.0f;
            ray.maxt = HUGE_VALF;
        } else {
            ray.mint = ray.maxt = HUGE_VALF;
        }
        rays[id] = ray;
    }
}

void port_a10_lightRender(void* pois_, void* rays_, float* acu, const float* L, uint total_rays) {   /* A10:600-629 */
    Poi* pois = (Poi*)pois_;
    Ray* rays = (Ray*)rays_;
    f3 light_pos = arg3(L), light_normal = arg3(L + 3), irradiance = normalize(arg3(L + 6));
    float light_radius = L[9];
#pragma omp parallel for schedule(dynamic, 2048)
    for (long long id = 0; id < (long long)total_rays; id++) {
        Ray ray = rays[id];
        if (ray.mint == ray.maxt) continue;
        Inter inter = interLight(ray, light_pos, light_normal, light_radius);
        if (!inter.v || inter.t >= ray.maxt) continue;
        ray.mint = ray.maxt = HUGE_VALF;
        rays[id] = ray;
        pois[id].matId = -1;
        acu[4 * id] += irradiance.x; acu[4 * id + 1] += irradiance.y; acu[4 * id + 2] += irradiance.z; acu[4 * id + 3] += 1.0f;
    }
}

void port_a10_initShadowTrace(void* shadow_, void* pois_, uint total_rays, const float* L, int* seeds) {   /* A10:631-673 */
    Ray* shadow = (Ray*)shadow_;
    Poi* pois = (Poi*)pois_;
    f3 light_pos0 = arg3(L), T = arg3(L + 3), B = arg3(L + 6);
    float light_radius = L[9];
#pragma omp parallel for schedule(dynamic, 2048)
    for (long long id = 0; id < (long long)total_rays; id++) {
        Poi poi = pois[id];
        if (poi.matId < 0) {
            Ray dead;
            memset(&dead, 0, sizeof dead);
            dead.mint = dead.maxt = HUGE_VALF;
            shadow[id] = dead;
            continue;
        }
        poi.p = add(poi.p, lcs(poi.normal, 0.001f));
        f2 xy;
        xy.x = getRand(seeds, (size_t)id);
        xy.y = getRand(seeds, (size_t)id);
        xy = concentric_distort(xy);
        xy.x = xy.x * light_radius;
        xy.y = xy.y * light_radius;
        f3 light_pos = add(light_pos0, add(scl(xy.x, T), scl(xy.y, B)));
        shadow[id] = makeRay(poi.p, light_pos);
    }
}

/* closest hit: sphereTrace A10:675-800, triangleTrace :802-935, meshTrace :937-1070.  On a hit the kernel stores a Poi
 * whose atte member was never assigned; the contract (quirk Q1) is "a hit preserves pois[id].atte". */
static void closestTrace(uint total_rays, Poi* pois, Ray* rays, const float* prim, const float* normals, const uint* matid, uint scalar_matid,
                         const uint* box, AABB bound, uint n_slabs, int kind) {
#pragma omp parallel for schedule(dynamic, 2048)
    for (long long id = 0; id < (long long)total_rays; id++) {
        Ray ray = rays[id];
        if (ray.mint == ray.maxt) continue;
        AabbInter binter = interAABB(ray, bound);
        if (!binter.v) continue;
        Champ ch = walk(ray, binter, prim, box, bound, n_slabs, kind, 0);
        if (ch.i == 0xFFFFFFFFu) continue;
        rays[id].maxt = ch.t;
        f3 p = getPoint(ray, ch.t);
        f3 nrm;
        int m;
        if (kind == 0) {
            const float* s = prim + 4 * (size_t)ch.i;
            nrm = normalize(sub(p, v3(s[0], s[1], s[2])));
            m = (int)matid[ch.i];
        } else {
            const float* q = normals + 12 * (size_t)ch.i;
            nrm = normalize(interp(ch.beta, ch.gamma, v3(q[0], q[1], q[2]), v3(q[4], q[5], q[6]), v3(q[8], q[9], q[10])));
            m = matid ? (int)matid[ch.i] : (int)scalar_matid;
        }
        pois[id].p = p;
        pois[id].normal = nrm;
        pois[id].matId = m;
    }
}
void port_a10_sphereTrace(uint total_rays, void* pois, void* rays, const float* spheres, const uint* s_matid, const uint* s_box, const float* bound,
                          uint n_slabs) {
    closestTrace(total_rays, (Poi*)pois, (Ray*)rays, spheres, 0, s_matid, 0, s_box, toAABB(bound), n_slabs, 0);
}
void port_a10_triangleTrace(uint total_rays, void* pois, void* rays, const float* t_pos, const float* t_normal, const uint* t_matid, const uint* t_box,
                            const float* bound, uint n_slabs) {
    closestTrace(total_rays, (Poi*)pois, (Ray*)rays, t_pos, t_normal, t_matid, 0, t_box, toAABB(bound), n_slabs, 1);
}
void port_a10_meshTrace(uint total_rays, void* pois, void* rays, const float* t_pos, const float* t_normal, const uint* t_box, uint t_matid,
                        const float* bound, uint n_slabs) {
    closestTrace(total_rays, (Poi*)pois, (Ray*)rays, t_pos, t_normal, 0, t_matid, t_box, toAABB(bound), n_slabs, 1);
}

/* any hit: sphereShadowTrace A10:1073-1193, triangleShadowTrace :1195-1321 */
static void anyTrace(uint total_rays, Ray* shadow, const float* prim, const uint* box, AABB bound, uint n_slabs, int kind) {
#pragma omp parallel for schedule(dynamic, 2048)
    for (long long id = 0; id < (long long)total_rays; id++) {
        Ray ray = shadow[id];
        if (ray.mint == ray.maxt) continue;
        AabbInter binter = interAABB(ray, bound);
        if (!binter.v) continue;
        Champ ch = walk(ray, binter, prim, box, bound, n_slabs, kind, 1);
        if (ch.i < 0xFFFFFFFFu) { shadow[id].maxt = ch.t; shadow[id].mint = ch.t; }
        else shadow[id].maxt = ch.t;
    }
}
void port_a10_sphereShadowTrace(uint total_rays, void* shadow, const float* spheres, const uint* s_box, const float* bound, uint n_slabs) {
    anyTrace(total_rays, (Ray*)shadow, spheres, s_box, toAABB(bound), n_slabs, 0);
}
void port_a10_triangleShadowTrace(uint total_rays, void* shadow, const float* t_pos, const uint* t_box, const float* bound, uint n_slabs) {
    anyTrace(total_rays, (Ray*)shadow, t_pos, t_box, toAABB(bound), n_slabs, 1);
}

void port_a10_sceneRender(float* acu, void* pois_, void* shadow_, const float* material, const float* L, uint total_rays) {   /* A10:1323-1364 */
    Poi* pois = (Poi*)pois_;
    const Ray* shadow = (const Ray*)shadow_;
    f3 lpos = arg3(L), lnor = arg3(L + 3), es = arg3(L + 6);
    float area = L[9];
#pragma omp parallel for schedule(dynamic, 2048)
    for (long long id = 0; id < (long long)total_rays; id++) {
        Poi poi = pois[id];
        if (poi.matId < 0) continue;
        f3 shade = v3(0.0f, 0.0f, 0.0f);
        Ray sr = shadow[id];
        if (sr.maxt != sr.mint) {
            float r = length3(sub(poi.p, lpos));
            float cosx = cl_clamp(dot(sr.d, poi.normal), 0.0f, 1.0f);
            float cosy = cl_clamp(dot(neg(sr.d), lnor), 0.0f, 1.0f);
            shade = scl(area * ((cosx * cosy) / (r * r)), es);
        }
        f3 color = arg3(material + 4 * (size_t)poi.matId);
        pois[id].atte = mul(pois[id].atte, color);   /* per light (Q3) */
        color = mul(color, poi.atte);
        color = mul(color, shade);
        acu[4 * id] += color.x; acu[4 * id + 1] += color.y; acu[4 * id + 2] += color.z; acu[4 * id + 3] += 1.0f;
    }
}

void port_a10_copyToPixel(unsigned char* pixel, const float* acu, float m, uint pixels, uint rpp) {   /* A10:1366-1386 */
#pragma omp parallel for
    for (long long id = 0; id < (long long)pixels; id++) {
        const float* a = acu + 4 * (size_t)id * rpp;
        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint i = 0; i < rpp; i++)
            for (int k = 0; k < 4; k++) c[k] += a[4 * i + k];
        for (int k = 0; k < 3; k++) {
            c[k] *= 255.0f * m;
            c[k] *= 1.8f;
            c[k] = cl_clamp(c[k], 0.0f, 255.0f);
            pixel[4 * id + k] = (unsigned char)c[k];
        }
        pixel[4 * id + 3] = 255;
    }
}
