// rt_kernels_a10.cu -- one CUDA kernel + C-ABI launcher per OpenCL kernel of
// Assign10-Path_Tracing/code.cl (the superset assignment), same names, argument order and
// buffer layouts, one thread per ray slot.  This is the kernel-by-kernel drop-in layer; the
// fused wavefront renderer (rt_frame.cu) is built from the same device functions.
#include "rt_device.cuh"
#include "rt_internal.h"

using namespace rt;

namespace {

constexpr unsigned kBlock = 256;


// ---------------------------------------------------------------------------- initAcu
__global__ void k_initAcu(float4* acu, unsigned total_rays) {   // A10/code.cl:448-456
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
    acu[id] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------- initTrace
// A10/code.cl:458-543.  The reference runs one work-item per pixel looping over the
// rays_per_pixel slots; here one thread per slot recomputes the pixel's focal point and
// replays the fp32 `coord += delta` accumulation up to its own (i, j), which yields the
// same coordinates bit for bit.
__global__ void k_initTrace_strat(Ray* rays, Poi10* pois, AabbArg bound_a, CamArg fcam, float focal_length, float lens_rad,
                                  unsigned rays_per_pixel, unsigned long long total) {
    unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    Camera cam = floatToCamera(fcam.v);
    AABB bound = toAABB(bound_a);
    unsigned pix = (unsigned)(id / rays_per_pixel);
    unsigned k = (unsigned)(id % rays_per_pixel);
    unsigned col = pix % cam.cols, row = pix / cam.cols;
    unsigned side = (unsigned)sqrtf((float)rays_per_pixel);
    // Poi reset for every slot (A10/code.cl:538-542): matId = -1, atte = 1; p/normal untouched.
    float4* pq = reinterpret_cast<float4*>(pois + id);
    pq[2] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    pois[id].matId = -1;
    if (k >= side * side) return;   // slots beyond the stratified grid are never written by the reference
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
    AabbHit inter = interAABB(ray.o, ray.d, bound);
    if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
    else { ray.mint = RT_INF; ray.maxt = RT_INF; }
    storeRay(rays + id, ray);
}

// rays_per_pixel == 1: the reference draws from seeds[get_global_id(0)] = seeds[col] in a
// 2-D launch, i.e. every row of a column races on one seed (quirk Q7).  The only defined
// outcome is a serial one; we implement row-major order: column `col` hands draws
// (2*row, 2*row+1) of its stream to row `row` (y first, then x -- A10/code.cl:513-515).
__global__ void k_initTrace_rpp1_coords(int* seeds, float2* coords, unsigned cols, unsigned rows) {
    unsigned col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    int seed = seeds[col];
    for (unsigned row = 0; row < rows; row++) {
        float y = nextRand(seed);
        float x = nextRand(seed);
        coords[(size_t)row * cols + col] = make_float2(x, y);
    }
    seeds[col] = seed;
}
__global__ void k_initTrace_rpp1(const float2* coords, Ray* rays, Poi10* pois, AabbArg bound_a, CamArg fcam, float focal_length,
                                 float lens_rad, unsigned total) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    Camera cam = floatToCamera(fcam.v);
    AABB bound = toAABB(bound_a);
    unsigned col = id % cam.cols, row = id / cam.cols;
    f3 focal_point = getFocalPoint(cam, (float)col, (float)row, focal_length);
    float2 c = coords[id];
    f2 coord; coord.x = c.x; coord.y = c.y;
    RayR ray;
    getThinLensRay(cam, focal_point, lens_rad, coord, ray.o, ray.d);
    AabbHit inter = interAABB(ray.o, ray.d, bound);
    if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
    else { ray.mint = RT_INF; ray.maxt = RT_INF; }
    storeRay(rays + id, ray);
    float4* pq = reinterpret_cast<float4*>(pois + id);
    pq[2] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    pois[id].matId = -1;
}

// ---------------------------------------------------------------------------- bouncePaths
__global__ void k_bouncePaths(const Poi10* pois, Ray* rays, int* seeds, unsigned total_rays) {   // A10/code.cl:581-598
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
    const float4* pq = reinterpret_cast<const float4*>(pois + id);
    int matId = pois[id].matId;
    if (matId >= 0) {
        float4 p = pq[0], n = pq[1];
        int seed = seeds[id];
        RayR ray;
        getHemisphereRay(mk3(p.x, p.y, p.z), mk3(n.x, n.y, n.z), seed, ray.o, ray.d);
        ray.mint = 0.0f;
        ray.maxt = RT_INF;
        seeds[id] = seed;
        storeRay(rays + id, ray);
    } else {
        storeDeadRay(rays + id);
    }
}

// ---------------------------------------------------------------------------- lightRender
__global__ void k_lightRender(Poi10* pois, Ray* rays, float4* acu, LightArg L, unsigned total_rays) {   // A10/code.cl:600-629
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
    RayR ray = loadRay(rays + id);
    if (ray.mint == ray.maxt) return;
    f3 light_pos = mk3(L.v[0], L.v[1], L.v[2]);
    f3 light_normal = mk3(L.v[3], L.v[4], L.v[5]);
    f3 irradiance = normalize(mk3(L.v[6], L.v[7], L.v[8]));
    float light_radius = L.v[9];
    float t;
    if (!interLight(ray.o, ray.d, light_pos, light_normal, light_radius, t) || t >= ray.maxt) return;
    reinterpret_cast<float4*>(rays + id)[2] = make_float4(RT_INF, RT_INF, 0.f, 0.f);   // o, d kept
    pois[id].matId = -1;
    float4 a = acu[id];
    acu[id] = make_float4(a.x + irradiance.x, a.y + irradiance.y, a.z + irradiance.z, a.w + 1.0f);
}

// ---------------------------------------------------------------------------- initShadowTrace
__global__ void k_initShadowTrace(Ray* shadow_rays, const Poi10* pois, unsigned total_rays, LightArg L, int* seeds) {   // A10/code.cl:631-673
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
    if (pois[id].matId < 0) {
        storeDeadRay(shadow_rays + id);
        return;
    }
    const float4* pq = reinterpret_cast<const float4*>(pois + id);
    float4 p = pq[0], n = pq[1];
    int seed = seeds[id];
    RayR sr = makeShadowRay(mk3(p.x, p.y, p.z), mk3(n.x, n.y, n.z), L, seed);
    seeds[id] = seed;
    storeRay(shadow_rays + id, sr);
}

// ---------------------------------------------------------------------------- closest-hit traces
// sphereTrace / triangleTrace / meshTrace, A10/code.cl:675-800, 802-935, 937-1070.
template <int PRIM, bool STATS>
__global__ void k_closestTrace(unsigned total_rays, Poi10* pois, Ray* rays, GridView g, const float4* normals, const unsigned* matid,
                               unsigned scalar_matid, StatPtrs sp) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
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
    Hit h = gridWalk<PRIM, false, true, STATS>(ray.o, ray.d, ray.maxt, g, binter, &ws);
    if (STATS) {
        tallyWalk(sp.totals, 1, ws, h.i != 0xFFFFFFFFu);
        if (sp.hit) sp.hit[id] = h.i;
        if (sp.cells) sp.cells[id] = (unsigned)ws.cells;
        if (sp.tests) sp.tests[id] = (unsigned)ws.tests;
    }
    if (h.i == 0xFFFFFFFFu) return;
    rays[id].maxt = h.t;
    f3 p = getPoint(ray.o, ray.d, h.t);
    f3 nrm;
    int m;
    if (PRIM == PRIM_SPHERE) {
        float4 s = __ldg(g.prim + h.i);
        nrm = normalize(p - mk3(s.x, s.y, s.z));
        m = (int)__ldg(matid + h.i);
    } else {
        float4 n0 = __ldg(normals + 3 * h.i), n1 = __ldg(normals + 3 * h.i + 1), n2 = __ldg(normals + 3 * h.i + 2);
        nrm = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
        m = matid ? (int)__ldg(matid + h.i) : (int)scalar_matid;
    }
    // Quirk Q1: the reference stores a Poi whose atte was never assigned; the contract is
    // "a hit preserves atte" -- write p, normal and matId only.
    float4* pq = reinterpret_cast<float4*>(pois + id);
    pq[0] = make_float4(p.x, p.y, p.z, 0.f);
    pq[1] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
    pois[id].matId = m;
}

// sphereShadowTrace / triangleShadowTrace, A10/code.cl:1073-1193, 1195-1321.
template <int PRIM, bool STATS>
__global__ void k_anyTrace(unsigned total_rays, Ray* shadow_rays, GridView g, StatPtrs sp) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
    if (STATS) {
        if (sp.hit) sp.hit[id] = 0xFFFFFFFFu;
        if (sp.cells) sp.cells[id] = 0;
        if (sp.tests) sp.tests[id] = 0;
    }
    RayR ray = loadRay(shadow_rays + id);
    if (ray.mint == ray.maxt) return;
    AabbHit binter = interAABB(ray.o, ray.d, g.bound);
    WalkStats ws = {0, 0, 0};
    if (!binter.v) { if (STATS) tallyWalk(sp.totals, 0, ws, false); return; }
    Hit h = gridWalk<PRIM, true, true, STATS>(ray.o, ray.d, ray.maxt, g, binter, &ws);
    if (STATS) {
        tallyWalk(sp.totals, 1, ws, h.i != 0xFFFFFFFFu);
        if (sp.hit) sp.hit[id] = h.i;
        if (sp.cells) sp.cells[id] = (unsigned)ws.cells;
        if (sp.tests) sp.tests[id] = (unsigned)ws.tests;
    }
    if (h.i != 0xFFFFFFFFu) {
        shadow_rays[id].maxt = h.t;
        shadow_rays[id].mint = h.t;
    } else {
        shadow_rays[id].maxt = h.t;   // "way is free": champ_t is the unchanged stored maxt
    }
}

// ---------------------------------------------------------------------------- sceneRender
__global__ void k_sceneRender(float4* acu, Poi10* pois, const Ray* shadow_rays, const float4* material, LightArg L,
                              unsigned total_rays) {   // A10/code.cl:1323-1364
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total_rays) return;
    int matId = pois[id].matId;
    if (matId < 0) return;
    float4* pq = reinterpret_cast<float4*>(pois + id);
    float4 p = pq[0], n = pq[1], at = pq[2];
    RayR sr = loadRay(shadow_rays + id);
    f3 shade = neeShade(mk3(p.x, p.y, p.z), mk3(n.x, n.y, n.z), sr.d, sr.maxt != sr.mint, L);
    float4 color = __ldg(material + matId);
    f3 c = mk3(color.x, color.y, color.z);
    f3 atte = mk3(at.x, at.y, at.z);
    f3 na = atte * c;                 // pois[id].atte *= color.s012  (per light -- quirk Q3)
    pq[2] = make_float4(na.x, na.y, na.z, at.w);
    f3 contrib = (c * atte) * shade;  // color.s012 *= poi.atte; color.s012 *= shade
    float4 a = acu[id];
    acu[id] = make_float4(a.x + contrib.x, a.y + contrib.y, a.z + contrib.z, a.w + 1.0f);
}

// ---------------------------------------------------------------------------- copyToPixel
__global__ void k_copyToPixel(uchar4* pixel, const float4* acu, float m, unsigned pixels, unsigned rays_per_pixel) {   // A10/code.cl:1366-1386
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pixels) return;
    const float4* a = acu + (size_t)id * rays_per_pixel;
    float4 color = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned i = 0; i < rays_per_pixel; i++) {
        float4 v = a[i];
        color.x += v.x; color.y += v.y; color.z += v.z; color.w += v.w;
    }
    float s = 255.0f * m;
    color.x *= s; color.y *= s; color.z *= s;
    color.x *= 1.8f; color.y *= 1.8f; color.z *= 1.8f;
    color.x = cl_clamp(color.x, 0.0f, 255.0f);
    color.y = cl_clamp(color.y, 0.0f, 255.0f);
    color.z = cl_clamp(color.z, 0.0f, 255.0f);
    pixel[id] = make_uchar4((unsigned char)color.x, (unsigned char)color.y, (unsigned char)color.z, 255);
}

AabbArg mkAabb(const float* b) { AabbArg a; memcpy(a.v, b, sizeof a.v); return a; }
CamArg mkCam(const float* c) { CamArg a; memcpy(a.v, c, sizeof a.v); return a; }
LightArg mkLight(const float* c) { LightArg a; memcpy(a.v, c, sizeof a.v); return a; }

GridView mkGrid(const void* prim, const void* box, const float* bound, unsigned n) {
    GridView g;
    g.prim = (const float4*)prim;
    g.box = (const unsigned*)box;
    g.occ = nullptr;
    AabbArg a = mkAabb(bound);
    g.bound.pmin.x = a.v[0]; g.bound.pmin.y = a.v[1]; g.bound.pmin.z = a.v[2];
    g.bound.pmax.x = a.v[4]; g.bound.pmax.y = a.v[5]; g.bound.pmax.z = a.v[6];
    g.n = n;
    return g;
}

template <int PRIM>
int launchClosest(rt_ctx* ctx, unsigned total_rays, void* pois, void* rays, GridView g, const void* normals, const void* matid,
                  unsigned scalar_matid, const char* name) {
    if (!total_rays) return RT_OK;
    StatPtrs sp = {ctx->st_hit, ctx->st_cells, ctx->st_tests, ctx->st_totals};
    bool stats = sp.hit || sp.cells || sp.tests || sp.totals;
    dim3 grid(rt_blocks(total_rays, kBlock));
    if (stats)
        k_closestTrace<PRIM, true><<<grid, kBlock, 0, ctx->stream>>>(total_rays, (Poi10*)pois, (Ray*)rays, g, (const float4*)normals,
                                                                      (const unsigned*)matid, scalar_matid, sp);
    else
        k_closestTrace<PRIM, false><<<grid, kBlock, 0, ctx->stream>>>(total_rays, (Poi10*)pois, (Ray*)rays, g, (const float4*)normals,
                                                                       (const unsigned*)matid, scalar_matid, sp);
    RT_LAUNCH_CHECK(ctx, name);
    return RT_OK;
}

template <int PRIM>
int launchAny(rt_ctx* ctx, unsigned total_rays, void* shadow, GridView g, const char* name) {
    if (!total_rays) return RT_OK;
    StatPtrs sp = {ctx->st_hit, ctx->st_cells, ctx->st_tests, ctx->st_totals};
    bool stats = sp.hit || sp.cells || sp.tests || sp.totals;
    dim3 grid(rt_blocks(total_rays, kBlock));
    if (stats)
        k_anyTrace<PRIM, true><<<grid, kBlock, 0, ctx->stream>>>(total_rays, (Ray*)shadow, g, sp);
    else
        k_anyTrace<PRIM, false><<<grid, kBlock, 0, ctx->stream>>>(total_rays, (Ray*)shadow, g, sp);
    RT_LAUNCH_CHECK(ctx, name);
    return RT_OK;
}

}  // namespace

extern "C" {

int rt_a10_initAcu(rt_ctx* ctx, void* acu, unsigned total_rays) {
    RT_CHECK_CTX(ctx);
    if (!acu) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_initAcu<<<rt_blocks(total_rays, kBlock), kBlock, 0, ctx->stream>>>((float4*)acu, total_rays);
    RT_LAUNCH_CHECK(ctx, "initAcu");
    return RT_OK;
}

int rt_a10_initTrace(rt_ctx* ctx, void* seeds, void* rays, void* pois, const float* bound, const float* fcam, float focal_length,
                     float lens_rad, unsigned rays_per_pixel) {
    RT_CHECK_CTX(ctx);
    if (!rays || !pois || !bound || !fcam || !rays_per_pixel) return RT_ERR_INVALID;
    unsigned cols = (unsigned)fcam[14], rows = (unsigned)fcam[15];
    unsigned long long total = (unsigned long long)cols * rows * rays_per_pixel;
    if (!total) return RT_OK;
    if (total > 0xFFFFFFFFull) return rt_fail(ctx, RT_ERR_INVALID, "initTrace: total_rays exceeds the reference's uint range");
    if (rays_per_pixel > 1) {
        k_initTrace_strat<<<rt_blocks(total, kBlock), kBlock, 0, ctx->stream>>>((Ray*)rays, (Poi10*)pois, mkAabb(bound), mkCam(fcam),
                                                                                 focal_length, lens_rad, rays_per_pixel, total);
        RT_LAUNCH_CHECK(ctx, "initTrace");
    } else {
        if (!seeds) return RT_ERR_INVALID;
        float2* coords = nullptr;
        RT_CUDA(ctx, rt_scratch_alloc(ctx, (void**)&coords, sizeof(float2) * (size_t)total));
        k_initTrace_rpp1_coords<<<rt_blocks(cols, 64), 64, 0, ctx->stream>>>((int*)seeds, coords, cols, rows);
        RT_LAUNCH_CHECK(ctx, "initTrace(seeds)");
        k_initTrace_rpp1<<<rt_blocks(total, kBlock), kBlock, 0, ctx->stream>>>(coords, (Ray*)rays, (Poi10*)pois, mkAabb(bound), mkCam(fcam),
                                                                                focal_length, lens_rad, (unsigned)total);
        RT_LAUNCH_CHECK(ctx, "initTrace");
        RT_CUDA(ctx, cudaFreeAsync(coords, ctx->stream));
    }
    return RT_OK;
}

int rt_a10_bouncePaths(rt_ctx* ctx, void* pois, void* rays, void* seeds, unsigned total_rays) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !seeds) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_bouncePaths<<<rt_blocks(total_rays, kBlock), kBlock, 0, ctx->stream>>>((const Poi10*)pois, (Ray*)rays, (int*)seeds, total_rays);
    RT_LAUNCH_CHECK(ctx, "bouncePaths");
    return RT_OK;
}

int rt_a10_lightRender(rt_ctx* ctx, void* pois, void* rays, void* acu, const float* light_info, unsigned total_rays) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !acu || !light_info) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_lightRender<<<rt_blocks(total_rays, kBlock), kBlock, 0, ctx->stream>>>((Poi10*)pois, (Ray*)rays, (float4*)acu, mkLight(light_info), total_rays);
    RT_LAUNCH_CHECK(ctx, "lightRender");
    return RT_OK;
}

int rt_a10_initShadowTrace(rt_ctx* ctx, void* shadow_rays, void* pois, unsigned total_rays, const float* light_info, void* seeds) {
    RT_CHECK_CTX(ctx);
    if (!shadow_rays || !pois || !light_info || !seeds) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_initShadowTrace<<<rt_blocks(total_rays, kBlock), kBlock, 0, ctx->stream>>>((Ray*)shadow_rays, (const Poi10*)pois, total_rays,
                                                                                  mkLight(light_info), (int*)seeds);
    RT_LAUNCH_CHECK(ctx, "initShadowTrace");
    return RT_OK;
}

int rt_a10_sphereTrace(rt_ctx* ctx, unsigned total_rays, void* pois, void* rays, const void* spheres, const void* s_matid,
                       const void* s_box_size, const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !spheres || !s_matid || !s_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    return launchClosest<PRIM_SPHERE>(ctx, total_rays, pois, rays, mkGrid(spheres, s_box_size, bound, n_slabs), nullptr, s_matid, 0, "sphereTrace");
}

int rt_a10_triangleTrace(rt_ctx* ctx, unsigned total_rays, void* pois, void* rays, const void* t_pos, const void* t_normal,
                         const void* t_matid, const void* t_box_size, const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !t_pos || !t_normal || !t_matid || !t_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    return launchClosest<PRIM_TRIANGLE>(ctx, total_rays, pois, rays, mkGrid(t_pos, t_box_size, bound, n_slabs), t_normal, t_matid, 0, "triangleTrace");
}

int rt_a10_meshTrace(rt_ctx* ctx, unsigned total_rays, void* pois, void* rays, const void* t_pos, const void* t_normal,
                     const void* t_box_size, unsigned t_matid, const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!pois || !rays || !t_pos || !t_normal || !t_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    return launchClosest<PRIM_TRIANGLE>(ctx, total_rays, pois, rays, mkGrid(t_pos, t_box_size, bound, n_slabs), t_normal, nullptr, t_matid, "meshTrace");
}

int rt_a10_sphereShadowTrace(rt_ctx* ctx, unsigned total_rays, void* shadow_rays, const void* spheres, const void* s_box_size,
                             const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!shadow_rays || !spheres || !s_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    return launchAny<PRIM_SPHERE>(ctx, total_rays, shadow_rays, mkGrid(spheres, s_box_size, bound, n_slabs), "sphereShadowTrace");
}

int rt_a10_triangleShadowTrace(rt_ctx* ctx, unsigned total_rays, void* shadow_rays, const void* t_pos, const void* t_box_size,
                               const float* bound, unsigned n_slabs) {
    RT_CHECK_CTX(ctx);
    if (!shadow_rays || !t_pos || !t_box_size || !bound || !n_slabs) return RT_ERR_INVALID;
    return launchAny<PRIM_TRIANGLE>(ctx, total_rays, shadow_rays, mkGrid(t_pos, t_box_size, bound, n_slabs), "triangleShadowTrace");
}

int rt_a10_sceneRender(rt_ctx* ctx, void* acu, void* pois, const void* shadow_rays, const void* material, const float* light_info,
                       unsigned total_rays) {
    RT_CHECK_CTX(ctx);
    if (!acu || !pois || !shadow_rays || !material || !light_info) return RT_ERR_INVALID;
    if (!total_rays) return RT_OK;
    k_sceneRender<<<rt_blocks(total_rays, kBlock), kBlock, 0, ctx->stream>>>((float4*)acu, (Poi10*)pois, (const Ray*)shadow_rays,
                                                                              (const float4*)material, mkLight(light_info), total_rays);
    RT_LAUNCH_CHECK(ctx, "sceneRender");
    return RT_OK;
}

int rt_a10_copyToPixel(rt_ctx* ctx, void* pixel, const void* acu, float m, unsigned pixels, unsigned rays_per_pixel) {
    RT_CHECK_CTX(ctx);
    if (!pixel || !acu) return RT_ERR_INVALID;
    if (!pixels) return RT_OK;
    k_copyToPixel<<<rt_blocks(pixels, kBlock), kBlock, 0, ctx->stream>>>((uchar4*)pixel, (const float4*)acu, m, pixels, rays_per_pixel);
    RT_LAUNCH_CHECK(ctx, "copyToPixel");
    return RT_OK;
}

}  // extern "C"
