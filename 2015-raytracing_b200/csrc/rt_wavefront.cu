// rt_wavefront.cu -- fused path-tracing kernels for the Assignment-10 frame.
//
// The reference enqueues ~90 OpenCL kernels per pass, each streaming 48-112 B of per-slot
// state through global memory (executeRender, Assign10-Path_Tracing/code.js:1806-1854).  Every
// one of those kernels touches only its own slot, so a whole pass of a slot can run in ONE
// thread with the ray, the hit record, the shadow ray, the RNG state and the accumulator in
// registers: per slot and pass only the seed (4 B) and the accumulator (16 B) are read and
// written.  The kernel below does that ("megakernel", mode 2).  It executes, per slot, the
// same device functions in the same order as the kernel-by-kernel path, so it is bit-exact
// with it (tests/test_gpu_a10.py).
#include "rt_frame.h"

using namespace rt;

namespace {

constexpr int kMaxSets = 8;
constexpr int kMaxLights = 8;

struct SetDev {
    GridView g;
    const float4* normals;
    const unsigned* matid;     // per-reference material (spheres, scene triangles) or null
    unsigned scalar_matid;     // meshes
    int kind;                  // PRIM_SPHERE / PRIM_TRIANGLE
    int use_occ;
};
struct LightDev { LightArg shadow, scene, light; };
struct SceneDev {
    SetDev sets[kMaxSets];
    LightDev lights[kMaxLights];
    int n_sets, n_lights;
    const float4* materials;
    AabbArg bound;
};

struct PathArgs {
    CamArg cam;
    float focal_length, lens_rad;
    unsigned rays_per_pixel, slot_begin, slots_pp, depth;
    size_t pixel_base;
    unsigned n_local;
    const float2* rpp1_coords;
    int* seeds;               // tile-local
    float4* acu;              // tile-local
    unsigned long long* counters;
    unsigned long long* profile;
};

struct PoiR { f3 p, n, atte; int matId; };

// closest hit of one ray against one geometry set; updates ray.maxt and the hit record
// (sphereTrace / triangleTrace / meshTrace, A10/code.cl:675-1070)
template <bool OCC, bool STATS>
RT_DEV void closestSet(const SetDev& s, RayR& ray, PoiR& poi, unsigned* prof) {
    if (ray.mint == ray.maxt) return;
    if (STATS) prof[0]++;
    AabbHit binter = interAABB(ray.o, ray.d, s.g.bound);
    if (!binter.v) return;
    Hit h;
    WalkStats ws = {0, 0};
    if (s.kind == PRIM_SPHERE) h = gridWalk<PRIM_SPHERE, false, true, STATS, OCC>(ray.o, ray.d, ray.maxt, s.g, binter, &ws);
    else h = gridWalk<PRIM_TRIANGLE, false, true, STATS, OCC>(ray.o, ray.d, ray.maxt, s.g, binter, &ws);
    if (STATS) {
        prof[1]++;
        prof[2] += (unsigned)ws.cells;
        prof[s.kind == PRIM_SPHERE ? 3 : 4] += (unsigned)ws.tests;
        if (h.i != 0xFFFFFFFFu) prof[s.kind == PRIM_SPHERE ? 5 : (s.matid ? 6 : 7)]++;
    }
    if (h.i == 0xFFFFFFFFu) return;
    ray.maxt = h.t;
    poi.p = getPoint(ray.o, ray.d, h.t);
    if (s.kind == PRIM_SPHERE) {
        float4 sp = __ldg(s.g.prim + h.i);
        poi.n = normalize(poi.p - mk3(sp.x, sp.y, sp.z));
        poi.matId = (int)__ldg(s.matid + h.i);
    } else {
        float4 n0 = __ldg(s.normals + 3 * h.i), n1 = __ldg(s.normals + 3 * h.i + 1), n2 = __ldg(s.normals + 3 * h.i + 2);
        poi.n = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
        poi.matId = s.matid ? (int)__ldg(s.matid + h.i) : (int)s.scalar_matid;
    }
}

// any hit of a shadow ray against one set (sphereShadowTrace / triangleShadowTrace, :1073-1321)
template <bool OCC, bool STATS>
RT_DEV void anySet(const SetDev& s, RayR& sr, unsigned* prof) {
    if (sr.mint == sr.maxt) return;
    if (STATS) prof[8]++;
    AabbHit binter = interAABB(sr.o, sr.d, s.g.bound);
    if (!binter.v) return;
    Hit h;
    WalkStats ws = {0, 0};
    if (s.kind == PRIM_SPHERE) h = gridWalk<PRIM_SPHERE, true, true, STATS, OCC>(sr.o, sr.d, sr.maxt, s.g, binter, &ws);
    else h = gridWalk<PRIM_TRIANGLE, true, true, STATS, OCC>(sr.o, sr.d, sr.maxt, s.g, binter, &ws);
    if (STATS) {
        prof[9]++;
        prof[10] += (unsigned)ws.cells;
        prof[s.kind == PRIM_SPHERE ? 11 : 12] += (unsigned)ws.tests;
        if (h.i != 0xFFFFFFFFu) prof[13]++;
    }
    if (h.i != 0xFFFFFFFFu) { sr.maxt = h.t; sr.mint = h.t; }
    else sr.maxt = h.t;
}

template <bool STATS>
RT_DEV void closestAllSets(const SceneDev& sc, RayR& ray, PoiR& poi, unsigned* prof) {
    for (int s = 0; s < sc.n_sets; s++) {
        if (sc.sets[s].use_occ) closestSet<true, STATS>(sc.sets[s], ray, poi, prof);
        else closestSet<false, STATS>(sc.sets[s], ray, poi, prof);
    }
}

// one light: initShadowTrace + shadow traces + sceneRender (A10/code.cl:631-673, 1073-1364)
template <bool STATS>
RT_DEV void shadeLight(const SceneDev& sc, const LightDev& L, PoiR& poi, int& seed, float4& acu, unsigned& n_any, unsigned* prof) {
    if (poi.matId < 0) return;
    RayR sr = makeShadowRay(poi.p, poi.n, L.shadow, seed);
    if (sr.mint != sr.maxt) n_any++;
    for (int s = 0; s < sc.n_sets; s++) {
        if (sc.sets[s].use_occ) anySet<true, STATS>(sc.sets[s], sr, prof);
        else anySet<false, STATS>(sc.sets[s], sr, prof);
    }
    f3 shade = neeShade(poi.p, poi.n, sr.d, sr.maxt != sr.mint, L.scene);
    float4 color = __ldg(sc.materials + poi.matId);
    f3 c = mk3(color.x, color.y, color.z);
    f3 contrib = (c * poi.atte) * shade;
    poi.atte = poi.atte * c;   // per light (Q3)
    acu = make_float4(acu.x + contrib.x, acu.y + contrib.y, acu.z + contrib.z, acu.w + 1.0f);
}

template <bool STATS>
__global__ void __launch_bounds__(128) k_pathMega(const __grid_constant__ SceneDev sc, const __grid_constant__ PathArgs a) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_closest = 0, n_any = 0;
    unsigned prof[16];
    if (STATS) {
        for (int q = 0; q < 16; q++) prof[q] = 0;
        if (id < a.n_local) prof[14] = 1;
    }
    if (id < a.n_local) {
        // ---- initTrace (A10/code.cl:458-543)
        Camera cam = floatToCamera(a.cam.v);
        AABB bound = toAABB(sc.bound);
        size_t pix = a.pixel_base + id / a.slots_pp;
        unsigned k = a.slot_begin + id % a.slots_pp;
        unsigned col = (unsigned)(pix % cam.cols), row = (unsigned)(pix / cam.cols);
        RayR ray;
        ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 0.f);
        ray.mint = RT_INF; ray.maxt = RT_INF;
        PoiR poi;
        poi.p = mk3(0.f, 0.f, 0.f); poi.n = mk3(0.f, 0.f, 0.f);
        poi.atte = mk3(1.0f, 1.0f, 1.0f);
        poi.matId = -1;
        bool have_ray = true;
        f2 coord;
        if (a.rays_per_pixel > 1) {
            unsigned side = (unsigned)sqrtf((float)a.rays_per_pixel);
            if (k >= side * side) have_ray = false;
            unsigned i = k / side, j = k % side;
            float delta = 1.0f / (float)side;
            coord.y = delta / 2.0f;
            for (unsigned q = 0; q < i; q++) coord.y += delta;
            coord.x = delta / 2.0f;
            for (unsigned q = 0; q < j; q++) coord.x += delta;
        } else {
            float2 c = a.rpp1_coords[pix];
            coord.x = c.x; coord.y = c.y;
        }
        if (have_ray) {
            f3 focal_point = getFocalPoint(cam, (float)col, (float)row, a.focal_length);
            getThinLensRay(cam, focal_point, a.lens_rad, coord, ray.o, ray.d);
            AabbHit inter = interAABB(ray.o, ray.d, bound);
            if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
        }
        int seed = a.seeds[id];
        float4 acu = a.acu[id];
        // ---- primary segment
        if (ray.mint != ray.maxt) n_closest++;
        closestAllSets<STATS>(sc, ray, poi, prof);
        for (int l = 0; l < sc.n_lights; l++) {   // lightRender, A10/code.cl:600-629
            if (ray.mint == ray.maxt) continue;
            const LightArg& L = sc.lights[l].light;
            f3 irradiance = normalize(mk3(L.v[6], L.v[7], L.v[8]));
            float t;
            if (!interLight(ray.o, ray.d, mk3(L.v[0], L.v[1], L.v[2]), mk3(L.v[3], L.v[4], L.v[5]), L.v[9], t) || t >= ray.maxt) continue;
            ray.mint = RT_INF; ray.maxt = RT_INF;
            poi.matId = -1;
            acu = make_float4(acu.x + irradiance.x, acu.y + irradiance.y, acu.z + irradiance.z, acu.w + 1.0f);
        }
        for (int l = 0; l < sc.n_lights; l++) shadeLight<STATS>(sc, sc.lights[l], poi, seed, acu, n_any, prof);
        // ---- bounces (A10/code.js:1829-1846)
        for (unsigned j = 0; j < a.depth; j++) {
            if (poi.matId >= 0) {   // bouncePaths, A10/code.cl:581-598
                getHemisphereRay(poi.p, poi.n, seed, ray.o, ray.d);
                ray.mint = 0.0f; ray.maxt = RT_INF;
                n_closest++;
            } else {
                ray.mint = RT_INF; ray.maxt = RT_INF;
            }
            closestAllSets<STATS>(sc, ray, poi, prof);
            for (int l = 0; l < sc.n_lights; l++) shadeLight<STATS>(sc, sc.lights[l], poi, seed, acu, n_any, prof);
        }
        a.seeds[id] = seed;
        a.acu[id] = acu;
    }
    // ray counters: warp reduce, one atomic per warp
    for (int d = 16; d > 0; d >>= 1) {
        n_closest += __shfl_down_sync(0xffffffffu, n_closest, d);
        n_any += __shfl_down_sync(0xffffffffu, n_any, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_closest) atomicAdd(a.counters + 0, (unsigned long long)n_closest);
        if (n_any) atomicAdd(a.counters + 1, (unsigned long long)n_any);
    }
    if (STATS) {
        for (int q = 0; q < 16; q++) {
            unsigned v = prof[q];
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(a.profile + q, (unsigned long long)v);
        }
    }
}

int buildSceneDev(rt_render* r, SceneDev& sc) {
    rt_scene* s = r->scene;
    if ((int)s->sets.size() > kMaxSets || (int)s->lights.size() > kMaxLights)
        return rt_fail(r->ctx, RT_ERR_INVALID, "fused path: more than 8 geometry sets or 8 lights");
    memset(&sc, 0, sizeof sc);
    sc.n_sets = (int)s->sets.size();
    sc.n_lights = (int)s->lights.size();
    for (int i = 0; i < sc.n_sets; i++) {
        const SceneSet& in = s->sets[i];
        SetDev& d = sc.sets[i];
        d.g.prim = (const float4*)in.grid.prim;
        d.g.box = (const unsigned*)in.grid.box_size;
        d.g.occ = (const unsigned*)in.grid.occupancy;
        d.g.bound.pmin = f3{in.bound[0], in.bound[1], in.bound[2]};
        d.g.bound.pmax = f3{in.bound[4], in.bound[5], in.bound[6]};
        d.g.n = in.grid.n_slabs;
        d.normals = (const float4*)in.grid.normal;
        d.matid = in.is_mesh ? nullptr : (const unsigned*)in.grid.matid;
        d.scalar_matid = in.mesh_matid;
        d.kind = in.grid.kind == 0 ? PRIM_SPHERE : PRIM_TRIANGLE;
        d.use_occ = (in.grid.occupancy != nullptr && in.grid.n_slabs > 2) ? 1 : 0;
    }
    for (int i = 0; i < sc.n_lights; i++) {
        memcpy(sc.lights[i].shadow.v, s->lights[i].shadow, 64);
        memcpy(sc.lights[i].scene.v, s->lights[i].scene, 64);
        memcpy(sc.lights[i].light.v, s->lights[i].light, 64);
    }
    sc.materials = (const float4*)s->materials;
    memcpy(sc.bound.v, s->bound, sizeof sc.bound.v);
    return RT_OK;
}

}  // namespace

int rt_fused_tile(rt_render* r, const float* fcam, size_t slot0, unsigned n, const float2* rpp1_coords) {
    rt_ctx* ctx = r->ctx;
    const rt_render_opts& o = r->o;
    SceneDev sc;
    int rc = buildSceneDev(r, sc);
    if (rc) return rc;
    PathArgs a;
    memcpy(a.cam.v, fcam, sizeof a.cam.v);
    a.focal_length = o.focal_length;
    a.lens_rad = o.lens_rad;
    a.rays_per_pixel = o.rays_per_pixel;
    a.slot_begin = o.slot_begin;
    a.slots_pp = r->slots_pp;
    a.depth = o.depth;
    a.pixel_base = slot0 / r->slots_pp;
    a.n_local = n;
    a.rpp1_coords = rpp1_coords;
    a.seeds = r->seeds + slot0;
    a.acu = r->acu + slot0;
    a.counters = r->d_counters;
    a.profile = r->d_profile;
    if (r->profile) k_pathMega<true><<<rt_blocks(n, 128), 128, 0, ctx->stream>>>(sc, a);
    else k_pathMega<false><<<rt_blocks(n, 128), 128, 0, ctx->stream>>>(sc, a);
    RT_LAUNCH_CHECK(ctx, "pathMega");
    return RT_OK;
}
