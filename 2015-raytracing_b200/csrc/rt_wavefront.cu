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
static_assert(kMaxSets == RT_MAX_SETS, "profile rows");
#define RT_TRY_W(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

struct SetDev {
    GridView g;
    const float4* normals;
    const unsigned* matid;     // per-reference material (spheres, scene triangles) or null
    unsigned scalar_matid;     // meshes
    int kind;                  // PRIM_SPHERE / PRIM_TRIANGLE
    int use_occ;
    const float4* pre_ng;      // heavy triangle sets: face vector per reference
    const float4* pre_pe;      // heavy triangle sets: (p0, e1, e2) per reference
    const unsigned* macro_occ; // heavy sets: coarse occupancy (<= 64^3 bits)
    unsigned macro_shift, macro_n;
    int far_ok;                // 1-cell sets: the cell's exit planes are bit for bit the box's far planes (see farPlanesShared)
    const unsigned char* macro_dist;   // heavy sets: chessboard distance of every cell of the distance field to the nearest occupied one (or null)
    float dist_inv[3];         // world -> field-index scale per axis
    float org_unit[3];         // world -> unit cube of the set's box (the queue partition's origin blocks)
    unsigned dist_shift, dist_n;       // field cells are (2^dist_shift)^3 grid cells, dist_n per axis
    unsigned n_refs;
    int wall_ok;               // 1-cell triangle sets made of axis-aligned planar triangles ("walls"): see shadowClearsWalls
    float wall_lo[3], wall_hi[3], wall_scale;
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
    WalkStats ws = {0, 0, 0};
    if (s.g.n == 1 && (s.kind == PRIM_SPHERE || s.pre_ng)) {
        if (s.kind == PRIM_SPHERE) h = singleCellWalk<PRIM_SPHERE, false, STATS>(ray.o, ray.d, ray.maxt, s.g, nullptr, nullptr, binter, &ws);
        else h = singleCellWalk<PRIM_TRIANGLE, false, STATS>(ray.o, ray.d, ray.maxt, s.g, s.pre_ng, s.pre_pe, binter, &ws);
    } else if (s.kind == PRIM_SPHERE) h = gridWalk<PRIM_SPHERE, false, true, STATS, OCC>(ray.o, ray.d, ray.maxt, s.g, binter, &ws);
    else h = gridWalk<PRIM_TRIANGLE, false, true, STATS, OCC>(ray.o, ray.d, ray.maxt, s.g, binter, &ws);
    if (STATS) {
        prof[1]++;
        prof[2] += (unsigned)ws.cells;
        prof[s.kind == PRIM_SPHERE ? 3 : 4] += (unsigned)ws.tests;
        prof[15] += (unsigned)ws.front;
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
    WalkStats ws = {0, 0, 0};
    if (s.g.n == 1 && (s.kind == PRIM_SPHERE || s.pre_ng)) {
        if (s.kind == PRIM_SPHERE) h = singleCellWalk<PRIM_SPHERE, true, STATS>(sr.o, sr.d, sr.maxt, s.g, nullptr, nullptr, binter, &ws);
        else h = singleCellWalk<PRIM_TRIANGLE, true, STATS>(sr.o, sr.d, sr.maxt, s.g, s.pre_ng, s.pre_pe, binter, &ws);
    } else if (s.kind == PRIM_SPHERE) h = gridWalk<PRIM_SPHERE, true, true, STATS, OCC>(sr.o, sr.d, sr.maxt, s.g, binter, &ws);
    else h = gridWalk<PRIM_TRIANGLE, true, true, STATS, OCC>(sr.o, sr.d, sr.maxt, s.g, binter, &ws);
    if (STATS) {
        prof[9]++;
        prof[10] += (unsigned)ws.cells;
        prof[s.kind == PRIM_SPHERE ? 11 : 12] += (unsigned)ws.tests;
        prof[15] += (unsigned)ws.front;
        if (h.i != 0xFFFFFFFFu) prof[13]++;
    }
    if (h.i != 0xFFFFFFFFu) { sr.maxt = h.t; sr.mint = h.t; }
    else sr.maxt = h.t;
}

// Profile build only: add this thread's counters [lo,hi) to row `set` of the global table and clear them.  The lanes that
// happen to execute this together may stand in DIFFERENT iterations of the caller's loop over sets (after divergence the
// hardware regroups lanes by program counter, not by iteration), so they are grouped by `set` first: one atomic per counter
// per group.  (Round 1 reduced over __activemask() alone and credited some counts to a neighbouring set's row -- column
// sums were right, rows were off by ~0.3 %; tests/test_gpu_gates.py checks the rows against the instrumented oracle now.)
RT_DEV void flushProfile(unsigned* prof, unsigned long long* table, int set, int lo, int hi) {
    const unsigned m = __match_any_sync(__activemask(), set);
    const int leader = __ffs(m) - 1;
    for (int q = lo; q < hi; q++) {
        unsigned v = __reduce_add_sync(m, prof[q]);
        if ((int)(threadIdx.x & 31) == leader && v) atomicAdd(table + set * 16 + q, (unsigned long long)v);
        prof[q] = 0;
    }
}

// The same two functions for the per-slot stage kernels, where every inline set is a ONE-cell grid (multi-cell
// sets are "heavy" and go to the queue walkers): only singleCellWalk is compiled in.
// A small sphere set is cheaper to rule out by its spheres than by its box: interSphere's first rejection is dis < 0
// (A10/code.cl:199-215), computed from the ray alone with the same operations here; when it fires for every sphere of the set the
// reference finds nothing whether or not the ray hits the set's box, so the six IEEE divisions of the slab test are skipped.
RT_DEV bool missesAllSpheres(const SetDev& s, f3 o, f3 d) {
    if (s.kind != PRIM_SPHERE || s.n_refs == 0 || s.n_refs > 2) return false;
    const float a = dot(d, d);
    for (unsigned i = 0; i < s.n_refs; i++) {
        const float4 sp = __ldg(s.g.prim + i);
        const f3 omc = o - mk3(sp.x, sp.y, sp.z);
        const float b = 2.0f * dot(omc, d);
        const float c = dot(omc, omc) - sp.w;
        if (!(fmaf(-4.0f * c, a, b * b) < 0.0f)) return false;
    }
    return true;
}

RT_DEV void closestSet1(const SetDev& s, RayR& ray, PoiR& poi) {
    if (ray.mint == ray.maxt) return;
    if (missesAllSpheres(s, ray.o, ray.d)) return;
    AabbFar binter = interAABBFar(ray.o, ray.d, s.g.bound);
    if (!binter.v) return;
    Hit h;
    if (s.kind == PRIM_SPHERE) h = singleCellWalk<PRIM_SPHERE, false, false>(ray.o, ray.d, ray.maxt, s.g, nullptr, nullptr, binter, nullptr, s.far_ok);
    else h = singleCellWalk<PRIM_TRIANGLE, false, false>(ray.o, ray.d, ray.maxt, s.g, s.pre_ng, s.pre_pe, binter, nullptr, s.far_ok);
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
RT_DEV void anySet1(const SetDev& s, RayR& sr) {
    if (sr.mint == sr.maxt) return;
    if (missesAllSpheres(s, sr.o, sr.d)) return;
    AabbFar binter = interAABBFar(sr.o, sr.d, s.g.bound);
    if (!binter.v) return;
    Hit h;
    if (s.kind == PRIM_SPHERE) h = singleCellWalk<PRIM_SPHERE, true, false>(sr.o, sr.d, sr.maxt, s.g, nullptr, nullptr, binter, nullptr, s.far_ok);
    else h = singleCellWalk<PRIM_TRIANGLE, true, false>(sr.o, sr.d, sr.maxt, s.g, s.pre_ng, s.pre_pe, binter, nullptr, s.far_ok);
    if (h.i != 0xFFFFFFFFu) { sr.maxt = h.t; sr.mint = h.t; }
    else sr.maxt = h.t;
}

template <bool STATS>
RT_DEV void closestAllSets(const SceneDev& sc, RayR& ray, PoiR& poi, unsigned* prof, unsigned long long* table) {
    for (int s = 0; s < sc.n_sets; s++) {
        if (sc.sets[s].use_occ) closestSet<true, STATS>(sc.sets[s], ray, poi, prof);
        else closestSet<false, STATS>(sc.sets[s], ray, poi, prof);
        if (STATS) { flushProfile(prof, table, s, 0, 8); flushProfile(prof, table, s, 15, 16); }
    }
}

// one light: initShadowTrace + shadow traces + sceneRender (A10/code.cl:631-673, 1073-1364)
template <bool STATS>
RT_DEV void shadeLight(const SceneDev& sc, const LightDev& L, PoiR& poi, int& seed, float4& acu, unsigned& n_any, unsigned* prof,
                       unsigned long long* table) {
    if (poi.matId < 0) return;
    RayR sr = makeShadowRay(poi.p, poi.n, L.shadow, seed);
    if (sr.mint != sr.maxt) n_any++;
    for (int s = 0; s < sc.n_sets; s++) {
        if (sc.sets[s].use_occ) anySet<true, STATS>(sc.sets[s], sr, prof);
        else anySet<false, STATS>(sc.sets[s], sr, prof);
        if (STATS) { flushProfile(prof, table, s, 8, 14); flushProfile(prof, table, s, 15, 16); }
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
        closestAllSets<STATS>(sc, ray, poi, prof, a.profile);
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
        for (int l = 0; l < sc.n_lights; l++) shadeLight<STATS>(sc, sc.lights[l], poi, seed, acu, n_any, prof, a.profile);
        // ---- bounces (A10/code.js:1829-1846)
        for (unsigned j = 0; j < a.depth; j++) {
            if (poi.matId >= 0) {   // bouncePaths, A10/code.cl:581-598
                getHemisphereRay(poi.p, poi.n, seed, ray.o, ray.d);
                ray.mint = 0.0f; ray.maxt = RT_INF;
                n_closest++;
            } else {
                ray.mint = RT_INF; ray.maxt = RT_INF;
            }
            closestAllSets<STATS>(sc, ray, poi, prof, a.profile);
            for (int l = 0; l < sc.n_lights; l++) shadeLight<STATS>(sc, sc.lights[l], poi, seed, acu, n_any, prof, a.profile);
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
    if (STATS) {   // slots processed (the per-set counters were flushed as they were produced)
        unsigned v = prof[14];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(a.profile + 14, (unsigned long long)v);
    }
}

// With n_slabs = 1 the DDA's first "next plane" of an axis is (pmin + (d >= 0 ? 1 : 0) * delta - o) / d with
// delta = (pmax - pmin) / 1.0f (A10/code.cl:696-707).  That is bit for bit the far-plane quotient the slab test of
// interAABB has just computed for the same box -- (pmax - o) / d for d >= 0, (pmin - o) / d for d < 0 (A10/code.cl:344-352)
// -- when, in fp32, pmin + (pmax - pmin) == pmax exactly, the extent is finite, and pmin is not -0 (so that
// pmin + 0 * delta keeps its bits).  Checked here, on the host, with the same fp32 operations; the stage kernels then
// reuse the three quotients instead of dividing again (RT_FAR_SHARE).  Boxes that fail the check take the plain path.
#ifndef RT_FAR_SHARE
#define RT_FAR_SHARE 1
#endif
bool farPlanesShared(const float* b8) {
    for (int a = 0; a < 3; a++) {
        volatile float lo = b8[a], hi = b8[4 + a];
        volatile float delta = (hi - lo) / 1.0f;
        volatile float one = 1.0f * delta, zero = 0.0f * delta;
        volatile float up = lo + one, down = lo + zero;
        unsigned lo_bits, down_bits;
        float lo_v = lo, down_v = down;
        memcpy(&lo_bits, &lo_v, 4);
        memcpy(&down_bits, &down_v, 4);
        if (!(delta == delta) || delta > 3.0e38f || delta < -3.0e38f) return false;
        if (!(up == hi) || lo_bits != down_bits) return false;
    }
    return true;
}

// RT2015_NO_SKIP=1 (environment, read once) switches the empty-walk proof off (A/B runs; results are identical either way).
bool skipEmptyWalks() {
    static const bool v = !(getenv("RT2015_NO_SKIP") && atoi(getenv("RT2015_NO_SKIP")) != 0);
    return v;
}

// RT2015_NO_WALL_SKIP=1 (environment, read once) switches shadowClearsWalls off (A/B runs; results are identical either way).
bool skipWallTests() {
    static const bool v = !(getenv("RT2015_NO_WALL_SKIP") && atoi(getenv("RT2015_NO_WALL_SKIP")) != 0);
    return v;
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
        d.use_occ = (in.grid.occupancy != nullptr && in.grid.n_slabs > 1) ? 1 : 0;
        d.pre_ng = in.pre_ng;
        d.pre_pe = in.pre_pe;
        d.macro_occ = in.macro_occ;
        d.macro_shift = in.macro_shift;
        d.macro_n = in.macro_n;
        d.far_ok = (RT_FAR_SHARE && in.grid.n_slabs == 1 && farPlanesShared(in.bound)) ? 1 : 0;
        d.macro_dist = skipEmptyWalks() ? in.macro_dist : nullptr;
        for (int a = 0; a < 3; a++) d.dist_inv[a] = in.dist_inv[a];
        for (int a = 0; a < 3; a++) d.org_unit[a] = in.dist_n ? in.dist_inv[a] / (float)in.dist_n : 0.f;
        d.dist_shift = in.dist_shift;
        d.dist_n = in.dist_n;
        d.n_refs = in.grid.n_refs;
        d.wall_ok = (in.wall_ok && skipWallTests()) ? 1 : 0;
        for (int a = 0; a < 3; a++) { d.wall_lo[a] = in.wall_lo[a]; d.wall_hi[a] = in.wall_hi[a]; }
        d.wall_scale = in.wall_scale;
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


// =======================================================================================
// Wavefront path (mode 0): the pass as a short sequence of per-slot "stage" kernels and
// queue-driven persistent "walk" kernels.
//
// Why: in the megakernel a warp that enters the DDA of a big grid runs until its slowest lane
// is done while most lanes idle (rays that miss the grid's AABB, short walks) -- ncu measured
// 4.6 of 32 lanes active per instruction (profiles/).  Here the cheap, uniform work (ray
// generation, the small 1-cell sets, shading) stays one-thread-per-slot, while every ray that
// must walk a BIG grid ("heavy" set) is pushed -- warp ballot + popc prefix + one atomicAdd per
// warp -- into a compact queue of slot ids, and a persistent kernel walks the queue, refilling
// idle lanes one cell step at a time, so its warps stay full regardless of which slots are
// live, blocked, or missed the box.  Per-slot results are bit-identical to the other modes:
// the same device functions run in the same per-slot order.
// =======================================================================================
constexpr int kMaxStages = 64;

struct WaveState {
    float4* ray;      // [2n]
    float4* poi;      // [2n]
    float4* atte;     // [n]
    float4* sh;       // [lights][2n]: light l's shadow ray of slot id = sh[2l*n + id] (o, mint), sh[(2l+1)*n + id] (d, maxt)
    unsigned* queue;  // [n * lights]: closest-hit walks queue slot ids, any-hit walks queue light * n + slot
    unsigned* qctr;   // {count, head} per walk stage
};

// What one per-slot stage kernel does, in this order (all fields uniform across the grid):
//   shade            : sceneRender for EVERY light, in light order, with that light's current shadow ray
//   gen == 1/2       : initTrace / bouncePaths -> new ray
//   light_render     : lightRender for every light (primary segment, after the closest pass)
//   shadow           : initShadowTrace for EVERY light, in light order -> one new shadow ray per light
//   inline sets [set_lo, set_hi) against the current ray (kind 0 = closest) or against every light's shadow ray (kind 1)
//   push_set >= 0    : AABB-test the ray(s) against that heavy set and enqueue the slot (kind 1: one entry per light)
//
// All lights of a segment travel TOGETHER (one shadow ray buffer per light): the random draws of a slot do not depend on the
// shadow results -- initShadowTrace draws iff matId >= 0 (A10/code.cl:641-651) and neither the shadow traces nor sceneRender
// touch matId -- so drawing light 1's sample right after light 0's is the reference's draw order (quirk Q8), and the
// sceneRender calls still run in light order afterwards (atte *= colour per light, quirk Q3).  Per tile that is 1 + 6 * 2
// stage launches and 6 any-hit walks per heavy set instead of 1 + 6 * (1 + L) and 6 * L, with L-times deeper any-hit queues.
struct StageOp {
    int shade, gen, light_render, shadow, kind, set_lo, set_hi, push_set, qslot;
    int shade_lo, shade_hi;   // lights whose sceneRender runs when `shade`
    int light_lo, light_hi;   // lights whose shadow rays this stage generates / traces (kind 1)
};

// ---------------------------------------------------------------------------------------
// Shadow rays against a room.  Two thirds of all inline triangle tests of a path-traced pass are shadow rays tested
// against the walls of the room they live in (12 triangles x lights x 6 segments per slot) -- and a segment between two
// points INSIDE a room cannot hit its walls.  When a 1-cell triangle set consists of axis-aligned planar triangles only,
// the host derives the open box between those planes that contains the set's centre (rt_scene_add_set); a shadow
// segment whose origin and end point both lie inside that box -- with margins that dominate the rounding of the
// reference's test -- is rejected by every triangle of the set in the reference too, so the whole set is skipped for it.
// Why the reference rejects (A10/code.cl:250-288, t = dot(cross(s, e2), e1) / -div): for a triangle in the plane x_a = c the
// exact parameter is t* = h / v with h = c - o_a the origin's distance to the plane and v = d_a; the computed t carries a
// relative error E <= 6e-7 (|s| / (h sin(theta)) + 1 / |v|) (10 roundings of 2^-24 in the triple product and in div; theta =
// angle at p0, >= 14.5 degrees required on the host).  (i) Plane behind or at the origin side, h v < 0: t* < 0 and stays
// negative for E < 1, guaranteed by h >= 2e-5 scale -- or, once |v| is so small that the sign of div is in doubt, |t| >= h
// / |v| exceeds any segment length.  (ii) Plane beyond the end point by g: t* - maxt = g / |v| and t* <= 2 scale / |v|, so
// (t* - maxt) / t* >= g / (2 scale) must exceed E <= 8.4e-6 scale / g + 6e-7 / |v| (h >= g, |s| <= 3.5 scale): required with a
// factor 6 to spare as g >= 1e-2 scale and g |v| >= 4e-6 scale.  In both cases
// `t < champ_t = maxt` or `t >= mint >= 0` fails, whatever beta and gamma are.  Checked against the reference kernels on
// adversarial shadow rays (tests/test_gpu_gates.py); RT2015_NO_WALL_SKIP=1 switches it off.
// ---------------------------------------------------------------------------------------
RT_DEV bool shadowClearsWalls(const SetDev& s, const RayR& sr) {
    if (!(sr.maxt < RT_INF)) return false;
    const float sc = s.wall_scale;
    const float o[3] = {sr.o.x, sr.o.y, sr.o.z}, d[3] = {sr.d.x, sr.d.y, sr.d.z};
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float q = o[a] + sr.maxt * d[a];            // end point of the segment
        const float v = fabsf(d[a]);
        // origin strictly inside, by more than the rounding of its own distance to either plane
        ok = ok && (o[a] - s.wall_lo[a] >= 2e-5f * sc) && (s.wall_hi[a] - o[a] >= 2e-5f * sc);
        // end point inside with a margin that grows as the ray gets parallel to the plane
        const float g = fminf(q - s.wall_lo[a], s.wall_hi[a] - q);
        ok = ok && (g >= 1e-2f * sc) && (g * v >= 4e-6f * sc);
    }
    return ok;   // NaNs anywhere fail the comparisons
}

// Loads / stores of the per-slot wavefront state (rays, hit records, shadow rays, queues, seeds, accumulators): every
// byte is touched once per kernel and the tile is gigabytes, so nothing of it survives in the L2 until the next kernel
// anyway.  RT_STREAM_CS = 1 marks these accesses evict-first (ld/st.global.cs) so that they stop displacing the scene data
// (cell table, occupancy bits, face vectors) the walkers re-read all the time.
#ifndef RT_STREAM_CS
#define RT_STREAM_CS 0
#endif
template <class T> RT_DEV T ldS(const T* p) {
#if RT_STREAM_CS
    return __ldcs(p);
#else
    return *p;
#endif
}
template <class T> RT_DEV void stS(T* p, T v) {
#if RT_STREAM_CS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// ---------------------------------------------------------------------------------------
// Empty-walk proof.  A third of the rays that enter a heavy set's box cross EMPTY cells only (they cut the corners of
// the box around a mesh) and account for almost half of all cell steps (instrumented oracle, config 5: 34 % of the walks,
// 45-49 % of the cells).  Such a walk tests nothing and changes nothing, so it may be skipped -- if that can be PROVEN
// without doing it.  The proof marches the exact ray through a distance field of the occupancy grid (the grid's own cells
// up to n = 128, blocks of 2^k cells above): dist[c] = D says every field cell closer than D (chessboard metric) to c is empty,
// so from a point in c the ray may advance D - 3.5 field cells and every cell within 2 field cells of that stretch is still empty.  The 2 is the safety margin
// for the difference between the exact ray and the cells the reference's fp32 DDA (A10/code.cl:694-786) really visits:
// its entry slab is one division + truncation (<= 1 fine cell off), its per-axis crossing times are the exact first
// crossing plus up to n fp32 additions (relative error 2^-24 each) -- bounded below per ray by `e` fine cells and
// required to be <= 1 -- so it stays within 2 fine cells (<= 2 field cells) of the exact ray; 1.5 more field cells
// cover point-to-cell quantisation and the rounding of this march itself.  Rays with a zero / non-finite direction
// component or a non-finite box interval (where the reference's float-equality stepping degenerates, quirk Q12) are
// never skipped.  A skipped ray is exactly a ray whose walk would have visited no reference: checked against the
// instrumented reference kernels on millions of random and adversarial rays (tests/test_gpu_gates.py).
// Where it runs: NOT in the per-slot stage kernels -- only ~13 % of the slots of a stage hit the heavy set's box, so a
// warp there pays its slowest lane's march for 4 useful lanes (measured: stage 606 -> 724 ms, more than the walkers
// gained) -- but in k_filter_mark, a pass over the compact queue between stage and walk where every lane has a ray to prove.
// ---------------------------------------------------------------------------------------
#ifndef RT_MAX_MARCH
#define RT_MAX_MARCH 16
#endif
constexpr int kMaxMarch = RT_MAX_MARCH;   // undecided after that many steps = kept
RT_DEV bool walkProvablyEmpty(f3 o, f3 d, float tmin, float tmax, const SetDev& s) {
    const unsigned char* dist = s.macro_dist;
    if (!dist) return false;
    const f3 du = mk3(d.x * s.dist_inv[0], d.y * s.dist_inv[1], d.z * s.dist_inv[2]);   // coarse cells per unit t
    const float dmax = fmaxf(fmaxf(fabsf(du.x), fabsf(du.y)), fabsf(du.z));
    if (d.x == 0.0f || d.y == 0.0f || d.z == 0.0f || !(dmax < RT_INF) || !(tmin >= 0.0f) || !(tmax < RT_INF)) return false;
    const float fcell = (float)(1u << s.dist_shift);
    const float e = (float)s.g.n * 1.2e-7f * tmax * dmax * fcell;   // drift bound of the reference's t_next sequences, in fine cells
    if (!(e <= 1.0f)) return false;
    const f3 u0 = mk3((o.x - s.g.bound.pmin.x) * s.dist_inv[0], (o.y - s.g.bound.pmin.y) * s.dist_inv[1], (o.z - s.g.bound.pmin.z) * s.dist_inv[2]);
    const float rd = 1.0f / dmax;
    const int nm = (int)s.dist_n;
    float t = tmin;
    for (int it = 0; it < kMaxMarch; it++) {
        int cx = (int)floorf(u0.x + t * du.x), cy = (int)floorf(u0.y + t * du.y), cz = (int)floorf(u0.z + t * du.z);
        cx = min(max(cx, 0), nm - 1); cy = min(max(cy, 0), nm - 1); cz = min(max(cz, 0), nm - 1);
        const unsigned D = __ldg(dist + ((size_t)cz * nm + cy) * nm + cx);
        if (D < 4u) return false;
        t += ((float)D - 3.5f) * rd;
        if (t >= tmax) return true;
    }
    return false;
}

RT_DEV void pushTask(bool want, unsigned id, unsigned* queue, unsigned* count) {
    unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) stS(queue + base + __popc(m & ((1u << lane) - 1u)), id);
}

#ifndef RT_STAGE_MINB
#define RT_STAGE_MINB 5
#endif
#ifndef RT_STAGE_MINB_BOUNCE   // the shade + bounce + closest-hit stage spills 186 B at 48 registers (5 blocks); at 64 (4 blocks):
#define RT_STAGE_MINB_BOUNCE 4  // stage class 500.5 -> 491.7 ms.  4 blocks for every stage shape: 497; 3 blocks: 546
#endif
// GEN / LR / SHADE / SHADOW / KIND mirror StageOp's gen, light_render, shade, shadow and kind; they are template
// parameters so that each of the few stage shapes a pass is made of gets its own register allocation (the
// all-in-one kernel needed 80 registers or spilled ~400 B at 64).  The set ranges stay run-time values.
template <int GEN, bool LR, bool SHADE, bool SHADOW, int KIND>
__global__ void __launch_bounds__(256, (GEN == 2 && SHADE) ? RT_STAGE_MINB_BOUNCE : RT_STAGE_MINB) k_stage(const __grid_constant__ SceneDev sc, const __grid_constant__ PathArgs a,
                                               const __grid_constant__ WaveState w, const __grid_constant__ StageOp op) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inb = id < a.n_local;
    const unsigned n = a.n_local;
    const int nl = sc.n_lights;
    unsigned n_closest = 0, n_any = 0;
    bool want_push = false;
    PoiR poi;
    poi.p = mk3(0.f, 0.f, 0.f); poi.n = mk3(0.f, 0.f, 0.f); poi.atte = mk3(1.f, 1.f, 1.f); poi.matId = -1;
    int seed = 0;
    bool seed_loaded = false, seed_dirty = false;
    if (inb) {
        RayR ray;
        bool poi_dirty = false, ray_dirty = false, atte_dirty = false;
        // ---- load what this stage needs
        if (GEN != 1 && (SHADE || GEN == 2 || LR || SHADOW || KIND == 0)) {
            float4 p0 = ldS(w.poi + id), p1 = ldS(w.poi + n + id);
            poi.p = mk3(p0.x, p0.y, p0.z); poi.matId = __float_as_int(p0.w);
            poi.n = mk3(p1.x, p1.y, p1.z);
        }
        bool need_ray = (GEN == 0) && (LR || (KIND == 0 && (op.set_hi > op.set_lo || op.push_set >= 0)));
        if (need_ray) {
            float4 r0 = ldS(w.ray + id), r1 = ldS(w.ray + n + id);
            ray.o = mk3(r0.x, r0.y, r0.z); ray.mint = r0.w;
            ray.d = mk3(r1.x, r1.y, r1.z); ray.maxt = r1.w;
        }
        float4 acu;
        bool acu_loaded = false;
        // ---- sceneRender of the previous shadow rays, light by light (A10/code.cl:1323-1364)
        if (SHADE && poi.matId >= 0) {
            float4 at = ldS(w.atte + id);
            poi.atte = mk3(at.x, at.y, at.z);
            acu = ldS(a.acu + id); acu_loaded = true;
            float4 color = __ldg(sc.materials + poi.matId);
            f3 c = mk3(color.x, color.y, color.z);
            for (int l = op.shade_lo; l < op.shade_hi; l++) {
                float4 s0 = ldS(w.sh + (size_t)(2 * l) * n + id), s1 = ldS(w.sh + (size_t)(2 * l + 1) * n + id);
                f3 shade = neeShade(poi.p, poi.n, mk3(s1.x, s1.y, s1.z), s1.w != s0.w, sc.lights[l].scene);
                f3 contrib = (c * poi.atte) * shade;
                poi.atte = poi.atte * c;   // per light (quirk Q3)
                acu = make_float4(acu.x + contrib.x, acu.y + contrib.y, acu.z + contrib.z, acu.w + 1.0f);
            }
            atte_dirty = op.shade_hi > op.shade_lo;
        }
        // ---- new ray
        if (GEN == 1) {   // initTrace (A10/code.cl:458-543)
            Camera cam = floatToCamera(a.cam.v);
            AABB bound = toAABB(sc.bound);
            size_t pix = a.pixel_base + id / a.slots_pp;
            unsigned k = a.slot_begin + id % a.slots_pp;
            unsigned col = (unsigned)(pix % cam.cols), row = (unsigned)(pix / cam.cols);
            ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 0.f);
            ray.mint = RT_INF; ray.maxt = RT_INF;
            poi_dirty = atte_dirty = true;
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
            ray_dirty = true;
            if (ray.mint != ray.maxt) n_closest++;
        } else if (GEN == 2) {   // bouncePaths (A10/code.cl:581-598)
            if (poi.matId >= 0) {
                seed = ldS(a.seeds + id); seed_loaded = true;
                getHemisphereRay(poi.p, poi.n, seed, ray.o, ray.d);
                seed_dirty = true;
                ray.mint = 0.0f; ray.maxt = RT_INF;
                n_closest++;
            } else {
                ray.o = mk3(0.f, 0.f, 0.f); ray.d = mk3(0.f, 0.f, 0.f);
                ray.mint = RT_INF; ray.maxt = RT_INF;
            }
            ray_dirty = true;
        }
        // ---- lightRender for every light (A10/code.cl:600-629)
        if (LR) {
            for (int l = 0; l < nl; l++) {
                if (ray.mint == ray.maxt) continue;
                const LightArg& L = sc.lights[l].light;
                f3 irradiance = normalize(mk3(L.v[6], L.v[7], L.v[8]));
                float t;
                if (!interLight(ray.o, ray.d, mk3(L.v[0], L.v[1], L.v[2]), mk3(L.v[3], L.v[4], L.v[5]), L.v[9], t) || t >= ray.maxt) continue;
                ray.mint = RT_INF; ray.maxt = RT_INF;
                ray_dirty = true;
                poi.matId = -1;
                poi_dirty = true;
                if (!acu_loaded) { acu = ldS(a.acu + id); acu_loaded = true; }
                acu = make_float4(acu.x + irradiance.x, acu.y + irradiance.y, acu.z + irradiance.z, acu.w + 1.0f);
            }
        }
        // ---- small sets inline against the ray, in set order
        if (KIND == 0) {
            for (int s = op.set_lo; s < op.set_hi; s++) {
                float m0 = ray.maxt; int id0 = poi.matId;
                closestSet1(sc.sets[s], ray, poi);
                if (ray.maxt != m0 || poi.matId != id0) { ray_dirty = true; poi_dirty = true; }
            }
            if (op.push_set >= 0 && ray.mint != ray.maxt) want_push = interAABB(ray.o, ray.d, sc.sets[op.push_set].g.bound).v;
        }
        // ---- write back
        if (ray_dirty) {
            stS(w.ray + (id), make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint));
            stS(w.ray + (n + id), make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt));
        }
        if (poi_dirty) {
            stS(w.poi + (id), make_float4(poi.p.x, poi.p.y, poi.p.z, __int_as_float(poi.matId)));
            stS(w.poi + (n + id), make_float4(poi.n.x, poi.n.y, poi.n.z, 0.f));
        }
        if (atte_dirty) stS(w.atte + id, make_float4(poi.atte.x, poi.atte.y, poi.atte.z, 0.f));
        if (acu_loaded) stS(a.acu + id, acu);
    }
    if (KIND == 0) {
        if (op.push_set >= 0) pushTask(want_push, id, w.queue, w.qctr + 2 * op.qslot);
    } else {
        // ---- every light's shadow ray: new ray (A10/code.cl:631-673), small sets inline, push to the heavy set.  The loop
        // is uniform (pushTask is a warp collective); draws happen in light order from the slot's one seed.
        const bool trace = op.set_hi > op.set_lo || op.push_set >= 0;
        for (int l = op.light_lo; l < op.light_hi; l++) {
            bool want = false;
            if (inb && (SHADOW || trace)) {
                RayR sr;
                bool sr_dirty = false;
                float4* s0p = w.sh + (size_t)(2 * l) * n + id;
                float4* s1p = w.sh + (size_t)(2 * l + 1) * n + id;
                if (SHADOW) {
                    if (poi.matId >= 0) {
                        if (!seed_loaded) { seed = ldS(a.seeds + id); seed_loaded = true; }
                        sr = makeShadowRay(poi.p, poi.n, sc.lights[l].shadow, seed);
                        seed_dirty = true;
                        if (sr.mint != sr.maxt) n_any++;
                    } else {
                        sr.o = mk3(0.f, 0.f, 0.f); sr.d = mk3(0.f, 0.f, 0.f);
                        sr.mint = RT_INF; sr.maxt = RT_INF;
                    }
                    sr_dirty = true;
                } else {
                    float4 s0 = ldS(s0p), s1 = ldS(s1p);
                    sr.o = mk3(s0.x, s0.y, s0.z); sr.mint = s0.w;
                    sr.d = mk3(s1.x, s1.y, s1.z); sr.maxt = s1.w;
                }
                for (int s = op.set_lo; s < op.set_hi; s++) {
                    float m0 = sr.mint, x0 = sr.maxt;
                    if (sc.sets[s].wall_ok && shadowClearsWalls(sc.sets[s], sr)) continue;   // a segment inside the room misses its walls
                    anySet1(sc.sets[s], sr);
                    if (sr.mint != m0 || sr.maxt != x0) sr_dirty = true;
                }
                if (op.push_set >= 0 && sr.mint != sr.maxt) want = interAABB(sr.o, sr.d, sc.sets[op.push_set].g.bound).v;
                if (sr_dirty) {
                    stS(s0p, make_float4(sr.o.x, sr.o.y, sr.o.z, sr.mint));
                    stS(s1p, make_float4(sr.d.x, sr.d.y, sr.d.z, sr.maxt));
                }
            }
            if (op.push_set >= 0) pushTask(want, (unsigned)l * n + id, w.queue, w.qctr + 2 * op.qslot);
        }
    }
    if (inb && seed_dirty) stS(a.seeds + id, seed);
    // ray counters: only a stage that generates rays of a kind can have counted any (one redux + one atomic per warp)
    if (GEN != 0) {
        const unsigned t = __reduce_add_sync(0xffffffffu, n_closest);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(a.counters + 0, (unsigned long long)t);
    }
    if (SHADOW) {
        const unsigned t = __reduce_add_sync(0xffffffffu, n_any);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(a.counters + 1, (unsigned long long)t);
    }
}

// Persistent queue walkers.  Each lane owns at most one ray; a lane whose walk ended becomes idle; when at
// least kRefill lanes of the warp are idle (or none has work) the warp pops that many slot ids with ONE
// atomicAdd and the idle lanes set up their DDA.
#ifndef RT_REFILL
#define RT_REFILL 24
#endif
constexpr int kRefill = RT_REFILL;
#ifndef RT_STEP_BURST
#define RT_STEP_BURST 8
#endif
#ifndef RT_WALK_MINB
#define RT_WALK_MINB 4
#endif
// RT_SKIP2 = 1 makes the pair walker leave EMPTY 2x2x2 blocks of cells in one exact step (flatAdvance in
// rt_device.cuh; bit-exact, GPU suite green with it).  Measured on B200 at the full config: 5391 vs 5393 Mrays/s --
// fewer but costlier steps, no gain -- so it stays off; kept because the exactness argument is tested.
#ifndef RT_SKIP2
#define RT_SKIP2 0
#endif
constexpr int kStepBurst = RT_STEP_BURST;   // empty-cell steps per outer iteration
constexpr int kWalkWarps = 8;               // warps per block
constexpr int kWalkMinBlocks = RT_WALK_MINB; // resident blocks per SM the register budget is tuned for

// ---------------------------------------------------------------------------------------
// Pair-list queue walker.  Every lane owns one ray and steps its own DDA through empty cells; the
// primitive tests of ALL lanes that stand in a non-empty cell are done together: the (owner lane,
// reference) pairs of those cells are laid out as one list (warp prefix sum of the cell
// populations), and the warp walks that list 32 pairs at a time, so the face-vector cull runs on
// dense lanes whatever the cell populations are (two earlier designs -- one lane per ray testing its
// own cell, and the whole warp testing one lane's cell at a time -- left 3/4 of the lanes idle on the
// 1 M-triangle mesh and were removed after the A/B of round 1, profiles/README.md).  Survivors are compacted into a candidate buffer and the full
// test again runs 32 candidates at a time.  Each lane fetches its pair's ray from the owner lane
// with indexed shuffles.  Accepted hits (rare) are reduced per owner with one shared-memory
// atomicMin on the key (t bits, reference index) [closest: min t, ties -> lowest index] or
// (reference index) [any hit: first accepted reference in list order] -- exactly the winner of
// the reference's sequential loop (A10/code.cl:882-897, 1272-1287), whose floats the owner then
// recomputes with one more test of the winning reference.
// ---------------------------------------------------------------------------------------
constexpr int kCandCap = 64;

// Software prefetch (exactness-neutral).  ncu attributes most of the walkers' stall samples to the first use of a
// face-vector / (p0,e1,e2) record that was requested only when the pair list reached it (profiles/), so two hints
// were tried on B200 at the full config (base 6076 Mrays/s):
//   RT_PF_PE: when a reference survives the cull, request its (p0,e1,e2) record -- it is read by ANOTHER lane once
//             32 survivors have been collected.                         L1: 6119, L2: 6115 Mrays/s  -> on (L1)
//   RT_PF_NG: when a lane enters a non-empty cell, request the lines of that cell's face vectors right away.
//             L1: 5895, L2: 5897 Mrays/s (the extra instructions in the step loop cost more than the hint saves) -> off
//   On the partitioned queue (round 2: the rays of a warp visit neighbouring cells, the lines are in L1 anyway): RT_PF_PE off 8 630 vs
//   on 8 590 Mrays/s (any-hit walks 372 vs 377 ms); RT_PF_NG on 8 307 -> both off.
// 0 = off, 1 = into L1, 2 = into L2.
#ifndef RT_PF_NG
#define RT_PF_NG 0
#endif
#ifndef RT_PF_PE
#define RT_PF_PE 0
#endif
template <int LEVEL>
RT_DEV void prefetchLine(const void* p) {
    if (LEVEL == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    else if (LEVEL == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// A07 = true: Assignment 7's exclusive triangle range test (quirk Q9); the hit record carries the hit's cell instead of a
// material (its parity colours the pixel, A07/code.cl:462-468).
template <int PRIM, bool ANY, bool A07 = false>
__global__ void __launch_bounds__(kWalkWarps * 32, kWalkMinBlocks) k_walk_pairs(const __grid_constant__ SetDev set, const __grid_constant__ WaveState w,
                                                                               unsigned n, int qslot) {
    // 64^3 bits.  (Reading the coarse bitmap through the L1 instead -- 10 KB of shared memory per block, 128 KB more L1 per SM --
    // was measured slower: 7548 vs 7813 Mrays/s.)
    __shared__ unsigned s_macro[8192];
    __shared__ unsigned s_plist[kWalkWarps][32];   // pending lanes in lane order ...
    __shared__ unsigned s_poff[kWalkWarps][32];    // ... and where their pairs start in the list
    __shared__ unsigned s_cref[kWalkWarps][kCandCap];
    __shared__ unsigned s_cown[kWalkWarps][kCandCap];
    __shared__ float s_cdiv[kWalkWarps][kCandCap];
    __shared__ unsigned long long s_best[kWalkWarps][32];
    for (unsigned i = threadIdx.x; i < 8192; i += blockDim.x) s_macro[i] = set.macro_occ[i];
    __syncthreads();
    const unsigned mshift = set.macro_shift, mn = set.macro_n;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    const unsigned long long NONE = ~0ull;
    const unsigned count = w.qctr[2 * qslot];
    unsigned* head = w.qctr + 2 * qslot + 1;
    // any-hit entries are light * n + slot (one shadow ray buffer per light); closest-hit entries are plain slot ids
    auto rayBase = [&](unsigned e) -> size_t {
        if (!ANY) return e;
        const unsigned l = e / n;
        return (size_t)(2 * l) * n + (e - l * n);
    };
    const GridView& g = set.g;
    FlatWalker f;
    f.i = f.end = 0;
    unsigned slot = 0;
    bool have = false;
    bool fin = false;       // walk ended with a hit that still has to be written back
    bool drained = false;
#if RT_SKIP2
    // exact skipping of empty 2x2x2 blocks (flatAdvance): needs 2-cell coarse blocks, an even n and a ray without
    // zero direction components (finite t_next / delta_t)
    const bool can_skip = (mshift == 1) && ((g.n & 1u) == 0);
    bool skip = false, finite = false;
#endif

    // full test of candidates [0, cnt) of the buffer, one per lane; accepted hits go to s_best[owner]
    auto testCandidates = [&](unsigned cnt) {
        const bool act = lane < cnt;
        unsigned r = 0, own = lane;
        float dv = 0.f;
        if (act) { r = s_cref[wid][lane]; own = s_cown[wid][lane]; dv = s_cdiv[wid][lane]; }
        const f3 o = mk3(__shfl_sync(FULL, f.w.o.x, own), __shfl_sync(FULL, f.w.o.y, own), __shfl_sync(FULL, f.w.o.z, own));
        const f3 d = mk3(__shfl_sync(FULL, f.w.d.x, own), __shfl_sync(FULL, f.w.d.y, own), __shfl_sync(FULL, f.w.d.z, own));
        const float mint = __shfl_sync(FULL, f.mint, own), maxt = __shfl_sync(FULL, f.maxt, own);
        const float champ = __shfl_sync(FULL, f.w.h.t, own);
        const float a_dd = (PRIM == PRIM_SPHERE) ? __shfl_sync(FULL, f.w.a_dd, own) : 0.f;
        if (act) {
            float ti, be, ga;
            bool v;
            if (PRIM == PRIM_TRIANGLE) {
                float4 q0 = __ldg(set.pre_pe + 3 * r), q1 = __ldg(set.pre_pe + 3 * r + 1), q2 = __ldg(set.pre_pe + 3 * r + 2);
                v = interTriangleFast<!A07>(o, d, mint, maxt, dv, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), be, ga, ti);
            } else {
                v = interSphere(o, d, a_dd, mint, maxt, __ldg(g.prim + r), ti);
            }
            if (v && ti < champ) {
                // accepted t >= mint >= 0 (interAABB clamps tmin at 0), so the bit pattern of t + 0.0f orders like t
                unsigned long long key = ANY ? (unsigned long long)r : (((unsigned long long)__float_as_uint(ti + 0.0f) << 32) | r);
                atomicMin(&s_best[wid][own], key);
            }
        }
    };

    // Write the result of a finished walk (A10/code.cl:921-934 closest hit, :1185-1192 any hit).  Deferred to the refill
    // step so that the lanes that finished since the last refill do it together instead of one at a time inside
    // the stepping loop.
    auto writeBack = [&]() {
        const Hit& h = f.w.h;
        if (ANY) {
            float4* sp = w.sh + rayBase(slot);
            float4 s0 = ldS(sp);
            stS(sp, make_float4(s0.x, s0.y, s0.z, h.t));
            float4 s1 = ldS(sp + n);
            stS(sp + n, make_float4(s1.x, s1.y, s1.z, h.t));
        } else {
            f3 p = getPoint(f.w.o, f.w.d, h.t);
            f3 nrm;
            int m;
            if (PRIM == PRIM_SPHERE) {
                float4 sp = __ldg(g.prim + h.i);
                nrm = normalize(p - mk3(sp.x, sp.y, sp.z));
                m = A07 ? 0 : (int)__ldg(set.matid + h.i);
            } else {
                float4 n0 = __ldg(set.normals + 3 * h.i), n1 = __ldg(set.normals + 3 * h.i + 1), n2 = __ldg(set.normals + 3 * h.i + 2);
                nrm = normalize(interp(h.beta, h.gamma, mk3(n0.x, n0.y, n0.z), mk3(n1.x, n1.y, n1.z), mk3(n2.x, n2.y, n2.z)));
                m = A07 ? 0 : (set.matid ? (int)__ldg(set.matid + h.i) : (int)set.scalar_matid);
            }
            // the walk stops in the cell that produced the hit (flatLeave returns before stepping): the slabs are the hit's cell
            if (A07) m = (int)((unsigned)f.w.ax.slab | ((unsigned)f.w.ay.slab << 10) | ((unsigned)f.w.az.slab << 20));
            float4 r1 = ldS(w.ray + n + slot);
            stS(w.ray + (n + slot), make_float4(r1.x, r1.y, r1.z, h.t));
            stS(w.poi + (slot), make_float4(p.x, p.y, p.z, __int_as_float(m)));
            stS(w.poi + (n + slot), make_float4(nrm.x, nrm.y, nrm.z, 0.f));
        }
    };

    // face vectors of the cell just entered: at most 4 lines of 8 records
    auto prefetchCell = [&]() {
        if (PRIM == PRIM_TRIANGLE && RT_PF_NG) {
            unsigned a = f.i & ~7u;
#pragma unroll
            for (int k = 0; k < 4; k++, a += 8)
                if (a < f.end) prefetchLine<RT_PF_NG>(set.pre_ng + a);
        }
    };

    while (true) {
        // ---- refill idle lanes from the queue (one atomicAdd per warp)
        unsigned idle = __ballot_sync(FULL, !have);
        if (__popc(idle) >= kRefill || idle == FULL) {
            if (fin) { writeBack(); fin = false; }
        }
        if (!drained && (__popc(idle) >= kRefill || idle == FULL)) {
            unsigned base = 0;
            int leader = __ffs(idle) - 1;
            if ((int)lane == leader) base = atomicAdd(head, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            const unsigned limit = count;
            if (!have) {
                unsigned idx = base + __popc(idle & lt);
                if (idx < limit) {
                    slot = ldS(w.queue + idx);
                    const float4* src = (ANY ? w.sh : w.ray) + rayBase(slot);
                    float4 r0 = ldS(src), r1 = ldS(src + n);
                    f3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r1.x, r1.y, r1.z);
                    AabbHit binter = interAABB(o, d, g.bound);
                    walkInit(f.w, PRIM, o, d, r1.w, g, binter);
#if RT_SKIP2
                    finite = can_skip && isfinite(f.w.ax.t_next) && isfinite(f.w.ay.t_next) && isfinite(f.w.az.t_next) &&
                             isfinite(f.w.ax.delta_t) && isfinite(f.w.ay.delta_t) && isfinite(f.w.az.delta_t);
                    skip = flatEnterSkip(f, g, s_macro, mshift, mn) && finite;
#else
                    flatEnterMacro(f, g, s_macro, mshift, mn);
#endif
                    prefetchCell();
                    have = true;
                }
            }
            if (base + __popc(idle) >= count) drained = true;
            idle = __ballot_sync(FULL, !have);
        }
        if (idle == FULL) break;
        // ---- per-lane stepping through empty / finished cells
        for (int k = 0; k < kStepBurst; k++) {
            bool stepping = have && f.i >= f.end;
            if (!__any_sync(FULL, stepping)) break;
            if (stepping) {
#if RT_SKIP2
                if (flatAdvance(f, skip)) {
#else
                if (flatLeave(f)) {
#endif
                    have = false;
                    fin = f.w.h.i != 0xFFFFFFFFu;   // result written back in a batch, see the refill step
                } else {
#if RT_SKIP2
                    skip = flatEnterSkip(f, g, s_macro, mshift, mn) && finite;
#else
                    flatEnterMacro(f, g, s_macro, mshift, mn);
#endif
                    prefetchCell();
                }
            }
        }
        // ---- test the references of every pending non-empty cell, as one list of (owner, reference) pairs
        const bool pending = have && f.i < f.end;
        if (!__any_sync(FULL, pending)) continue;
        const unsigned c = pending ? f.end - f.i : 0u;
        unsigned incl = c;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            unsigned v = __shfl_up_sync(FULL, incl, dd);
            if ((int)lane >= dd) incl += v;
        }
        const unsigned P = __shfl_sync(FULL, incl, 31);
        const unsigned off = incl - c;                       // start of this lane's pairs in the list
        const unsigned pmask = __ballot_sync(FULL, pending);
        if (pending) {
            s_plist[wid][__popc(pmask & lt)] = lane;
            s_poff[wid][__popc(pmask & lt)] = off;
        }
        s_best[wid][lane] = NONE;
        __syncwarp();
        unsigned ncand = 0;
        // owner of pair p = the last pending lane whose list offset is <= p.  Rank it instead of searching:
        // pending lanes that start before this batch are counted with one ballot, the ones that start inside
        // it set a bit at their start position (offsets of pending lanes are distinct), and a popc of the
        // bits up to this lane's position gives the rank into the ordered list of pending lanes.
        auto pairOf = [&](unsigned base, unsigned& own, unsigned& ref) -> bool {
            const unsigned p = base + lane;
            const unsigned before = __popc(__ballot_sync(FULL, pending && off <= base));
            const unsigned starts = __reduce_or_sync(FULL, (pending && off > base && off < base + 32) ? (1u << (off - base)) : 0u);
            const unsigned rank = before + __popc(starts & (0xFFFFFFFFu >> (31 - lane)));   // >= 1 for valid pairs
            own = s_plist[wid][(rank ? rank : 1u) - 1u];
            ref = __shfl_sync(FULL, f.i, own) + (p - s_poff[wid][(rank ? rank : 1u) - 1u]);
            return p < P;
        };
        for (unsigned base = 0; base < P; base += 32) {
            unsigned own, ref;
            const bool valid = pairOf(base, own, ref);
            bool pass = valid;
            float dv = 0.f;
            if (PRIM == PRIM_TRIANGLE) {
                const f3 d = mk3(__shfl_sync(FULL, f.w.d.x, own), __shfl_sync(FULL, f.w.d.y, own), __shfl_sync(FULL, f.w.d.z, own));
                if (valid) {
                    float4 q = ldKeep(set.pre_ng + ref);
                    dv = dot(mk3(q.x, q.y, q.z), d);
                    pass = dv > 0;   // the reference's first rejection (div <= 0), on the precomputed face vector
                }
            }
            const unsigned m = __ballot_sync(FULL, pass);
            if (pass) {
                unsigned pos = ncand + __popc(m & lt);
                s_cref[wid][pos] = ref;
                s_cown[wid][pos] = own;
                s_cdiv[wid][pos] = dv;
                if (PRIM == PRIM_TRIANGLE && RT_PF_PE) {
                    prefetchLine<RT_PF_PE>(set.pre_pe + 3 * ref);
                    prefetchLine<RT_PF_PE>(set.pre_pe + 3 * ref + 2);
                }
            }
            ncand += __popc(m);
            __syncwarp();
            if (ncand >= 32) {
                testCandidates(32);
                // keep the (< 32) leftovers at the front of the buffer
                unsigned tr = 0, to = 0;
                float td = 0.f;
                const bool mv = lane + 32 < ncand;
                if (mv) { tr = s_cref[wid][lane + 32]; to = s_cown[wid][lane + 32]; td = s_cdiv[wid][lane + 32]; }
                __syncwarp();
                if (mv) { s_cref[wid][lane] = tr; s_cown[wid][lane] = to; s_cdiv[wid][lane] = td; }
                ncand -= 32;
                __syncwarp();
            }
        }
        if (ncand) testCandidates(ncand);
        __syncwarp();
        if (pending) {
            const unsigned long long key = s_best[wid][lane];
            if (key != NONE) {   // recompute the winner's floats on the owner lane (same operations, same bits)
                const unsigned r = (unsigned)key;
                float ti = 0.f, be = 0.f, ga = 0.f;
                if (PRIM == PRIM_TRIANGLE) {
                    float4 q = __ldg(set.pre_ng + r);
                    float dv = dot(mk3(q.x, q.y, q.z), f.w.d);
                    float4 q0 = __ldg(set.pre_pe + 3 * r), q1 = __ldg(set.pre_pe + 3 * r + 1), q2 = __ldg(set.pre_pe + 3 * r + 2);
                    interTriangleFast<!A07>(f.w.o, f.w.d, f.mint, f.maxt, dv, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), be, ga, ti);
                } else {
                    interSphere(f.w.o, f.w.d, f.w.a_dd, f.mint, f.maxt, __ldg(g.prim + r), ti);
                }
                f.w.h.t = ti;
                f.w.h.i = r;
                f.w.h.beta = be;
                f.w.h.gamma = ga;
            }
            f.i = f.end;   // cell done; flatLeave ends the walk if it produced the hit
        }
        __syncwarp();
    }
}

// Queue filter: drops the entries whose walk is provably empty (walkProvablyEmpty) or cannot produce an accepted hit at all
// (the box is entered beyond the stored maxt: every cell interval starts at or after tmin, every acceptance needs
// t < champ_t = maxt; 0.01 % of slack for the rounding of the first cell boundaries) and writes the survivors, warp-compacted, to the
// queue the walker reads.  Entry order across warps follows the atomics; the walk of a ray does not depend on it.
// Two kernels.  k_filter_mark decides: one warp per 32 entries, no block-wide synchronisation, so a warp only waits for the
// slowest march among its own 32 rays and the SM hides the chain of dependent loads behind its other 63 warps; the decisions
// go out as one 32-bit mask per warp.  k_filter_pack compacts: rounds of 256 entries, warp ballots -> shared prefix -> ONE
// atomicAdd per block and round (all launches of a stage add to one counter and same-address atomics retire at about one per
// nanosecond: with one atomic per warp the first version spent 3.7 ms per launch on 3.3 M of them, three times its own work;
// with the march inside the synchronised rounds it still took 2.7 ms).
constexpr int kFilterBlocks = 8;   // per SM
// RT_OCT_SORT (default on): the survivors leave the filter PARTITIONED, so that the 32 rays a walker warp pops together start in the
// same part of the grid and step the same way through it -- a queue in slot order is coherent for primary rays only; bounce rays
// of neighbouring slots leave in random hemisphere directions, and from the second bounce on they also start anywhere.  The order
// of a queue never changes a result (every ray is independent and writes its own slot), so this is a pure scheduling choice.
//   major key: DIRECTION CLASS -- the cube-map cell of the direction: face of the dominant component (6) x 4 x 4 = 96 classes;
//   minor key (RT_ORG_SORT = G): the G^3 BLOCK OF THE SET'S BOX THE RAY ENTERS IN, Morton order.
// mark writes both codes per entry and the class totals; two scatter passes follow, origin block first (pushed queue -> filtered
// queue, the direction code travels along), direction class second (back into the pushed queue's buffer, which the walker then
// reads).  The second pass is stable round-wise only (rounds of 4096 reach a class's cursor in arrival order): a class receives
// ~43 consecutive entries of one origin block at a time, more than a warp pops.
// Measured on B200, full config, same box each line (Mrays/s; any-hit / closest walks, stage + filter class, ms per step):
//   no partition 7 814 (460 / 253); octant (8 classes) 7 977; octant x dominant axis (24) 8 004-8 048;
//   cube map 4 x 4 (96) 8 033-8 070 (425 / 245); cube map 8 x 8 (384) 8 042 (419 / 246, stage + 4);
//   + origin blocks, on top of cube map 4 x 4 = 8 100 (426 / 246, 500): 4^3 8 104 (409 / 236, 527); 8^3 8 169 (403 / 232, 528);
//   16^3 with rounds of 4096 in the second pass 8 165 -> with a multiplication instead of three divisions in the block key
//   8 324 together with RT_STAGE_MINB_BOUNCE = 4 (395 / 227, 519); per launch: mark 3.5 / 1.9 ms (any / closest; 2.7 / 1.5 without
//   the origin key), origin pass 0.79 ms, direction pass 0.21 ms.
//   Counting the classes with one shared-memory atomic per class present in the warp (match.any) instead of one per entry: slower
//   (stage class 500 -> 508 ms).
#ifndef RT_OCT_SORT
#define RT_OCT_SORT 1
#endif
#if RT_OCT_SORT
constexpr int kDirGrid = 4;   // cube-map cells per face and axis
constexpr int kDirBins = 128;   // >= 6 * kDirGrid^2
typedef unsigned char DirCode;
RT_DEV unsigned dirClass(f3 d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    float m, u, v;
    unsigned face;
    if (ax >= ay && ax >= az) { m = ax; u = d.y; v = d.z; face = d.x < 0.f ? 1u : 0u; }
    else if (ay >= az) { m = ay; u = d.x; v = d.z; face = d.y < 0.f ? 3u : 2u; }
    else { m = az; u = d.x; v = d.y; face = d.z < 0.f ? 5u : 4u; }
    const float h = 0.5f * (float)kDirGrid;
    const float r = m > 0.f ? h / m : 0.f;   // (-1, 1) -> (0, kDirGrid)
    const int iu = min(max((int)(u * r + h), 0), kDirGrid - 1), iv = min(max((int)(v * r + h), 0), kDirGrid - 1);
    return face * (unsigned)(kDirGrid * kDirGrid) + (unsigned)(iu * kDirGrid + iv);
}
#ifndef RT_ORG_SORT
#define RT_ORG_SORT 16   // 0 = direction classes only; <= 16 (Morton key of 4 bits per axis)
#endif
#ifndef RT_PACK2_PER   // round size of the second pass / 256: a class gets ROUND / 96 consecutive entries of one origin block
#define RT_PACK2_PER 16
#endif
constexpr int kOrgGrid = RT_ORG_SORT;
constexpr int kOrgBins = RT_ORG_SORT ? RT_ORG_SORT * RT_ORG_SORT * RT_ORG_SORT : 1;
#if RT_ORG_SORT > 4
typedef unsigned short OrgCode;
#else
typedef unsigned char OrgCode;
#endif
constexpr int kSortCtrs = 2 * kDirBins + 2 * kOrgBins;   // per walk stage: class totals + cursors of either pass
RT_DEV unsigned spread3(unsigned v) {   // 4 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}
RT_DEV unsigned orgClass(f3 o, f3 d, float tmin, const SetDev& set) {
    // the block of the entry point in a kOrgGrid^3 partition of the set's box; org_unit = 1 / extent (a binning key only)
    const AABB& b = set.g.bound;
    const f3 p = o + tmin * d;
    const int ix = min(max((int)((p.x - b.pmin.x) * (set.org_unit[0] * (float)kOrgGrid)), 0), kOrgGrid - 1);
    const int iy = min(max((int)((p.y - b.pmin.y) * (set.org_unit[1] * (float)kOrgGrid)), 0), kOrgGrid - 1);
    const int iz = min(max((int)((p.z - b.pmin.z) * (set.org_unit[2] * (float)kOrgGrid)), 0), kOrgGrid - 1);
    return spread3((unsigned)ix) | (spread3((unsigned)iy) << 1) | (spread3((unsigned)iz) << 2);
}
// codes: [0, cap) DirCode in queue order; with RT_ORG_SORT also [cap, 2 cap) DirCode in origin order and, behind them, OrgCode[cap]
template <bool ANY>
__global__ void __launch_bounds__(256) k_filter_mark(const __grid_constant__ SetDev set, const __grid_constant__ WaveState w, unsigned* masks,
                                                     unsigned n, size_t cap, int qslot) {
    __shared__ unsigned s_hist[kDirBins + kOrgBins];
    for (int k = threadIdx.x; k < kDirBins + kOrgBins; k += 256) s_hist[k] = 0;
    __syncthreads();
    DirCode* codes = (DirCode*)masks;
    OrgCode* ocodes = (OrgCode*)(codes + 2 * cap);
    const unsigned count = w.qctr[2 * qslot];
    const unsigned lane = threadIdx.x & 31;
    const unsigned words = (count + 31u) / 32u;
    for (unsigned word = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; word < words; word += (gridDim.x * blockDim.x) >> 5) {
        const unsigned idx = word * 32u + lane;
        if (idx < count) {
            const unsigned e = ldS(w.queue + idx);
            size_t base = e;
            if (ANY) { const unsigned l = e / n; base = (size_t)(2 * l) * n + (e - l * n); }
            const float4* src = (ANY ? w.sh : w.ray) + base;
            const float4 r0 = ldS(src), r1 = ldS(src + n);
            const f3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r1.x, r1.y, r1.z);
            const AabbHit bi = interAABB(o, d, set.g.bound);
            const bool keep = !(bi.tmin * 0.9999f > r1.w) && !walkProvablyEmpty(o, d, bi.tmin, bi.tmax, set);
            const unsigned cls = dirClass(d);
            // (one shared-memory atomic per entry: aggregating per warp with match.any was measured slower, stage class 500 -> 508 ms)
            if (RT_ORG_SORT) {
                const unsigned oc = orgClass(o, d, bi.tmin, set);
                codes[idx] = (DirCode)(1u + cls);
                ocodes[idx] = keep ? (OrgCode)(1u + oc) : (OrgCode)0;
                if (keep) atomicAdd(&s_hist[kDirBins + oc], 1u);
            } else {
                codes[idx] = keep ? (DirCode)(1u + cls) : (DirCode)0;
            }
            if (keep) atomicAdd(&s_hist[cls], 1u);
        }
    }
    __syncthreads();
    unsigned* ctr = w.qctr + 4 * kMaxStages + kSortCtrs * qslot;
    for (int k = threadIdx.x; k < kDirBins; k += 256)
        if (s_hist[k]) atomicAdd(ctr + k, s_hist[k]);
    if (RT_ORG_SORT)
        for (int k = threadIdx.x; k < kOrgBins; k += 256)
            if (s_hist[kDirBins + k]) atomicAdd(ctr + 2 * kDirBins + k, s_hist[kDirBins + k]);
}

// One scatter pass: the entries of qin with a non-zero code go to their class's segment of qout (classes in index order);
// ctr = BINS class totals followed by BINS cursors.  CARRY moves a byte per entry along (the other pass's code).
template <typename CodeT, int BINS, bool CARRY, int PER>   // PER = entries per thread and round
__global__ void __launch_bounds__(256) k_filter_pack(const unsigned* __restrict__ qin, unsigned* __restrict__ qout, const CodeT* __restrict__ codes,
                                                     const DirCode* __restrict__ carry_in, DirCode* __restrict__ carry_out,
                                                     const unsigned* __restrict__ count_ptr, unsigned* __restrict__ ctr, unsigned* out_count) {
    constexpr unsigned ROUND = 256u * PER;
    __shared__ unsigned s_loc[BINS], s_glob[BINS], s_base[BINS];
    const unsigned count = *count_ptr;
    const unsigned* tot = ctr;
    unsigned* pos = ctr + BINS;
    {   // class segment starts = exclusive scan of the totals, by the whole block: a thread sums its run of bins, the runs are scanned
        // through shared memory, the thread writes its bins' starts.  (One thread walking 4096 totals took ~0.15 ms per launch -- every
        // block of every launch: 2.5 % of a pass at N = 8.)
        constexpr int RUN = (BINS + 255) / 256;
        __shared__ unsigned s_run_small[BINS >= 256 ? 1 : 256];
        unsigned* s_run = BINS >= 256 ? s_loc : s_run_small;   // the round counters are not in use yet
        unsigned v[RUN], sum = 0;
#pragma unroll
        for (int k = 0; k < RUN; k++) { const int b = threadIdx.x * RUN + k; v[k] = b < BINS ? tot[b] : 0u; sum += v[k]; }
        s_run[threadIdx.x] = sum;
        __syncthreads();
        if (threadIdx.x < 32) {   // 256 run sums: 8 per lane, warp scan of the lane sums
            unsigned r[8], ls = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { r[k] = s_run[threadIdx.x * 8 + k]; ls += r[k]; }
            unsigned inc = ls;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += t; }
            unsigned acc = inc - ls;
#pragma unroll
            for (int k = 0; k < 8; k++) { s_run[threadIdx.x * 8 + k] = acc; acc += r[k]; }
            if (threadIdx.x == 31 && blockIdx.x == 0 && out_count) *out_count = inc;   // what the walker will find in the filtered queue
        }
        __syncthreads();
        unsigned acc = s_run[threadIdx.x];
#pragma unroll
        for (int k = 0; k < RUN; k++) { const int b = threadIdx.x * RUN + k; if (b < BINS) s_base[b] = acc; acc += v[k]; }
        __syncthreads();   // s_run (= s_loc) is free again
    }
    const unsigned rounds = (count + ROUND - 1u) / ROUND;
    for (unsigned round = blockIdx.x; round < rounds; round += gridDim.x) {
        for (int k = threadIdx.x; k < BINS; k += 256) s_loc[k] = 0;
        __syncthreads();
        unsigned code[PER], rank[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const unsigned idx = round * ROUND + k * 256u + threadIdx.x;
            code[k] = idx < count ? codes[idx] : 0u;
            rank[k] = code[k] ? atomicAdd(&s_loc[code[k] - 1u], 1u) : 0u;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < BINS; k += 256) s_glob[k] = s_loc[k] ? atomicAdd(pos + k, s_loc[k]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++)
            if (code[k]) {
                const unsigned o = code[k] - 1u;
                const unsigned idx = round * ROUND + k * 256u + threadIdx.x;
                const unsigned dst = s_base[o] + s_glob[o] + rank[k];
                stS(qout + dst, ldS(qin + idx));
                if (CARRY) carry_out[dst] = carry_in[idx];
            }
        __syncthreads();
    }
}
#else
template <bool ANY>
__global__ void __launch_bounds__(256) k_filter_mark(const __grid_constant__ SetDev set, const __grid_constant__ WaveState w, unsigned* masks,
                                                     unsigned n, int qslot) {
    const unsigned count = w.qctr[2 * qslot];
    const unsigned lane = threadIdx.x & 31;
    const unsigned words = (count + 31u) / 32u;
    for (unsigned word = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; word < words; word += (gridDim.x * blockDim.x) >> 5) {
        const unsigned idx = word * 32u + lane;
        bool keep = false;
        if (idx < count) {
            const unsigned e = ldS(w.queue + idx);
            size_t base = e;
            if (ANY) { const unsigned l = e / n; base = (size_t)(2 * l) * n + (e - l * n); }
            const float4* src = (ANY ? w.sh : w.ray) + base;
            const float4 r0 = ldS(src), r1 = ldS(src + n);
            const f3 o = mk3(r0.x, r0.y, r0.z), d = mk3(r1.x, r1.y, r1.z);
            const AabbHit bi = interAABB(o, d, set.g.bound);
            keep = !(bi.tmin * 0.9999f > r1.w) && !walkProvablyEmpty(o, d, bi.tmin, bi.tmax, set);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) masks[word] = m;
    }
}

__global__ void __launch_bounds__(256) k_filter_pack(const __grid_constant__ WaveState w, const __grid_constant__ WaveState wf,
                                                     const unsigned* masks, int qslot) {
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned s_base;
    const unsigned count = w.qctr[2 * qslot];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned rounds = (count + 255u) / 256u;
    for (unsigned round = blockIdx.x; round < rounds; round += gridDim.x) {
        const unsigned idx = round * 256u + threadIdx.x;
        const unsigned m = (round * 8u + wid) * 32u < count ? masks[round * 8u + wid] : 0u;
        const bool keep = (m >> lane) & 1u;
        if (lane == 0) s_cnt[wid] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int k = 0; k < 8; k++) { const unsigned c = s_cnt[k]; s_cnt[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(wf.qctr + 2 * qslot, tot) : 0u;
        }
        __syncthreads();
        if (keep) stS(wf.queue + s_base + s_cnt[wid] + __popc(m & ((1u << lane) - 1u)), ldS(w.queue + idx));
        __syncthreads();   // s_cnt / s_base are reused by the next round
    }
}

#endif   // RT_OCT_SORT
#if RT_OCT_SORT
constexpr size_t kQctrWords = (size_t)(4 + kSortCtrs) * kMaxStages;
constexpr size_t kCodeBytes = RT_ORG_SORT ? 2 * sizeof(DirCode) + sizeof(OrgCode) : sizeof(DirCode);
#else
constexpr size_t kQctrWords = (size_t)4 * kMaxStages;
constexpr size_t kCodeBytes = 0;
#endif

// Builds the stage list for a scene (host) and runs one tile through it.
struct Stage { bool is_walk; StageOp op; int set; bool any; int qslot; };

// every multi-cell grid is walked by the queue walkers; 1-cell grids (all XML sets of A10) are intersected inline
bool isHeavy(const SceneSet& s) { return s.grid.n_slabs > 1 && s.grid.occupancy != nullptr && s.macro_occ != nullptr; }

// Appends the stages that trace the current ray kind against all sets, in set order.  `first` is
// a per-slot op already holding the work that precedes the trace (shade / gen / shadow gen).
void appendTrace(std::vector<Stage>& st, const rt_scene* sc, StageOp first, int kind, int& qslot) {
    StageOp cur = first;
    cur.kind = kind;
    bool cur_used = true;   // `first` must be emitted even if empty of set work
    int nset = (int)sc->sets.size();
    int s = 0;
    StageOp blank = {0, 0, 0, 0, kind, 0, 0, -1, 0, 0, 0, first.light_lo, first.light_hi};
    while (s < nset) {
        if (!isHeavy(sc->sets[s])) {
            int e = s;
            while (e < nset && !isHeavy(sc->sets[e])) e++;
            cur.set_lo = s; cur.set_hi = e;
            cur_used = true;
            s = e;
        } else {
            cur.push_set = s;
            cur.qslot = qslot;
            st.push_back({false, cur, -1, false, 0});
            st.push_back({true, blank, s, kind == 1, qslot});
            qslot++;
            cur = blank;
            cur_used = false;
            s++;
        }
    }
    if (cur_used) st.push_back({false, cur, -1, false, 0});
}

// RT2015_SPLIT_LIGHTS=1 (environment, read once) restores round 1's schedule -- one shadow stage + one any-hit walk per light --
// for A/B runs; the default sends all lights of a segment through one stage and one walk.
bool splitLights() {
    static const bool v = getenv("RT2015_SPLIT_LIGHTS") && atoi(getenv("RT2015_SPLIT_LIGHTS")) != 0;
    return v;
}

int buildStages(const rt_render* r, std::vector<Stage>& st) {
    const rt_scene* sc = r->scene;
    const int nl = (int)sc->lights.size();
    const int group = splitLights() ? 1 : (nl > 0 ? nl : 1);   // lights per shadow stage
    int qslot = 0;
    const StageOp blank = {0, 0, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0};
    int shade_lo = 0, shade_hi = 0;   // lights whose sceneRender still has to run
    for (unsigned seg = 0; seg <= r->o.depth; seg++) {
        StageOp g = blank;
        g.shade = shade_hi > shade_lo; g.shade_lo = shade_lo; g.shade_hi = shade_hi;
        shade_lo = shade_hi = 0;
        g.gen = seg == 0 ? 1 : 2;
        appendTrace(st, sc, g, 0, qslot);
        for (int l0 = 0; l0 < nl; l0 += group) {   // lightRender (primary segment), then this group's shadow rays at once
            StageOp h = blank;
            h.shade = shade_hi > shade_lo; h.shade_lo = shade_lo; h.shade_hi = shade_hi;
            h.light_render = seg == 0 && l0 == 0;
            h.shadow = 1;
            h.light_lo = l0; h.light_hi = l0 + group < nl ? l0 + group : nl;
            appendTrace(st, sc, h, 1, qslot);
            shade_lo = h.light_lo; shade_hi = h.light_hi;
        }
    }
    if (shade_hi > shade_lo) {
        StageOp f = blank;
        f.shade = 1; f.shade_lo = shade_lo; f.shade_hi = shade_hi;
        st.push_back({false, f, -1, false, 0});
    }
    // merge adjacent per-slot stages where the second one carries no set work of its own that
    // depends on a walk in between (they are adjacent only when no walk separates them)
    std::vector<Stage> out;
    for (const Stage& s : st) {
        if (!out.empty() && !s.is_walk && !out.back().is_walk) {
            StageOp& a = out.back().op;
            const StageOp& b = s.op;
            bool a_traces = a.set_hi > a.set_lo || a.push_set >= 0;
            // b's shade/gen/shadow come after a's set work in program order: only merge when a has none
            if (!a_traces && !a.shadow && a.gen == 0 && !a.light_render && !b.shade) {
                StageOp m = b;
                m.shade = a.shade; m.shade_lo = a.shade_lo; m.shade_hi = a.shade_hi;
                a = m;
                continue;
            }
        }
        out.push_back(s);
    }
    st.swap(out);
    if (qslot > kMaxStages) return RT_ERR_INVALID;
    return RT_OK;
}

int ensureWaveBuffers(rt_render* r) {
    if (r->w_ray) return RT_OK;
    rt_ctx* ctx = r->ctx;
    size_t n = r->tile_slots;
    int rc = RT_OK;
    auto A = [&](void** p, size_t bytes) { if (!rc) rc = rt_buffer_create(ctx, bytes, p); };
    A((void**)&r->w_ray, sizeof(float4) * 2 * n);
    A((void**)&r->w_poi, sizeof(float4) * 2 * n);
    A((void**)&r->w_atte, sizeof(float4) * n);
    const size_t nl = r->scene->lights.size() ? r->scene->lights.size() : 1;
    A((void**)&r->w_sh, sizeof(float4) * 2 * n * nl);      // one shadow ray per light and slot
    A((void**)&r->w_queue, sizeof(unsigned) * n * nl);     // any-hit walks queue (light, slot) entries
    A((void**)&r->w_qctr, sizeof(unsigned) * kQctrWords);  // {count, head} per walk stage: [0, 2K) as pushed, [2K, 4K) filtered; behind them the class totals / cursors of the filter's partition
    if (skipEmptyWalks()) {
        A((void**)&r->w_queue_f, sizeof(unsigned) * n * nl);              // what the queue filter leaves for the walker
        A((void**)&r->w_masks, kCodeBytes ? kCodeBytes * (n * nl + 64) : sizeof(unsigned) * ((n * nl + 31) / 32 + 8));   // its decisions: a class code (or one keep bit) per entry
    }
    return rc;
}

// Picks the k_stage instantiation for a stage shape.  gen is 0/1/2, the rest are flags: 48 shapes exist on paper,
// a pass uses about six of them; all are instantiated through one recursive dispatch.
template <int GEN, bool LR, bool SHADE, bool SHADOW, int KIND>
void launchStageT(rt_ctx* ctx, const SceneDev& sc, const PathArgs& a, const WaveState& w, const StageOp& op, unsigned n) {
    k_stage<GEN, LR, SHADE, SHADOW, KIND><<<rt_blocks(n, 256), 256, 0, ctx->stream>>>(sc, a, w, op);
}
template <int GEN, bool LR, bool SHADE>
void launchStage2(rt_ctx* ctx, const SceneDev& sc, const PathArgs& a, const WaveState& w, const StageOp& op, unsigned n) {
    const bool shadow = op.shadow != 0;
    if (op.kind == 0) {
        if (shadow) launchStageT<GEN, LR, SHADE, true, 0>(ctx, sc, a, w, op, n); else launchStageT<GEN, LR, SHADE, false, 0>(ctx, sc, a, w, op, n);
    } else {
        if (shadow) launchStageT<GEN, LR, SHADE, true, 1>(ctx, sc, a, w, op, n); else launchStageT<GEN, LR, SHADE, false, 1>(ctx, sc, a, w, op, n);
    }
}
template <int GEN>
void launchStage1(rt_ctx* ctx, const SceneDev& sc, const PathArgs& a, const WaveState& w, const StageOp& op, unsigned n) {
    const bool shade = op.shade != 0;
    if (op.light_render) {
        if (shade) launchStage2<GEN, true, true>(ctx, sc, a, w, op, n); else launchStage2<GEN, true, false>(ctx, sc, a, w, op, n);
    } else {
        if (shade) launchStage2<GEN, false, true>(ctx, sc, a, w, op, n); else launchStage2<GEN, false, false>(ctx, sc, a, w, op, n);
    }
}
void launchStage(rt_ctx* ctx, const SceneDev& sc, const PathArgs& a, const WaveState& w, const StageOp& op, unsigned n) {
    if (op.gen == 1) launchStage1<1>(ctx, sc, a, w, op, n);
    else if (op.gen == 2) launchStage1<2>(ctx, sc, a, w, op, n);
    else launchStage1<0>(ctx, sc, a, w, op, n);
}

int waveTile(rt_render* r, const SceneDev& sc, const PathArgs& a) {
    rt_ctx* ctx = r->ctx;
    int rc = ensureWaveBuffers(r);
    if (rc) return rc;
    std::vector<Stage> stages;
    rc = buildStages(r, stages);
    if (rc) {   // more walk stages than queue counters: run the tile through the megakernel instead
        RT_TRY_W(rt_seeds_ready(r));
        RT_TRY_W(rt_time_mark(r, 5));
        k_pathMega<false><<<rt_blocks(a.n_local, 128), 128, 0, ctx->stream>>>(sc, a);
        RT_LAUNCH_CHECK(ctx, "pathMega");
        return RT_OK;
    }
    WaveState w = {r->w_ray, r->w_poi, r->w_atte, r->w_sh, r->w_queue, r->w_qctr};
    WaveState wf = w;   // the filtered queue (k_filter) the walkers read when the empty-walk proof is on
    wf.queue = r->w_queue_f;
    wf.qctr = r->w_qctr + 2 * kMaxStages;
    WaveState wg = wf;   // what the walkers read: with the two-pass partition the second pass lands in the pushed queue's buffer
#if RT_OCT_SORT
    if (RT_ORG_SORT) wg.queue = r->w_queue;
#endif
    RT_CUDA(ctx, cudaMemsetAsync(r->w_qctr, 0, sizeof(unsigned) * kQctrWords, ctx->stream));
    const unsigned n = a.n_local;
    const WaveState& w0 = w;
    for (const Stage& s : stages) {
        if (!s.is_walk) {
            // a seed upload still in flight (rt_render_write_local_seeds_async) is waited for here, in front of the
            // first stage that draws random numbers: bouncePaths / initShadowTrace (initTrace only when rpp == 1,
            // and then rt_render_execute has waited already)
            if (s.op.gen == 2 || s.op.shadow) RT_TRY_W(rt_seeds_ready(r));
            RT_TRY_W(rt_time_mark(r, 0));
            launchStage(ctx, sc, a, w, s.op, n);
            RT_LAUNCH_CHECK(ctx, "wave_stage");
        } else {
            const SetDev& set = sc.sets[s.set];
            const int walk_blocks = ctx->prop.multiProcessorCount * kWalkMinBlocks;   // persistent: one resident wave
            const bool filtered = set.macro_dist != nullptr && r->w_queue_f != nullptr;
            if (filtered) {
                RT_TRY_W(rt_time_mark(r, 0));   // charged to the stage class: it is queue preparation
                const int fb = ctx->prop.multiProcessorCount * kFilterBlocks;
#if RT_OCT_SORT
                const size_t cap = (size_t)n * (r->scene->lights.size() ? r->scene->lights.size() : 1) + 64;
                DirCode* dc = (DirCode*)r->w_masks;
                unsigned* ctr = r->w_qctr + 4 * kMaxStages + kSortCtrs * s.qslot;
                if (s.any) k_filter_mark<true><<<fb, 256, 0, ctx->stream>>>(set, w, r->w_masks, n, cap, s.qslot);
                else k_filter_mark<false><<<fb, 256, 0, ctx->stream>>>(set, w, r->w_masks, n, cap, s.qslot);
                RT_LAUNCH_CHECK(ctx, "wave_filter_mark");
                if (RT_ORG_SORT) {   // origin block first (into the filtered queue), direction class second (back into the pushed queue)
                    k_filter_pack<OrgCode, kOrgBins, true, (kOrgBins > 512 ? 16 : 4)><<<fb, 256, 0, ctx->stream>>>(w.queue, wf.queue, (const OrgCode*)(dc + 2 * cap), dc, dc + cap,
                                                                                       w.qctr + 2 * s.qslot, ctr + 2 * kDirBins, wf.qctr + 2 * s.qslot);
                    RT_LAUNCH_CHECK(ctx, "wave_filter_pack_origin");
                    k_filter_pack<DirCode, kDirBins, false, RT_PACK2_PER><<<fb, 256, 0, ctx->stream>>>(wf.queue, w.queue, dc + cap, nullptr, nullptr,
                                                                                        wf.qctr + 2 * s.qslot, ctr, nullptr);
                } else {
                    k_filter_pack<DirCode, kDirBins, false, 4><<<fb, 256, 0, ctx->stream>>>(w.queue, wf.queue, dc, nullptr, nullptr, w.qctr + 2 * s.qslot, ctr,
                                                                                        wf.qctr + 2 * s.qslot);
                }
#else
                if (s.any) k_filter_mark<true><<<fb, 256, 0, ctx->stream>>>(set, w, r->w_masks, n, s.qslot);
                else k_filter_mark<false><<<fb, 256, 0, ctx->stream>>>(set, w, r->w_masks, n, s.qslot);
                RT_LAUNCH_CHECK(ctx, "wave_filter_mark");
                k_filter_pack<<<fb, 256, 0, ctx->stream>>>(w, wf, r->w_masks, s.qslot);
#endif
                RT_LAUNCH_CHECK(ctx, "wave_filter_pack");
            }
            const WaveState& w = filtered ? wg : *&w0;
            RT_TRY_W(rt_time_mark(r, (set.kind == PRIM_SPHERE ? 1 : 3) + (s.any ? 1 : 0)));
            if (set.kind == PRIM_SPHERE) {
                if (s.any) k_walk_pairs<PRIM_SPHERE, true><<<walk_blocks, kWalkWarps * 32, 0, ctx->stream>>>(set, w, n, s.qslot);
                else k_walk_pairs<PRIM_SPHERE, false><<<walk_blocks, kWalkWarps * 32, 0, ctx->stream>>>(set, w, n, s.qslot);
            } else {
                if (s.any) k_walk_pairs<PRIM_TRIANGLE, true><<<walk_blocks, kWalkWarps * 32, 0, ctx->stream>>>(set, w, n, s.qslot);
                else k_walk_pairs<PRIM_TRIANGLE, false><<<walk_blocks, kWalkWarps * 32, 0, ctx->stream>>>(set, w, n, s.qslot);
            }
            RT_LAUNCH_CHECK(ctx, "wave_walk");
        }
    }
    return RT_OK;
}

}  // namespace

namespace {
// ---------------------------------------------------------------------------------------
// Assignment-7 molTrace / meshTrace of a library-built grid through the queue walker (config 3).  The reference kernel is one
// work-item per pixel (A07/code.cl:337-626); one thread per ray on the GPU leaves 0.58 M coherent rays of a 1080p frame
// waiting on dependent loads (1.5 ms at 1 M triangles, 4.6 x the cost per ray of the path tracer's incoherent walks).
// Here: prepare (Ray records -> the walker's SoA form + a queue of the rays that hit the grid's box), walk (k_walk_pairs with
// the exclusive triangle test), finish (ray.maxt back into the Ray records, the debug colour by cell parity into the pixels).
// Grids with up to 1023 slabs per axis (the hit's cell travels in 3 x 10 bits).
// ---------------------------------------------------------------------------------------
__global__ void k_a07_prepare(const Ray* rays, unsigned n, const __grid_constant__ SetDev set, const __grid_constant__ WaveState w) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    if (id < n) {
        const RayR ray = loadRay(rays + id);
        w.ray[id] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.mint);
        w.ray[n + id] = make_float4(ray.d.x, ray.d.y, ray.d.z, ray.maxt);
        w.poi[id] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (ray.mint != ray.maxt) want = interAABB(ray.o, ray.d, set.g.bound).v;
    }
    pushTask(want, id, w.queue, w.qctr);
}

template <int PRIM>
__global__ void k_a07_finish(uchar4* pixels, CamArg fcam, Ray* rays, unsigned n, const __grid_constant__ WaveState w) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const float4 p0 = w.poi[id];
    const int cell = __float_as_int(p0.w);
    if (cell < 0) return;   // no hit: the reference writes nothing
    const float4 p1 = w.poi[n + id];
    rays[id].maxt = w.ray[n + id].w;
    const Camera cam = floatToCamera(fcam.v);
    const float shade = cl_clamp(dot(cam.W, mk3(p1.x, p1.y, p1.z)), 0.0f, 1.0f);
    const float s = shade * 127.0f;
    const int cx = cell & 1023, cy = (cell >> 10) & 1023, cz = (cell >> 20) & 1023;
    const float r = (float)((cx % 2) + 1) * s, g = (float)((cy % 2) + 1) * s, b = (float)((cz % 2) + 1) * s;
    pixels[id] = make_uchar4((unsigned char)(int)r, (unsigned char)(int)g, (unsigned char)(int)b, 255);   // the conversion of cellParityColor (rt_kernels_a0x.cu)
}
}  // namespace

int rt_walk_a07(rt_ctx* ctx, rt_ctx::GridAux& g, int prim, void* pixels, const float* fcam, void* rays, const void* normals, const float* bound,
                size_t npix) {
    RT_TRY_W(rt_grid_aux_build(ctx, g));
    const unsigned n = (unsigned)npix;
    SetDev set;
    memset(&set, 0, sizeof set);
    set.g.prim = (const float4*)g.prim;
    set.g.box = (const unsigned*)g.box_size;
    set.g.occ = g.occupancy;
    set.g.bound.pmin = f3{bound[0], bound[1], bound[2]};
    set.g.bound.pmax = f3{bound[4], bound[5], bound[6]};
    set.g.n = g.n_slabs;
    set.normals = (const float4*)normals;
    set.kind = prim;
    set.use_occ = 1;
    set.pre_ng = g.pre_ng;
    set.pre_pe = g.pre_pe;
    set.macro_occ = g.macro_occ;
    set.macro_shift = g.macro_shift;
    set.macro_n = g.macro_n;
    set.n_refs = g.n_refs;
    // scratch: ray (2n) + hit record (2n) float4, queue (n), counters -- from the context's pool (kept between frames)
    char* scratch = nullptr;
    const size_t bytes = sizeof(float4) * 4 * (size_t)n + sizeof(unsigned) * ((size_t)n + 8);
    RT_CUDA(ctx, rt_scratch_alloc(ctx, (void**)&scratch, bytes));
    WaveState w;
    memset(&w, 0, sizeof w);
    w.ray = (float4*)scratch;
    w.poi = w.ray + 2 * (size_t)n;
    w.queue = (unsigned*)(w.poi + 2 * (size_t)n);
    w.qctr = w.queue + n;
    RT_CUDA(ctx, cudaMemsetAsync(w.qctr, 0, sizeof(unsigned) * 8, ctx->stream));
    CamArg cam;
    memcpy(cam.v, fcam, sizeof cam.v);
    k_a07_prepare<<<rt_blocks(n, 256), 256, 0, ctx->stream>>>((const Ray*)rays, n, set, w);
    RT_LAUNCH_CHECK(ctx, "A07 prepare");
    const int walk_blocks = ctx->prop.multiProcessorCount * kWalkMinBlocks;
    if (prim == PRIM_SPHERE) k_walk_pairs<PRIM_SPHERE, false, true><<<walk_blocks, kWalkWarps * 32, 0, ctx->stream>>>(set, w, n, 0);
    else k_walk_pairs<PRIM_TRIANGLE, false, true><<<walk_blocks, kWalkWarps * 32, 0, ctx->stream>>>(set, w, n, 0);
    RT_LAUNCH_CHECK(ctx, "A07 walk");
    if (prim == PRIM_SPHERE) k_a07_finish<PRIM_SPHERE><<<rt_blocks(n, 256), 256, 0, ctx->stream>>>((uchar4*)pixels, cam, (Ray*)rays, n, w);
    else k_a07_finish<PRIM_TRIANGLE><<<rt_blocks(n, 256), 256, 0, ctx->stream>>>((uchar4*)pixels, cam, (Ray*)rays, n, w);
    RT_LAUNCH_CHECK(ctx, "A07 finish");
    RT_CUDA(ctx, cudaFreeAsync(scratch, ctx->stream));
    return RT_OK;
}

namespace {
__global__ void k_skipProbe(const __grid_constant__ SetDev set, const Ray* rays, unsigned n, unsigned char* out) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    RayR ray = loadRay(rays + id);
    unsigned char flag = 0;
    if (ray.mint != ray.maxt) {
        if (set.g.n > 1) {
            const AabbHit bi = interAABB(ray.o, ray.d, set.g.bound);
            if (bi.v && walkProvablyEmpty(ray.o, ray.d, bi.tmin, bi.tmax, set)) flag = 1;
        } else if (set.wall_ok && shadowClearsWalls(set, ray)) {
            flag = 1;
        }
    }
    out[id] = flag;
}
}  // namespace

// Diagnostic: which rays of a Ray buffer would the wavefront path NOT send to the queue walker of geometry set `set_index`
// because their walk provably crosses empty cells only (walkProvablyEmpty)?  out_flags: one byte per ray.
extern "C" int rt_scene_probe_empty_walks(rt_scene* s, unsigned set_index, const void* rays, unsigned n, void* out_flags_u8) {
    if (!s || !rays || !out_flags_u8 || set_index >= s->sets.size()) return RT_ERR_INVALID;
    rt_ctx* ctx = s->ctx;
    if (!n) return RT_OK;
    rt_render tmp;
    tmp.ctx = ctx;
    tmp.scene = s;
    SceneDev sc;
    int rc = buildSceneDev(&tmp, sc);
    if (rc) return rc;
    k_skipProbe<<<rt_blocks(n, 256), 256, 0, ctx->stream>>>(sc.sets[set_index], (const Ray*)rays, n, (unsigned char*)out_flags_u8);
    RT_LAUNCH_CHECK(ctx, "skipProbe");
    return RT_OK;
}

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
    if (r->profile) {   // work counters: the instrumented megakernel does the same per-slot work
        RT_TRY_W(rt_time_mark(r, 5));
        k_pathMega<true><<<rt_blocks(n, 128), 128, 0, ctx->stream>>>(sc, a);
        RT_LAUNCH_CHECK(ctx, "pathMega(profile)");
        return RT_OK;
    }
    bool any_heavy = false;
    for (const SceneSet& s : r->scene->sets) any_heavy = any_heavy || isHeavy(s);
    if (o.mode == 2 || !any_heavy) {   // megakernel: explicit, or nothing would be queued anyway
        RT_TRY_W(rt_seeds_ready(r));
        RT_TRY_W(rt_time_mark(r, 5));
        k_pathMega<false><<<rt_blocks(n, 128), 128, 0, ctx->stream>>>(sc, a);
        RT_LAUNCH_CHECK(ctx, "pathMega");
        return RT_OK;
    }
    return waveTile(r, sc, a);
}
