// rt_frame.cu -- scene object and progressive path-tracing renderer behind the C ABI.
// Mirrors preRender / executeRender / postRender of Assign10-Path_Tracing/code.js:1784-1859.
//
// Ray slots are independent (every reference kernel indexes only its own slot id), so the
// frame is processed in TILES of consecutive slots (a quarter of device memory by default, see
// rt_render_create: the queue walkers want deep queues, the state streams through HBM either way);
// only the per-slot seed and accumulation buffers persist across passes, exactly as in the
// reference (acu is never cleared between passes, A10/code.js:1078-1099).
#include <math.h>

#include "rt_frame.h"

using namespace rt;

namespace {

constexpr unsigned kBlock = 256;

// Tile-aware initTrace (A10/code.cl:458-543): local slot id -> (pixel, global k).
__global__ void f_initTrace(Ray* rays, Poi10* pois, AabbArg bound_a, CamArg fcam, float focal_length, float lens_rad,
                            unsigned rays_per_pixel, unsigned slot_begin, unsigned slots_pp, size_t pixel_base, unsigned n_local,
                            const float2* rpp1_coords) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_local) return;
    Camera cam = floatToCamera(fcam.v);
    AABB bound = toAABB(bound_a);
    size_t pix = pixel_base + id / slots_pp;
    unsigned k = slot_begin + id % slots_pp;
    unsigned col = (unsigned)(pix % cam.cols), row = (unsigned)(pix / cam.cols);
    float4* pq = reinterpret_cast<float4*>(pois + id);
    pq[2] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    pois[id].matId = -1;
    f2 coord;
    if (rays_per_pixel > 1) {
        unsigned side = (unsigned)sqrtf((float)rays_per_pixel);
        if (k >= side * side) { storeDeadRay(rays + id); return; }
        unsigned i = k / side, j = k % side;
        float delta = 1.0f / (float)side;
        coord.y = delta / 2.0f;
        for (unsigned a = 0; a < i; a++) coord.y += delta;
        coord.x = delta / 2.0f;
        for (unsigned a = 0; a < j; a++) coord.x += delta;
    } else {
        float2 c = rpp1_coords[pix];
        coord.x = c.x;
        coord.y = c.y;
    }
    f3 focal_point = getFocalPoint(cam, (float)col, (float)row, focal_length);
    RayR ray;
    getThinLensRay(cam, focal_point, lens_rad, coord, ray.o, ray.d);
    AabbHit inter = interAABB(ray.o, ray.d, bound);
    if (inter.v) { ray.mint = inter.tmin; ray.maxt = inter.tmax; }
    else { ray.mint = RT_INF; ray.maxt = RT_INF; }
    storeRay(rays + id, ray);
}

// rpp == 1 seeds (quirk Q7, see rt_kernels_a10.cu): column `col` serves its rows in order.
__global__ void f_rpp1_coords(int* seeds, float2* coords, unsigned cols, unsigned rows) {
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

__global__ void f_countValid(const Ray* rays, unsigned n, unsigned long long* counter) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    if (id < n) {
        float4 c = reinterpret_cast<const float4*>(rays + id)[2];
        valid = c.x != c.y;
    }
    unsigned b = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(counter, (unsigned long long)__popc(b));
}

// accum[pixel] = sum_k acu[pixel][k] in k order (the summation of copyToPixel, A10/code.cl:1376-1380)
__global__ void f_sumSlots(const float4* acu, float4* accum, size_t pixels, unsigned slots_pp) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pixels) return;
    const float4* a = acu + id * slots_pp;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned i = 0; i < slots_pp; i++) {
        float4 v = a[i];
        c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
    }
    accum[id] = c;
}

// copyToPixel tail on a per-pixel sum (A10/code.cl:1381-1384).
__global__ void f_accumToPixel(uchar4* pixel, const float4* accum, float m, unsigned pixels) {
    unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pixels) return;
    float4 color = accum[id];
    float s = 255.0f * m;
    color.x *= s; color.y *= s; color.z *= s;
    color.x *= 1.8f; color.y *= 1.8f; color.z *= 1.8f;
    color.x = cl_clamp(color.x, 0.0f, 255.0f);
    color.y = cl_clamp(color.y, 0.0f, 255.0f);
    color.z = cl_clamp(color.z, 0.0f, 255.0f);
    pixel[id] = make_uchar4((unsigned char)color.x, (unsigned char)color.y, (unsigned char)color.z, 255);
}

// Face vector and edges of every triangle reference, with exactly the operations of
// interTriangle (A10/code.cl:252-256): e1 = p1 - p0, e2 = p2 - p0, ng = cross(e2, e1).
__global__ void f_precomputeTriangles(const float4* prim, unsigned n_refs, float4* ng, float4* pe) {
    unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_refs) return;
    float4 q0 = prim[3 * r], q1 = prim[3 * r + 1], q2 = prim[3 * r + 2];
    f3 p0 = mk3(q0.x, q0.y, q0.z), p1 = mk3(q1.x, q1.y, q1.z), p2 = mk3(q2.x, q2.y, q2.z);
    f3 e1 = p1 - p0, e2 = p2 - p0;
    f3 g = cross(e2, e1);
    ng[r] = make_float4(g.x, g.y, g.z, 0.f);
    pe[3 * r] = make_float4(p0.x, p0.y, p0.z, 0.f);
    pe[3 * r + 1] = make_float4(e1.x, e1.y, e1.z, 0.f);
    pe[3 * r + 2] = make_float4(e2.x, e2.y, e2.z, 0.f);
}

// Coarse occupancy: bit (mz,my,mx) = OR of the fine bits of its (2^sh)^3 cells.
__global__ void f_macroOccupancy(const unsigned* occ, unsigned n, unsigned sh, unsigned nm, unsigned* macro) {
    size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t mcells = (size_t)nm * nm * nm;
    bool any = false;
    if (m < mcells) {
        unsigned mx = (unsigned)(m % nm), my = (unsigned)((m / nm) % nm), mz = (unsigned)(m / ((size_t)nm * nm));
        unsigned f = 1u << sh;
        for (unsigned z = mz << sh; z < min(n, (mz << sh) + f) && !any; z++)
            for (unsigned y = my << sh; y < min(n, (my << sh) + f) && !any; y++)
                for (unsigned x = mx << sh; x < min(n, (mx << sh) + f); x++) {
                    size_t c = ((size_t)z * n + y) * n + x;
                    if ((occ[c >> 5] >> (c & 31)) & 1u) { any = true; break; }
                }
    }
    unsigned bits = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && (m >> 5) < 8192) macro[m >> 5] = bits;
}

// Occupancy bits of a cell table that came without them (a grid the caller built): bit c = cell c is non-empty.
__global__ void f_occupancyFromTable(const unsigned* box, size_t cells, unsigned* occ) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool full = c < cells && box[c + 1] > box[c];
    unsigned bits = __ballot_sync(0xffffffffu, full);
    if ((threadIdx.x & 31) == 0 && c < cells) occ[c >> 5] = bits;
}

#define RT_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc) return _rc;      \
    } while (0)

int closestAll(rt_render* r, unsigned n, Poi10* pois, Ray* rays) {
    rt_ctx* ctx = r->ctx;
    f_countValid<<<rt_blocks(n, kBlock), kBlock, 0, ctx->stream>>>(rays, n, r->d_counters + 0);
    RT_LAUNCH_CHECK(ctx, "countValid");
    for (const SceneSet& s : r->scene->sets) {
        if (s.grid.kind == 0)
            RT_TRY(rt_a10_sphereTrace(ctx, n, pois, rays, s.grid.prim, s.grid.matid, s.grid.box_size, s.bound, s.grid.n_slabs));
        else if (!s.is_mesh)
            RT_TRY(rt_a10_triangleTrace(ctx, n, pois, rays, s.grid.prim, s.grid.normal, s.grid.matid, s.grid.box_size, s.bound, s.grid.n_slabs));
        else
            RT_TRY(rt_a10_meshTrace(ctx, n, pois, rays, s.grid.prim, s.grid.normal, s.grid.box_size, s.mesh_matid, s.bound, s.grid.n_slabs));
    }
    return RT_OK;
}

int shadeAll(rt_render* r, unsigned n, Poi10* pois, Ray* shadow, float4* acu, int* seeds) {
    rt_ctx* ctx = r->ctx;
    for (const SceneLight& L : r->scene->lights) {
        RT_TRY(rt_a10_initShadowTrace(ctx, shadow, pois, n, L.shadow, seeds));
        f_countValid<<<rt_blocks(n, kBlock), kBlock, 0, ctx->stream>>>(shadow, n, r->d_counters + 1);
        RT_LAUNCH_CHECK(ctx, "countValid");
        for (const SceneSet& s : r->scene->sets) {
            if (s.grid.kind == 0)
                RT_TRY(rt_a10_sphereShadowTrace(ctx, n, shadow, s.grid.prim, s.grid.box_size, s.bound, s.grid.n_slabs));
            else
                RT_TRY(rt_a10_triangleShadowTrace(ctx, n, shadow, s.grid.prim, s.grid.box_size, s.bound, s.grid.n_slabs));
        }
        RT_TRY(rt_a10_sceneRender(ctx, acu, pois, shadow, r->scene->materials, L.scene, n));
    }
    return RT_OK;
}

// One tile through the reference's kernel sequence (executeRender, A10/code.js:1806-1846).
int tileReferenceSchedule(rt_render* r, const float* fcam, size_t slot0, unsigned n, const float2* rpp1_coords) {
    rt_ctx* ctx = r->ctx;
    const rt_render_opts& o = r->o;
    AabbArg ba; memcpy(ba.v, r->scene->bound, sizeof ba.v);
    CamArg ca; memcpy(ca.v, fcam, sizeof ca.v);
    f_initTrace<<<rt_blocks(n, kBlock), kBlock, 0, ctx->stream>>>(r->rays, r->pois, ba, ca, o.focal_length, o.lens_rad, o.rays_per_pixel,
                                                                   o.slot_begin, r->slots_pp, slot0 / r->slots_pp, n, rpp1_coords);
    RT_LAUNCH_CHECK(ctx, "initTrace");
    float4* acu = r->acu + slot0;
    int* seeds = r->seeds + slot0;
    RT_TRY(closestAll(r, n, r->pois, r->rays));
    for (const SceneLight& L : r->scene->lights) RT_TRY(rt_a10_lightRender(ctx, r->pois, r->rays, acu, L.light, n));
    RT_TRY(shadeAll(r, n, r->pois, r->shadow, acu, seeds));
    for (unsigned j = 0; j < o.depth; j++) {
        RT_TRY(rt_a10_bouncePaths(ctx, r->pois, r->rays, seeds, n));
        RT_TRY(closestAll(r, n, r->pois, r->rays));
        RT_TRY(shadeAll(r, n, r->pois, r->shadow, acu, seeds));
    }
    return RT_OK;
}

}  // namespace

int rt_seeds_ready(rt_render* r) {
    if (!r->seeds_in_flight) return RT_OK;
    RT_CUDA(r->ctx, cudaStreamWaitEvent(r->ctx->stream, r->ev_seeds, 0));
    r->seeds_in_flight = false;
    return RT_OK;
}

int rt_time_mark(rt_render* r, int cls) {
    if (!r->timing) return RT_OK;
    rt_ctx* ctx = r->ctx;
    if (r->tev_used == r->tev.size()) {
        cudaEvent_t e;
        RT_CUDA(ctx, cudaEventCreate(&e));
        r->tev.push_back(e);
        r->tcls.push_back(0);
    }
    r->tcls[r->tev_used] = cls;
    RT_CUDA(ctx, cudaEventRecord(r->tev[r->tev_used], ctx->stream));
    r->tev_used++;
    return RT_OK;
}

// Face vectors / edge form (triangle grids) and the coarse occupancy of a registered multi-cell grid, for the queue-walker route of
// the Assignment-7 launchers: built once, on first use, released with the grid.
int rt_grid_aux_build(rt_ctx* ctx, rt_ctx::GridAux& g) {
    if (g.aux_ready) return RT_OK;
    if (!g.occupancy || g.n_slabs < 2 || g.dims != 3) return RT_ERR_INVALID;
    unsigned sh = 0;
    while (((g.n_slabs + (1u << sh) - 1) >> sh) > 64) sh++;
    g.macro_shift = sh;
    g.macro_n = (g.n_slabs + (1u << sh) - 1) >> sh;
    const size_t mcells = (size_t)g.macro_n * g.macro_n * g.macro_n;
    RT_TRY(rt_buffer_create(ctx, sizeof(unsigned) * 8192, (void**)&g.macro_occ));
    RT_TRY(rt_buffer_fill(ctx, g.macro_occ, 0, sizeof(unsigned) * 8192));
    f_macroOccupancy<<<rt_blocks((mcells + 31) / 32 * 32, kBlock), kBlock, 0, ctx->stream>>>(g.occupancy, g.n_slabs, sh, g.macro_n, g.macro_occ);
    RT_LAUNCH_CHECK(ctx, "macroOccupancy");
    if (g.kind == 1 && g.n_refs > 0) {
        RT_TRY(rt_buffer_create(ctx, sizeof(float4) * g.n_refs, (void**)&g.pre_ng));
        RT_TRY(rt_buffer_create(ctx, sizeof(float4) * 3 * (size_t)g.n_refs, (void**)&g.pre_pe));
        f_precomputeTriangles<<<rt_blocks(g.n_refs, kBlock), kBlock, 0, ctx->stream>>>((const float4*)g.prim, g.n_refs, g.pre_ng, g.pre_pe);
        RT_LAUNCH_CHECK(ctx, "precomputeTriangles");
    }
    g.aux_ready = true;
    return RT_OK;
}

extern "C" {

int rt_scene_create(rt_ctx* ctx, rt_scene** out) {
    RT_CHECK_CTX(ctx);
    if (!out) return RT_ERR_INVALID;
    rt_scene* s = new rt_scene();
    s->ctx = ctx;
    *out = s;
    return RT_OK;
}

int rt_scene_destroy(rt_scene* s) {
    if (!s) return RT_ERR_INVALID;
    if (s->materials) rt_buffer_release(s->ctx, s->materials);
    for (SceneSet& st : s->sets) {
        if (st.pre_ng) rt_buffer_release(s->ctx, st.pre_ng);
        if (st.pre_pe) rt_buffer_release(s->ctx, st.pre_pe);
        if (st.macro_occ) rt_buffer_release(s->ctx, st.macro_occ);
        if (st.own_occ) rt_buffer_release(s->ctx, st.own_occ);
        if (st.macro_dist) rt_buffer_release(s->ctx, st.macro_dist);
    }
    delete s;
    return RT_OK;
}

int rt_scene_set_bounds(rt_scene* s, const float bound[8]) {
    if (!s || !bound) return RT_ERR_INVALID;
    memcpy(s->bound, bound, sizeof s->bound);
    return RT_OK;
}

int rt_scene_set_materials(rt_scene* s, const float* rgba, unsigned n_materials) {
    if (!s || (!rgba && n_materials)) return RT_ERR_INVALID;
    if (s->materials) { rt_buffer_release(s->ctx, s->materials); s->materials = nullptr; }
    RT_TRY(rt_buffer_create(s->ctx, sizeof(float) * 4 * (n_materials ? n_materials : 1), &s->materials));
    if (n_materials) RT_TRY(rt_buffer_write(s->ctx, s->materials, 0, sizeof(float) * 4 * n_materials, rgba));
    s->n_materials = n_materials;
    return RT_OK;
}

int rt_scene_add_set(rt_scene* s, const rt_grid* grid, const float bound[8], int is_mesh, unsigned mesh_matid) {
    if (!s || !grid || !bound || !grid->box_size || !grid->prim) return RT_ERR_INVALID;
    if (grid->kind == 1 && !grid->normal) return RT_ERR_INVALID;
    SceneSet st;
    st.grid = *grid;
    memcpy(st.bound, bound, sizeof st.bound);
    st.is_mesh = is_mesh;
    st.mesh_matid = mesh_matid;
    if (grid->n_slabs > 1 && !grid->occupancy) {
        // a caller-built grid (the drop-in ABI accepts any box_size / prim buffers): the fused paths walk multi-cell
        // grids through the occupancy bits, so derive them from the cell table here (owned by the scene)
        rt_ctx* ctx = s->ctx;
        size_t cells = (size_t)grid->n_slabs * grid->n_slabs * grid->n_slabs;
        RT_TRY(rt_buffer_create(ctx, sizeof(unsigned) * ((cells + 31) / 32), (void**)&st.own_occ));
        f_occupancyFromTable<<<rt_blocks((cells + 31) / 32 * 32, kBlock), kBlock, 0, ctx->stream>>>((const unsigned*)grid->box_size, cells, st.own_occ);
        RT_LAUNCH_CHECK(ctx, "occupancyFromTable");
        st.grid.occupancy = st.own_occ;
    }
    grid = &st.grid;
    if (grid->n_slabs > 1) {   // multi-cell ("heavy") set: coarse occupancy for the queue walkers
        rt_ctx* ctx = s->ctx;
        unsigned sh = 0;
        while (((grid->n_slabs + (1u << sh) - 1) >> sh) > 64) sh++;
        st.macro_shift = sh;
        st.macro_n = (grid->n_slabs + (1u << sh) - 1) >> sh;
        size_t mcells = (size_t)st.macro_n * st.macro_n * st.macro_n;
        RT_TRY(rt_buffer_create(ctx, sizeof(unsigned) * 8192, (void**)&st.macro_occ));
        RT_TRY(rt_buffer_fill(ctx, st.macro_occ, 0, sizeof(unsigned) * 8192));
        f_macroOccupancy<<<rt_blocks((mcells + 31) / 32 * 32, kBlock), kBlock, 0, ctx->stream>>>((const unsigned*)grid->occupancy, grid->n_slabs, sh,
                                                                                                 st.macro_n, st.macro_occ);
        RT_LAUNCH_CHECK(ctx, "macroOccupancy");
        // Distance field for the empty-walk proof (rt_wavefront.cu): chessboard (L-infinity) distance of every cell of a
        // <= 128^3 grid (the fine cells themselves up to n = 128, 2 x 2 x 2 blocks up to 256, ...) to the nearest occupied one, on
        // the host: two raster passes with the 13 already-visited neighbours each are exact for this metric.
        unsigned dsh = 0;
        while (((grid->n_slabs + (1u << dsh) - 1) >> dsh) > 128) dsh++;
        const int nd = (int)((grid->n_slabs + (1u << dsh) - 1) >> dsh);
        const size_t dcells = (size_t)nd * nd * nd, fine = (size_t)grid->n_slabs * grid->n_slabs * grid->n_slabs;
        std::vector<unsigned> occ((fine + 31) / 32);
        RT_TRY(rt_buffer_read(ctx, grid->occupancy, 0, sizeof(unsigned) * occ.size(), occ.data()));
        std::vector<unsigned char> dist(dcells, 255);
        {
            const unsigned nn = grid->n_slabs;
            for (size_t w = 0; w < occ.size(); w++) {
                unsigned bits = occ[w];
                while (bits) {
                    const size_t c = w * 32 + (size_t)(__builtin_ctz(bits));
                    bits &= bits - 1;
                    if (c >= fine) break;
                    const unsigned x = (unsigned)(c % nn), y = (unsigned)((c / nn) % nn), z = (unsigned)(c / ((size_t)nn * nn));
                    dist[((size_t)(z >> dsh) * nd + (y >> dsh)) * nd + (x >> dsh)] = 0;
                }
            }
        }
        auto at = [&](int z, int y, int x) -> unsigned char& { return dist[((size_t)z * nd + y) * nd + x]; };
        for (int pass = 0; pass < 2; pass++) {
            const int lo = pass ? nd - 1 : 0, hi = pass ? -1 : nd, dir = pass ? -1 : 1;
            for (int z = lo; z != hi; z += dir)
                for (int y = lo; y != hi; y += dir)
                    for (int x = lo; x != hi; x += dir) {
                        unsigned best = at(z, y, x);
                        if (best == 0) continue;
                        for (int dz = -1; dz <= 0; dz++)
                            for (int dy = -1; dy <= (dz < 0 ? 1 : 0); dy++)
                                for (int dx = -1; dx <= ((dz < 0 || dy < 0) ? 1 : -1); dx++) {
                                    const int zz = z + dir * dz, yy = y + dir * dy, xx = x + dir * dx;
                                    if (zz < 0 || yy < 0 || xx < 0 || zz >= nd || yy >= nd || xx >= nd) continue;
                                    const unsigned v = at(zz, yy, xx) + 1u;
                                    if (v < best) best = v;
                                }
                        at(z, y, x) = (unsigned char)(best > 255 ? 255 : best);
                    }
        }
        RT_TRY(rt_buffer_create(ctx, dcells, (void**)&st.macro_dist));
        RT_TRY(rt_buffer_write(ctx, st.macro_dist, 0, dcells, dist.data()));
        st.dist_shift = dsh;
        st.dist_n = (unsigned)nd;
        const float fcell = (float)(1u << dsh), nf = (float)grid->n_slabs;
        for (int a = 0; a < 3; a++) st.dist_inv[a] = nf / ((bound[4 + a] - bound[a]) * fcell);
    }
    if (grid->kind == 1 && grid->n_slabs == 1 && grid->n_refs > 0 && grid->n_refs <= 4096) {
        // "walls": every triangle planar in one coordinate (three equal fp32 values) and not a sliver -> the room between the planes
        rt_ctx* ctx = s->ctx;
        std::vector<float> tp((size_t)grid->n_refs * 12);
        RT_TRY(rt_buffer_read(ctx, grid->prim, 0, sizeof(float) * tp.size(), tp.data()));
        const float inf = __builtin_inff();
        float lo[3] = {-inf, -inf, -inf}, hi[3] = {inf, inf, inf}, scale = 0.f;
        const float mid[3] = {0.5f * (bound[0] + bound[4]), 0.5f * (bound[1] + bound[5]), 0.5f * (bound[2] + bound[6])};
        bool ok = true;
        for (unsigned r = 0; r < grid->n_refs && ok; r++) {
            const float* p0 = &tp[(size_t)r * 12];
            const float* p1 = p0 + 4;
            const float* p2 = p0 + 8;
            int axis = -1;
            for (int a = 0; a < 3; a++) if (p0[a] == p1[a] && p1[a] == p2[a]) axis = a;
            if (axis < 0) { ok = false; break; }
            double e1[3], e2[3], n[3];
            for (int a = 0; a < 3; a++) { e1[a] = (double)p1[a] - p0[a]; e2[a] = (double)p2[a] - p0[a]; }
            n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
            const double l1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]), l2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
            const double ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (!(ln >= 0.25 * l1 * l2) || !(ln > 0)) { ok = false; break; }   // sin of the angle at p0 >= 0.25, finite
            const float c = p0[axis];
            if (c <= mid[axis]) lo[axis] = lo[axis] > c ? lo[axis] : c; else hi[axis] = hi[axis] < c ? hi[axis] : c;
            for (int k = 0; k < 3; k++) for (int a = 0; a < 3; a++) { const float m = fabsf(p0[4 * k + a]); if (m > scale) scale = m; }
        }
        for (int a = 0; a < 8 && ok; a++) if (a != 3 && a != 7) { const float m = fabsf(bound[a]); if (m > scale) scale = m; }
        if (ok && scale > 0.f && scale < 1e30f) {
            st.wall_ok = 1;
            for (int a = 0; a < 3; a++) { st.wall_lo[a] = lo[a]; st.wall_hi[a] = hi[a]; }
            st.wall_scale = scale;
        }
    }
    if (grid->kind == 1 && grid->n_refs > 0) {   // every triangle set: precomputed face vector / edges (queue walkers and 1-cell sets)
        rt_ctx* ctx = s->ctx;
        RT_TRY(rt_buffer_create(ctx, sizeof(float4) * grid->n_refs, (void**)&st.pre_ng));
        RT_TRY(rt_buffer_create(ctx, sizeof(float4) * 3 * (size_t)grid->n_refs, (void**)&st.pre_pe));
        f_precomputeTriangles<<<rt_blocks(grid->n_refs, kBlock), kBlock, 0, ctx->stream>>>((const float4*)grid->prim, grid->n_refs, st.pre_ng, st.pre_pe);
        RT_LAUNCH_CHECK(ctx, "precomputeTriangles");
    }
    s->sets.push_back(st);
    return RT_OK;
}

int rt_scene_add_light(rt_scene* s, const float shadow_info[16], const float scene_info[16], const float light_info[16]) {
    if (!s || !shadow_info || !scene_info || !light_info) return RT_ERR_INVALID;
    SceneLight L;
    memcpy(L.shadow, shadow_info, sizeof L.shadow);
    memcpy(L.scene, scene_info, sizeof L.scene);
    memcpy(L.light, light_info, sizeof L.light);
    s->lights.push_back(L);
    return RT_OK;
}

int rt_render_create(rt_ctx* ctx, rt_scene* scene, const rt_render_opts* opts, rt_render** out) {
    RT_CHECK_CTX(ctx);
    if (!scene || !opts || !out || !opts->cols || !opts->rows || !opts->rays_per_pixel) return RT_ERR_INVALID;
    if (!scene->materials) return rt_fail(ctx, RT_ERR_STATE, "render: scene has no materials");
    rt_render* r = new rt_render();
    r->ctx = ctx;
    r->scene = scene;
    r->o = *opts;
    if (r->o.slot_count == 0) {   // 0,0 = every slot; an EMPTY range at slot_begin > 0 must not silently become "all slots"
        if (r->o.slot_begin != 0) { delete r; return rt_fail(ctx, RT_ERR_INVALID, "render: empty slot range (this rank has no slots: skip the render, contribute a zero image)"); }
        r->o.slot_count = r->o.rays_per_pixel;
    }
    if (r->o.mode > 2) { delete r; return rt_fail(ctx, RT_ERR_INVALID, "render: mode must be 0 (wavefront), 1 (reference schedule) or 2 (megakernel)"); }
    if (r->o.slot_begin + r->o.slot_count > r->o.rays_per_pixel) { delete r; return rt_fail(ctx, RT_ERR_INVALID, "render: slot range exceeds rays_per_pixel"); }
    if (r->o.rays_per_pixel > 1) {
        unsigned side = (unsigned)sqrtf((float)r->o.rays_per_pixel);
        if (side * side != r->o.rays_per_pixel) { delete r; return rt_fail(ctx, RT_ERR_INVALID, "render: rays_per_pixel must be a perfect square"); }
    }
    r->slots_pp = r->o.slot_count;
    r->pixels = (size_t)r->o.cols * r->o.rows;
    r->local_slots = r->pixels * r->slots_pp;
    if ((unsigned long long)r->pixels * r->o.rays_per_pixel > 0xFFFFFFFFull) { delete r; return rt_fail(ctx, RT_ERR_INVALID, "render: total_rays exceeds the reference's uint range"); }
    // default tile: as many slots as a quarter of the device memory holds (wavefront state per slot: ray 32 + hit 32 +
    // throughput 16 bytes, and per light a 32-byte shadow ray and two 4-byte queue entries; 180 GB of HBM3e on B200 -> ~300 Mi
    // slots with two lights): the persistent queue walkers need a DEEP queue -- with 4 Mi-slot tiles a walk launch got
    // ~0.5 M rays for 151 k lanes and spent a quarter of its time in the drain tail
    const size_t nlq = scene->lights.size() ? scene->lights.size() : 1;
    const size_t slot_bytes = 80 + 44 * nlq;   // + a second (filtered) queue entry and the filter's codes (<= 4 B) per light
    size_t want = r->o.tile_slots ? r->o.tile_slots : (size_t)(ctx->prop.totalGlobalMem / 4 / slot_bytes);
    if (want < ((size_t)1 << 22)) want = (size_t)1 << 22;
    if (want * nlq > 0xFFFFFFFFull) want = 0xFFFFFFFFull / nlq;   // queue entries are light * tile_slots + slot in 32 bits
    size_t px_per_tile = want / r->slots_pp;
    if (px_per_tile == 0) px_per_tile = 1;
    if (px_per_tile > r->pixels) px_per_tile = r->pixels;
    r->tile_slots = px_per_tile * r->slots_pp;
    int rc = RT_OK;
    auto A = [&](void** p, size_t bytes) { if (!rc) rc = rt_buffer_create(ctx, bytes, p); };
    A((void**)&r->seeds, sizeof(int) * r->local_slots);
    A((void**)&r->acu, sizeof(float4) * r->local_slots);
    A((void**)&r->accum, sizeof(float4) * r->pixels);
    A((void**)&r->pixel, sizeof(uchar4) * r->pixels);
    // kept for the life of the render: a per-pass cudaMallocAsync of this buffer cost ~27 ms a pass at 1080p (the
    // stream-ordered pool hands its memory back at every synchronisation) -- three times the pass's kernels
    if (r->o.rays_per_pixel == 1) A((void**)&r->rpp1_coords, sizeof(float2) * r->pixels);
    if (r->o.mode == 1) {   // the kernel-by-kernel schedule keeps the reference's AoS state per tile
        A((void**)&r->rays, sizeof(Ray) * r->tile_slots);
        A((void**)&r->pois, sizeof(Poi10) * r->tile_slots);
        A((void**)&r->shadow, sizeof(Ray) * r->tile_slots);
    }
    A((void**)&r->d_counters, sizeof(unsigned long long) * 2);
    A((void**)&r->d_profile, sizeof(unsigned long long) * 16 * RT_MAX_SETS);
    if (!rc) rc = rt_buffer_fill(ctx, r->d_profile, 0, sizeof(unsigned long long) * 16 * RT_MAX_SETS);
    if (!rc && (cudaEventCreate(&r->ev0) != cudaSuccess || cudaEventCreate(&r->ev1) != cudaSuccess)) rc = RT_ERR_CUDA;
    // prepareInitAcu (A10/code.js:1078-1099): zero once, never again between passes
    if (!rc) rc = rt_buffer_fill(ctx, r->acu, 0, sizeof(float4) * r->local_slots);
    if (!rc && r->pois) rc = rt_buffer_fill(ctx, r->pois, 0, sizeof(Poi10) * r->tile_slots);
    if (!rc && r->rays) rc = rt_buffer_fill(ctx, r->rays, 0, sizeof(Ray) * r->tile_slots);
    if (rc) { rt_render_destroy(r); return rc; }
    *out = r;
    return RT_OK;
}

int rt_render_destroy(rt_render* r) {
    if (!r) return RT_ERR_INVALID;
    rt_ctx* ctx = r->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    void* bufs[] = {r->seeds, r->acu, r->accum, r->pixel, r->rpp1_coords, r->rays, r->pois, r->shadow, r->d_counters, r->d_profile,
                    r->w_ray, r->w_poi, r->w_atte, r->w_sh, r->w_queue, r->w_qctr, r->w_queue_f, r->w_masks};
    for (void* b : bufs) if (b) cudaFree(b);
    if (r->copy_stream) { cudaStreamSynchronize(r->copy_stream); cudaStreamDestroy(r->copy_stream); }
    if (r->ev_seeds) cudaEventDestroy(r->ev_seeds);
    if (r->ev_main) cudaEventDestroy(r->ev_main);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    for (cudaEvent_t e : r->tev) cudaEventDestroy(e);
    delete r;
    return RT_OK;
}

namespace {
// global seed array [pixel][k] -> this context's [pixel][k_local]
__global__ void f_sliceSeeds(const int* global_seeds, int* local, size_t pixels, unsigned rpp, unsigned slot_begin, unsigned slots_pp) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pixels * slots_pp) return;
    size_t pix = id / slots_pp;
    unsigned k = slot_begin + (unsigned)(id % slots_pp);
    local[id] = global_seeds[pix * rpp + k];
}
}  // namespace

int rt_render_set_seeds(rt_render* r, const int* seeds, size_t count, int on_device) {
    if (!r || !seeds) return RT_ERR_INVALID;
    rt_ctx* ctx = r->ctx;
    const rt_render_opts& o = r->o;
    size_t total = r->pixels * o.rays_per_pixel;
    if (count != total) return rt_fail(ctx, RT_ERR_INVALID, "set_seeds: count must be cols*rows*rays_per_pixel");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_TRY(rt_seeds_ready(r));
    if (r->slots_pp == o.rays_per_pixel) {
        RT_CUDA(ctx, cudaMemcpyAsync(r->seeds, seeds, sizeof(int) * total, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    } else if (on_device) {
        f_sliceSeeds<<<rt_blocks(r->local_slots, kBlock), kBlock, 0, ctx->stream>>>(seeds, r->seeds, r->pixels, o.rays_per_pixel, o.slot_begin, r->slots_pp);
        RT_LAUNCH_CHECK(ctx, "sliceSeeds");
    } else {
        // strided host -> device copy: one 2-D copy, row = pixel, width = this context's slot range
        RT_CUDA(ctx, cudaMemcpy2DAsync(r->seeds, sizeof(int) * r->slots_pp, seeds + o.slot_begin, sizeof(int) * o.rays_per_pixel,
                                       sizeof(int) * r->slots_pp, r->pixels, cudaMemcpyHostToDevice, ctx->stream));
    }
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    r->have_seeds = true;
    return RT_OK;
}

int rt_render_write_local_seeds(rt_render* r, const int* host_seeds, size_t count) {
    if (!r || !host_seeds) return RT_ERR_INVALID;
    rt_ctx* ctx = r->ctx;
    if (count != r->local_slots) return rt_fail(ctx, RT_ERR_INVALID, "write_local_seeds: count must be cols*rows*slot_count");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_TRY(rt_seeds_ready(r));
    RT_CUDA(ctx, cudaMemcpyAsync(r->seeds, host_seeds, sizeof(int) * count, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    r->have_seeds = true;
    return RT_OK;
}

// Non-blocking form (the reference's own uploads are enqueueWriteBuffer(buf, false, ...), A10/code.js:1149): the copy
// is issued on a side stream behind everything already queued on the context's stream and returns at once; the
// next pass starts immediately and waits for it only in front of its first kernel that reads seeds -- with
// rays_per_pixel > 1 that is the first shadow-ray stage, so the upload overlaps ray generation and the primary traversal.
int rt_render_write_local_seeds_async(rt_render* r, const int* host_seeds, size_t count) {
    if (!r || !host_seeds) return RT_ERR_INVALID;
    rt_ctx* ctx = r->ctx;
    if (count != r->local_slots) return rt_fail(ctx, RT_ERR_INVALID, "write_local_seeds_async: count must be cols*rows*slot_count");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!r->copy_stream) {
        RT_CUDA(ctx, cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
        RT_CUDA(ctx, cudaEventCreateWithFlags(&r->ev_seeds, cudaEventDisableTiming));
        RT_CUDA(ctx, cudaEventCreateWithFlags(&r->ev_main, cudaEventDisableTiming));
    }
    RT_CUDA(ctx, cudaEventRecord(r->ev_main, ctx->stream));             // after the passes that still read / write the old seeds
    RT_CUDA(ctx, cudaStreamWaitEvent(r->copy_stream, r->ev_main, 0));
    RT_CUDA(ctx, cudaMemcpyAsync(r->seeds, host_seeds, sizeof(int) * count, cudaMemcpyHostToDevice, r->copy_stream));
    RT_CUDA(ctx, cudaEventRecord(r->ev_seeds, r->copy_stream));
    r->seeds_in_flight = true;
    r->have_seeds = true;
    return RT_OK;
}

int rt_render_set_profile(rt_render* r, int on) {
    if (!r) return RT_ERR_INVALID;
    r->profile = on != 0;
    return rt_buffer_fill(r->ctx, r->d_profile, 0, sizeof(unsigned long long) * 16 * RT_MAX_SETS);
}

int rt_render_read_profile_sets(rt_render* r, unsigned long long out[RT_MAX_SETS * 16]) {
    if (!r || !out) return RT_ERR_INVALID;
    return rt_buffer_read(r->ctx, r->d_profile, 0, sizeof(unsigned long long) * 16 * RT_MAX_SETS, out);
}

int rt_render_read_profile(rt_render* r, unsigned long long out[16]) {
    if (!r || !out) return RT_ERR_INVALID;
    unsigned long long all[RT_MAX_SETS * 16];
    RT_TRY(rt_render_read_profile_sets(r, all));
    for (int q = 0; q < 16; q++) {
        out[q] = 0;
        for (int s = 0; s < RT_MAX_SETS; s++) out[q] += all[s * 16 + q];
    }
    return RT_OK;
}

int rt_render_set_timing(rt_render* r, int on) {
    if (!r) return RT_ERR_INVALID;
    r->timing = on != 0;
    for (int c = 0; c < RT_TIMING_CLASSES; c++) { r->class_ms[c] = 0.f; r->class_launches[c] = 0; }
    return RT_OK;
}

int rt_render_read_timing(rt_render* r, float ms[RT_TIMING_CLASSES], unsigned launches[RT_TIMING_CLASSES]) {
    if (!r) return RT_ERR_INVALID;
    for (int c = 0; c < RT_TIMING_CLASSES; c++) {
        if (ms) ms[c] = r->class_ms[c];
        if (launches) launches[c] = r->class_launches[c];
    }
    return RT_OK;
}

int rt_render_execute(rt_render* r, const float fcam[16], unsigned char* host_pixels) {
    if (!r || !fcam) return RT_ERR_INVALID;
    rt_ctx* ctx = r->ctx;
    const rt_render_opts& o = r->o;
    if (!r->have_seeds) return rt_fail(ctx, RT_ERR_STATE, "render_execute: seeds not set (rt_render_set_seeds)");
    if ((unsigned)fcam[14] != o.cols || (unsigned)fcam[15] != o.rows) return rt_fail(ctx, RT_ERR_INVALID, "render_execute: camera cols/rows differ from the render's");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long launches0 = ctx->launches;
    RT_CUDA(ctx, cudaMemsetAsync(r->d_counters, 0, sizeof(unsigned long long) * 2, ctx->stream));
    RT_CUDA(ctx, cudaEventRecord(r->ev0, ctx->stream));
    r->tev_used = 0;
    float2* coords = nullptr;
    // the wavefront path of a stratified pass (rays_per_pixel > 1) first needs seeds at its first shadow stage and waits
    // there (rt_wavefront.cu); every other path reads them from its first kernel on
    if (o.rays_per_pixel == 1 || o.mode != 0 || r->profile) RT_TRY(rt_seeds_ready(r));
    if (o.rays_per_pixel == 1) {
        coords = r->rpp1_coords;
        f_rpp1_coords<<<rt_blocks(o.cols, 64), 64, 0, ctx->stream>>>(r->seeds, coords, o.cols, o.rows);
        RT_LAUNCH_CHECK(ctx, "initTrace(seeds)");
    }
    for (size_t slot0 = 0; slot0 < r->local_slots; slot0 += r->tile_slots) {
        size_t rem = r->local_slots - slot0;
        unsigned n = (unsigned)(rem < r->tile_slots ? rem : r->tile_slots);
        int rc;
        if (o.mode == 1) { rc = rt_time_mark(r, 6); if (!rc) rc = tileReferenceSchedule(r, fcam, slot0, n, coords); }
        else rc = rt_fused_tile(r, fcam, slot0, n, coords);
        if (rc) return rc;
    }
    RT_TRY(rt_time_mark(r, 7));
    f_sumSlots<<<rt_blocks(r->pixels, kBlock), kBlock, 0, ctx->stream>>>(r->acu, r->accum, r->pixels, r->slots_pp);
    RT_LAUNCH_CHECK(ctx, "sumSlots");
    RT_TRY(rt_time_mark(r, -1));
    RT_CUDA(ctx, cudaEventRecord(r->ev1, ctx->stream));
    if (host_pixels) {
        // executeCopyToPixel(passes): m = 1/(rays_per_pixel*passes) rounded to fp32 (A10/code.js:1410-1415)
        float m = (float)(1.0 / ((double)o.rays_per_pixel * (double)r->passes));
        f_accumToPixel<<<rt_blocks(r->pixels, kBlock), kBlock, 0, ctx->stream>>>(r->pixel, r->accum, m, (unsigned)r->pixels);
        RT_LAUNCH_CHECK(ctx, "copyToPixel");
        RT_CUDA(ctx, cudaMemcpyAsync(host_pixels, r->pixel, sizeof(uchar4) * r->pixels, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RT_CUDA(ctx, cudaMemcpyAsync(r->h_counters, r->d_counters, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RT_CUDA(ctx, cudaEventElapsedTime(&r->last_ms, r->ev0, r->ev1));
    for (size_t i = 0; i + 1 < r->tev_used; i++) {
        int c = r->tcls[i];
        if (c < 0 || c >= RT_TIMING_CLASSES) continue;
        float ms = 0.f;
        RT_CUDA(ctx, cudaEventElapsedTime(&ms, r->tev[i], r->tev[i + 1]));
        r->class_ms[c] += ms;
        r->class_launches[c]++;
    }
    r->last_launches = (unsigned)(ctx->launches - launches0);
    r->passes++;
    return RT_OK;
}

int rt_render_accum_image(rt_render* r, void** dptr_float4) {
    if (!r || !dptr_float4) return RT_ERR_INVALID;
    *dptr_float4 = r->accum;
    return RT_OK;
}

int rt_render_read_accum(rt_render* r, float* host_float4) {
    if (!r || !host_float4) return RT_ERR_INVALID;
    return rt_buffer_read(r->ctx, r->accum, 0, sizeof(float4) * r->pixels, host_float4);
}

int rt_render_read_seeds(rt_render* r, int* host_seeds, size_t count) {
    if (!r || !host_seeds) return RT_ERR_INVALID;
    if (count != r->local_slots) return rt_fail(r->ctx, RT_ERR_INVALID, "read_seeds: count must be cols*rows*slot_count");
    RT_TRY(rt_seeds_ready(r));
    return rt_buffer_read(r->ctx, r->seeds, 0, sizeof(int) * count, host_seeds);
}

int rt_render_export_state(rt_render* r, float* host_acu, int* host_seeds, unsigned* passes) {
    if (!r) return RT_ERR_INVALID;
    RT_TRY(rt_seeds_ready(r));
    if (host_acu) RT_TRY(rt_buffer_read(r->ctx, r->acu, 0, sizeof(float4) * r->local_slots, host_acu));
    if (host_seeds) RT_TRY(rt_buffer_read(r->ctx, r->seeds, 0, sizeof(int) * r->local_slots, host_seeds));
    if (passes) *passes = r->passes;
    return RT_OK;
}

int rt_render_import_state(rt_render* r, const float* host_acu, const int* host_seeds, unsigned passes) {
    if (!r || passes == 0) return RT_ERR_INVALID;
    RT_TRY(rt_seeds_ready(r));
    if (host_acu) RT_TRY(rt_buffer_write(r->ctx, r->acu, 0, sizeof(float4) * r->local_slots, host_acu));
    if (host_seeds) {
        RT_TRY(rt_buffer_write(r->ctx, r->seeds, 0, sizeof(int) * r->local_slots, host_seeds));
        r->have_seeds = true;
    }
    r->passes = passes;
    return RT_OK;
}

int rt_accum_to_pixel(rt_ctx* ctx, void* pixel, const void* accum_float4, float m, unsigned pixels) {
    RT_CHECK_CTX(ctx);
    if (!pixel || !accum_float4) return RT_ERR_INVALID;
    if (!pixels) return RT_OK;
    f_accumToPixel<<<rt_blocks(pixels, kBlock), kBlock, 0, ctx->stream>>>((uchar4*)pixel, (const float4*)accum_float4, m, pixels);
    RT_LAUNCH_CHECK(ctx, "accumToPixel");
    return RT_OK;
}

int rt_render_stats(rt_render* r, unsigned long long* closest_rays, unsigned long long* any_rays, unsigned* launches, float* device_ms) {
    if (!r) return RT_ERR_INVALID;
    if (closest_rays) *closest_rays = r->h_counters[0];
    if (any_rays) *any_rays = r->h_counters[1];
    if (launches) *launches = r->last_launches;
    if (device_ms) *device_ms = r->last_ms;
    return RT_OK;
}

}  // extern "C"
