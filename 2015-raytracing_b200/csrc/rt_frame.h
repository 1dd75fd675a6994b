// rt_frame.h -- scene / render objects shared by rt_frame.cu (driver, reference schedule) and
// rt_wavefront.cu (fused paths).
#pragma once
#include "rt_device.cuh"
#include "rt_internal.h"

struct SceneSet {
    rt_grid grid;
    float bound[8];
    int is_mesh;
    unsigned mesh_matid;
    // internal, triangle sets walked by the queue walkers: per reference the face vector
    // ng = cross(p2-p0, p1-p0) and (p0, e1 = p1-p0, e2 = p2-p0), computed ONCE with the same fp32
    // operations interTriangle performs per test (bit-identical values, A10/code.cl:252-256)
    float4* pre_ng = nullptr;    // [n_refs]
    float4* pre_pe = nullptr;    // [3*n_refs]
    // internal: coarse occupancy, 1 bit per (2^macro_shift)^3 block of cells, at most 64^3 bits = 32 KB,
    // small enough to sit in shared memory of the queue walkers (the fine bitmap is n^3 bits)
    unsigned* macro_occ = nullptr;
    unsigned macro_shift = 0, macro_n = 0;
    // internal: chessboard distance (in field cells, capped at 255) from every cell of a <= 128^3 field to the nearest occupied one --
    // lets the stage kernels PROVE that a ray's whole walk crosses empty cells only and skip the walk (rt_wavefront.cu)
    unsigned char* macro_dist = nullptr;
    float dist_inv[3] = {0.f, 0.f, 0.f};   // world -> distance-field index scale per axis
    unsigned dist_shift = 0, dist_n = 0;   // the field has dist_n^3 cells of (2^dist_shift)^3 grid cells each (dist_n <= 128)
    // internal, 1-cell triangle sets whose triangles are all axis-aligned and planar (the walls of a room): the open box between
    // those planes around the set's centre -- shadow segments inside it are not tested against the set (rt_wavefront.cu)
    int wall_ok = 0;
    float wall_lo[3] = {0.f, 0.f, 0.f}, wall_hi[3] = {0.f, 0.f, 0.f}, wall_scale = 0.f;
    unsigned* own_occ = nullptr;   // occupancy bits derived here for a caller-built grid that came without (rt_scene_add_set)
};
struct SceneLight { float shadow[16], scene[16], light[16]; };

struct rt_scene {
    rt_ctx* ctx = nullptr;
    float bound[8] = {0};
    std::vector<SceneSet> sets;
    std::vector<SceneLight> lights;
    void* materials = nullptr;
    unsigned n_materials = 0;
};

struct rt_render {
    rt_ctx* ctx = nullptr;
    rt_scene* scene = nullptr;
    rt_render_opts o{};
    unsigned slots_pp = 0;           // slots per pixel handled by this context
    size_t pixels = 0, local_slots = 0;
    size_t tile_slots = 0;           // multiple of slots_pp
    // persistent
    int* seeds = nullptr;            // [pixel][k_local]
    float4* acu = nullptr;           // [pixel][k_local]
    float4* accum = nullptr;         // [pixel] sum over k_local
    uchar4* pixel = nullptr;
    float2* rpp1_coords = nullptr;   // rays_per_pixel == 1 only: the lens coordinates initTrace draws, [pixel] (quirk Q7)
    bool have_seeds = false;
    // non-blocking seed upload (rt_render_write_local_seeds_async): side stream + the event the pass waits on
    // right before its first seed-consuming kernel
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_seeds = nullptr, ev_main = nullptr;
    bool seeds_in_flight = false;
    unsigned passes = 1;             // the reference starts at 1 and divides by it (A10/code.js:416,1850)
    // per tile
    rt::Ray* rays = nullptr;
    rt::Poi10* pois = nullptr;
    rt::Ray* shadow = nullptr;
    // wavefront path (rt_wavefront.cu): per-tile SoA state, task queues, queue counters
    float4* w_ray = nullptr;         // [2*n]: (o.xyz, mint) (d.xyz, maxt)
    float4* w_poi = nullptr;         // [2*n]: (p.xyz, matId bits) (normal.xyz, 0)
    float4* w_atte = nullptr;        // [n]
    float4* w_sh = nullptr;          // [2*n]: shadow ray (o.xyz, mint) (d.xyz, maxt)
    unsigned* w_queue = nullptr;     // [n] slot ids of the current heavy-set walk
    unsigned* w_qctr = nullptr;      // [4*kMaxStages]: per walk stage {count, head}, as pushed and as filtered
    unsigned* w_queue_f = nullptr;   // [n * lights] the queue after the filter (empty-walk proof)
    unsigned* w_masks = nullptr;     // [n * lights] the filter's decisions (direction-class code per queue entry)
    // stats
    unsigned long long* d_counters = nullptr;   // [0] closest rays, [1] any rays
    unsigned long long* d_profile = nullptr;    // [RT_MAX_SETS][16] work counters per geometry set (rt_render_read_profile*)
    bool profile = false;
    unsigned long long h_counters[2] = {0, 0};
    unsigned last_launches = 0;
    float last_ms = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // optional per-kernel-class timing (rt_render_set_timing): one event is recorded in front of
    // every launch; the interval up to the next event is charged to that launch's class
    bool timing = false;
    std::vector<cudaEvent_t> tev;
    std::vector<int> tcls;
    size_t tev_used = 0;
    float class_ms[RT_TIMING_CLASSES] = {0};
    unsigned class_launches[RT_TIMING_CLASSES] = {0};
};

// Orders the context's stream after a seed upload that is still in flight (no-op otherwise).  Called in front of
// the first kernel of a pass that reads the seed buffer, and by everything else that touches it.
int rt_seeds_ready(rt_render* r);

// Marks the start of a launch of timing class `cls` (no-op unless timing is on).
int rt_time_mark(rt_render* r, int cls);


// Fused paths (rt_wavefront.cu).  `mode` as in rt_render_opts.
int rt_fused_tile(rt_render* r, const float* fcam, size_t slot0, unsigned n, const float2* rpp1_coords);

// Assignment-7 traces of a grid the library built, through the persistent pair-list queue walker (rt_wavefront.cu): same hits, same
// floats as the one-thread-per-ray kernel, ~3x faster on big grids.  prim = PRIM_SPHERE / PRIM_TRIANGLE.
int rt_walk_a07(rt_ctx* ctx, rt_ctx::GridAux& g, int prim, void* pixels, const float* fcam, void* rays, const void* normals, const float* bound,
                size_t npix);
