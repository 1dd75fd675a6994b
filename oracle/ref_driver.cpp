// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product.
//
// NDRange driver for the parity oracle: includes the reference's kernel text (one
// namespace per assignment; text generated into oracle/_ref/ by cl2cpp.py from the
// read-only /root/reference tree, never committed) and runs every work-item of a launch
// as one loop iteration on the host cores (OpenMP).  Exports one extern "C" function per
// reference kernel: ref_<assignment>_<kernel>(...), pointer arguments in the kernel's
// own order, by-value structs (AABB = 8 floats, float16 = 16 floats, float3 = 4 floats)
// passed as const float*.
//
// The single sanctioned deviation from "run the text as is" is quirk Q1:
// A10's closest-hit kernels store a hit record whose `atte` member was never assigned.
// The contract is "a hit preserves pois[id].atte"; the driver enforces it by saving
// atte before the work-item and restoring it afterwards.
#include "clshim.h"
#include <omp.h>

thread_local size_t cl_gid[3];

// ---- optional instrumentation (REF_INSTRUMENT build only; hook macros are inserted by
// cl2cpp.py --hooks at non-arithmetic places) -------------------------------------------
thread_local unsigned long long ref_tl_cells, ref_tl_tests;
thread_local unsigned int ref_tl_hit;
#ifdef REF_INSTRUMENT
#define REF_HOOK_CELL ref_tl_cells++;
#define REF_HOOK_TEST ref_tl_tests++;
#define REF_HOOK_HIT(i) ref_tl_hit = (unsigned int)(i);
#else
#define REF_HOOK_CELL
#define REF_HOOK_TEST
#define REF_HOOK_HIT(i)
#endif

// The kernel text: _ref/aNN_code.inc, or -- instrumented build -- _ref/instr/_ref/aNN_code.inc with the hook macros.
// (A quoted #include "_ref/..." would find the un-hooked file next to this source before any -I directory: round 1's
// "instrumented" library was built that way and never counted anything.)
#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#ifdef REF_INSTRUMENT
#define REF_CODE(n) REF_STR(_ref/instr/_ref/n##_code.inc)
#else
#define REF_CODE(n) REF_STR(_ref/n##_code.inc)
#endif

namespace a01 {
#include REF_CODE(a01)
}
namespace a02 {
#include REF_CODE(a02)
}
namespace a03 {
#include REF_CODE(a03)
}
namespace a04 {
#include REF_CODE(a04)
}
namespace a05 {
#include REF_CODE(a05)
}
namespace a06 {
#include REF_CODE(a06)
}
namespace a07 {
#include REF_CODE(a07)
}
namespace a08 {
#include REF_CODE(a08)
}
namespace a09 {
#include REF_CODE(a09)
}
namespace a10 {
#include REF_CODE(a10)
}

static_assert(sizeof(a10::Ray) == 48 && sizeof(a10::Poi) == 64, "A10 layouts (SURVEY 8)");
static_assert(sizeof(a08::Ray) == 48 && sizeof(a08::Poi) == 48, "A08 layouts (SURVEY 8)");
static_assert(sizeof(a09::Ray) == 48 && sizeof(a09::Poi) == 48, "A09 layouts (SURVEY 8)");
static_assert(sizeof(a07::Ray) == 48 && sizeof(a10::AABB) == 32, "A07 layouts (SURVEY 8)");
static_assert(sizeof(a04::Ray) == 48 && sizeof(a05::Ray) == 48 && sizeof(a06::Ray) == 48 && sizeof(a06::AABB) == 32, "A04-A06 layouts");

template <typename AABB_T>
static inline AABB_T mk_aabb(const float* b) {
    AABB_T r;
    r.pmin = float3(b[0], b[1], b[2]);
    r.pmax = float3(b[4], b[5], b[6]);
    return r;
}
static inline float16 mk_f16(const float* c) {
    float16 r;
    memcpy(r.v, c, sizeof r.v);
    return r;
}
static inline float3 mk_f3(const float* c) { return float3(c[0], c[1], c[2]); }

// Per-ray statistics sink for instrumented builds (all optional / may be NULL).
struct RefStats {
    unsigned int* hit_id;             // per work-item champ_i, 0xFFFFFFFF if none
    unsigned long long* cells;        // per work-item cells visited
    unsigned long long* tests;        // per work-item primitive tests
};
static RefStats g_stats = {nullptr, nullptr, nullptr};
extern "C" void ref_set_stats(unsigned int* hit_id, unsigned long long* cells, unsigned long long* tests) {
    g_stats.hit_id = hit_id;
    g_stats.cells = cells;
    g_stats.tests = tests;
}
extern "C" int ref_is_instrumented() {
#ifdef REF_INSTRUMENT
    return 1;
#else
    return 0;
#endif
}
extern "C" int ref_num_threads() { return omp_get_max_threads(); }
// launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline wants all host cores
extern "C" void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

#define STAT_BEGIN() \
    ref_tl_cells = 0; ref_tl_tests = 0; ref_tl_hit = 0xFFFFFFFFu;
#define STAT_END(i)                                      \
    if (g_stats.hit_id) g_stats.hit_id[i] = ref_tl_hit;  \
    if (g_stats.cells) g_stats.cells[i] = ref_tl_cells;  \
    if (g_stats.tests) g_stats.tests[i] = ref_tl_tests;

// 1-D launch over n work-items.
#define ND1(n, body)                                              \
    _Pragma("omp parallel for schedule(dynamic, 2048)")           \
    for (long long _i = 0; _i < (long long)(n); _i++) {           \
        cl_gid[0] = (size_t)_i; cl_gid[1] = 0; cl_gid[2] = 0;     \
        STAT_BEGIN();                                             \
        body;                                                     \
        STAT_END(_i);                                             \
    }
// 2-D launch: gid0 = col, gid1 = row (reference: A10/code.cl:469-470).
#define ND2(cols, rows, body)                                     \
    _Pragma("omp parallel for schedule(dynamic, 4)")              \
    for (long long _r = 0; _r < (long long)(rows); _r++)          \
        for (long long _c = 0; _c < (long long)(cols); _c++) {    \
            cl_gid[0] = (size_t)_c; cl_gid[1] = (size_t)_r; cl_gid[2] = 0; \
            long long _i = _r * (long long)(cols) + _c;           \
            STAT_BEGIN();                                         \
            body;                                                 \
            STAT_END(_i);                                         \
        }
// Serial 2-D launch in row-major order (used for the racy rpp==1 initTrace, Q7).
#define ND2_SERIAL(cols, rows, body)                              \
    for (long long _r = 0; _r < (long long)(rows); _r++)          \
        for (long long _c = 0; _c < (long long)(cols); _c++) {    \
            cl_gid[0] = (size_t)_c; cl_gid[1] = (size_t)_r; cl_gid[2] = 0; \
            body;                                                 \
        }

extern "C" {

// ======================================================================== A01
void ref_a01_raytrace(void* pixels, const float* cam, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a01::raytrace((uchar4*)pixels, c));
}

// ======================================================================== A02
void ref_a02_raytrace(void* pixels, const float* cam, uint s_size, void* s_atoms, void* s_colors, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a02::raytrace((uchar4*)pixels, c, s_size, (float4*)s_atoms, (float4*)s_colors));
}

// ======================================================================== A03
uint ref_a03_sizeofRay() { uint s; a03::sizeofRay(&s); return s; }
void ref_a03_initTrace(void* pixels, const float* cam, void* rays, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a03::initTrace((uchar4*)pixels, c, (a03::Ray*)rays));
}
void ref_a03_molTrace(void* pixels, const float* cam, void* rays, uint s_size, void* s_atoms, void* s_colors, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a03::molTrace((uchar4*)pixels, c, (a03::Ray*)rays, s_size, (float4*)s_atoms, (float4*)s_colors));
}

// ======================================================================== A04 (brute force, spheres + triangle mesh)
uint ref_a04_sizeofRay() { uint s; a04::sizeofRay(&s); return s; }
void ref_a04_initTrace(void* pixels, const float* cam, void* rays, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a04::initTrace((uchar4*)pixels, c, (a04::Ray*)rays));
}
void ref_a04_molTrace(void* pixels, const float* cam, void* rays, uint s_size, void* s_atoms, void* s_colors, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a04::molTrace((uchar4*)pixels, c, (a04::Ray*)rays, s_size, (float4*)s_atoms, (float4*)s_colors));
}
void ref_a04_meshTrace(void* pixels, const float* cam, void* rays, uint t_size, void* t_pos, void* t_normal, uint* t_mindex, void* m_color,
                       uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a04::meshTrace((uchar4*)pixels, c, (a04::Ray*)rays, t_size, (float3*)t_pos, (float3*)t_normal, t_mindex, (float4*)m_color));
}
void ref_a04_raytrace(void* pixels, const float* cam, uint s_size, void* s_atoms, void* s_colors, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    ND2(cols, rows, a04::raytrace((uchar4*)pixels, c, s_size, (float4*)s_atoms, (float4*)s_colors));
}

// ======================================================================== A05 (+ bounding boxes)
uint ref_a05_sizeofRay() { uint s; a05::sizeofRay(&s); return s; }
void ref_a05_initTrace(void* pixels, const float* cam, void* rays, const float* bound, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a05::AABB b = mk_aabb<a05::AABB>(bound);
    ND2(cols, rows, a05::initTrace((uchar4*)pixels, c, (a05::Ray*)rays, b));
}
void ref_a05_molTrace(void* pixels, const float* cam, void* rays, uint s_size, void* s_atoms, void* s_colors, const float* bound,
                      uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a05::AABB b = mk_aabb<a05::AABB>(bound);
    ND2(cols, rows, a05::molTrace((uchar4*)pixels, c, (a05::Ray*)rays, s_size, (float4*)s_atoms, (float4*)s_colors, b));
}
void ref_a05_meshTrace(void* pixels, const float* cam, void* rays, uint t_size, void* t_pos, void* t_normal, uint* t_mindex, void* m_color,
                       const float* bound, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a05::AABB b = mk_aabb<a05::AABB>(bound);
    ND2(cols, rows, a05::meshTrace((uchar4*)pixels, c, (a05::Ray*)rays, t_size, (float3*)t_pos, (float3*)t_normal, t_mindex, (float4*)m_color, b));
}

// ======================================================================== A06 (1-D slabs along x)
uint ref_a06_sizeofRay() { uint s; a06::sizeofRay(&s); return s; }
void ref_a06_initTrace(void* pixels, const float* cam, void* rays, const float* bound, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a06::AABB b = mk_aabb<a06::AABB>(bound);
    ND2(cols, rows, a06::initTrace((uchar4*)pixels, c, (a06::Ray*)rays, b));
}
void ref_a06_molTrace(void* pixels, const float* cam, void* rays, uint s_size, void* s_atoms, void* s_colors, const float* bound,
                      uint n_slabs, uint* slab_size, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a06::AABB b = mk_aabb<a06::AABB>(bound);
    ND2(cols, rows, a06::molTrace((uchar4*)pixels, c, (a06::Ray*)rays, s_size, (float4*)s_atoms, (float4*)s_colors, b, n_slabs, slab_size));
}
void ref_a06_meshTrace(void* pixels, const float* cam, void* rays, uint t_size, void* t_pos, void* t_normal, uint* t_mindex, void* m_color,
                       const float* bound, uint n_slabs, uint* slab_size, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a06::AABB b = mk_aabb<a06::AABB>(bound);
    ND2(cols, rows, a06::meshTrace((uchar4*)pixels, c, (a06::Ray*)rays, t_size, (float3*)t_pos, (float3*)t_normal, t_mindex, (float4*)m_color, b, n_slabs, slab_size));
}

// ======================================================================== A07
uint ref_a07_sizeofRay() { uint s; a07::sizeofRay(&s); return s; }
void ref_a07_initTrace(void* pixels, const float* cam, void* rays, const float* bound, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a07::AABB b = mk_aabb<a07::AABB>(bound);
    ND2(cols, rows, a07::initTrace((uchar4*)pixels, c, (a07::Ray*)rays, b));
}
void ref_a07_molTrace(void* pixels, const float* cam, void* rays, uint s_size, void* s_atoms, uint* s_mindex, void* m_color,
                      const float* bound, uint n_slabs, uint* slab_size, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a07::AABB b = mk_aabb<a07::AABB>(bound);
    ND2(cols, rows, a07::molTrace((uchar4*)pixels, c, (a07::Ray*)rays, s_size, (float4*)s_atoms, s_mindex, (float4*)m_color, b, n_slabs, slab_size));
}
void ref_a07_meshTrace(void* pixels, const float* cam, void* rays, uint t_size, void* t_pos, void* t_normal, uint* t_mindex,
                       void* m_color, const float* bound, uint n_slabs, uint* slab_size, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a07::AABB b = mk_aabb<a07::AABB>(bound);
    ND2(cols, rows, a07::meshTrace((uchar4*)pixels, c, (a07::Ray*)rays, t_size, (float3*)t_pos, (float3*)t_normal, t_mindex, (float4*)m_color, b, n_slabs, slab_size));
}

// ======================================================================== A08 (2-D NDRange, uint2 cols_rows)
uint ref_a08_sizeofRay() { uint s; a08::sizeofRay(&s); return s; }
uint ref_a08_sizeofPoi() { uint s; a08::sizeofPoi(&s); return s; }
void ref_a08_initTrace(void* acu, void* rays, void* pois, const float* bound, const float* cam, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a08::AABB b = mk_aabb<a08::AABB>(bound);
    ND2(cols, rows, a08::initTrace((float4*)acu, (a08::Ray*)rays, (a08::Poi*)pois, b, c));
}
void ref_a08_initShadowTrace(void* shadow, void* pois, uint cols, uint rows, const float* light_pos) {
    float3 lp = mk_f3(light_pos);
    uint2 cr(cols, rows);
    ND2(cols, rows, a08::initShadowTrace((a08::Ray*)shadow, (a08::Poi*)pois, cr, lp));
}
void ref_a08_sphereTrace(uint cols, uint rows, void* pois, void* rays, void* spheres, uint* s_matid, uint* s_box, const float* bound, uint n_slabs) {
    a08::AABB b = mk_aabb<a08::AABB>(bound);
    uint2 cr(cols, rows);
    ND2(cols, rows, a08::sphereTrace(cr, (a08::Poi*)pois, (a08::Ray*)rays, (float4*)spheres, s_matid, s_box, b, n_slabs));
}
void ref_a08_triangleTrace(uint cols, uint rows, void* pois, void* rays, void* t_pos, void* t_normal, uint* t_matid, uint* t_box, const float* bound, uint n_slabs) {
    a08::AABB b = mk_aabb<a08::AABB>(bound);
    uint2 cr(cols, rows);
    ND2(cols, rows, a08::triangleTrace(cr, (a08::Poi*)pois, (a08::Ray*)rays, (float3*)t_pos, (float3*)t_normal, t_matid, t_box, b, n_slabs));
}
void ref_a08_sphereShadowTrace(uint cols, uint rows, void* shadow, void* spheres, uint* s_box, const float* bound, uint n_slabs) {
    a08::AABB b = mk_aabb<a08::AABB>(bound);
    uint2 cr(cols, rows);
    ND2(cols, rows, a08::sphereShadowTrace(cr, (a08::Ray*)shadow, (float4*)spheres, s_box, b, n_slabs));
}
void ref_a08_triangleShadowTrace(uint cols, uint rows, void* shadow, void* t_pos, uint* t_box, const float* bound, uint n_slabs) {
    a08::AABB b = mk_aabb<a08::AABB>(bound);
    uint2 cr(cols, rows);
    ND2(cols, rows, a08::triangleShadowTrace(cr, (a08::Ray*)shadow, (float3*)t_pos, t_box, b, n_slabs));
}
void ref_a08_sceneRender(void* acu, void* pois, void* shadow, void* material, uint pixels) {
    ND1(pixels, a08::sceneRender((float4*)acu, (a08::Poi*)pois, (a08::Ray*)shadow, (float4*)material, pixels));
}
void ref_a08_copyToPixel(void* pixel, void* acu, float m, uint pixels) {
    ND1(pixels, a08::copyToPixel((uchar4*)pixel, (float4*)acu, m, pixels));
}

// ======================================================================== A09 (1-D over total_rays)
uint ref_a09_sizeofRay() { uint s; a09::sizeofRay(&s); return s; }
uint ref_a09_sizeofPoi() { uint s; a09::sizeofPoi(&s); return s; }
void ref_a09_initTrace(void* acu, void* rays, void* pois, const float* bound, const float* cam, float focal_length, float lens_rad,
                       uint rays_per_pixel, uint cols, uint rows) {
    float16 c = mk_f16(cam);
    a09::AABB b = mk_aabb<a09::AABB>(bound);
    ND2(cols, rows, a09::initTrace((float4*)acu, (a09::Ray*)rays, (a09::Poi*)pois, b, c, focal_length, lens_rad, rays_per_pixel));
}
void ref_a09_initShadowTrace(void* shadow, void* pois, uint total_rays, const float* light_pos) {
    float3 lp = mk_f3(light_pos);
    ND1(total_rays, a09::initShadowTrace((a09::Ray*)shadow, (a09::Poi*)pois, total_rays, lp));
}
void ref_a09_sphereTrace(uint total_rays, void* pois, void* rays, void* spheres, uint* s_matid, uint* s_box, const float* bound, uint n_slabs) {
    a09::AABB b = mk_aabb<a09::AABB>(bound);
    ND1(total_rays, a09::sphereTrace(total_rays, (a09::Poi*)pois, (a09::Ray*)rays, (float4*)spheres, s_matid, s_box, b, n_slabs));
}
void ref_a09_triangleTrace(uint total_rays, void* pois, void* rays, void* t_pos, void* t_normal, uint* t_matid, uint* t_box, const float* bound, uint n_slabs) {
    a09::AABB b = mk_aabb<a09::AABB>(bound);
    ND1(total_rays, a09::triangleTrace(total_rays, (a09::Poi*)pois, (a09::Ray*)rays, (float3*)t_pos, (float3*)t_normal, t_matid, t_box, b, n_slabs));
}
void ref_a09_sphereShadowTrace(uint total_rays, void* shadow, void* spheres, uint* s_box, const float* bound, uint n_slabs) {
    a09::AABB b = mk_aabb<a09::AABB>(bound);
    ND1(total_rays, a09::sphereShadowTrace(total_rays, (a09::Ray*)shadow, (float4*)spheres, s_box, b, n_slabs));
}
void ref_a09_triangleShadowTrace(uint total_rays, void* shadow, void* t_pos, uint* t_box, const float* bound, uint n_slabs) {
    a09::AABB b = mk_aabb<a09::AABB>(bound);
    ND1(total_rays, a09::triangleShadowTrace(total_rays, (a09::Ray*)shadow, (float3*)t_pos, t_box, b, n_slabs));
}
void ref_a09_sceneRender(void* acu, void* pois, void* shadow, void* material, uint total_rays) {
    ND1(total_rays, a09::sceneRender((float4*)acu, (a09::Poi*)pois, (a09::Ray*)shadow, (float4*)material, total_rays));
}
void ref_a09_copyToPixel(void* pixel, void* acu, float m, uint pixels, uint rays_per_pixel) {
    ND1(pixels, a09::copyToPixel((uchar4*)pixel, (float4*)acu, m, pixels, rays_per_pixel));
}

// ======================================================================== A10
uint ref_a10_sizeofRay() { uint s; a10::sizeofRay(&s); return s; }
uint ref_a10_sizeofPoi() { uint s; a10::sizeofPoi(&s); return s; }
void ref_a10_initAcu(void* acu, uint total_rays) {
    ND1(total_rays, a10::initAcu((float4*)acu, total_rays));
}
// serial != 0 runs the launch single-threaded in row-major order: the only defined
// outcome of the reference's seeds[col] race when rays_per_pixel == 1 (Q7).
void ref_a10_initTrace(int* seeds, void* rays, void* pois, const float* bound, const float* cam, float focal_length, float lens_rad,
                       uint rays_per_pixel, uint cols, uint rows, int serial) {
    float16 c = mk_f16(cam);
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    if (serial) {
        ND2_SERIAL(cols, rows, a10::initTrace(seeds, (a10::Ray*)rays, (a10::Poi*)pois, b, c, focal_length, lens_rad, rays_per_pixel));
    } else {
        ND2(cols, rows, a10::initTrace(seeds, (a10::Ray*)rays, (a10::Poi*)pois, b, c, focal_length, lens_rad, rays_per_pixel));
    }
}
// Row tile [row0, row0+nrows) of the same launch: `rays`/`pois` (and `seeds`) hold only the
// tile's slots; the kernel is handed pointers shifted so that its own
// `(cols*row+col)*rays_per_pixel` indexing lands inside them.  Exact because every work-item
// touches only its own pixel's slots.  rays_per_pixel must be > 1 (no RNG in initTrace).
void ref_a10_initTrace_rows(int* seeds, void* rays, void* pois, const float* bound, const float* cam, float focal_length, float lens_rad,
                            uint rays_per_pixel, uint cols, uint rows, uint row0, uint nrows) {
    float16 c = mk_f16(cam);
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    (void)rows;
    long long shift = (long long)row0 * cols * rays_per_pixel;
    a10::Ray* r0 = (a10::Ray*)rays - shift;
    a10::Poi* p0 = (a10::Poi*)pois - shift;
    _Pragma("omp parallel for schedule(dynamic, 1)")
    for (long long _r = row0; _r < (long long)row0 + nrows; _r++)
        for (long long _c = 0; _c < (long long)cols; _c++) {
            cl_gid[0] = (size_t)_c; cl_gid[1] = (size_t)_r; cl_gid[2] = 0;
            a10::initTrace(seeds, r0, p0, b, c, focal_length, lens_rad, rays_per_pixel);
        }
}
void ref_a10_bouncePaths(void* pois, void* rays, int* seeds, uint total_rays) {
    ND1(total_rays, a10::bouncePaths((a10::Poi*)pois, (a10::Ray*)rays, seeds, total_rays));
}
void ref_a10_lightRender(void* pois, void* rays, void* acu, const float* light, uint total_rays) {
    float16 l = mk_f16(light);
    ND1(total_rays, a10::lightRender((a10::Poi*)pois, (a10::Ray*)rays, (float4*)acu, l, total_rays));
}
void ref_a10_initShadowTrace(void* shadow, void* pois, uint total_rays, const float* light, int* seeds) {
    float16 l = mk_f16(light);
    ND1(total_rays, a10::initShadowTrace((a10::Ray*)shadow, (a10::Poi*)pois, total_rays, l, seeds));
}
// Q1: keep pois[id].atte across a closest-hit work-item.
#define KEEP_ATTE(call)                                  \
    {                                                    \
        a10::Poi* _p = (a10::Poi*)pois + _i;             \
        float _a0 = _p->atte.x, _a1 = _p->atte.y, _a2 = _p->atte.z, _a3 = _p->atte.v[3]; \
        call;                                            \
        _p->atte.x = _a0; _p->atte.y = _a1; _p->atte.z = _a2; _p->atte.v[3] = _a3; \
    }
void ref_a10_sphereTrace(uint total_rays, void* pois, void* rays, void* spheres, uint* s_matid, uint* s_box, const float* bound, uint n_slabs) {
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    ND1(total_rays, KEEP_ATTE(a10::sphereTrace(total_rays, (a10::Poi*)pois, (a10::Ray*)rays, (float4*)spheres, s_matid, s_box, b, n_slabs)));
}
void ref_a10_triangleTrace(uint total_rays, void* pois, void* rays, void* t_pos, void* t_normal, uint* t_matid, uint* t_box, const float* bound, uint n_slabs) {
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    ND1(total_rays, KEEP_ATTE(a10::triangleTrace(total_rays, (a10::Poi*)pois, (a10::Ray*)rays, (float3*)t_pos, (float3*)t_normal, t_matid, t_box, b, n_slabs)));
}
void ref_a10_meshTrace(uint total_rays, void* pois, void* rays, void* t_pos, void* t_normal, uint* t_box, uint t_matid, const float* bound, uint n_slabs) {
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    ND1(total_rays, KEEP_ATTE(a10::meshTrace(total_rays, (a10::Poi*)pois, (a10::Ray*)rays, (float3*)t_pos, (float3*)t_normal, t_box, t_matid, b, n_slabs)));
}
void ref_a10_sphereShadowTrace(uint total_rays, void* shadow, void* spheres, uint* s_box, const float* bound, uint n_slabs) {
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    ND1(total_rays, a10::sphereShadowTrace(total_rays, (a10::Ray*)shadow, (float4*)spheres, s_box, b, n_slabs));
}
void ref_a10_triangleShadowTrace(uint total_rays, void* shadow, void* t_pos, uint* t_box, const float* bound, uint n_slabs) {
    a10::AABB b = mk_aabb<a10::AABB>(bound);
    ND1(total_rays, a10::triangleShadowTrace(total_rays, (a10::Ray*)shadow, (float3*)t_pos, t_box, b, n_slabs));
}
void ref_a10_sceneRender(void* acu, void* pois, void* shadow, void* material, const float* light, uint total_rays) {
    float16 l = mk_f16(light);
    ND1(total_rays, a10::sceneRender((float4*)acu, (a10::Poi*)pois, (a10::Ray*)shadow, (float4*)material, l, total_rays));
}
void ref_a10_copyToPixel(void* pixel, void* acu, float m, uint pixels, uint rays_per_pixel) {
    ND1(pixels, a10::copyToPixel((uchar4*)pixel, (float4*)acu, m, pixels, rays_per_pixel));
}

}  // extern "C"
