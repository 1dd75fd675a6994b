/* rt2015.h -- C ABI of librt2015.so, the B200 (sm_100a) implementation of the data-parallel
 * hot path of eaymerich/2015-RayTracing.
 *
 * This is the drop-in boundary: it replaces the WebCL object model the reference's host
 * code drives (context / queue / program / kernel / buffer; call sites
 * Assign10-Path_Tracing/code.js:576-608, 1047-1099, 1101-1552) with plain C entry points
 * a Node N-API addon, a cgo/JNI stub or Python ctypes can bind directly.  There are no C++
 * or torch types in any signature: device memory is an opaque `void*` (a CUDA device
 * pointer; memory allocated elsewhere in the same CUDA context, e.g. by a tensor library,
 * may be passed as well), small by-value kernel arguments of the reference (camera
 * `float16`, light `float16`, `AABB`) are host `const float*`.
 *
 * Three layers, each citing what it replaces (paths relative to /root/reference,
 * A07/A08/A09/A10 = the assignment directories):
 *   1. context + buffers      = webcl.createContext / createBuffer / enqueue{Write,Read}Buffer
 *   2. one launcher per kernel = createKernel + setArg + enqueueNDRangeKernel, argument
 *                                order and struct layouts of code.cl kept
 *   3. grid build + scene + render = split*Data / preRender / executeRender / postRender
 *
 * Conventions: every function returns 0 on success or a negative rt_status; launchers are
 * asynchronous on the context's single in-order stream (the reference uses one in-order
 * command queue, A10/code.js:592); rt_finish / rt_buffer_read synchronise.  A context is
 * not thread-safe (neither is the reference); contexts on different GPUs are independent.
 */
#ifndef RT2015_H
#define RT2015_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RT_OK = 0,
    RT_ERR_INVALID = -1,      /* bad argument */
    RT_ERR_CUDA = -2,         /* CUDA runtime error, see rt_last_error_string */
    RT_ERR_NOMEM = -3,
    RT_ERR_NO_DEVICE = -4,    /* no usable CUDA device: there is NO CPU fallback */
    RT_ERR_STATE = -5         /* call order violated (e.g. execute before seeds) */
} rt_status;

typedef struct rt_ctx rt_ctx;
typedef struct rt_scene rt_scene;
typedef struct rt_render rt_render;
typedef struct rt_comm rt_comm;

/* ---- 1. context and buffers ------------------------------------------------------------ */
/* createCLBasicResources, A10/code.js:576-608 */
int rt_ctx_create(int device_ordinal, rt_ctx** out);
int rt_ctx_destroy(rt_ctx* ctx);                                  /* releaseCLResources :1539-1552 */
const char* rt_last_error_string(rt_ctx* ctx);                    /* program build log / exceptions */
int rt_finish(rt_ctx* ctx);                                       /* cmdQueue.finish() */
void* rt_ctx_stream(rt_ctx* ctx);                                 /* the cudaStream_t launches go to */
int rt_device_info(rt_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes, size_t* total_mem);

int rt_buffer_create(rt_ctx* ctx, size_t bytes, void** dptr);     /* ctx.createBuffer */
int rt_buffer_release(rt_ctx* ctx, void* dptr);                   /* buffer.release() */
int rt_buffer_write(rt_ctx* ctx, void* dptr, size_t offset, size_t bytes, const void* host);   /* enqueueWriteBuffer */
int rt_buffer_read(rt_ctx* ctx, const void* dptr, size_t offset, size_t bytes, void* host);    /* enqueueReadBuffer + finish */
int rt_buffer_fill(rt_ctx* ctx, void* dptr, int byte_value, size_t bytes);

/* sizeofRay / sizeofPoi probe kernels, A10/code.cl:440-446, A10/code.js:1064-1076.
 * name = "Ray" | "Poi"; assignment = 3..10.  Returns 0 for an unknown struct. */
unsigned rt_struct_size(const char* name, int assignment);

/* ---- 2. one launcher per reference kernel ------------------------------------------------
 * Pointer arguments are device pointers; `bound` = 8 floats (pmin.xyz,1,pmax.xyz,1 as
 * bounds2AABB packs them, A10/code.js:610-621); `fcam`/`light_info` = 16 floats. */

/* Assignment 10 (A10/code.cl) */
int rt_a10_initAcu(rt_ctx*, void* acu, unsigned total_rays);                                           /* :448-456 */
int rt_a10_initTrace(rt_ctx*, void* seeds, void* rays, void* pois, const float* bound, const float* fcam,
                     float focal_length, float lens_rad, unsigned rays_per_pixel);                     /* :458-543 */
int rt_a10_bouncePaths(rt_ctx*, void* pois, void* rays, void* seeds, unsigned total_rays);             /* :581-598 */
int rt_a10_lightRender(rt_ctx*, void* pois, void* rays, void* acu, const float* light_info, unsigned total_rays);   /* :600-629 */
int rt_a10_initShadowTrace(rt_ctx*, void* shadow_rays, void* pois, unsigned total_rays, const float* light_info,
                           void* seeds);                                                               /* :631-673 */
int rt_a10_sphereTrace(rt_ctx*, unsigned total_rays, void* pois, void* rays, const void* spheres, const void* s_matid,
                       const void* s_box_size, const float* bound, unsigned n_slabs);                  /* :675-800 */
int rt_a10_triangleTrace(rt_ctx*, unsigned total_rays, void* pois, void* rays, const void* t_pos, const void* t_normal,
                         const void* t_matid, const void* t_box_size, const float* bound, unsigned n_slabs);   /* :802-935 */
int rt_a10_meshTrace(rt_ctx*, unsigned total_rays, void* pois, void* rays, const void* t_pos, const void* t_normal,
                     const void* t_box_size, unsigned t_matid, const float* bound, unsigned n_slabs);  /* :937-1070 */
int rt_a10_sphereShadowTrace(rt_ctx*, unsigned total_rays, void* shadow_rays, const void* spheres, const void* s_box_size,
                             const float* bound, unsigned n_slabs);                                    /* :1073-1193 */
int rt_a10_triangleShadowTrace(rt_ctx*, unsigned total_rays, void* shadow_rays, const void* t_pos, const void* t_box_size,
                               const float* bound, unsigned n_slabs);                                  /* :1195-1321 */
int rt_a10_sceneRender(rt_ctx*, void* acu, void* pois, const void* shadow_rays, const void* material,
                       const float* light_info, unsigned total_rays);                                  /* :1323-1364 */
int rt_a10_copyToPixel(rt_ctx*, void* pixel, const void* acu, float m, unsigned pixels, unsigned rays_per_pixel);   /* :1366-1386 */

/* Earlier assignments (BASELINE.json configs 1-4).  Same conventions; the 2-D NDRange kernels take
 * the canvas size from the camera (fcam[14], fcam[15]) like the reference kernels do, A08's
 * `uint2 cols_rows` argument is passed as (cols, rows).  `light_pos` = 4 floats (x,y,z,1), the
 * Vec3.toFloat32Array of A08/code.js:17-19. */
int rt_a01_raytrace(rt_ctx*, void* pixels, const float* fcam);                                          /* A01/code.cl:116-147 (fcam packs rows, cols) */
int rt_a02_raytrace(rt_ctx*, void* pixels, const float* fcam, unsigned s_size, const void* s_atoms,
                    const void* s_colors);                                                              /* A02/code.cl:158-232 */
int rt_a03_initTrace(rt_ctx*, void* pixels, const float* fcam, void* rays);                             /* A03/code.cl:132-143 */
int rt_a03_molTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms,
                    const void* s_colors);                                                              /* A03/code.cl:145-187 */
/* A04-A06: brute force over spheres and over a triangle soup sharing one ray buffer, the
 * same with bounding boxes, and 1-D slabs along x.  t_pos / t_normal = 3 x float4 per triangle (w ignored). */
int rt_a04_initTrace(rt_ctx*, void* pixels, const float* fcam, void* rays);                             /* A04/code.cl:204-215 */
int rt_a04_molTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms,
                    const void* s_colors);                                                              /* A04/code.cl:217-259 */
int rt_a04_meshTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos,
                     const void* t_normal, const void* t_mindex, const void* m_color);                  /* A04/code.cl:261-315 */
int rt_a04_raytrace(rt_ctx*, void* pixels, const float* fcam, unsigned s_size, const void* s_atoms,
                    const void* s_colors);                                                              /* A04/code.cl:317-364 (= A02's kernel) */
int rt_a05_initTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, const float* bound);         /* A05/code.cl:304-328 */
int rt_a05_molTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms,
                    const void* s_colors, const float* bound);                                          /* A05/code.cl:330-384 */
int rt_a05_meshTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos,
                     const void* t_normal, const void* t_mindex, const void* m_color, const float* bound);   /* A05/code.cl:386-452 */
int rt_a06_initTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, const float* bound);         /* A06/code.cl:310-334 */
int rt_a06_molTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms,
                    const void* s_colors, const float* bound, unsigned n_slabs, const void* slab_size); /* A06/code.cl:336-426; s_atoms.w = RADIUS */
int rt_a06_meshTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos,
                     const void* t_normal, const void* t_mindex, const void* m_color, const float* bound,
                     unsigned n_slabs, const void* slab_size);                                          /* A06/code.cl:428-533 */
int rt_a07_initTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, const float* bound);         /* A07/code.cl:311-335 */
int rt_a07_molTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned s_size, const void* s_atoms,
                    const void* s_mindex, const void* m_color, const float* bound, unsigned n_slabs,
                    const void* slab_size);                                                             /* A07/code.cl:337-473 */
int rt_a07_meshTrace(rt_ctx*, void* pixels, const float* fcam, void* rays, unsigned t_size, const void* t_pos,
                     const void* t_normal, const void* t_mindex, const void* m_color, const float* bound,
                     unsigned n_slabs, const void* slab_size);                                          /* A07/code.cl:475-626 */
int rt_a08_initTrace(rt_ctx*, void* acu, void* rays, void* pois, const float* bound, const float* fcam);   /* A08/code.cl:331-363 */
int rt_a08_initShadowTrace(rt_ctx*, void* shadow_rays, void* pois, unsigned cols, unsigned rows, const float* light_pos);   /* :365-390 */
int rt_a08_sphereTrace(rt_ctx*, unsigned cols, unsigned rows, void* pois, void* rays, const void* spheres, const void* s_matid,
                       const void* s_box_size, const float* bound, unsigned n_slabs);                   /* A08/code.cl:392-517 */
int rt_a08_triangleTrace(rt_ctx*, unsigned cols, unsigned rows, void* pois, void* rays, const void* t_pos, const void* t_normal,
                         const void* t_matid, const void* t_box_size, const float* bound, unsigned n_slabs);
int rt_a08_sphereShadowTrace(rt_ctx*, unsigned cols, unsigned rows, void* shadow_rays, const void* spheres, const void* s_box_size,
                             const float* bound, unsigned n_slabs);
int rt_a08_triangleShadowTrace(rt_ctx*, unsigned cols, unsigned rows, void* shadow_rays, const void* t_pos, const void* t_box_size,
                               const float* bound, unsigned n_slabs);
int rt_a08_sceneRender(rt_ctx*, void* acu, void* pois, const void* shadow_rays, const void* material, unsigned pixels);   /* :916-939 */
int rt_a08_copyToPixel(rt_ctx*, void* pixel, const void* acu, float m, unsigned pixels);                /* A08/code.cl:941-951 */
int rt_a09_initTrace(rt_ctx*, void* acu, void* rays, void* pois, const float* bound, const float* fcam, float focal_length,
                     float lens_rad, unsigned rays_per_pixel);                                          /* A09/code.cl:400-461 */
int rt_a09_initShadowTrace(rt_ctx*, void* shadow_rays, void* pois, unsigned total_rays, const float* light_pos);   /* A09/code.cl:463-485 */
int rt_a09_sphereTrace(rt_ctx*, unsigned total_rays, void* pois, void* rays, const void* spheres, const void* s_matid,
                       const void* s_box_size, const float* bound, unsigned n_slabs);
int rt_a09_triangleTrace(rt_ctx*, unsigned total_rays, void* pois, void* rays, const void* t_pos, const void* t_normal,
                         const void* t_matid, const void* t_box_size, const float* bound, unsigned n_slabs);
int rt_a09_sphereShadowTrace(rt_ctx*, unsigned total_rays, void* shadow_rays, const void* spheres, const void* s_box_size,
                             const float* bound, unsigned n_slabs);
int rt_a09_triangleShadowTrace(rt_ctx*, unsigned total_rays, void* shadow_rays, const void* t_pos, const void* t_box_size,
                               const float* bound, unsigned n_slabs);
int rt_a09_sceneRender(rt_ctx*, void* acu, void* pois, const void* shadow_rays, const void* material, unsigned total_rays);
int rt_a09_copyToPixel(rt_ctx*, void* pixel, const void* acu, float m, unsigned pixels, unsigned rays_per_pixel);   /* A09/code.cl:1023-1040 */

/* The whole deterministic frame of Assignment 8 / 9 in ONE launch (ours: render() of A08/code.js:1194-1232 and
 * A09/code.js:1256-1294 without the per-kernel round trips of 48-byte Ray / Poi records through memory): primary ray
 * (pinhole, or thin lens when `thin_lens`), closest hit over the sphere and triangle grids, per point light a shadow ray
 * with any-hit traces and sceneRender -- the accumulator of every slot, bit for bit what the launcher sequence leaves in
 * `acu`.  Follow it with rt_a08_copyToPixel / rt_a09_copyToPixel.  Absent sets: spheres / t_pos = NULL.  All pointers are
 * device pointers except light_pos (host, 4 floats per light, at most 16 lights). */
typedef struct {
    const void* spheres; const void* s_matid; const void* s_box_size; float s_bound[8]; unsigned s_n_slabs;
    const void* t_pos; const void* t_normal; const void* t_matid; const void* t_box_size; float t_bound[8]; unsigned t_n_slabs;
    float t_shadow_bound[8];      /* bound handed to triangleShadowTrace: the SPHERE bounds in A08 (A08/code.js:918), t_bound in A09 */
    const void* material;
    const float* light_pos; unsigned n_lights;
    float bound[8]; float fcam[16];
    float focal_length, lens_rad; unsigned rays_per_pixel; unsigned thin_lens;
} rt_a089_frame;
/* out_matid (int per slot) / out_maxt (float per slot): optional hit record outputs, may be NULL */
int rt_a089_render_frame(rt_ctx*, const rt_a089_frame* frame, void* acu, void* out_matid, void* out_maxt);

/* Optional per-work-item statistics for the grid-walk launchers -- the five of Assignment 10 (sphereTrace,
 * triangleTrace, meshTrace, sphereShadowTrace, triangleShadowTrace), their Assignment 8 / 9 twins and molTrace /
 * meshTrace of Assignment 7 (all
 * device pointers, one uint per work-item, any may be NULL; pass all NULL to switch off): winning reference index
 * champ_i (0xFFFFFFFF = none; A10/code.cl:882-897), cells visited, primitive tests.  Used for the hit-primitive-id
 * parity gate and for the algorithmic byte count of the roofline.  With occupancy bits present the
 * walk skips the table loads of empty cells but still counts them as visited, like the reference's loop. */
int rt_set_walk_stats(rt_ctx*, void* hit_id_u32, void* cells_u32, void* tests_u32);
/* The same counters as running totals over every grid-walk launch until switched off (NULL): a device array of 8 x uint64
 * the launches ADD to -- [0] rays that entered a walk kernel alive (mint != maxt), [1] of those, walks started (the set's box was
 * hit), [2] cells visited, [3] primitive tests, [4] hits, [5] triangle tests that passed the face cull; [6], [7] reserved.
 * Feeds the algorithmic byte count 48 + 8 C + B T + H (...) per ray of BASELINE.md section 3 for configs 3 and 4. */
int rt_set_walk_totals(rt_ctx*, void* totals_u64x8);

/* ---- 3. grid build, scene, render ---------------------------------------------------------
 * Uniform-grid build = splitSphereData / splitTriangleData / splitMeshData
 * (A10/code.js:1554-1641, 1643-1772, 899-1041; A07/code.js:889-1122) as integer CUDA kernels
 * (count -> exclusive scan -> order-preserving scatter), binning in float64 like the
 * JavaScript.  Inputs are HOST arrays of doubles (JS Numbers).  Outputs are device buffers
 * owned by the returned rt_grid (release with rt_grid_release). */
typedef struct {
    void* prim;          /* spheres: float4 (cx,cy,cz,r*r) per ref; triangles: 3 x float4 (w=0) per ref */
    void* normal;        /* triangles: 3 x float4 per ref; spheres: NULL */
    void* matid;         /* uint per ref (material / atom index) or NULL */
    void* box_size;      /* uint[n^3 + 1] exclusive prefix sums, cell = z*n*n + y*n + x */
    void* occupancy;     /* internal: 1 bit per cell (non-empty), used by the fused render path */
    unsigned n_refs;
    unsigned n_slabs;
    unsigned kind;       /* 0 = sphere, 1 = triangle */
    unsigned _reserved;
} rt_grid;

/* Post-split position transform of Mesh.normalize/scale/translate (A10/code.js:114-169),
 * applied in float64 to the cell-ordered positions before the fp32 store:
 *   p = (p - center) * maxdim   (if do_normalize)   ;   p *= scale   ;   p += translate */
typedef struct {
    int do_normalize;
    double center[3];
    double maxdim;
    double scale[3];
    double translate[3];
} rt_mesh_xform;

/* xyzr: n x 4 doubles (cx,cy,cz,radius); id: n uints (material or atom index) or NULL. */
int rt_grid_build_spheres(rt_ctx*, const double* xyzr, const unsigned* id, unsigned n, const double bmin[3],
                          const double bmax[3], unsigned n_slabs, rt_grid* out);
/* pos9/nor9: n x 9 doubles (three vertices); id: n uints or NULL; xform may be NULL. */
int rt_grid_build_triangles(rt_ctx*, const double* pos9, const double* nor9, const unsigned* id, unsigned n,
                            const double bmin[3], const double bmax[3], unsigned n_slabs, const rt_mesh_xform* xform,
                            rt_grid* out);
/* 1-D slabs along x = the two splitters of Assignment 6 (prepareMolTrace A06/code.js:456-520, splitData
 * A06/code.js:936-1043): same binning rule on the x extent only.  The returned rt_grid has box_size[n_slabs + 1];
 * sphere records keep the RADIUS in w (A06's interSphere squares it itself). */
int rt_slab_build_spheres(rt_ctx*, const double* xyzr, const unsigned* id, unsigned n, double x_min, double x_max,
                          unsigned n_slabs, rt_grid* out);
int rt_slab_build_triangles(rt_ctx*, const double* pos9, const double* nor9, const unsigned* id, unsigned n, double x_min,
                            double x_max, unsigned n_slabs, rt_grid* out);
int rt_grid_release(rt_ctx*, rt_grid* g);

/* Native fast paths of the two loaders that feed the grid build (host code; the library allocates the arrays, release
 * them with the matching *_free).  Results equal the JavaScript loaders number for number, including gl-matrix's
 * Float32Array rounding of transformed vertices (A10/lib/gl-matrix.js:79-80). */
typedef struct {
    unsigned n_triangles, n_materials;
    double* positions;            /* 9 per triangle */
    double* normals;              /* 9 per triangle */
    unsigned* material_indices;   /* 1 per triangle */
    double* materials;            /* 4 per material (diffuseReflectance) */
    double bounds_min[3], bounds_max[3];
} rt_mesh_data;
int rt_parse_mesh_json(const char* text, size_t len, rt_mesh_data* out);    /* parseMeshJSON, A10/tri/meshDataVersion1.js:12-78 */
void rt_mesh_data_free(rt_mesh_data* d);
typedef struct {
    unsigned size;                /* atoms.length = largest serial (may exceed n_records: quirk Q13) */
    unsigned n_records, n_elements;
    double* atom_data;            /* 4 per record: element index, x, y, z */
    double* color_data;           /* 4 per element */
    double* radius_data;          /* 1 per element */
    double bounds_min[3], bounds_max[3];
} rt_mol_data;
int rt_parse_pdb(const char* text, size_t len, rt_mol_data* out);           /* parsePDB, A10/mol/pdbParserV1.js:2-85 */
void rt_mol_data_free(rt_mol_data* d);

/* Scene = what preRender uploads (A10/code.js:1784-1804). */
int rt_scene_create(rt_ctx*, rt_scene** out);
int rt_scene_destroy(rt_scene*);
int rt_scene_set_bounds(rt_scene*, const float bound[8]);                              /* scene.bounds */
int rt_scene_set_materials(rt_scene*, const float* rgba, unsigned n_materials);        /* splitMaterialData :1774-1782 */
/* Geometry sets are traced in the order added: spheres, scene triangles, meshes
 * (executeRender, A10/code.js:1809-1813).  `matid` is used for meshes only (scalar
 * t_matid of meshTrace); per-reference ids come from grid->matid otherwise. */
int rt_scene_add_set(rt_scene*, const rt_grid* grid, const float bound[8], int is_mesh, unsigned mesh_matid);
/* Light.toShadowInfo / toSceneRenderInfo / toLightRenderInfo, A10/code.js:323-352 */
int rt_scene_add_light(rt_scene*, const float shadow_info[16], const float scene_info[16], const float light_info[16]);

/* Diagnostic of the wavefront path (ours): which rays of a Ray buffer (48-byte records) would NOT be sent to the queue
 * walker of geometry set `set_index` because their whole walk through that set's grid provably crosses empty cells only and is
 * skipped (a third of the walks of BASELINE config 5); for a 1-cell set of wall triangles: which shadow segments (unit direction,
 * maxt = length) are not tested against it because both ends lie inside the room.  out_flags: one byte per ray on the device, 1 = skipped.  The parity tests
 * check every flagged ray against the instrumented reference kernels (it must visit no reference there).  RT2015_NO_SKIP=1 in the
 * environment switches the skip off. */
int rt_scene_probe_empty_walks(rt_scene*, unsigned set_index, const void* rays, unsigned n, void* out_flags_u8);

typedef struct {
    unsigned cols, rows;          /* canvas size */
    unsigned rays_per_pixel;      /* slots per pixel; > 1 must be a perfect square (stratified lens grid) */
    unsigned depth;               /* bounces after the primary hit; the reference hard-codes 5 (A10/code.js:1829) */
    float focal_length;           /* scene.focal_length */
    float lens_rad;               /* scene.lens_diameter / 2 */
    unsigned slot_begin, slot_count;   /* multi-GPU: this context renders slots k in [slot_begin, slot_begin+slot_count)
                                          of every pixel; 0,0 = all rays_per_pixel slots.  slot_count == 0 with
                                          slot_begin > 0 (a rank left without slots when world > rays_per_pixel) is
                                          rejected: such a rank renders nothing and contributes a zero image */
    unsigned mode;                /* 0 = wavefront path (default), 1 = reference kernel-by-kernel schedule, 2 = megakernel */
    unsigned tile_slots;          /* wavefront tile size in ray slots, 0 = auto (a quarter of device memory) */
} rt_render_opts;

/* preRender: allocates ray/hit/accumulation state (A10/code.js:1078-1138, 1417-1442). */
int rt_render_create(rt_ctx*, rt_scene*, const rt_render_opts* opts, rt_render** out);
int rt_render_destroy(rt_render*);                                                      /* postRender */
/* prepareInitSeeds (A10/code.js:1140-1154): `seeds` holds cols*rows*rays_per_pixel ints indexed by
 * the GLOBAL slot id (pixel*rays_per_pixel + k); host memory unless `on_device` != 0. */
int rt_render_set_seeds(rt_render*, const int* seeds, size_t count, int on_device);
/* executeRender (A10/code.js:1806-1854): one progressive pass -- primary rays, lights, `depth`
 * bounces with next-event estimation, accumulation.  `host_pixels` (cols*rows*4 bytes, RGBA)
 * receives copyToPixel's image when non-NULL (executeCopyToPixel + sendImagetoHTML). */
int rt_render_execute(rt_render*, const float fcam[16], unsigned char* host_pixels);
/* Per-pixel float accumulation image: sum over this context's slots of acu (float4 per pixel,
 * .w = contribution count).  Device pointer, valid until the next execute/destroy. */
int rt_render_accum_image(rt_render*, void** dptr_float4);
int rt_render_read_accum(rt_render*, float* host_float4);                               /* cols*rows*4 floats */
int rt_render_read_seeds(rt_render*, int* host_seeds, size_t count);                    /* seed state after the pass */
/* Same as rt_render_set_seeds for a caller that already holds only THIS context's slots, laid out
 * [pixel][k_local] (count = cols*rows*slot_count host ints). */
int rt_render_write_local_seeds(rt_render*, const int* host_seeds, size_t count);
/* Non-blocking form of the same upload, like the reference's own enqueueWriteBuffer(buf, false, ...) (A10/code.js:1149):
 * issued on a side stream behind the work already queued, returns at once; the next rt_render_execute overlaps it with
 * ray generation and the primary traversal and waits for it in front of its first kernel that draws random numbers.
 * `host_seeds` (pinned memory for a truly asynchronous copy) must stay valid until that rt_render_execute has returned. */
int rt_render_write_local_seeds_async(rt_render*, const int* host_seeds, size_t count);
/* Progressive state of a render = what lives only on the device between passes in the reference: the per-slot
 * accumulators, the per-slot seeds and the pass counter (A10/code.js:416, 1078-1099, 1140-1154, 1853).  Export after
 * any pass, import into a render created with the same scene/options to resume bit-exactly where it stopped
 * (checkpoint / resume; the reference loses this state on stopRender).  acu: cols*rows*slot_count float4,
 * seeds: cols*rows*slot_count ints, both [pixel][k_local]; either pointer may be NULL to skip that part. */
int rt_render_export_state(rt_render*, float* host_acu_float4, int* host_seeds, unsigned* passes);
int rt_render_import_state(rt_render*, const float* host_acu_float4, const int* host_seeds, unsigned passes);
/* copyToPixel on an accumulation image (after a multi-GPU reduce): m = 1/(rays_per_pixel*passes). */
int rt_accum_to_pixel(rt_ctx*, void* pixel, const void* accum_float4, float m, unsigned pixels);
/* Multi-GPU (ours; the reference drives one device, A10/code.js:576-608): one process and one context per GPU, each
 * rendering its own slot range (rt_render_opts.slot_begin / slot_count), then ONE sum-reduce of the per-pixel
 * accumulation images into `root` -- ncclReduce(sum, fp32) over NVLink, issued on the context's stream, asynchronous like
 * the launchers.  NCCL is bound at run time (libnccl.so.2; RT2015_NCCL_LIB overrides the path); a host that never creates
 * a communicator never loads it.  Rank 0 obtains an id, ships the RT_COMM_ID_BYTES bytes to the other processes by any
 * means (file, socket, env), every rank then calls rt_comm_create (collective).  After rt_render_reduce the root
 * follows with rt_accum_to_pixel on rt_render_accum_image. */
#define RT_COMM_ID_BYTES 128
int rt_comm_unique_id(unsigned char id[RT_COMM_ID_BYTES]);
int rt_comm_create(rt_ctx*, int world, int rank, const unsigned char id[RT_COMM_ID_BYTES], rt_comm** out);
int rt_comm_destroy(rt_comm*);
int rt_render_reduce(rt_render*, rt_comm*, int root);
/* Counters of the last execute: valid closest-hit and any-hit queries ("rays"),
 * kernels launched, device milliseconds between the first and last launch. */
int rt_render_stats(rt_render*, unsigned long long* closest_rays, unsigned long long* any_rays, unsigned* launches,
                    float* device_ms);

/* Work profile of the fused path (for the roofline's ALGORITHMIC byte count): when
 * switched on, the next executes run an instrumented build of the same kernel and accumulate
 *   [0] closest (ray,set) queries entered alive   [1] of those, walks started (set AABB hit)
 *   [2] cells visited   [3] sphere tests   [4] triangle tests
 *   [5] hits on sphere sets   [6] hits on triangle sets with a matid array   [7] hits on meshes
 *   [8] any-hit (ray,set) queries entered alive   [9] walks started   [10] cells   [11] sphere tests
 *   [12] triangle tests   [13] blocked   [14] slots processed
 *   [15] triangle tests (closest + any hit) that pass the face cull and run the full barycentric test
 * Timing taken with the profile on is not a benchmark number. */
int rt_render_set_profile(rt_render*, int on);
int rt_render_read_profile(rt_render*, unsigned long long out[16]);
/* Same counters per geometry set, in the order the sets were added (row s = set s; [14] is kept in
 * row 0 only).  rt_render_read_profile returns the column sums. */
#define RT_MAX_SETS 8
int rt_render_read_profile_sets(rt_render*, unsigned long long out[RT_MAX_SETS * 16]);

/* Per-kernel-class device time of the executes since timing was switched on (CUDA events on the
 * context's stream, one in front of every launch): the dominant kernel's average launch duration for
 * the roofline.  Classes:
 *   0 per-slot stage kernels (ray generation, small sets, shading)   1 queue walker, spheres, closest hit
 *   2 queue walker, spheres, any hit    3 queue walker, triangles, closest hit
 *   4 queue walker, triangles, any hit  5 megakernel   6 reference-schedule kernels   7 sum/copyToPixel */
#define RT_TIMING_CLASSES 8
int rt_render_set_timing(rt_render*, int on);
int rt_render_read_timing(rt_render*, float ms[RT_TIMING_CLASSES], unsigned launches[RT_TIMING_CLASSES]);

#ifdef __cplusplus
}
#endif
#endif /* RT2015_H */
