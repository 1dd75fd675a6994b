/* rt2015_napi.c -- N-API addon that lets a headless Node.js host drive librt2015.so in place of
 * the WebCL object model the reference's code.js uses (Assign10-Path_Tracing/code.js:576-608,
 * 1047-1552).  Thin by design: TypedArray / ArrayBuffer in, opaque handles (BigInt) and thrown
 * Errors out; no computation here.  EVERY entry point of include/rt2015.h is exported under its C
 * name without the rt_ prefix (rt.a10_meshTrace, rt.render_execute, ...):
 *   - the launchers and most of the render object have uniform signatures (handles, host arrays,
 *     scalars) and are generated from the header by gen_napi.py into rt2015_napi_gen.inc;
 *   - the entry points with out-parameters or struct arguments are written out below.
 *
 * Node.js and node_api.h are absent from the build image, so this file is
 * compile-checked against the hand-declared subset of the stable N-API ABI in napi_min.h
 *   gcc -std=c11 -Wall -Wextra -fsyntax-only -I../include -I. rt2015_napi.c
 * (tests/test_abi.py::test_napi_addon_compiles).  With a real toolchain build it as usual:
 *   gcc -shared -fPIC -DRT_HAVE_NODE_API -I$NODE/include/node -I../include rt2015_napi.c -L.. -lrt2015 -o rt2015.node
 * rt2015.js (same directory) is the code.js-shaped module on top of it.
 */
#ifdef RT_HAVE_NODE_API
#include <node_api.h>
#else
#include "napi_min.h"
#endif
#include <stdint.h>
#include <string.h>

#include "rt2015.h"

#define NAPI_CALL(env, call)                                               \
    do {                                                                   \
        if ((call) != napi_ok) {                                           \
            napi_throw_error((env), NULL, "rt2015: N-API call failed");    \
            return NULL;                                                   \
        }                                                                  \
    } while (0)

static rt_ctx* g_last_ctx;   /* for error text only */

static napi_value rt_throw(napi_env env, int rc) {
    char msg[600];
    const char* detail = g_last_ctx ? rt_last_error_string(g_last_ctx) : "";
    const char* name = rc == RT_ERR_NO_DEVICE ? "no CUDA device (there is no CPU fallback)" : rc == RT_ERR_NOMEM ? "out of device memory"
                       : rc == RT_ERR_INVALID ? "invalid argument" : rc == RT_ERR_STATE ? "call order violated" : "CUDA error";
    size_t n = strlen(name);
    memcpy(msg, name, n);
    msg[n++] = ':'; msg[n++] = ' ';
    strncpy(msg + n, detail ? detail : "", sizeof msg - n - 1);
    msg[sizeof msg - 1] = 0;
    napi_throw_error(env, NULL, msg);   /* the reference alert()s the build log, A10/code.js:598-606 */
    return NULL;
}

static int get_args(napi_env env, napi_callback_info info, size_t want, napi_value* argv) {
    size_t argc = want;
    if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < want) {
        napi_throw_type_error(env, NULL, "rt2015: wrong number of arguments");
        return 0;
    }
    return 1;
}

static void* get_handle(napi_env env, napi_value v) {   /* BigInt -> pointer (null / undefined -> NULL) */
    uint64_t u = 0;
    bool lossless = false;
    if (napi_get_value_bigint_uint64(env, v, &u, &lossless) != napi_ok) return NULL;
    return (void*)(uintptr_t)u;
}

static napi_value make_handle(napi_env env, const void* p) {
    napi_value v = NULL;
    NAPI_CALL(env, napi_create_bigint_uint64(env, (uint64_t)(uintptr_t)p, &v));
    return v;
}

static void* typed_data(napi_env env, napi_value v, size_t* bytes) {   /* TypedArray / ArrayBuffer -> host pointer */
    bool is_ta = false;
    void* data = NULL;
    size_t len = 0;
    if (napi_is_typedarray(env, v, &is_ta) == napi_ok && is_ta) {
        napi_typedarray_type t;
        napi_value ab;
        size_t off = 0;
        static const size_t esz[] = {1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
        if (napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) != napi_ok) return NULL;
        if (bytes) *bytes = len * esz[t];
        return data;
    }
    if (napi_get_arraybuffer_info(env, v, &data, &len) != napi_ok) return NULL;
    if (bytes) *bytes = len;
    return data;
}

static void* typed_or_null(napi_env env, napi_value v) {   /* null / undefined -> NULL (optional host arrays) */
    napi_valuetype vt = napi_undefined;
    if (napi_typeof(env, v, &vt) != napi_ok || vt == napi_null || vt == napi_undefined) return NULL;
    return typed_data(env, v, NULL);
}

/* ---- small object helpers ----------------------------------------------------------------- */
static int set_num(napi_env env, napi_value obj, const char* key, double v) {
    napi_value n;
    return napi_create_double(env, v, &n) == napi_ok && napi_set_named_property(env, obj, key, n) == napi_ok;
}
static int set_handle(napi_env env, napi_value obj, const char* key, const void* p) {
    napi_value n;
    return napi_create_bigint_uint64(env, (uint64_t)(uintptr_t)p, &n) == napi_ok && napi_set_named_property(env, obj, key, n) == napi_ok;
}
static double get_num(napi_env env, napi_value obj, const char* key, double dflt) {
    bool has = false;
    napi_value v;
    double d = dflt;
    if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return dflt;
    if (napi_get_named_property(env, obj, key, &v) != napi_ok || napi_get_value_double(env, v, &d) != napi_ok) return dflt;
    return d;
}
static void* get_handle_prop(napi_env env, napi_value obj, const char* key) {
    napi_value v;
    if (napi_get_named_property(env, obj, key, &v) != napi_ok) return NULL;
    return get_handle(env, v);
}
/* a fresh TypedArray holding a copy of `count` elements */
static napi_value copy_out(napi_env env, napi_typedarray_type t, size_t elem, const void* src, size_t count) {
    void* dst = NULL;
    napi_value ab, ta;
    NAPI_CALL(env, napi_create_arraybuffer(env, count * elem, &dst, &ab));
    if (count) memcpy(dst, src, count * elem);
    NAPI_CALL(env, napi_create_typedarray(env, t, count, ab, 0, &ta));
    return ta;
}

/* rt_grid <-> { prim, normal, matid, box_size, occupancy: BigInt, n_refs, n_slabs, kind: Number } */
static napi_value grid_to_js(napi_env env, const rt_grid* g) {
    napi_value o;
    NAPI_CALL(env, napi_create_object(env, &o));
    if (!set_handle(env, o, "prim", g->prim) || !set_handle(env, o, "normal", g->normal) || !set_handle(env, o, "matid", g->matid) ||
        !set_handle(env, o, "box_size", g->box_size) || !set_handle(env, o, "occupancy", g->occupancy) || !set_num(env, o, "n_refs", g->n_refs) ||
        !set_num(env, o, "n_slabs", g->n_slabs) || !set_num(env, o, "kind", g->kind)) {
        napi_throw_error(env, NULL, "rt2015: N-API call failed");
        return NULL;
    }
    return o;
}
static void grid_from_js(napi_env env, napi_value o, rt_grid* g) {
    memset(g, 0, sizeof *g);
    g->prim = get_handle_prop(env, o, "prim");
    g->normal = get_handle_prop(env, o, "normal");
    g->matid = get_handle_prop(env, o, "matid");
    g->box_size = get_handle_prop(env, o, "box_size");
    g->occupancy = get_handle_prop(env, o, "occupancy");
    g->n_refs = (unsigned)get_num(env, o, "n_refs", 0);
    g->n_slabs = (unsigned)get_num(env, o, "n_slabs", 0);
    g->kind = (unsigned)get_num(env, o, "kind", 0);
}

/* ---- context / buffers: webcl.createContext, createBuffer, enqueue{Write,Read}Buffer ---------- */
static napi_value js_ctx_create(napi_env env, napi_callback_info info) {   /* ctx_create(device_ordinal) -> ctx */
    napi_value a[1];
    int32_t dev = 0;
    rt_ctx* ctx = NULL;
    if (!get_args(env, info, 1, a)) return NULL;
    NAPI_CALL(env, napi_get_value_int32(env, a[0], &dev));
    int rc = rt_ctx_create(dev, &ctx);
    if (rc) return rt_throw(env, rc);
    g_last_ctx = ctx;
    return make_handle(env, ctx);
}

static napi_value js_last_error_string(napi_env env, napi_callback_info info) {
    napi_value a[1], s;
    if (!get_args(env, info, 1, a)) return NULL;
    const char* t = rt_last_error_string((rt_ctx*)get_handle(env, a[0]));
    NAPI_CALL(env, napi_create_string_utf8(env, t ? t : "", NAPI_AUTO_LENGTH, &s));
    return s;
}

static napi_value js_ctx_stream(napi_env env, napi_callback_info info) {
    napi_value a[1];
    if (!get_args(env, info, 1, a)) return NULL;
    return make_handle(env, rt_ctx_stream((rt_ctx*)get_handle(env, a[0])));
}

static napi_value js_device_info(napi_env env, napi_callback_info info) {   /* -> { sm_count, cc_major, cc_minor, l2_bytes, total_mem } */
    napi_value a[1], o;
    int sm = 0, maj = 0, min = 0;
    size_t l2 = 0, mem = 0;
    if (!get_args(env, info, 1, a)) return NULL;
    int rc = rt_device_info((rt_ctx*)get_handle(env, a[0]), &sm, &maj, &min, &l2, &mem);
    if (rc) return rt_throw(env, rc);
    NAPI_CALL(env, napi_create_object(env, &o));
    if (!set_num(env, o, "sm_count", sm) || !set_num(env, o, "cc_major", maj) || !set_num(env, o, "cc_minor", min) ||
        !set_num(env, o, "l2_bytes", (double)l2) || !set_num(env, o, "total_mem", (double)mem))
        return NULL;
    return o;
}

static napi_value js_buffer_create(napi_env env, napi_callback_info info) {   /* buffer_create(ctx, bytes) -> device pointer */
    napi_value a[2];
    void* d = NULL;
    int64_t bytes = 0;
    if (!get_args(env, info, 2, a)) return NULL;
    NAPI_CALL(env, napi_get_value_int64(env, a[1], &bytes));
    int rc = rt_buffer_create((rt_ctx*)get_handle(env, a[0]), (size_t)bytes, &d);
    return rc ? rt_throw(env, rc) : make_handle(env, d);
}

static napi_value js_buffer_write(napi_env env, napi_callback_info info) {   /* buffer_write(ctx, buf, typedArray[, byteOffset]) */
    napi_value a[4];
    size_t bytes = 0, argc = 4;
    int64_t off = 0;
    if (napi_get_cb_info(env, info, &argc, a, NULL, NULL) != napi_ok || argc < 3) {
        napi_throw_type_error(env, NULL, "rt2015: wrong number of arguments");
        return NULL;
    }
    if (argc > 3) NAPI_CALL(env, napi_get_value_int64(env, a[3], &off));
    void* host = typed_data(env, a[2], &bytes);
    int rc = rt_buffer_write((rt_ctx*)get_handle(env, a[0]), get_handle(env, a[1]), (size_t)off, bytes, host);
    return rc ? rt_throw(env, rc) : NULL;
}

static napi_value js_buffer_read(napi_env env, napi_callback_info info) {    /* buffer_read(ctx, buf, typedArray[, byteOffset]); synchronises */
    napi_value a[4];
    size_t bytes = 0, argc = 4;
    int64_t off = 0;
    if (napi_get_cb_info(env, info, &argc, a, NULL, NULL) != napi_ok || argc < 3) {
        napi_throw_type_error(env, NULL, "rt2015: wrong number of arguments");
        return NULL;
    }
    if (argc > 3) NAPI_CALL(env, napi_get_value_int64(env, a[3], &off));
    void* host = typed_data(env, a[2], &bytes);
    int rc = rt_buffer_read((rt_ctx*)get_handle(env, a[0]), get_handle(env, a[1]), (size_t)off, bytes, host);
    return rc ? rt_throw(env, rc) : NULL;
}

static napi_value js_struct_size(napi_env env, napi_callback_info info) {   /* getStructSize("Ray"|"Poi"), A10/code.js:1064-1076 */
    napi_value a[2], out = NULL;
    char name[8];
    size_t len = 0;
    int32_t assignment = 10;
    if (!get_args(env, info, 2, a)) return NULL;
    NAPI_CALL(env, napi_get_value_string_utf8(env, a[0], name, sizeof name, &len));
    NAPI_CALL(env, napi_get_value_int32(env, a[1], &assignment));
    NAPI_CALL(env, napi_create_uint32(env, rt_struct_size(name, assignment), &out));
    return out;
}

/* ---- grid build: split*Data (A10/code.js:899-1041, 1554-1772), A06's slab splitters ------------ */
/* grid_build_spheres(ctx, Float64Array xyzr, Uint32Array id|null, n, Float64Array bmin[3], Float64Array bmax[3], n_slabs) -> grid */
static napi_value js_grid_build_spheres(napi_env env, napi_callback_info info) {
    napi_value a[7];
    uint32_t n = 0, ns = 0;
    rt_grid g;
    if (!get_args(env, info, 7, a)) return NULL;
    NAPI_CALL(env, napi_get_value_uint32(env, a[3], &n));
    NAPI_CALL(env, napi_get_value_uint32(env, a[6], &ns));
    int rc = rt_grid_build_spheres((rt_ctx*)get_handle(env, a[0]), (const double*)typed_or_null(env, a[1]), (const unsigned*)typed_or_null(env, a[2]), n,
                                   (const double*)typed_or_null(env, a[4]), (const double*)typed_or_null(env, a[5]), ns, &g);
    return rc ? rt_throw(env, rc) : grid_to_js(env, &g);
}
/* grid_build_triangles(ctx, pos9, nor9, id|null, n, bmin, bmax, n_slabs, Float64Array xform[11]|null) -> grid
 * xform = [do_normalize, center xyz, maxdim, scale xyz, translate xyz]  (Mesh.normalize/scale/translate, A10/code.js:114-169) */
static napi_value js_grid_build_triangles(napi_env env, napi_callback_info info) {
    napi_value a[9];
    uint32_t n = 0, ns = 0;
    rt_grid g;
    rt_mesh_xform xf;
    if (!get_args(env, info, 9, a)) return NULL;
    NAPI_CALL(env, napi_get_value_uint32(env, a[4], &n));
    NAPI_CALL(env, napi_get_value_uint32(env, a[7], &ns));
    const double* x = (const double*)typed_or_null(env, a[8]);
    if (x) {
        xf.do_normalize = x[0] != 0.0;
        memcpy(xf.center, x + 1, sizeof xf.center);
        xf.maxdim = x[4];
        memcpy(xf.scale, x + 5, sizeof xf.scale);
        memcpy(xf.translate, x + 8, sizeof xf.translate);
    }
    int rc = rt_grid_build_triangles((rt_ctx*)get_handle(env, a[0]), (const double*)typed_or_null(env, a[1]), (const double*)typed_or_null(env, a[2]),
                                     (const unsigned*)typed_or_null(env, a[3]), n, (const double*)typed_or_null(env, a[5]),
                                     (const double*)typed_or_null(env, a[6]), ns, x ? &xf : NULL, &g);
    return rc ? rt_throw(env, rc) : grid_to_js(env, &g);
}
/* slab_build_spheres(ctx, xyzr, id|null, n, x_min, x_max, n_slabs) -> grid   (A06/code.js:456-520) */
static napi_value js_slab_build_spheres(napi_env env, napi_callback_info info) {
    napi_value a[7];
    uint32_t n = 0, ns = 0;
    double x0 = 0, x1 = 0;
    rt_grid g;
    if (!get_args(env, info, 7, a)) return NULL;
    NAPI_CALL(env, napi_get_value_uint32(env, a[3], &n));
    NAPI_CALL(env, napi_get_value_double(env, a[4], &x0));
    NAPI_CALL(env, napi_get_value_double(env, a[5], &x1));
    NAPI_CALL(env, napi_get_value_uint32(env, a[6], &ns));
    int rc = rt_slab_build_spheres((rt_ctx*)get_handle(env, a[0]), (const double*)typed_or_null(env, a[1]), (const unsigned*)typed_or_null(env, a[2]), n,
                                   x0, x1, ns, &g);
    return rc ? rt_throw(env, rc) : grid_to_js(env, &g);
}
/* slab_build_triangles(ctx, pos9, nor9, id|null, n, x_min, x_max, n_slabs) -> grid   (A06/code.js:936-1043) */
static napi_value js_slab_build_triangles(napi_env env, napi_callback_info info) {
    napi_value a[8];
    uint32_t n = 0, ns = 0;
    double x0 = 0, x1 = 0;
    rt_grid g;
    if (!get_args(env, info, 8, a)) return NULL;
    NAPI_CALL(env, napi_get_value_uint32(env, a[4], &n));
    NAPI_CALL(env, napi_get_value_double(env, a[5], &x0));
    NAPI_CALL(env, napi_get_value_double(env, a[6], &x1));
    NAPI_CALL(env, napi_get_value_uint32(env, a[7], &ns));
    int rc = rt_slab_build_triangles((rt_ctx*)get_handle(env, a[0]), (const double*)typed_or_null(env, a[1]), (const double*)typed_or_null(env, a[2]),
                                     (const unsigned*)typed_or_null(env, a[3]), n, x0, x1, ns, &g);
    return rc ? rt_throw(env, rc) : grid_to_js(env, &g);
}
static napi_value js_grid_release(napi_env env, napi_callback_info info) {   /* grid_release(ctx, grid) */
    napi_value a[2];
    rt_grid g;
    if (!get_args(env, info, 2, a)) return NULL;
    grid_from_js(env, a[1], &g);
    int rc = rt_grid_release((rt_ctx*)get_handle(env, a[0]), &g);
    return rc ? rt_throw(env, rc) : NULL;
}

/* ---- native loaders: parseMeshJSON (A10/tri/meshDataVersion1.js:12-78), parsePDB (A10/mol/pdbParserV1.js:2-85) --
 * parse_mesh_json(Uint8Array text) -> { nTriangles, nMaterials, positions, normals: Float64Array, materialIndices: Uint32Array,
 *                                       materials: Float64Array, boundsMin, boundsMax: Float64Array(3) }
 * The native arrays are copied into fresh TypedArrays and released here (rt_mesh_data_free / rt_mol_data_free). */
static napi_value js_parse_mesh_json(napi_env env, napi_callback_info info) {
    napi_value a[1], o, v;
    size_t bytes = 0;
    rt_mesh_data d;
    if (!get_args(env, info, 1, a)) return NULL;
    const char* text = (const char*)typed_data(env, a[0], &bytes);
    int rc = rt_parse_mesh_json(text, bytes, &d);
    if (rc) return rt_throw(env, rc);
    NAPI_CALL(env, napi_create_object(env, &o));
    int ok = set_num(env, o, "nTriangles", d.n_triangles) && set_num(env, o, "nMaterials", d.n_materials);
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.positions, 9u * (size_t)d.n_triangles)) && napi_set_named_property(env, o, "positions", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.normals, 9u * (size_t)d.n_triangles)) && napi_set_named_property(env, o, "normals", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_uint32_array, 4, d.material_indices, d.n_triangles)) && napi_set_named_property(env, o, "materialIndices", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.materials, 4u * (size_t)d.n_materials)) && napi_set_named_property(env, o, "materials", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.bounds_min, 3)) && napi_set_named_property(env, o, "boundsMin", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.bounds_max, 3)) && napi_set_named_property(env, o, "boundsMax", v) == napi_ok;
    rt_mesh_data_free(&d);
    return ok ? o : NULL;
}
/* parse_pdb(Uint8Array text) -> { size, nRecords, nElements, atomData, colorData, radiusData: Float64Array, boundsMin, boundsMax } */
static napi_value js_parse_pdb(napi_env env, napi_callback_info info) {
    napi_value a[1], o, v;
    size_t bytes = 0;
    rt_mol_data d;
    if (!get_args(env, info, 1, a)) return NULL;
    const char* text = (const char*)typed_data(env, a[0], &bytes);
    int rc = rt_parse_pdb(text, bytes, &d);
    if (rc) return rt_throw(env, rc);
    NAPI_CALL(env, napi_create_object(env, &o));
    int ok = set_num(env, o, "size", d.size) && set_num(env, o, "nRecords", d.n_records) && set_num(env, o, "nElements", d.n_elements);
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.atom_data, 4u * (size_t)d.n_records)) && napi_set_named_property(env, o, "atomData", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.color_data, 4u * (size_t)d.n_elements)) && napi_set_named_property(env, o, "colorData", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.radius_data, d.n_elements)) && napi_set_named_property(env, o, "radiusData", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.bounds_min, 3)) && napi_set_named_property(env, o, "boundsMin", v) == napi_ok;
    ok = ok && (v = copy_out(env, napi_float64_array, 8, d.bounds_max, 3)) && napi_set_named_property(env, o, "boundsMax", v) == napi_ok;
    rt_mol_data_free(&d);
    return ok ? o : NULL;
}

/* ---- scene / render: preRender, executeRender, postRender (A10/code.js:1784-1859) ------------------ */
static napi_value js_scene_create(napi_env env, napi_callback_info info) {   /* scene_create(ctx) -> scene */
    napi_value a[1];
    rt_scene* s = NULL;
    if (!get_args(env, info, 1, a)) return NULL;
    int rc = rt_scene_create((rt_ctx*)get_handle(env, a[0]), &s);
    return rc ? rt_throw(env, rc) : make_handle(env, s);
}
/* comm_create(ctx, world, rank, Uint8Array id[128]) -> comm   (multi-GPU: one process per GPU, id from comm_unique_id on rank 0) */
static napi_value js_comm_create(napi_env env, napi_callback_info info) {
    napi_value a[4];
    rt_comm* c = NULL;
    int32_t world = 0, rank = 0;
    if (!get_args(env, info, 4, a)) return NULL;
    NAPI_CALL(env, napi_get_value_int32(env, a[1], &world));
    NAPI_CALL(env, napi_get_value_int32(env, a[2], &rank));
    int rc = rt_comm_create((rt_ctx*)get_handle(env, a[0]), world, rank, (const unsigned char*)typed_or_null(env, a[3]), &c);
    return rc ? rt_throw(env, rc) : make_handle(env, c);
}
/* scene_add_set(scene, grid, Float32Array bound[8], is_mesh, mesh_matid) */
static napi_value js_scene_add_set(napi_env env, napi_callback_info info) {
    napi_value a[5];
    rt_grid g;
    int32_t is_mesh = 0;
    uint32_t matid = 0;
    if (!get_args(env, info, 5, a)) return NULL;
    grid_from_js(env, a[1], &g);
    NAPI_CALL(env, napi_get_value_int32(env, a[3], &is_mesh));
    NAPI_CALL(env, napi_get_value_uint32(env, a[4], &matid));
    int rc = rt_scene_add_set((rt_scene*)get_handle(env, a[0]), &g, (const float*)typed_or_null(env, a[2]), is_mesh, matid);
    return rc ? rt_throw(env, rc) : NULL;
}
/* render_create(ctx, scene, { cols, rows, rays_per_pixel, depth, focal_length, lens_rad, slot_begin, slot_count, mode, tile_slots }) -> render */
static napi_value js_render_create(napi_env env, napi_callback_info info) {
    napi_value a[3];
    rt_render_opts o;
    rt_render* r = NULL;
    if (!get_args(env, info, 3, a)) return NULL;
    memset(&o, 0, sizeof o);
    o.cols = (unsigned)get_num(env, a[2], "cols", 0);
    o.rows = (unsigned)get_num(env, a[2], "rows", 0);
    o.rays_per_pixel = (unsigned)get_num(env, a[2], "rays_per_pixel", 1);
    o.depth = (unsigned)get_num(env, a[2], "depth", 5);   /* the reference hard-codes 5 bounces, A10/code.js:1829 */
    o.focal_length = (float)get_num(env, a[2], "focal_length", 1.0);
    o.lens_rad = (float)get_num(env, a[2], "lens_rad", 0.0);
    o.slot_begin = (unsigned)get_num(env, a[2], "slot_begin", 0);
    o.slot_count = (unsigned)get_num(env, a[2], "slot_count", 0);
    o.mode = (unsigned)get_num(env, a[2], "mode", 0);
    o.tile_slots = (unsigned)get_num(env, a[2], "tile_slots", 0);
    int rc = rt_render_create((rt_ctx*)get_handle(env, a[0]), (rt_scene*)get_handle(env, a[1]), &o, &r);
    return rc ? rt_throw(env, rc) : make_handle(env, r);
}
static napi_value js_render_accum_image(napi_env env, napi_callback_info info) {   /* -> device pointer of the float4 image */
    napi_value a[1];
    void* d = NULL;
    if (!get_args(env, info, 1, a)) return NULL;
    int rc = rt_render_accum_image((rt_render*)get_handle(env, a[0]), &d);
    return rc ? rt_throw(env, rc) : make_handle(env, d);
}
static napi_value js_render_stats(napi_env env, napi_callback_info info) {   /* -> { closest_rays, any_rays: BigInt, launches, device_ms } */
    napi_value a[1], o, v;
    unsigned long long c = 0, y = 0;
    unsigned launches = 0;
    float ms = 0.f;
    if (!get_args(env, info, 1, a)) return NULL;
    int rc = rt_render_stats((rt_render*)get_handle(env, a[0]), &c, &y, &launches, &ms);
    if (rc) return rt_throw(env, rc);
    NAPI_CALL(env, napi_create_object(env, &o));
    NAPI_CALL(env, napi_create_bigint_uint64(env, c, &v));
    NAPI_CALL(env, napi_set_named_property(env, o, "closest_rays", v));
    NAPI_CALL(env, napi_create_bigint_uint64(env, y, &v));
    NAPI_CALL(env, napi_set_named_property(env, o, "any_rays", v));
    if (!set_num(env, o, "launches", launches) || !set_num(env, o, "device_ms", ms)) return NULL;
    return o;
}

/* a089_render_frame(ctx, frame, acu, out_matid|null, out_maxt|null): the whole Assignment 8 / 9 frame in one launch.
 * frame = { spheres: grid|null, s_bound: Float32Array(8), triangles: grid|null, t_bound, t_shadow_bound: Float32Array(8),
 *           material: BigInt, light_pos: Float32Array(4 * n_lights), n_lights, bound: Float32Array(8), fcam: Float32Array(16),
 *           focal_length, lens_rad, rays_per_pixel, thin_lens }  (grids as returned by grid_build_*) */
static int copy_floats(napi_env env, napi_value obj, const char* key, float* dst, size_t n) {
    napi_value v;
    size_t bytes = 0;
    if (napi_get_named_property(env, obj, key, &v) != napi_ok) return 0;
    const float* src = (const float*)typed_data(env, v, &bytes);
    if (!src || bytes < n * sizeof(float)) return 0;
    memcpy(dst, src, n * sizeof(float));
    return 1;
}
static int grid_prop(napi_env env, napi_value obj, const char* key, rt_grid* g) {   /* 1 = present */
    napi_value v;
    napi_valuetype vt = napi_undefined;
    memset(g, 0, sizeof *g);
    if (napi_get_named_property(env, obj, key, &v) != napi_ok || napi_typeof(env, v, &vt) != napi_ok || vt != napi_object) return 0;
    grid_from_js(env, v, g);
    return g->prim != NULL;
}
static napi_value js_a089_render_frame(napi_env env, napi_callback_info info) {
    napi_value a[5], lp;
    rt_a089_frame f;
    rt_grid gs, gt;
    if (!get_args(env, info, 5, a)) return NULL;
    memset(&f, 0, sizeof f);
    if (grid_prop(env, a[1], "spheres", &gs)) {
        f.spheres = gs.prim; f.s_matid = gs.matid; f.s_box_size = gs.box_size; f.s_n_slabs = gs.n_slabs;
        if (!copy_floats(env, a[1], "s_bound", f.s_bound, 8)) return rt_throw(env, RT_ERR_INVALID);
    }
    if (grid_prop(env, a[1], "triangles", &gt)) {
        f.t_pos = gt.prim; f.t_normal = gt.normal; f.t_matid = gt.matid; f.t_box_size = gt.box_size; f.t_n_slabs = gt.n_slabs;
        if (!copy_floats(env, a[1], "t_bound", f.t_bound, 8) || !copy_floats(env, a[1], "t_shadow_bound", f.t_shadow_bound, 8))
            return rt_throw(env, RT_ERR_INVALID);
    }
    f.material = get_handle_prop(env, a[1], "material");
    f.n_lights = (unsigned)get_num(env, a[1], "n_lights", 0);
    f.light_pos = napi_get_named_property(env, a[1], "light_pos", &lp) == napi_ok ? (const float*)typed_or_null(env, lp) : NULL;
    if (!copy_floats(env, a[1], "bound", f.bound, 8) || !copy_floats(env, a[1], "fcam", f.fcam, 16)) return rt_throw(env, RT_ERR_INVALID);
    f.focal_length = (float)get_num(env, a[1], "focal_length", 1.0);
    f.lens_rad = (float)get_num(env, a[1], "lens_rad", 0.0);
    f.rays_per_pixel = (unsigned)get_num(env, a[1], "rays_per_pixel", 1);
    f.thin_lens = (unsigned)get_num(env, a[1], "thin_lens", 0);
    int rc = rt_a089_render_frame((rt_ctx*)get_handle(env, a[0]), &f, get_handle(env, a[2]), get_handle(env, a[3]), get_handle(env, a[4]));
    return rc ? rt_throw(env, rc) : NULL;
}

/* ---- everything with a uniform signature (launchers, buffers, scene setters, render object) ---- */
#include "rt2015_napi_gen.inc"

napi_value rt2015_init(napi_env env, napi_value exports) {
    static const struct { const char* name; napi_callback fn; } table[] = {
        {"ctx_create", js_ctx_create}, {"last_error_string", js_last_error_string}, {"ctx_stream", js_ctx_stream},
        {"device_info", js_device_info}, {"buffer_create", js_buffer_create}, {"buffer_write", js_buffer_write},
        {"buffer_read", js_buffer_read}, {"struct_size", js_struct_size}, {"grid_build_spheres", js_grid_build_spheres},
        {"grid_build_triangles", js_grid_build_triangles}, {"slab_build_spheres", js_slab_build_spheres},
        {"slab_build_triangles", js_slab_build_triangles}, {"grid_release", js_grid_release}, {"parse_mesh_json", js_parse_mesh_json},
        {"parse_pdb", js_parse_pdb}, {"scene_create", js_scene_create}, {"scene_add_set", js_scene_add_set},
        {"render_create", js_render_create}, {"render_accum_image", js_render_accum_image}, {"render_stats", js_render_stats},
        {"a089_render_frame", js_a089_render_frame}, {"comm_create", js_comm_create},
        RT2015_GENERATED_EXPORTS
    };
    for (size_t i = 0; i < sizeof table / sizeof table[0]; i++) {
        napi_value fn;
        NAPI_CALL(env, napi_create_function(env, table[i].name, NAPI_AUTO_LENGTH, table[i].fn, NULL, &fn));
        NAPI_CALL(env, napi_set_named_property(env, exports, table[i].name, fn));
    }
    return exports;
}
#ifdef RT_HAVE_NODE_API
NAPI_MODULE(NODE_GYP_MODULE_NAME, rt2015_init)
#endif
