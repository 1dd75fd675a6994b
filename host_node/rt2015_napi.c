/* rt2015_napi.c -- N-API addon that lets a headless Node.js host drive librt2015.so in place of
 * the WebCL object model the reference's code.js uses (Assign10-Path_Tracing/code.js:576-608,
 * 1047-1552).  Thin by design: TypedArray / ArrayBuffer in, opaque handles (BigInt) and thrown
 * Errors out; no computation here.
 *
 * Node.js and node_api.h are absent from the build image, so this file is
 * compile-checked against the hand-declared subset of the stable N-API ABI in napi_min.h
 *   gcc -std=c11 -Wall -Wextra -fsyntax-only -I../include -I. rt2015_napi.c
 * (tests/test_abi.py::test_napi_addon_compiles).  With a real toolchain build it as usual:
 *   gcc -shared -fPIC -I$NODE/include/node -I../include rt2015_napi.c -L.. -lrt2015 -o rt2015.node
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
    strncpy(msg + n, detail, sizeof msg - n - 1);
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

static void* get_handle(napi_env env, napi_value v) {   /* BigInt -> pointer */
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

/* ---- context / buffers: webcl.createContext, createBuffer, enqueue{Write,Read}Buffer, finish ---- */
static napi_value js_ctx_create(napi_env env, napi_callback_info info) {
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

static napi_value js_ctx_destroy(napi_env env, napi_callback_info info) {
    napi_value a[1];
    if (!get_args(env, info, 1, a)) return NULL;
    rt_ctx* ctx = (rt_ctx*)get_handle(env, a[0]);
    if (ctx == g_last_ctx) g_last_ctx = NULL;
    int rc = rt_ctx_destroy(ctx);
    return rc ? rt_throw(env, rc) : NULL;
}

static napi_value js_buffer_create(napi_env env, napi_callback_info info) {
    napi_value a[2];
    void* d = NULL;
    int64_t bytes = 0;
    if (!get_args(env, info, 2, a)) return NULL;
    NAPI_CALL(env, napi_get_value_int64(env, a[1], &bytes));
    int rc = rt_buffer_create((rt_ctx*)get_handle(env, a[0]), (size_t)bytes, &d);
    return rc ? rt_throw(env, rc) : make_handle(env, d);
}

static napi_value js_buffer_write(napi_env env, napi_callback_info info) {   /* (ctx, buf, typedArray) */
    napi_value a[3];
    size_t bytes = 0;
    if (!get_args(env, info, 3, a)) return NULL;
    void* host = typed_data(env, a[2], &bytes);
    int rc = rt_buffer_write((rt_ctx*)get_handle(env, a[0]), get_handle(env, a[1]), 0, bytes, host);
    return rc ? rt_throw(env, rc) : NULL;
}

static napi_value js_buffer_read(napi_env env, napi_callback_info info) {    /* (ctx, buf, typedArray) */
    napi_value a[3];
    size_t bytes = 0;
    if (!get_args(env, info, 3, a)) return NULL;
    void* host = typed_data(env, a[2], &bytes);
    int rc = rt_buffer_read((rt_ctx*)get_handle(env, a[0]), get_handle(env, a[1]), 0, bytes, host);
    return rc ? rt_throw(env, rc) : NULL;
}

static napi_value js_buffer_release(napi_env env, napi_callback_info info) {
    napi_value a[2];
    if (!get_args(env, info, 2, a)) return NULL;
    int rc = rt_buffer_release((rt_ctx*)get_handle(env, a[0]), get_handle(env, a[1]));
    return rc ? rt_throw(env, rc) : NULL;
}

static napi_value js_finish(napi_env env, napi_callback_info info) {
    napi_value a[1];
    if (!get_args(env, info, 1, a)) return NULL;
    int rc = rt_finish((rt_ctx*)get_handle(env, a[0]));
    return rc ? rt_throw(env, rc) : NULL;
}

/* ---- one launcher shown in full; the others follow the same pattern ------------------------
 * meshTrace(ctx, total_rays, pois, rays, t_pos, t_normal, t_box_size, t_matid, Float32Array bound[8], n_slabs)
 * = createKernel("meshTrace") + 10 x setArg + enqueueNDRangeKernel, A10/code.js:1254-1303 */
static napi_value js_a10_meshTrace(napi_env env, napi_callback_info info) {
    napi_value a[10];
    uint32_t total = 0, matid = 0, n = 0;
    if (!get_args(env, info, 10, a)) return NULL;
    NAPI_CALL(env, napi_get_value_uint32(env, a[1], &total));
    NAPI_CALL(env, napi_get_value_uint32(env, a[7], &matid));
    NAPI_CALL(env, napi_get_value_uint32(env, a[9], &n));
    int rc = rt_a10_meshTrace((rt_ctx*)get_handle(env, a[0]), total, get_handle(env, a[2]), get_handle(env, a[3]), get_handle(env, a[4]),
                              get_handle(env, a[5]), get_handle(env, a[6]), matid, (const float*)typed_data(env, a[8], NULL), n);
    return rc ? rt_throw(env, rc) : NULL;
}

/* ---- render-frame entry: executeRender (A10/code.js:1806-1854) + sendImagetoHTML (:1530-1537) ----
 * renderExecute(render, Float32Array cam16, Uint8ClampedArray pixels|null) */
static napi_value js_render_execute(napi_env env, napi_callback_info info) {
    napi_value a[3];
    if (!get_args(env, info, 3, a)) return NULL;
    napi_valuetype vt;
    NAPI_CALL(env, napi_typeof(env, a[2], &vt));
    unsigned char* pix = vt == napi_null || vt == napi_undefined ? NULL : (unsigned char*)typed_data(env, a[2], NULL);
    int rc = rt_render_execute((rt_render*)get_handle(env, a[0]), (const float*)typed_data(env, a[1], NULL), pix);
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

napi_value rt2015_init(napi_env env, napi_value exports) {
    static const struct { const char* name; napi_callback fn; } table[] = {
        {"ctxCreate", js_ctx_create}, {"ctxDestroy", js_ctx_destroy}, {"bufferCreate", js_buffer_create}, {"bufferWrite", js_buffer_write},
        {"bufferRead", js_buffer_read}, {"bufferRelease", js_buffer_release}, {"finish", js_finish}, {"structSize", js_struct_size},
        {"a10_meshTrace", js_a10_meshTrace}, {"renderExecute", js_render_execute},
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
