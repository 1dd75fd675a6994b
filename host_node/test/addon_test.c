/* addon_test.c -- TEST INFRASTRUCTURE.  Drives the N-API addon (../rt2015_napi.c) through the in-process N-API
 * stand-in (napi_mock.c) exactly the way a JavaScript host would: rt2015_init(exports), then exports.name(args).
 *
 *   addon_test cpu            no device needed: export list, struct_size, the two native loaders, argument errors,
 *                             and the loud failure of ctx_create without a GPU
 *   addon_test gpu <outdir>   on a CUDA box: Assignment-1 frame and a small Assignment-10 render (grid builds, scene,
 *                             render object, two progressive passes) through the addon; inputs and outputs are
 *                             written to <outdir> so that tests/test_node_addon.py can replay the same inputs through
 *                             the ctypes binding and compare bit for bit.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "napi_mock.h"

napi_value rt2015_init(napi_env env, napi_value exports);

static struct napi_env__ g_env;
static napi_value g_exports;

static napi_value call(const char* name, size_t argc, napi_value* argv) {
    napi_value r = mock_call(&g_env, g_exports, name, argc, argv);
    if (!r) {
        fprintf(stderr, "%s threw: %s\n", name, g_env.message);
        exit(2);
    }
    return r;
}
#define CALL(name, ...) call(name, sizeof((napi_value[]){__VA_ARGS__}) / sizeof(napi_value), (napi_value[]){__VA_ARGS__})
#define NUM(x) mock_number((double)(x))
static napi_value f32(size_t n, const float* p) { return mock_typedarray(napi_float32_array, n, p); }
static napi_value f64(size_t n, const double* p) { return mock_typedarray(napi_float64_array, n, p); }
static napi_value u32(size_t n, const unsigned* p) { return mock_typedarray(napi_uint32_array, n, p); }
static napi_value bytes(size_t n, const void* p) { return mock_typedarray(napi_uint8_array, n, p); }

static void dump(const char* dir, const char* name, const void* p, size_t n) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(p, 1, n, f) != n) { perror(path); exit(3); }
    fclose(f);
}

static const char* PDB =
    "HEADER    TEST\n"
    "ATOM      1  C   RES A   1       1.000   2.000   3.000  1.00  0.00           C\n"
    "HETATM    2  O   RES A   1      -1.500   0.250   4.125  1.00  0.00           O\n"
    "TER       3\n"
    "ATOM      4  N   RES A   2       0.000  -2.000   1.000  1.00  0.00           N\n"
    "END\n";
static const char* MESH =
    "{\"meshes\":[{\"vertexPositions\":[0,0,0, 1,0,0, 0,1,0, 1,1,0],\"vertexNormals\":[0,0,1, 0,0,1, 0,0,1, 0,0,1],"
    "\"indices\":[0,1,2, 2,1,3],\"materialIndex\":0}],\"materials\":[{\"diffuseReflectance\":[0.5,0.25,0.125,1]}]}";

static int run_cpu(void) {
    size_t n = 0;
    for (mock_prop* p = g_exports->props; p; p = p->next) { printf("export %s\n", p->key); n++; }
    printf("exports %zu\n", n);
    printf("struct_size Ray10 %d Poi10 %d Poi8 %d Ray6 %d Nope %d\n", (int)CALL("struct_size", mock_string("Ray"), NUM(10))->num,
           (int)CALL("struct_size", mock_string("Poi"), NUM(10))->num, (int)CALL("struct_size", mock_string("Poi"), NUM(8))->num,
           (int)CALL("struct_size", mock_string("Ray"), NUM(6))->num, (int)CALL("struct_size", mock_string("Nope"), NUM(10))->num);
    napi_value mol = CALL("parse_pdb", bytes(strlen(PDB), PDB));
    napi_value ad = mock_get(mol, "atomData"), bmin = mock_get(mol, "boundsMin");
    printf("parse_pdb size %d records %d elements %d atomData.length %zu last %.3f %.3f %.3f boundsMin.x %.3f\n", (int)mock_get(mol, "size")->num,
           (int)mock_get(mol, "nRecords")->num, (int)mock_get(mol, "nElements")->num, ad->length, ((double*)ad->data)[9], ((double*)ad->data)[10],
           ((double*)ad->data)[11], ((double*)bmin->data)[0]);
    napi_value mesh = CALL("parse_mesh_json", bytes(strlen(MESH), MESH));
    napi_value pos = mock_get(mesh, "positions"), mats = mock_get(mesh, "materials");
    printf("parse_mesh_json triangles %d materials %d positions.length %zu p[3] %.1f p[16] %.1f material %.3f %.3f\n",
           (int)mock_get(mesh, "nTriangles")->num, (int)mock_get(mesh, "nMaterials")->num, pos->length, ((double*)pos->data)[3],
           ((double*)pos->data)[16], ((double*)mats->data)[0], ((double*)mats->data)[2]);
    /* argument errors become thrown TypeErrors, a missing device a thrown Error -- never a silent fallback */
    napi_value one[1] = {mock_string("Ray")};
    napi_value r = mock_call(&g_env, g_exports, "struct_size", 1, one);
    printf("struct_size(1 arg): %s\n", r ? "returned" : g_env.message);
    napi_value dev[1] = {NUM(0)};
    r = mock_call(&g_env, g_exports, "ctx_create", 1, dev);
    if (r) {
        printf("ctx_create: ok\n");
        CALL("ctx_destroy", r);
    } else {
        printf("ctx_create threw: %s\n", g_env.message);
    }
    return 0;
}

static int run_gpu(const char* dir) {
    napi_value ctx = CALL("ctx_create", NUM(0));
    napi_value info = CALL("device_info", ctx);
    printf("device sm_count %d cc %d.%d\n", (int)mock_get(info, "sm_count")->num, (int)mock_get(info, "cc_major")->num,
           (int)mock_get(info, "cc_minor")->num);
    /* ---- Assignment 1: one launcher, A01/code.js:166-269 */
    enum { N1 = 64 };
    const float cam1[16] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 2.66f, 2.0f, N1, N1};
    napi_value pix1 = CALL("buffer_create", ctx, NUM(N1 * N1 * 4));
    CALL("a01_raytrace", ctx, pix1, f32(16, cam1));
    napi_value host1 = bytes(N1 * N1 * 4, NULL);
    CALL("buffer_read", ctx, pix1, host1);
    dump(dir, "a01_pixels.bin", host1->data, N1 * N1 * 4);
    CALL("buffer_release", ctx, pix1);

    /* ---- Assignment 10: a floor quad, a sphere, one disk light; split*Data -> preRender -> 2 x executeRender */
    enum { COLS = 48, ROWS = 32, RPP = 4, TOTAL = COLS * ROWS * RPP };
    const double xyzr[4] = {0.0, -0.4, 0.0, 0.6};
    const unsigned sid[1] = {0};
    const double smin[3] = {-0.6, -1.0, -0.6}, smax[3] = {0.6, 0.2, 0.6};
    const double pos9[18] = {-2, -1, -2, -2, -1, 2, 2, -1, -2, 2, -1, -2, -2, -1, 2, 2, -1, 2};
    const double nor9[18] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};
    const unsigned tid[2] = {1, 1};
    const double tmin[3] = {-2, -1.1, -2}, tmax[3] = {2, -0.9, 2};
    const float sbound[8] = {-0.6f, -1.0f, -0.6f, 1, 0.6f, 0.2f, 0.6f, 1}, tbound[8] = {-2, -1.1f, -2, 1, 2, -0.9f, 2, 1};
    const float bound[8] = {-2, -1.1f, -2, 1, 2, 0.2f, 2, 1};
    const float materials[8] = {0.8f, 0.2f, 0.2f, 1, 0.6f, 0.6f, 0.6f, 1};
    const float area = (float)(3.14159265358979323846 * 0.5 * 0.5);
    const float l_shadow[16] = {0, 3, 0, 1, 0, 0, 0, 0, 1, 0.5f}, l_scene[16] = {0, 3, 0, 0, -1, 0, 12, 12, 12, area},
                l_light[16] = {0, 3, 0, 0, -1, 0, 12, 12, 12, 0.5f};
    const float h = 2.0f * (float)tan(0.5 * 60.0 * 3.14159265358979323846 / 180.0);
    const float cam[16] = {0, 0.5f, 5, 1, 0, 0, 0, 1, 0, 0, 0, 1, h * COLS / ROWS, h, COLS, ROWS};
    static int seeds[TOTAL];
    for (int i = 0; i < TOTAL; i++) seeds[i] = 1 + (int)(((unsigned)i * 2654435761u) % 2147483646u);
    dump(dir, "in_cam.bin", cam, sizeof cam);
    dump(dir, "in_seeds.bin", seeds, sizeof seeds);
    dump(dir, "in_xyzr.bin", xyzr, sizeof xyzr);
    dump(dir, "in_smin.bin", smin, sizeof smin);
    dump(dir, "in_smax.bin", smax, sizeof smax);
    dump(dir, "in_pos9.bin", pos9, sizeof pos9);
    dump(dir, "in_nor9.bin", nor9, sizeof nor9);
    dump(dir, "in_tmin.bin", tmin, sizeof tmin);
    dump(dir, "in_tmax.bin", tmax, sizeof tmax);
    dump(dir, "in_sbound.bin", sbound, sizeof sbound);
    dump(dir, "in_tbound.bin", tbound, sizeof tbound);
    dump(dir, "in_bound.bin", bound, sizeof bound);
    dump(dir, "in_materials.bin", materials, sizeof materials);
    dump(dir, "in_l_shadow.bin", l_shadow, sizeof l_shadow);
    dump(dir, "in_l_scene.bin", l_scene, sizeof l_scene);
    dump(dir, "in_l_light.bin", l_light, sizeof l_light);

    napi_value gs = CALL("grid_build_spheres", ctx, f64(4, xyzr), u32(1, sid), NUM(1), f64(3, smin), f64(3, smax), NUM(1));
    napi_value gt = CALL("grid_build_triangles", ctx, f64(18, pos9), f64(18, nor9), u32(2, tid), NUM(2), f64(3, tmin), f64(3, tmax), NUM(1), mock_null());
    printf("grids refs %d %d kind %d %d\n", (int)mock_get(gs, "n_refs")->num, (int)mock_get(gt, "n_refs")->num, (int)mock_get(gs, "kind")->num,
           (int)mock_get(gt, "kind")->num);
    napi_value scene = CALL("scene_create", ctx);
    CALL("scene_set_bounds", scene, f32(8, bound));
    CALL("scene_set_materials", scene, f32(8, materials), NUM(2));
    CALL("scene_add_set", scene, gs, f32(8, sbound), NUM(0), NUM(0));
    CALL("scene_add_set", scene, gt, f32(8, tbound), NUM(0), NUM(0));
    CALL("scene_add_light", scene, f32(16, l_shadow), f32(16, l_scene), f32(16, l_light));
    napi_value opts = mock_object();
    mock_set(opts, "cols", NUM(COLS));
    mock_set(opts, "rows", NUM(ROWS));
    mock_set(opts, "rays_per_pixel", NUM(RPP));
    mock_set(opts, "focal_length", NUM(5.0));
    mock_set(opts, "lens_rad", NUM(0.05));
    napi_value render = CALL("render_create", ctx, scene, opts);
    CALL("render_set_seeds", render, mock_typedarray(napi_int32_array, TOTAL, seeds), NUM(TOTAL), NUM(0));
    napi_value pix = mock_typedarray(napi_uint8_clamped_array, COLS * ROWS * 4, NULL);
    for (int pass = 0; pass < 2; pass++) CALL("render_execute", render, f32(16, cam), pix);
    napi_value acc = mock_typedarray(napi_float32_array, COLS * ROWS * 4, NULL);
    CALL("render_read_accum", render, acc);
    napi_value sd = mock_typedarray(napi_int32_array, TOTAL, NULL);
    CALL("render_read_seeds", render, sd, NUM(TOTAL));
    napi_value st = CALL("render_stats", render);
    printf("a10 closest_rays %llu any_rays %llu launches %d\n", (unsigned long long)mock_get(st, "closest_rays")->big,
           (unsigned long long)mock_get(st, "any_rays")->big, (int)mock_get(st, "launches")->num);
    dump(dir, "a10_pixels.bin", pix->data, COLS * ROWS * 4);
    dump(dir, "a10_accum.bin", acc->data, COLS * ROWS * 16);
    dump(dir, "a10_seeds.bin", sd->data, TOTAL * 4);
    CALL("render_destroy", render);
    CALL("scene_destroy", scene);
    CALL("grid_release", ctx, gt);
    CALL("grid_release", ctx, gs);
    CALL("finish", ctx);
    CALL("ctx_destroy", ctx);
    printf("gpu ok\n");
    return 0;
}

int main(int argc, char** argv) {
    g_exports = mock_object();
    if (!rt2015_init(&g_env, g_exports) || g_env.pending) {
        fprintf(stderr, "rt2015_init failed: %s\n", g_env.message);
        return 1;
    }
    if (argc >= 2 && !strcmp(argv[1], "cpu")) return run_cpu();
    if (argc >= 3 && !strcmp(argv[1], "gpu")) return run_gpu(argv[2]);
    fprintf(stderr, "usage: addon_test cpu | gpu <outdir>\n");
    return 64;
}
