// rt_loaders.cu -- native fast paths for the two on-disk formats on the step before the grid build
// (host code only; it lives in the same library so that a host needs one .so):
//   rt_parse_mesh_json  = parseMeshJSON, Assign10-Path_Tracing/tri/meshDataVersion1.js:12-78 -- assimp-style
//                         JSON (materials / meshes / nodes) -> per-triangle positions, normals, material index,
//                         bounds, with gl-matrix 2.2.1's Float32Array rounding points (lib/gl-matrix.js:79-80:
//                         matrices, the normal matrix and every transformed vector are rounded to fp32 on store,
//                         the arithmetic itself is double, left to right: :1063-1085, 2723-2760)
//   rt_parse_pdb        = parsePDB, mol/pdbParserV1.js:2-85 -- ATOM/HETATM records -> (element index, x, y, z),
//                         element colour / van-der-Waals radius tables, bounds, size = largest serial (quirk Q13)
// At 1 M triangles the JSON text is ~150 MB and JSON.parse / json.load dominates preRender; this parser is a
// single pass over the text that materialises only the five arrays the renderer needs.  Numbers go through
// strtod (correctly rounded, like JavaScript's number parsing).
#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rt2015.h"

namespace {

// ------------------------------------------------------------------------------- minimal JSON reader
struct Json {
    const char* p;
    const char* end;
    bool ok = true;

    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    void fail() { ok = false; p = end; }

    bool string(std::string* out) {
        ws();
        if (p >= end || *p != '"') { fail(); return false; }
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                if (out) out->push_back(p[1]);   // keys of this schema never contain escapes; values are skipped
                p += 2;
            } else {
                if (out) out->push_back(*p);
                p++;
            }
        }
        if (p >= end) { fail(); return false; }
        p++;
        return true;
    }
    // JSON number -> double, correctly rounded.  Fast path (Clinger): a decimal significand below 2^53 scaled by a
    // power of ten up to 10^22 is ONE exactly-rounded double operation on exactly representable operands; anything
    // else (long significands, big exponents) goes to strtod.
    bool number(double* out) {
        ws();
        if (p >= end) { fail(); return false; }
        const char* q = p;
        bool negative = false;
        if (*q == '-') { negative = true; q++; }
        unsigned long long mant = 0;
        int digits = 0, exp10 = 0;
        bool any = false;
        while (q < end && *q >= '0' && *q <= '9') { if (digits < 19) { mant = mant * 10 + (unsigned)(*q - '0'); digits += (mant != 0); } else exp10++; q++; any = true; }
        if (q < end && *q == '.') {
            q++;
            while (q < end && *q >= '0' && *q <= '9') { if (digits < 19) { mant = mant * 10 + (unsigned)(*q - '0'); digits += (mant != 0); exp10--; } q++; any = true; }
        }
        bool fast = any;
        if (q < end && (*q == 'e' || *q == 'E')) {
            const char* r = q + 1;
            bool eneg = false;
            if (r < end && (*r == '+' || *r == '-')) { eneg = *r == '-'; r++; }
            int ev = 0, ed = 0;
            while (r < end && *r >= '0' && *r <= '9' && ed < 6) { ev = ev * 10 + (*r - '0'); r++; ed++; }
            if (ed == 0 || (r < end && *r >= '0' && *r <= '9')) fast = false; else { exp10 += eneg ? -ev : ev; q = r; }
        }
        static const double kPow10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20,
                                        1e21, 1e22};
        if (fast && mant < (1ULL << 53) && exp10 >= -22 && exp10 <= 22) {
            double v = (double)mant;
            v = exp10 < 0 ? v / kPow10[-exp10] : v * kPow10[exp10];
            p = q;
            if (out) *out = negative ? -v : v;
            return true;
        }
        // slow path: strtod on a NUL-terminated copy of the token (the text need not be terminated, and strtod must
        // not read past `end`)
        const char* t = p;
        while (t < end && ((*t >= '0' && *t <= '9') || *t == '-' || *t == '+' || *t == '.' || *t == 'e' || *t == 'E')) t++;
        std::string tok(p, (size_t)(t - p));
        char* e = nullptr;
        double v = strtod(tok.c_str(), &e);
        if (e == tok.c_str()) { fail(); return false; }
        p += e - tok.c_str();
        if (out) *out = v;
        return true;
    }
    bool word(const char* w) {   // literal at p, bounded by end
        size_t n = strlen(w);
        if ((size_t)(end - p) < n || memcmp(p, w, n) != 0) return false;
        p += n;
        return true;
    }
    static constexpr int kMaxDepth = 256;   // nesting cap of skipped values (hostile input must not overflow the stack)
    void skip(int depth = 0) {   // any value
        ws();
        if (p >= end || depth > kMaxDepth) { fail(); return; }
        if (*p == '"') { string(nullptr); return; }
        if (*p == '{') {
            p++;
            if (eat('}')) return;
            do { string(nullptr); if (!eat(':')) { fail(); return; } skip(depth + 1); } while (ok && eat(','));
            if (!eat('}')) fail();
            return;
        }
        if (*p == '[') {
            p++;
            if (eat(']')) return;
            do { skip(depth + 1); } while (ok && eat(','));
            if (!eat(']')) fail();
            return;
        }
        if (word("true") || word("false") || word("null")) return;
        number(nullptr);
    }
    bool numbers(std::vector<double>& v) {   // [n, n, ...]
        if (!eat('[')) { fail(); return false; }
        if (eat(']')) return true;
        do { double x; if (!number(&x)) return false; v.push_back(x); } while (eat(','));
        if (!eat(']')) { fail(); return false; }
        return true;
    }
    // for each "key": value of an object calls f(key) which must consume the value
    template <class F> bool object(F f) {
        if (!eat('{')) { fail(); return false; }
        if (eat('}')) return true;
        do {
            std::string k;
            if (!string(&k) || !eat(':')) { fail(); return false; }
            f(k);
        } while (ok && eat(','));
        if (!eat('}')) { fail(); return false; }
        return ok;
    }
    template <class F> bool array(F f) {
        if (!eat('[')) { fail(); return false; }
        if (eat(']')) return true;
        do { f(); } while (ok && eat(','));
        if (!eat(']')) { fail(); return false; }
        return ok;
    }
};

struct JMesh { std::vector<double> pos, nor, idx; bool has_idx = false; double material = 0; };
struct JNode { std::vector<double> matrix, meshes; };

inline double f32(double v) { return (double)(float)v; }   // Float32Array store

// A JSON number used as an array index: a finite non-negative integer below `limit` (JavaScript would read
// `undefined` for anything else and the loader would throw a few lines later).
inline bool asIndex(double v, size_t limit, size_t* out) {
    if (!(v >= 0.0) || v >= 9007199254740992.0 || v != floor(v)) return false;
    size_t i = (size_t)v;
    if (i >= limit) return false;
    *out = i;
    return true;
}

// mat3.normalFromMat4 (lib/gl-matrix.js:2723-2760) on an fp32-valued matrix, result rounded to fp32
bool normalFromMat4(const double* a, double* out) {
    double a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3], a10 = a[4], a11 = a[5], a12 = a[6], a13 = a[7], a20 = a[8], a21 = a[9], a22 = a[10],
           a23 = a[11], a30 = a[12], a31 = a[13], a32 = a[14], a33 = a[15];
    double b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10, b03 = a01 * a12 - a02 * a11,
           b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12, b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30,
           b08 = a20 * a33 - a23 * a30, b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    double det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    if (!det) return false;
    det = 1.0 / det;
    out[0] = f32((a11 * b11 - a12 * b10 + a13 * b09) * det);
    out[1] = f32((a12 * b08 - a10 * b11 - a13 * b07) * det);
    out[2] = f32((a10 * b10 - a11 * b08 + a13 * b06) * det);
    out[3] = f32((a02 * b10 - a01 * b11 - a03 * b09) * det);
    out[4] = f32((a00 * b11 - a02 * b08 + a03 * b07) * det);
    out[5] = f32((a01 * b08 - a00 * b10 - a03 * b06) * det);
    out[6] = f32((a31 * b05 - a32 * b04 + a33 * b03) * det);
    out[7] = f32((a32 * b02 - a30 * b05 - a33 * b01) * det);
    out[8] = f32((a30 * b04 - a31 * b02 + a33 * b00) * det);
    return true;
}

template <class T> T* dup(const std::vector<T>& v) {
    T* p = (T*)malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

const double kMax = 1.7976931348623157e308;   // Number.MAX_VALUE (Bounds, lib/utilities.js:389-395)

}  // namespace

extern "C" {

int rt_parse_mesh_json(const char* text, size_t len, rt_mesh_data* out) {
    if (!text || !out) return RT_ERR_INVALID;
    memset(out, 0, sizeof *out);
    if (len >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) { text += 3; len -= 3; }
    Json j{text, text + len};
    std::vector<JMesh> meshes;
    std::vector<JNode> nodes;
    std::vector<double> materials;
    bool have_nodes = false;
    j.object([&](const std::string& k) {
        if (k == "materials") {
            j.array([&] {
                std::vector<double> dr;
                j.object([&](const std::string& mk) { if (mk == "diffuseReflectance") j.numbers(dr); else j.skip(); });
                for (int c = 0; c < 4; c++) materials.push_back(c < (int)dr.size() ? dr[c] : NAN);
            });
        } else if (k == "meshes") {
            j.array([&] {
                meshes.emplace_back();
                JMesh& m = meshes.back();
                j.object([&](const std::string& mk) {
                    if (mk == "vertexPositions") j.numbers(m.pos);
                    else if (mk == "vertexNormals") j.numbers(m.nor);
                    else if (mk == "indices") { j.numbers(m.idx); m.has_idx = !m.idx.empty(); }
                    else if (mk == "materialIndex") j.number(&m.material);
                    else j.skip();
                });
            });
        } else if (k == "nodes") {
            have_nodes = true;
            j.array([&] {
                nodes.emplace_back();
                JNode& n = nodes.back();
                j.object([&](const std::string& nk) {
                    if (nk == "modelMatrix") j.numbers(n.matrix);
                    else if (nk == "meshIndices") j.numbers(n.meshes);
                    else j.skip();
                });
            });
        } else {
            j.skip();
        }
    });
    if (!j.ok) return RT_ERR_INVALID;
    std::vector<double> P, N;
    std::vector<unsigned> M;
    double bmin[3] = {kMax, kMax, kMax}, bmax[3] = {-kMax, -kMax, -kMax};
    size_t n_nodes = have_nodes ? nodes.size() : 1;
    for (size_t k = 0; k < n_nodes; k++) {
        double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, nm[9];
        if (have_nodes) {
            if (nodes[k].matrix.size() < 16) return RT_ERR_INVALID;
            for (int i = 0; i < 16; i++) m[i] = f32(nodes[k].matrix[i]);   // mat4.copy into a Float32Array
        }
        if (!normalFromMat4(m, nm)) return RT_ERR_INVALID;   // the reference dereferences the null result and throws
        size_t n_meshes = have_nodes ? nodes[k].meshes.size() : meshes.size();
        for (size_t q = 0; q < n_meshes; q++) {
            size_t index = q;
            if (have_nodes && !asIndex(nodes[k].meshes[q], meshes.size(), &index)) return RT_ERR_INVALID;
            if (index >= meshes.size()) return RT_ERR_INVALID;
            const JMesh& mesh = meshes[index];
            const std::vector<double>& vp = mesh.pos;
            const std::vector<double>& vn = mesh.nor;
            for (size_t i = 0; i + 2 < vp.size(); i += 3) {   // bounds over ALL vertices of the mesh, referenced or not
                double x = vp[i], y = vp[i + 1], z = vp[i + 2];
                double v[3] = {f32(m[0] * x + m[4] * y + m[8] * z + m[12]), f32(m[1] * x + m[5] * y + m[9] * z + m[13]),
                               f32(m[2] * x + m[6] * y + m[10] * z + m[14])};
                for (int a = 0; a < 3; a++) { if (v[a] < bmin[a]) bmin[a] = v[a]; if (v[a] > bmax[a]) bmax[a] = v[a]; }
            }
            size_t mi = 0;
            if (!asIndex(mesh.material, 0xFFFFFFFFull, &mi)) return RT_ERR_INVALID;
            const unsigned material = (unsigned)mi;
            size_t nV = mesh.has_idx ? mesh.idx.size() : vp.size() / 3;
            size_t nT = nV / 3;
            for (size_t i = 0; i < nT; i++) {
                for (int c = 0; c < 3; c++) {
                    size_t vi = i * 3 + c;
                    if (mesh.has_idx && !asIndex(mesh.idx[vi], vp.size() / 3, &vi)) return RT_ERR_INVALID;
                    if (vi * 3 + 2 >= vp.size() || vi * 3 + 2 >= vn.size()) return RT_ERR_INVALID;
                    double x = vp[vi * 3], y = vp[vi * 3 + 1], z = vp[vi * 3 + 2];
                    P.push_back(f32(m[0] * x + m[4] * y + m[8] * z + m[12]));
                    P.push_back(f32(m[1] * x + m[5] * y + m[9] * z + m[13]));
                    P.push_back(f32(m[2] * x + m[6] * y + m[10] * z + m[14]));
                    x = vn[vi * 3]; y = vn[vi * 3 + 1]; z = vn[vi * 3 + 2];
                    N.push_back(f32(x * nm[0] + y * nm[3] + z * nm[6]));
                    N.push_back(f32(x * nm[1] + y * nm[4] + z * nm[7]));
                    N.push_back(f32(x * nm[2] + y * nm[5] + z * nm[8]));
                }
                M.push_back(material);
            }
        }
    }
    out->n_triangles = (unsigned)M.size();
    out->n_materials = (unsigned)(materials.size() / 4);
    out->positions = dup(P);
    out->normals = dup(N);
    out->material_indices = dup(M);
    out->materials = dup(materials);
    memcpy(out->bounds_min, bmin, sizeof bmin);
    memcpy(out->bounds_max, bmax, sizeof bmax);
    if (!out->positions || !out->normals || !out->material_indices || !out->materials) { rt_mesh_data_free(out); return RT_ERR_NOMEM; }
    return RT_OK;
}

void rt_mesh_data_free(rt_mesh_data* d) {
    if (!d) return;
    free(d->positions); free(d->normals); free(d->material_indices); free(d->materials);
    memset(d, 0, sizeof *d);
}

// ------------------------------------------------------------------------------- PDB
namespace {
struct Elem { const char* name; unsigned color; double radius; };
// colours (pdbParserV1.js:3-5) and van-der-Waals radii (:7-9); an element missing from a table makes the reference
// produce undefined/NaN entries -- those files are not valid inputs and are rejected here
const Elem kColors[] = {{"H", 0xCCCCCC, 0}, {"C", 0xAAAAAA, 0}, {"O", 0xCC0000, 0}, {"N", 0x0000CC, 0}, {"S", 0xCCCC00, 0}, {"P", 0x6622CC, 0},
                        {"F", 0x00CC00, 0}, {"CL", 0x00CC00, 0}, {"BR", 0x882200, 0}, {"I", 0x6600AA, 0}, {"FE", 0xCC6600, 0}, {"CA", 0x8888AA, 0}};
const Elem kRadii[] = {{"H", 0, 1.2}, {"Li", 0, 1.82}, {"Na", 0, 2.27}, {"K", 0, 2.75}, {"C", 0, 1.7}, {"N", 0, 1.55}, {"O", 0, 1.52}, {"F", 0, 1.47},
                       {"P", 0, 1.80}, {"S", 0, 1.80}, {"CL", 0, 1.75}, {"BR", 0, 1.85}, {"SE", 0, 1.90}, {"ZN", 0, 1.39}, {"CU", 0, 1.4}, {"NI", 0, 1.63}};

std::string strip_spaces(const char* s, size_t n) {
    std::string r;
    for (size_t i = 0; i < n; i++) if (s[i] != ' ') r.push_back(s[i]);
    return r;
}
// parseFloat(line.substr(a, n)): longest numeric prefix after leading white space, NaN if none
double parse_float(const char* s, size_t n) {
    std::string t(s, n);
    char* e = nullptr;
    const char* b = t.c_str();
    while (*b == ' ' || *b == '\t') b++;
    if (!((*b >= '0' && *b <= '9') || *b == '+' || *b == '-' || *b == '.')) return NAN;
    double v = strtod(b, &e);
    return e == b ? NAN : v;
}
}  // namespace

int rt_parse_pdb(const char* text, size_t len, rt_mol_data* out) {
    if (!text || !out) return RT_ERR_INVALID;
    memset(out, 0, sizeof *out);
    struct Atom { long serial; std::string elem; double x, y, z; };
    std::vector<Atom> atoms;
    long size = 0;
    const char* p = text;
    const char* end = text + len;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        const char* s = p;
        while (s < le && (*s == ' ' || *s == '\t' || *s == '\r')) s++;   // line.replace(/^\s*/, '')
        size_t n = (size_t)(le - s);
        while (n && s[n - 1] == '\r') n--;
        p = nl ? nl + 1 : end;
        if (n < 6 || (strncmp(s, "ATOM  ", 6) && strncmp(s, "HETATM", 6))) continue;
        char pad[81];
        memset(pad, ' ', sizeof pad);
        memcpy(pad, s, n < 80 ? n : 80);
        pad[80] = 0;
        if (n > 16 && pad[16] != ' ' && pad[16] != 'A') continue;   // altLoc
        if (n <= 16) { /* substr past the end is '' -> altLoc '' is neither ' ' nor 'A' in JS; such lines carry no coordinates anyway */ continue; }
        long serial = strtol(std::string(pad + 6, 5).c_str(), nullptr, 10);
        std::string elem = n > 76 ? strip_spaces(pad + 76, n >= 78 ? 2 : n - 76) : std::string();
        if (elem.empty()) elem = strip_spaces(pad + 12, 4);
        atoms.push_back({serial, elem, parse_float(pad + 30, 8), parse_float(pad + 38, 8), parse_float(pad + 46, 8)});
        if (serial > size) size = serial;
    }
    // atoms[serial - 1] = ...: later records with the same serial overwrite earlier ones; output in index order
    std::vector<long> order(atoms.size());
    for (size_t i = 0; i < atoms.size(); i++) order[i] = (long)i;
    std::vector<long> last((size_t)(size > 0 ? size : 0), -1);
    for (size_t i = 0; i < atoms.size(); i++) if (atoms[i].serial >= 1) last[(size_t)atoms[i].serial - 1] = (long)i;
    std::vector<std::string> used;
    std::vector<double> atomData, colorData, radiusData;
    double lo[3] = {kMax, kMax, kMax}, hi[3] = {-kMax, -kMax, -kMax};
    for (size_t si = 0; si < last.size(); si++) {
        if (last[si] < 0) continue;
        const Atom& a = atoms[(size_t)last[si]];
        size_t e = 0;
        for (; e < used.size(); e++) if (used[e] == a.elem) break;
        if (e == used.size()) {
            const Elem* c = nullptr; const Elem* r = nullptr;
            for (const Elem& k : kColors) if (a.elem == k.name) c = &k;
            for (const Elem& k : kRadii) if (a.elem == k.name) r = &k;
            if (!c || !r) return RT_ERR_INVALID;
            used.push_back(a.elem);
            colorData.push_back(((c->color >> 16) & 255) / 255.0);
            colorData.push_back(((c->color >> 8) & 255) / 255.0);
            colorData.push_back((c->color & 255) / 255.0);
            colorData.push_back(1.0);
            radiusData.push_back(r->radius);
        }
        double R = radiusData[e];
        atomData.push_back((double)e); atomData.push_back(a.x); atomData.push_back(a.y); atomData.push_back(a.z);
        const double v[3] = {a.x, a.y, a.z};
        for (int k = 0; k < 3; k++) { if (v[k] - R < lo[k]) lo[k] = v[k] - R; if (v[k] + R > hi[k]) hi[k] = v[k] + R; }
    }
    out->size = (unsigned)size;
    out->n_records = (unsigned)(atomData.size() / 4);
    out->n_elements = (unsigned)used.size();
    out->atom_data = dup(atomData);
    out->color_data = dup(colorData);
    out->radius_data = dup(radiusData);
    memcpy(out->bounds_min, lo, sizeof lo);
    memcpy(out->bounds_max, hi, sizeof hi);
    if (!out->atom_data || !out->color_data || !out->radius_data) { rt_mol_data_free(out); return RT_ERR_NOMEM; }
    return RT_OK;
}

void rt_mol_data_free(rt_mol_data* d) {
    if (!d) return;
    free(d->atom_data); free(d->color_data); free(d->radius_data);
    memset(d, 0, sizeof *d);
}

}  // extern "C"
