'use strict';
/* rt2015.js -- headless Node.js host layer over librt2015.so (SURVEY.md 8f rank 1).
 *
 * Keeps the surface of the reference's code.js (Assign10-Path_Tracing/code.js; citations below are relative to
 * /root/reference, A10 = Assign10-Path_Tracing, A07 = Assign07-3D_uniform_grid_acceleration): the mol/ tri/ scenes/
 * loaders, Bounds / Vec3 / Camera / Light, bounds2AABB, the split*Data grid builders, Mesh, and the
 * preRender / executeRender / postRender frame driver -- with the WebCL object model (A10/code.js:576-608,
 * 1047-1552) replaced by the C ABI of include/rt2015.h through the N-API addon rt2015_napi.c (rt2015.node).
 * JS Number arithmetic is kept in the reference's operation order wherever a value reaches the device
 * (camera, lights, bounds, mesh transforms), so the uploaded Float32Arrays are the ones the browser host made.
 *
 * NOT EXECUTED in the build image (it has no Node.js): the addon underneath is executed against an N-API stand-in
 * (host_node/test/, tests/test_node_addon.py); this file is the thin layer a maintainer would run on top.
 * The Python package 2015-raytracing_b200/host.py is the same layer in the language the image can run, and is the
 * one the parity tests drive.
 *
 *   const RT = require('./rt2015.js');
 *   const scene = RT.loadScene('scenes/cornell_teapot3.xml', 1920, 1080);
 *   const r = new RT.Renderer(scene, 1920, 1080, { raysPerPixel: 16 });
 *   r.preRender();                         // split*Data on the GPU, scene + render objects, seeds
 *   for (let p = 0; p < 64; p++) r.executeRender();      // one progressive pass each, returns the RGBA image
 *   RT.writePNG('out.png', r.image, 1920, 1080);
 *   r.postRender();
 */
const fs = require('fs');
const path = require('path');
const zlib = require('zlib');
const rt = require('./rt2015.node');

const MAX_VALUE = Number.MAX_VALUE;

// ------------------------------------------------------------------------------------------- small types
class Bounds {   // A10/lib/utilities.js:389-422
  constructor(min, max) {
    this.min = min ? [min[0], min[1], min[2]] : [MAX_VALUE, MAX_VALUE, MAX_VALUE];
    this.max = max ? [max[0], max[1], max[2]] : [-MAX_VALUE, -MAX_VALUE, -MAX_VALUE];
  }
  center() { return [(this.min[0] + this.max[0]) / 2, (this.min[1] + this.max[1]) / 2, (this.min[2] + this.max[2]) / 2]; }
  diagonal() {
    const d = [this.max[0] - this.min[0], this.max[1] - this.min[1], this.max[2] - this.min[2]];
    return Math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  }
  merge(b) {
    for (let a = 0; a < 3; a++) {
      this.min[a] = Math.min(this.min[a], b.min[a]);
      this.max[a] = Math.max(this.max[a], b.max[a]);
    }
  }
}

function bounds2AABB(bounds) {   // A10/code.js:610-621: pmin.xyz, 1, pmax.xyz, 1
  return new Float32Array([bounds.min[0], bounds.min[1], bounds.min[2], 1, bounds.max[0], bounds.max[1], bounds.max[2], 1]);
}

class Vec3 {   // A10/code.js:13-53
  constructor(x = 0, y = 0, z = 0) { this.x = x; this.y = y; this.z = z; }
  set(x, y, z) { this.x = x; this.y = y; this.z = z; }
  subtract(b) { return new Vec3(this.x - b.x, this.y - b.y, this.z - b.z); }
  cross(b) { return new Vec3(this.y * b.z - this.z * b.y, this.z * b.x - this.x * b.z, this.x * b.y - this.y * b.x); }
  normalize() {
    const n = Math.sqrt(this.x * this.x + this.y * this.y + this.z * this.z);
    this.x /= n; this.y /= n; this.z /= n;
  }
  toArray() { return [this.x, this.y, this.z]; }
}

class Camera {   // A10/code.js:175-277
  constructor() {
    this.eye = new Vec3(); this.U = new Vec3(); this.V = new Vec3(); this.W = new Vec3();
    this.width = 1; this.height = 1; this.cols = 0; this.rows = 0;
  }
  defaultInit() { this.eye = new Vec3(0, 0, 0); this.U = new Vec3(1, 0, 0); this.V = new Vec3(0, 1, 0); this.W = new Vec3(0, 0, 1); }
  _frustum(fov, cols, rows) {
    this.cols = cols; this.rows = rows;
    this.height = 2.0 * Math.tan(0.5 * fov * Math.PI / 180.0);
    this.width = this.height * (cols / rows);
  }
  set(bounds, cols, rows) {   // :185-201
    this._frustum(60, cols, rows);
    const c = bounds.center();
    this.eye.set(c[0], c[1], c[2] + bounds.diagonal());
  }
  lookAt(eye, lookat, vup, fov, cols, rows) {   // :203-217
    this._frustum(fov, cols, rows);
    this.eye = eye;
    this.W = eye.subtract(lookat); this.W.normalize();
    this.U = vup.cross(this.W); this.U.normalize();
    this.V = this.W.cross(this.U);
  }
  rotate(bounds, angle) {   // :219-245
    const c = bounds.center(), diag = bounds.diagonal(), rad = angle * Math.PI / 180.0;
    this.eye.set(c[0] + Math.sin(rad) * diag, c[1], c[2] + Math.cos(rad) * diag);
    this.W.set(this.eye.x - c[0], this.eye.y - c[1], this.eye.z - c[2]); this.W.normalize();
    this.U = this.V.cross(this.W);
  }
  toFloat32Array() {   // :250-258 -- the kernels' float16 camera
    return new Float32Array([...this.eye.toArray(), ...this.U.toArray(), ...this.V.toArray(), ...this.W.toArray(),
      this.width, this.height, this.cols, this.rows]);
  }
}

class Light {   // disk area light, A10/code.js:279-353
  constructor() {
    this.position = new Vec3(); this.normal = new Vec3(); this.T = new Vec3(); this.B = new Vec3(); this.irradiance = new Vec3();
    this.radius = 0; this.area = 0;
  }
  set(position, normal, irradiance, radius) {
    this.position = position; this.normal = normal; this.irradiance = irradiance; this.radius = radius;
    this.normal.normalize(); this.calculateArea(); this.calculateTBN();
  }
  calculateArea() { this.area = Math.PI * this.radius * this.radius; }
  calculateTBN() {   // :302-321: replace the smallest |component| by 1, orthonormalise
    const n = this.normal, m = [Math.abs(n.x), Math.abs(n.y), Math.abs(n.z)], v = new Vec3(n.x, n.y, n.z);
    const lo = Math.min(m[0], m[1], m[2]);
    if (lo === m[0]) v.x = 1; else if (lo === m[1]) v.y = 1; else v.z = 1;
    v.normalize();
    this.T = v.cross(n); this.T.normalize();
    this.B = n.cross(this.T); this.B.normalize();
  }
  _info(b, c, s) { return new Float32Array([...this.position.toArray(), ...b.toArray(), ...c.toArray(), s, 0, 0, 0, 0, 0, 0]); }
  toShadowInfo() { return this._info(this.T, this.B, this.radius); }                      // :323-331
  toSceneRenderInfo() { return this._info(this.normal, this.irradiance, this.area); }     // :333-342
  toLightRenderInfo() { return this._info(this.normal, this.irradiance, this.radius); }   // :344-352
}

// ------------------------------------------------------------------------------------------- loaders
// parseMeshJSON (A10/tri/meshDataVersion1.js:12-78) and parsePDB (A10/mol/pdbParserV1.js:2-85) through the library's
// native parsers: same numbers as the JavaScript loaders (gl-matrix's Float32Array rounding of transformed vertices,
// the sparse-serial `size` quirk), an order of magnitude faster on million-triangle files.
function parseMeshJSON(jsonFileNameOrBuffer) {
  const buf = Buffer.isBuffer(jsonFileNameOrBuffer) ? jsonFileNameOrBuffer : fs.readFileSync(jsonFileNameOrBuffer);
  const m = rt.parse_mesh_json(new Uint8Array(buf.buffer, buf.byteOffset, buf.length));
  return { nTriangles: m.nTriangles, nMaterials: m.nMaterials, materialIndices: m.materialIndices, materials: m.materials,
    bounds: new Bounds(m.boundsMin, m.boundsMax), positions: m.positions, normals: m.normals, tCoords: null };
}

function parsePDB(text) {
  const buf = Buffer.isBuffer(text) ? text : Buffer.from(text, 'latin1');
  const m = rt.parse_pdb(new Uint8Array(buf.buffer, buf.byteOffset, buf.length));
  return { size: m.size, atomData: m.atomData, colorData: m.colorData, radiusData: m.radiusData, bounds: new Bounds(m.boundsMin, m.boundsMax) };
}

// Minimal XML reader for the scene files (A10/code.js:706-721 uses DOMParser): elements and text only; comments,
// the UTF-8 BOM, <?...?> and <!...> are skipped, attributes ignored.  Returns { tag, text, children }.
function parseXML(src) {
  let s = src.charCodeAt(0) === 0xFEFF ? src.slice(1) : src;
  s = s.replace(/<!--[\s\S]*?-->/g, '').replace(/<\?[\s\S]*?\?>/g, '').replace(/<![^>]*>/g, '');
  const root = { tag: '#document', text: '', children: [] };
  const stack = [root];
  const re = /<(\/?)([A-Za-z_][\w.\-]*)([^>]*?)(\/?)>|([^<]+)/g;
  let m;
  while ((m = re.exec(s)) !== null) {
    const top = stack[stack.length - 1];
    if (m[5] !== undefined) { top.text += m[5]; continue; }
    if (m[1] === '/') {
      if (stack.length < 2 || top.tag !== m[2]) throw new Error('loadScene: mismatched </' + m[2] + '>');
      stack.pop();
    } else {
      const e = { tag: m[2], text: '', children: [] };
      top.children.push(e);
      if (m[4] !== '/') stack.push(e);
    }
  }
  if (stack.length !== 1) throw new Error('loadScene: unclosed <' + stack[stack.length - 1].tag + '>');
  return root;
}
function* descendants(e, name) {   // getElementsByTagName: descendant search in document order
  for (const c of e.children) {
    if (c.tag === name) yield c;
    yield* descendants(c, name);
  }
}
function first(e, name) {
  for (const c of descendants(e, name)) return c;
  throw new Error('loadScene: missing <' + name + '>');
}
const num = (e, name) => { const t = first(e, name).text.trim(); return t ? Number(t) : 0; };
const str = (e, name) => first(e, name).text;
const vec3 = (e, name) => { const v = first(e, name); return new Vec3(num(v, 'x'), num(v, 'y'), num(v, 'z')); };

// Mesh (A10/code.js:94-170).  normalize / scale / translate act on the cell-ordered positions AFTER the grid build, in
// float64, before the fp32 upload; the bounds are transformed alike.  Here the transform is recorded and handed to the
// GPU gather of rt_grid_build_triangles, which applies it at exactly that point.
class Mesh {
  constructor() {
    this.bounds = new Bounds(); this.ntriangles = 0; this.nslabs = 1; this.matId = 0; this.grid = null;
    this._jmesh = null; this._splitBounds = null;
    this._xform = { normalize: false, center: [0, 0, 0], maxdim: 1, scale: [1, 1, 1], translate: [0, 0, 0] };
  }
  loadFromJSON(jmesh, nslabs, matId) {
    this.bounds = new Bounds(jmesh.bounds.min, jmesh.bounds.max);
    this._splitBounds = new Bounds(jmesh.bounds.min, jmesh.bounds.max);
    this.ntriangles = jmesh.nTriangles; this.nslabs = nslabs | 0; this.matId = matId; this._jmesh = jmesh;
  }
  normalize() {   // :114-140
    const mn = this.bounds.min, mx = this.bounds.max;
    const c = [(mx[0] + mn[0]) / 2.0, (mx[1] + mn[1]) / 2.0, (mx[2] + mn[2]) / 2.0];
    const maxdim = 1.0 / Math.max(Math.max(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
    this._xform.normalize = true; this._xform.center = c; this._xform.maxdim = maxdim;
    this.bounds.min = mn.map((v, a) => (v - c[a]) * maxdim);
    this.bounds.max = mx.map((v, a) => (v - c[a]) * maxdim);
  }
  scale(s) {   // :142-155
    const f = s.toArray();
    this._xform.scale = f;
    this.bounds.min = this.bounds.min.map((v, a) => v * f[a]);
    this.bounds.max = this.bounds.max.map((v, a) => v * f[a]);
  }
  translate(t) {   // :157-169
    const f = t.toArray();
    this._xform.translate = f;
    this.bounds.min = this.bounds.min.map((v, a) => v + f[a]);
    this.bounds.max = this.bounds.max.map((v, a) => v + f[a]);
  }
  upload(ctx) {
    if (!this.grid) this.grid = splitMeshData(ctx, Object.assign({}, this._jmesh, { bounds: this._splitBounds }), this.nslabs, this._xform);
    return this.grid;
  }
}

// loadScene (A10/code.js:723-897).  `width` / `height` are the canvas globals; mesh files are resolved against the
// assignment directory (the page's base URL), i.e. the parent of scenes/.
function loadScene(sceneName, width, height, meshLoader) {
  const doc = parseXML(fs.readFileSync(sceneName, 'utf8'));
  const base = path.dirname(path.dirname(path.resolve(sceneName)));
  const xc = first(doc, 'camera');
  const camera = new Camera();
  camera.lookAt(vec3(xc, 'eye'), vec3(xc, 'lookAt'), vec3(xc, 'vup'), num(xc, 'fov'), width, height);
  const lights = [];
  for (const xl of descendants(doc, 'light')) {
    const lt = new Light();   // fields assigned directly: the light normal is NOT normalised (:751-757, quirk Q4)
    lt.position = vec3(xl, 'position'); lt.normal = vec3(xl, 'normal'); lt.irradiance = vec3(xl, 'irradiance');
    lt.radius = num(xl, 'radius');
    lt.calculateArea(); lt.calculateTBN();
    lights.push(lt);
  }
  const materials = [], lookup = {};
  for (const xm of descendants(doc, 'material')) {
    const col = first(xm, 'color');
    lookup[str(xm, 'id')] = materials.length;
    materials.push([num(col, 'r'), num(col, 'g'), num(col, 'b'), num(col, 'a')]);
  }
  const spheres = [], sphereBounds = new Bounds();
  for (const xs of descendants(doc, 'sphere')) {
    const c = vec3(xs, 'center'), r = num(xs, 'radius');
    spheres.push({ c, r, matId: lookup[str(xs, 'matId')] });
    sphereBounds.merge(new Bounds([c.x - r, c.y - r, c.z - r], [c.x + r, c.y + r, c.z + r]));
  }
  const triangles = [], triangleBounds = new Bounds();
  for (const xt of descendants(doc, 'triangle')) {
    const t = { matId: lookup[str(xt, 'matId')] };
    for (const k of ['p0', 'p1', 'p2', 'n0', 'n1', 'n2']) t[k] = vec3(xt, k);
    triangles.push(t);
    const p = [t.p0.toArray(), t.p1.toArray(), t.p2.toArray()];
    triangleBounds.merge(new Bounds([0, 1, 2].map(a => Math.min(Math.min(p[0][a], p[1][a]), p[2][a])),
      [0, 1, 2].map(a => Math.max(Math.max(p[0][a], p[1][a]), p[2][a]))));
  }
  for (let a = 0; a < 3; a++) {   // zero-thickness guard, :837-842
    if (triangleBounds.min[a] === triangleBounds.max[a]) { triangleBounds.min[a] -= 0.1; triangleBounds.max[a] += 0.1; }
  }
  const bounds = new Bounds(), meshes = [];
  for (const xm of descendants(doc, 'mesh')) {
    const file = str(xm, 'file');
    const jmesh = meshLoader ? meshLoader(file) : parseMeshJSON(path.join(base, file));
    const mesh = new Mesh();
    mesh.loadFromJSON(jmesh, num(xm, 'nslabs'), lookup[str(xm, 'matId')]);
    if (str(xm, 'normalize') === 'yes') mesh.normalize();
    mesh.scale(vec3(xm, 'scale'));
    mesh.translate(vec3(xm, 'translate'));
    meshes.push(mesh);
    bounds.merge(mesh.bounds);
  }
  bounds.merge(sphereBounds);   // empty sets merge their +-MAX_VALUE bounds as they are (:875-880)
  bounds.merge(triangleBounds);
  return { camera, focal_length: num(xc, 'focal_length'), lens_diameter: num(xc, 'lens_diameter'), lights, materials, bounds,
    spheres, sphereBounds, triangles, triangleBounds, meshes };
}

// ------------------------------------------------------------------------------------------- grid build (GPU)
// split*Data (A10/code.js:899-1041, 1554-1772; A07/code.js:889-978): count -> exclusive scan -> order-preserving
// scatter as integer CUDA kernels, float64 binning like the JavaScript; results are device buffers inside the grid.
const f64 = a => (a instanceof Float64Array ? a : Float64Array.from(a));
function buildTriangles(ctx, pos9, nor9, ids, bounds, nSlabs, xform) {
  const xf = xform ? Float64Array.from([xform.normalize ? 1 : 0, ...xform.center, xform.maxdim, ...xform.scale, ...xform.translate]) : null;
  const p = f64(pos9);
  return rt.grid_build_triangles(ctx, p, f64(nor9), ids ? Uint32Array.from(ids) : null, p.length / 9, f64(bounds.min), f64(bounds.max), nSlabs, xf);
}
function splitMeshData(ctx, meshData, nSlabs, xform) {
  return buildTriangles(ctx, meshData.positions, meshData.normals, meshData.materialIndices, meshData.bounds, nSlabs, xform || null);
}
function splitTriangleData(ctx, scene, nSlabs) {
  const pos9 = [], nor9 = [], ids = [];
  for (const t of scene.triangles) {
    pos9.push(...t.p0.toArray(), ...t.p1.toArray(), ...t.p2.toArray());
    nor9.push(...t.n0.toArray(), ...t.n1.toArray(), ...t.n2.toArray());
    ids.push(t.matId);
  }
  return buildTriangles(ctx, pos9, nor9, ids, scene.triangleBounds, nSlabs, null);
}
function splitSphereData(ctx, scene, nSlabs) {
  const xyzr = [], ids = [];
  for (const s of scene.spheres) { xyzr.push(s.c.x, s.c.y, s.c.z, s.r); ids.push(s.matId); }
  return rt.grid_build_spheres(ctx, f64(xyzr), Uint32Array.from(ids), ids.length, f64(scene.sphereBounds.min), f64(scene.sphereBounds.max), nSlabs);
}
function splitMolData(ctx, molData, nSlabs) {   // visits molData.size records; the ones past atomData are NaN spheres (quirk Q13)
  const n = molData.size, xyzr = new Float64Array(4 * n).fill(NaN), ids = new Uint32Array(n);
  const have = Math.min(n, molData.atomData.length / 4);
  for (let i = 0; i < have; i++) {
    const id = molData.atomData[4 * i];
    ids[i] = id;
    xyzr[4 * i] = molData.atomData[4 * i + 1]; xyzr[4 * i + 1] = molData.atomData[4 * i + 2]; xyzr[4 * i + 2] = molData.atomData[4 * i + 3];
    xyzr[4 * i + 3] = molData.radiusData[id];
  }
  return rt.grid_build_spheres(ctx, xyzr, ids, n, f64(molData.bounds.min), f64(molData.bounds.max), nSlabs);
}
function splitMaterialData(scene) {   // A10/code.js:1774-1782
  const out = new Float32Array(4 * scene.materials.length);
  scene.materials.forEach((m, i) => out.set(m, 4 * i));
  return out;
}

// ------------------------------------------------------------------------------------------- frame driver
// preRender / executeRender / postRender / render of A10/code.js:1784-1894.  One executeRender = one progressive pass
// (primary rays, lights, `depth` bounces with next-event estimation, accumulation, copyToPixel + read-back).
class Renderer {
  constructor(scene, width, height, opts = {}) {
    this.scene = scene; this.width = width | 0; this.height = height | 0;
    this.raysPerPixel = opts.raysPerPixel || 1;      // rays_per_pixel global; > 1 must be a perfect square
    this.nSlabs = opts.nSlabs || 1;                  // n_slabs global (1 in Assignment 10, A10/code.js:399)
    this.depth = opts.depth === undefined ? 5 : opts.depth;
    this.slots = opts.slots || [0, this.raysPerPixel];   // multi-GPU: this process renders slots [begin, begin+count) of every pixel
    this.mode = opts.mode || 0;
    this.ctx = opts.ctx || rt.ctx_create(opts.device || 0);
    this._ownCtx = !opts.ctx;
    this.hScene = null; this.hRender = null; this._grids = [];
    this.passes = 1;
    this.image = new Uint8ClampedArray(4 * this.width * this.height);   // imgData.data
  }
  preRender(seeds) {
    const sc = this.scene, ctx = this.ctx;
    this.hScene = rt.scene_create(ctx);
    rt.scene_set_bounds(this.hScene, bounds2AABB(sc.bounds));
    rt.scene_set_materials(this.hScene, splitMaterialData(sc), sc.materials.length);
    if (sc.spheres.length > 0) {   // geometry sets in the reference's trace order: spheres, scene triangles, meshes (:1809-1813)
      const g = splitSphereData(ctx, sc, this.nSlabs);
      this._grids.push(g);
      rt.scene_add_set(this.hScene, g, bounds2AABB(sc.sphereBounds), 0, 0);
    }
    if (sc.triangles.length > 0) {
      const g = splitTriangleData(ctx, sc, this.nSlabs);
      this._grids.push(g);
      rt.scene_add_set(this.hScene, g, bounds2AABB(sc.triangleBounds), 0, 0);
    }
    for (const mesh of sc.meshes) rt.scene_add_set(this.hScene, mesh.upload(ctx), bounds2AABB(mesh.bounds), 1, mesh.matId);
    for (const lt of sc.lights) rt.scene_add_light(this.hScene, lt.toShadowInfo(), lt.toSceneRenderInfo(), lt.toLightRenderInfo());
    this.hRender = rt.render_create(ctx, this.hScene, { cols: this.width, rows: this.height, rays_per_pixel: this.raysPerPixel,
      depth: this.depth, focal_length: Math.fround(sc.focal_length), lens_rad: Math.fround(sc.lens_diameter / 2.0),
      slot_begin: this.slots[0], slot_count: this.slots[1], mode: this.mode });
    this.setSeeds(seeds || Renderer.randomSeeds(this.width * this.height * this.raysPerPixel));
    this.passes = 1;
  }
  static randomSeeds(n) {   // prepareInitSeeds, A10/code.js:1140-1154: one seed in [1, 2^31 - 1] per ray slot
    const s = new Int32Array(n);
    for (let i = 0; i < n; i++) s[i] = 1 + Math.floor(Math.random() * 2147483646);
    return s;
  }
  setSeeds(seeds) { rt.render_set_seeds(this.hRender, seeds, seeds.length, 0); }
  // this renderer's own slots, [pixel][k_local]; nonBlocking = enqueueWriteBuffer(buf, false, ...): the copy overlaps the
  // head of the next executeRender, which keeps `seeds` alive until it has returned
  writeLocalSeeds(seeds, nonBlocking) {
    if (nonBlocking) { this._seedKeepAlive = seeds; rt.render_write_local_seeds_async(this.hRender, seeds, seeds.length); }
    else rt.render_write_local_seeds(this.hRender, seeds, seeds.length);
  }
  executeRender(camera) {   // + sendImagetoHTML (:1530-1537): this.image receives copyToPixel's RGBA
    rt.render_execute(this.hRender, (camera || this.scene.camera).toFloat32Array(), this.image);
    this._seedKeepAlive = null;
    this.passes++;
    return this.image;
  }
  accum() { const a = new Float32Array(4 * this.width * this.height); rt.render_read_accum(this.hRender, a); return a; }
  stats() { return rt.render_stats(this.hRender); }
  exportState() {   // checkpoint of what the reference keeps only on the device: acu, seeds, passes
    const n = this.width * this.height * this.slots[1];
    const acu = new Float32Array(4 * n), seeds = new Int32Array(n), passes = new Uint32Array(1);
    rt.render_export_state(this.hRender, acu, seeds, passes);
    return { acu, seeds, passes: passes[0] };
  }
  importState(st) { rt.render_import_state(this.hRender, st.acu, st.seeds, st.passes); this.passes = st.passes; }
  postRender() {   // releaseCLResources, LIFO (:1539-1552)
    if (this.hRender) { rt.render_destroy(this.hRender); this.hRender = null; }
    if (this.hScene) { rt.scene_destroy(this.hScene); this.hScene = null; }
    while (this._grids.length) rt.grid_release(this.ctx, this._grids.pop());
    for (const mesh of this.scene.meshes) if (mesh.grid) { rt.grid_release(this.ctx, mesh.grid); mesh.grid = null; }
    if (this._ownCtx) rt.ctx_destroy(this.ctx);
  }
  render(seeds) {   // :1883-1894
    this.preRender(seeds);
    try { return this.executeRender(); } finally { this.postRender(); }
  }
}

// compute / computeTri / computeBoth of Assignment 7 (A07/code.js:571-668), kernel by kernel through the launchers:
// initTrace against the (merged) bounds, then molTrace and/or meshTrace over one ray buffer.
function a07Compute(ctx, cols, rows, nSlabs, molData, meshData) {
  let bounds;
  if (molData && meshData) { bounds = new Bounds(); bounds.merge(molData.bounds); bounds.merge(meshData.bounds); } else bounds = (molData || meshData).bounds;
  const cam = new Camera();
  cam.defaultInit(); cam.set(bounds, cols, rows);
  const fcam = cam.toFloat32Array(), raySize = rt.struct_size('Ray', 7);
  const pix = rt.buffer_create(ctx, 4 * cols * rows), rays = rt.buffer_create(ctx, raySize * cols * rows);
  const release = [pix, rays], grids = [];
  try {
    rt.a07_initTrace(ctx, pix, fcam, rays, bounds2AABB(bounds));
    if (molData) {
      const g = splitMolData(ctx, molData, nSlabs);
      grids.push(g);
      const col = rt.buffer_create(ctx, 4 * molData.colorData.length);
      release.push(col);
      rt.buffer_write(ctx, col, Float32Array.from(molData.colorData));
      rt.a07_molTrace(ctx, pix, fcam, rays, molData.size, g.prim, g.matid, col, bounds2AABB(molData.bounds), nSlabs, g.box_size);
    }
    if (meshData) {
      const g = splitMeshData(ctx, meshData, nSlabs);
      grids.push(g);
      const col = rt.buffer_create(ctx, 4 * meshData.materials.length);
      release.push(col);
      rt.buffer_write(ctx, col, Float32Array.from(meshData.materials));
      rt.a07_meshTrace(ctx, pix, fcam, rays, meshData.nTriangles, g.prim, g.normal, g.matid, col, bounds2AABB(meshData.bounds), nSlabs, g.box_size);
    }
    const image = new Uint8ClampedArray(4 * cols * rows);
    rt.buffer_read(ctx, pix, image);   // enqueueReadBuffer + finish
    return image;
  } finally {
    while (grids.length) rt.grid_release(ctx, grids.pop());
    while (release.length) rt.buffer_release(ctx, release.pop());
  }
}

// canvas.putImageData for a headless host: 8-bit RGBA PNG
function writePNG(file, rgba, width, height) {
  const raw = Buffer.alloc((4 * width + 1) * height);
  for (let y = 0; y < height; y++) Buffer.from(rgba.buffer, rgba.byteOffset + 4 * width * y, 4 * width).copy(raw, (4 * width + 1) * y + 1);
  const crcTable = new Int32Array(256).map((_, n) => { let c = n; for (let k = 0; k < 8; k++) c = c & 1 ? 0xEDB88320 ^ (c >>> 1) : c >>> 1; return c; });
  const crc = b => { let c = -1; for (const v of b) c = crcTable[(c ^ v) & 255] ^ (c >>> 8); return (c ^ -1) >>> 0; };
  const chunk = (tag, data) => {
    const body = Buffer.concat([Buffer.from(tag, 'latin1'), data]), out = Buffer.alloc(body.length + 8);
    out.writeUInt32BE(data.length, 0); body.copy(out, 4); out.writeUInt32BE(crc(body), body.length + 4);
    return out;
  };
  const ihdr = Buffer.alloc(13);
  ihdr.writeUInt32BE(width, 0); ihdr.writeUInt32BE(height, 4); ihdr[8] = 8; ihdr[9] = 6;
  fs.writeFileSync(file, Buffer.concat([Buffer.from([0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A]), chunk('IHDR', ihdr),
    chunk('IDAT', zlib.deflateSync(raw)), chunk('IEND', Buffer.alloc(0))]));
}

module.exports = { addon: rt, Bounds, Vec3, Camera, Light, Mesh, Renderer, bounds2AABB, parseMeshJSON, parsePDB, parseXML, loadScene,
  splitMeshData, splitTriangleData, splitSphereData, splitMolData, splitMaterialData, a07Compute, writePNG };
