// rt_grid.cu -- uniform-grid build on the GPU: bin -> count -> exclusive scan -> scatter ->
// per-cell order restore -> gather.  Integer kernels, bit-exact against the reference's
// host-side JavaScript cell lists:
//   splitSphereData   Assign10-Path_Tracing/code.js:1554-1641
//   splitTriangleData Assign10-Path_Tracing/code.js:1643-1772
//   splitMeshData     Assign10-Path_Tracing/code.js:899-1041
//   splitMolData / splitMeshData  Assign07-3D_uniform_grid_acceleration/code.js:889-1122
// Contract reproduced: binning in float64 with floor((v - bmin) / box_width), low clamp on
// the min index only, high clamp on the max index only (so a primitive sitting exactly on the
// upper bound face can end up in no cell), NaN/inf make the range empty; a primitive goes to
// EVERY cell of its inclusive index box; inside a cell primitives keep input order; cells are
// emitted z-major, then y, then x; box_size is the exclusive prefix sum with n^3+1 entries.
#include "rt_internal.h"

namespace {

constexpr unsigned kBlock = 256;

struct CellBox { int lo[3]; int hi[3]; };   // empty when hi < lo on any axis

// JS: lo = Math.floor((vmin - bmin) / bw); if (lo < 0) lo = 0;   (no upper clamp)
//     hi = Math.floor((vmax - bmin) / bw); if (hi >= n) hi = n-1; (no lower clamp)
// for (i = lo; i <= hi; i++) -- NaN compares false, so NaN => empty range.
__device__ __forceinline__ void jsRange(double vmin, double vmax, double bmin, double bw, int n, int& lo_o, int& hi_o) {
    double lo = floor(__ddiv_rn(__dsub_rn(vmin, bmin), bw));
    double hi = floor(__ddiv_rn(__dsub_rn(vmax, bmin), bw));
    if (lo < 0) lo = 0;
    if (hi >= (double)n) hi = (double)(n - 1);
    if (!(lo <= hi)) { lo_o = 0; hi_o = -1; return; }   // covers NaN on either side
    // here 0 <= lo <= hi <= n-1 unless hi < 0 (then lo <= hi fails unless lo was clamped: lo=0 > hi) -- handled above
    lo_o = (int)lo;
    hi_o = (int)hi;
}

// kind 0: spheres, prim = (cx,cy,cz,r) -> box = c -+ r ;  kind 1: triangles, prim = 9 doubles.
// dims = 3: n x n x n cells; dims = 1: n slabs along x (Assignment 6), the y / z extents are not looked at.
__global__ void k_bin(const double* __restrict__ prim, unsigned n_prims, int kind, double bminx, double bminy, double bminz, double bwx,
                      double bwy, double bwz, int n, int dims, CellBox* __restrict__ boxes, unsigned* __restrict__ prim_refs) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prims) return;
    double mn[3], mx[3];
    if (kind == 0) {
        const double* p = prim + 4ull * i;
        double r = p[3];
        for (int a = 0; a < 3; a++) { mn[a] = __dsub_rn(p[a], r); mx[a] = __dadd_rn(p[a], r); }
    } else {
        const double* p = prim + 9ull * i;
        for (int a = 0; a < 3; a++) {
            // Math.min(Math.min(x0,x1),x2): NaN-propagating, unlike fmin
            double a0 = p[a], a1 = p[3 + a], a2 = p[6 + a];
            double m01 = (a0 != a0 || a1 != a1) ? (a0 + a1) : (a1 < a0 ? a1 : a0);
            mn[a] = (m01 != m01 || a2 != a2) ? (m01 + a2) : (a2 < m01 ? a2 : m01);
            double M01 = (a0 != a0 || a1 != a1) ? (a0 + a1) : (a1 > a0 ? a1 : a0);
            mx[a] = (M01 != M01 || a2 != a2) ? (M01 + a2) : (a2 > M01 ? a2 : M01);
        }
    }
    CellBox b;
    jsRange(mn[0], mx[0], bminx, bwx, n, b.lo[0], b.hi[0]);
    if (dims == 3) {
        jsRange(mn[1], mx[1], bminy, bwy, n, b.lo[1], b.hi[1]);
        jsRange(mn[2], mx[2], bminz, bwz, n, b.lo[2], b.hi[2]);
    } else {
        b.lo[1] = b.hi[1] = b.lo[2] = b.hi[2] = 0;
    }
    unsigned long long c = 1;
    for (int a = 0; a < 3; a++) c *= (b.hi[a] >= b.lo[a]) ? (unsigned long long)(b.hi[a] - b.lo[a] + 1) : 0ull;
    boxes[i] = b;
    prim_refs[i] = (unsigned)c;
}

__global__ void k_count(const CellBox* __restrict__ boxes, const unsigned* __restrict__ prim_refs, unsigned n_prims, unsigned n,
                        unsigned* __restrict__ cell_count) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prims || prim_refs[i] == 0) return;
    CellBox b = boxes[i];
    for (int z = b.lo[2]; z <= b.hi[2]; z++)
        for (int y = b.lo[1]; y <= b.hi[1]; y++)
            for (int x = b.lo[0]; x <= b.hi[0]; x++) atomicAdd(cell_count + ((size_t)z * n + y) * n + x, 1u);
}

// ---- exclusive scan (uint32), three-kernel hierarchical form -----------------------------
constexpr unsigned kScanBlock = 256;
constexpr unsigned kScanItems = 8;                          // per thread
constexpr unsigned kScanTile = kScanBlock * kScanItems;     // 2048 per block

__global__ void k_scan_tiles(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ tile_sums, size_t n) {
    __shared__ unsigned warp_sums[kScanBlock / 32];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned sum = 0;
#pragma unroll
    for (unsigned k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        sum += v[k];
    }
    // inclusive warp scan of per-thread sums
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = sum;
#pragma unroll
    for (unsigned d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = (lane < kScanBlock / 32) ? warp_sums[lane] : 0u;
        unsigned wi = w;
#pragma unroll
        for (unsigned d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < kScanBlock / 32) warp_sums[lane] = wi - w;   // exclusive warp offsets
        if (lane == kScanBlock / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    unsigned run = warp_sums[warp] + (incl - sum);
#pragma unroll
    for (unsigned k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}
__global__ void k_scan_add(unsigned* __restrict__ out, const unsigned* __restrict__ tile_offsets, size_t n) {
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    unsigned off = tile_offsets[blockIdx.x];
#pragma unroll
    for (unsigned k = 0; k < kScanItems; k++)
        if (base + k < n) out[base + k] += off;
}

// out[0..n) = exclusive scan of in[0..n); in and out may alias.  Scratch is allocated from
// the stream-ordered pool.
int exclusiveScan(rt_ctx* ctx, const unsigned* in, unsigned* out, size_t n) {
    if (n == 0) return RT_OK;
    size_t tiles = (n + kScanTile - 1) / kScanTile;
    unsigned* sums = nullptr;
    RT_CUDA(ctx, rt_scratch_alloc(ctx, (void**)&sums, sizeof(unsigned) * tiles));
    k_scan_tiles<<<(unsigned)tiles, kScanBlock, 0, ctx->stream>>>(in, out, sums, n);
    RT_LAUNCH_CHECK(ctx, "scan_tiles");
    if (tiles > 1) {
        int rc = exclusiveScan(ctx, sums, sums, tiles);
        if (rc) return rc;
        k_scan_add<<<(unsigned)tiles, kScanBlock, 0, ctx->stream>>>(out, sums, n);
        RT_LAUNCH_CHECK(ctx, "scan_add");
    }
    RT_CUDA(ctx, cudaFreeAsync(sums, ctx->stream));
    return RT_OK;
}

// box_size[cells] = total (closing entry) and the occupancy bitmap.
__global__ void k_finish_table(const unsigned* __restrict__ cell_count, unsigned* __restrict__ box_size, unsigned* __restrict__ occupancy,
                               size_t cells) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool occ = (c < cells) && cell_count[c] != 0;
    unsigned bits = __ballot_sync(0xffffffffu, occ);
    if ((threadIdx.x & 31) == 0 && (c >> 5) < (cells + 31) / 32) occupancy[c >> 5] = bits;
    if (c == cells - 1) box_size[cells] = box_size[c] + cell_count[c];
}

__global__ void k_scatter(const CellBox* __restrict__ boxes, const unsigned* __restrict__ prim_refs, unsigned n_prims, unsigned n,
                          const unsigned* __restrict__ box_size, unsigned* __restrict__ cursor, unsigned* __restrict__ ref_prim) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prims || prim_refs[i] == 0) return;
    CellBox b = boxes[i];
    for (int z = b.lo[2]; z <= b.hi[2]; z++)
        for (int y = b.lo[1]; y <= b.hi[1]; y++)
            for (int x = b.lo[0]; x <= b.hi[0]; x++) {
                size_t c = ((size_t)z * n + y) * n + x;
                unsigned pos = box_size[c] + atomicAdd(cursor + c, 1u);
                ref_prim[pos] = i;
            }
}

// Restore input order inside every cell: the atomic scatter filled each segment with the
// right SET of primitive ids in arbitrary order; ascending id == input order.
constexpr unsigned kSmallSeg = 48;      // thread-per-cell insertion sort up to here
constexpr unsigned kBlockSeg = 4096;    // block-per-cell bitonic sort in shared memory up to here

__global__ void k_sort_small(const unsigned* __restrict__ box_size, unsigned* __restrict__ ref_prim, size_t cells,
                             unsigned* __restrict__ big_cells, unsigned* __restrict__ n_big) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    unsigned b = box_size[c], e = box_size[c + 1];
    unsigned len = e - b;
    if (len < 2) return;
    if (len > kSmallSeg) {
        big_cells[atomicAdd(n_big, 1u)] = (unsigned)c;
        return;
    }
    for (unsigned i = b + 1; i < e; i++) {
        unsigned v = ref_prim[i];
        unsigned j = i;
        while (j > b && ref_prim[j - 1] > v) { ref_prim[j] = ref_prim[j - 1]; j--; }
        ref_prim[j] = v;
    }
}

// One block per big cell.  len <= kBlockSeg: bitonic sort in shared memory.  Larger: rebuild
// the segment as a stable compaction over ALL primitives (flag = "my box contains the cell",
// block-wide scan per chunk) -- O(n_prims) per such cell, there are at most n_refs/4096 of them.
__global__ void k_sort_big(const unsigned* __restrict__ big_cells, const unsigned* __restrict__ box_size, unsigned* __restrict__ ref_prim,
                           const CellBox* __restrict__ boxes, const unsigned* __restrict__ prim_refs, unsigned n_prims, unsigned n) {
    __shared__ unsigned sh[kBlockSeg];
    unsigned c = big_cells[blockIdx.x];
    unsigned b = box_size[c], e = box_size[c + 1];
    unsigned len = e - b;
    if (len <= kBlockSeg) {
        unsigned p2 = 1;
        while (p2 < len) p2 <<= 1;
        for (unsigned i = threadIdx.x; i < p2; i += blockDim.x) sh[i] = (i < len) ? ref_prim[b + i] : 0xFFFFFFFFu;
        __syncthreads();
        for (unsigned k = 2; k <= p2; k <<= 1)
            for (unsigned j = k >> 1; j > 0; j >>= 1) {
                for (unsigned i = threadIdx.x; i < p2; i += blockDim.x) {
                    unsigned l = i ^ j;
                    if (l > i) {
                        unsigned a = sh[i], d = sh[l];
                        bool up = ((i & k) == 0);
                        if ((a > d) == up) { sh[i] = d; sh[l] = a; }
                    }
                }
                __syncthreads();
            }
        for (unsigned i = threadIdx.x; i < len; i += blockDim.x) ref_prim[b + i] = sh[i];
        return;
    }
    int cx = (int)(c % n), cy = (int)((c / n) % n), cz = (int)(c / ((size_t)n * n));
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_prims; base += blockDim.x) {
        unsigned i = base + threadIdx.x;
        bool f = false;
        if (i < n_prims && prim_refs[i] != 0) {
            CellBox bx = boxes[i];
            f = cx >= bx.lo[0] && cx <= bx.hi[0] && cy >= bx.lo[1] && cy <= bx.hi[1] && cz >= bx.lo[2] && cz <= bx.hi[2];
        }
        unsigned bal = __ballot_sync(0xffffffffu, f);
        unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        unsigned off = running;
        for (unsigned w = 0; w < warp; w++) off += warp_tot[w];
        if (f) ref_prim[b + off + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (unsigned w = 0; w < (blockDim.x + 31) / 32; w++) t += warp_tot[w];
            running += t;
        }
        __syncthreads();
    }
}

struct XformArg {
    int enabled, do_normalize;
    double c[3], maxdim, s[3], t[3];
};

// Cell-ordered output buffers (what split*Data pushes, then `new Float32Array(...)` rounds).
__global__ void k_gather_spheres(const unsigned* __restrict__ ref_prim, unsigned n_refs, const double* __restrict__ xyzr,
                                 const unsigned* __restrict__ id, int square_radius, float4* __restrict__ out, unsigned* __restrict__ out_id) {
    unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_refs) return;
    unsigned i = ref_prim[r];
    const double* p = xyzr + 4ull * i;
    // w = rad*rad, A10/code.js:1602 (A07-A10); the plain radius for the 1-D slabs of A06 (A06/code.js:481)
    out[r] = make_float4((float)p[0], (float)p[1], (float)p[2], square_radius ? (float)__dmul_rn(p[3], p[3]) : (float)p[3]);
    if (out_id) out_id[r] = id ? id[i] : 0u;
}

__global__ void k_gather_triangles(const unsigned* __restrict__ ref_prim, unsigned n_refs, const double* __restrict__ pos9,
                                   const double* __restrict__ nor9, const unsigned* __restrict__ id, XformArg xf, float4* __restrict__ out_pos,
                                   float4* __restrict__ out_nor, unsigned* __restrict__ out_id) {
    unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (ref, vertex)
    if (t >= 3ull * n_refs) return;
    unsigned r = (unsigned)(t / 3), v = (unsigned)(t % 3);
    unsigned i = ref_prim[r];
    const double* p = pos9 + 9ull * i + 3 * v;
    double q[3] = {p[0], p[1], p[2]};
    if (xf.enabled) {   // Mesh.normalize / scale / translate, A10/code.js:114-169, applied in this order in float64
        for (int a = 0; a < 3; a++) {
            if (xf.do_normalize) q[a] = __dmul_rn(__dsub_rn(q[a], xf.c[a]), xf.maxdim);
            q[a] = __dmul_rn(q[a], xf.s[a]);
            q[a] = __dadd_rn(q[a], xf.t[a]);
        }
    }
    out_pos[t] = make_float4((float)q[0], (float)q[1], (float)q[2], 0.0f);
    if (out_nor) {
        const double* m = nor9 + 9ull * i + 3 * v;
        out_nor[t] = make_float4((float)m[0], (float)m[1], (float)m[2], 0.0f);
    }
    if (out_id && v == 0) out_id[r] = id ? id[i] : 0u;
}

struct Scratch {   // frees everything it owns (stream-ordered) on scope exit
    rt_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Scratch(rt_ctx* c) : ctx(c) {}
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
    }
    template <typename T>
    cudaError_t alloc(T** p, size_t count) {
        cudaError_t e = rt_scratch_alloc(ctx, (void**)p, sizeof(T) * (count ? count : 1));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

int buildGrid(rt_ctx* ctx, int kind, const double* prim_host, const double* nor_host, const unsigned* id_host, unsigned n_prims,
              const double bmin[3], const double bmax[3], unsigned n, const rt_mesh_xform* xform, rt_grid* out, int dims = 3) {
    RT_CHECK_CTX(ctx);
    if (!out || !bmin || !bmax || n == 0 || (n_prims && !prim_host)) return RT_ERR_INVALID;
    if (dims == 3 && (unsigned long long)n * n * n > 0x7FFFFFFFull) return rt_fail(ctx, RT_ERR_INVALID, "grid: n_slabs^3 too large");
    if (n > 0x7FFFFFFFu) return rt_fail(ctx, RT_ERR_INVALID, "grid: n_slabs too large");
    memset(out, 0, sizeof *out);
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cells = dims == 3 ? (size_t)n * n * n : (size_t)n;
    const unsigned per = kind == 0 ? 4 : 9;
    Scratch s(ctx);
    double *d_prim = nullptr, *d_nor = nullptr;
    unsigned *d_id = nullptr, *d_refs = nullptr, *d_count = nullptr, *d_cursor = nullptr, *d_refprim = nullptr, *d_big = nullptr, *d_nbig = nullptr;
    CellBox* d_boxes = nullptr;
    RT_CUDA(ctx, s.alloc(&d_prim, (size_t)per * n_prims));
    RT_CUDA(ctx, s.alloc(&d_boxes, n_prims));
    RT_CUDA(ctx, s.alloc(&d_refs, n_prims));
    RT_CUDA(ctx, s.alloc(&d_count, cells));
    RT_CUDA(ctx, s.alloc(&d_cursor, cells));
    RT_CUDA(ctx, s.alloc(&d_nbig, 1));
    if (n_prims) RT_CUDA(ctx, cudaMemcpyAsync(d_prim, prim_host, sizeof(double) * per * n_prims, cudaMemcpyHostToDevice, ctx->stream));
    if (nor_host && n_prims) {
        RT_CUDA(ctx, s.alloc(&d_nor, (size_t)9 * n_prims));
        RT_CUDA(ctx, cudaMemcpyAsync(d_nor, nor_host, sizeof(double) * 9 * n_prims, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (id_host && n_prims) {
        RT_CUDA(ctx, s.alloc(&d_id, n_prims));
        RT_CUDA(ctx, cudaMemcpyAsync(d_id, id_host, sizeof(unsigned) * n_prims, cudaMemcpyHostToDevice, ctx->stream));
    }
    RT_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned) * cells, ctx->stream));
    RT_CUDA(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(unsigned) * cells, ctx->stream));
    RT_CUDA(ctx, cudaMemsetAsync(d_nbig, 0, sizeof(unsigned), ctx->stream));

    // box_width = (bmax - bmin) / n in float64, exactly the JS expression (A10/code.js:907-909)
    double bw[3];
    for (int a = 0; a < 3; a++) bw[a] = (bmax[a] - bmin[a]) / (double)n;

    unsigned *box_size = nullptr, *occupancy = nullptr;
    RT_CUDA(ctx, cudaMalloc((void**)&box_size, sizeof(unsigned) * (cells + 1)));
    out->box_size = box_size;
    RT_CUDA(ctx, cudaMalloc((void**)&occupancy, sizeof(unsigned) * ((cells + 31) / 32)));
    out->occupancy = occupancy;
    out->n_slabs = n;
    out->kind = (unsigned)kind;

    if (n_prims) {
        k_bin<<<rt_blocks(n_prims, kBlock), kBlock, 0, ctx->stream>>>(d_prim, n_prims, kind, bmin[0], bmin[1], bmin[2], bw[0], bw[1], bw[2],
                                                                       (int)n, dims, d_boxes, d_refs);
        RT_LAUNCH_CHECK(ctx, "grid_bin");
        k_count<<<rt_blocks(n_prims, kBlock), kBlock, 0, ctx->stream>>>(d_boxes, d_refs, n_prims, n, d_count);
        RT_LAUNCH_CHECK(ctx, "grid_count");
    }
    int rc = exclusiveScan(ctx, d_count, box_size, cells);
    if (rc) return rc;
    k_finish_table<<<rt_blocks((cells + 31) / 32 * 32, kBlock), kBlock, 0, ctx->stream>>>(d_count, box_size, occupancy, cells);
    RT_LAUNCH_CHECK(ctx, "grid_finish_table");
    unsigned n_refs = 0;
    RT_CUDA(ctx, cudaMemcpyAsync(&n_refs, box_size + cells, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->n_refs = n_refs;

    RT_CUDA(ctx, s.alloc(&d_refprim, n_refs));
    RT_CUDA(ctx, s.alloc(&d_big, (size_t)n_refs / kSmallSeg + 1));
    if (n_refs) {
        k_scatter<<<rt_blocks(n_prims, kBlock), kBlock, 0, ctx->stream>>>(d_boxes, d_refs, n_prims, n, box_size, d_cursor, d_refprim);
        RT_LAUNCH_CHECK(ctx, "grid_scatter");
        k_sort_small<<<rt_blocks(cells, kBlock), kBlock, 0, ctx->stream>>>(box_size, d_refprim, cells, d_big, d_nbig);
        RT_LAUNCH_CHECK(ctx, "grid_sort_small");
        unsigned n_big = 0;
        RT_CUDA(ctx, cudaMemcpyAsync(&n_big, d_nbig, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (n_big) {
            k_sort_big<<<n_big, 1024, 0, ctx->stream>>>(d_big, box_size, d_refprim, d_boxes, d_refs, n_prims, n);
            RT_LAUNCH_CHECK(ctx, "grid_sort_big");
        }
    }

    size_t prim_bytes = (kind == 0 ? 16ull : 48ull) * n_refs;
    RT_CUDA(ctx, cudaMalloc(&out->prim, prim_bytes ? prim_bytes : 16));
    RT_CUDA(ctx, cudaMalloc(&out->matid, sizeof(unsigned) * (n_refs ? n_refs : 1)));
    if (kind == 1 && nor_host) RT_CUDA(ctx, cudaMalloc(&out->normal, prim_bytes ? prim_bytes : 16));
    if (n_refs) {
        if (kind == 0) {
            k_gather_spheres<<<rt_blocks(n_refs, kBlock), kBlock, 0, ctx->stream>>>(d_refprim, n_refs, d_prim, d_id, dims == 3 ? 1 : 0,
                                                                                     (float4*)out->prim, (unsigned*)out->matid);
        } else {
            XformArg xf;
            memset(&xf, 0, sizeof xf);
            if (xform) {
                xf.enabled = 1;
                xf.do_normalize = xform->do_normalize;
                xf.maxdim = xform->maxdim;
                for (int a = 0; a < 3; a++) { xf.c[a] = xform->center[a]; xf.s[a] = xform->scale[a]; xf.t[a] = xform->translate[a]; }
            }
            k_gather_triangles<<<rt_blocks(3ull * n_refs, kBlock), kBlock, 0, ctx->stream>>>(
                d_refprim, n_refs, d_prim, d_nor, d_id, xf, (float4*)out->prim, (float4*)out->normal, (unsigned*)out->matid);
        }
        RT_LAUNCH_CHECK(ctx, "grid_gather");
    }
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rt_ctx::GridAux aux;
    aux.box_size = out->box_size;
    aux.occupancy = (const unsigned*)out->occupancy;
    aux.prim = out->prim;
    aux.n_refs = out->n_refs;
    aux.n_slabs = out->n_slabs;
    aux.kind = out->kind;
    aux.dims = (unsigned)dims;
    if (rt_grid_wants_walker(aux)) {   // big grid: what the queue-walker route of molTrace / meshTrace needs, now rather than inside a frame
        int rc = rt_grid_aux_build(ctx, aux);
        if (rc) return rc;
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->grids.push_back(aux);
    return RT_OK;
}

}  // namespace

extern "C" {

int rt_grid_build_spheres(rt_ctx* ctx, const double* xyzr, const unsigned* id, unsigned n, const double bmin[3], const double bmax[3],
                          unsigned n_slabs, rt_grid* out) {
    return buildGrid(ctx, 0, xyzr, nullptr, id, n, bmin, bmax, n_slabs, nullptr, out);
}

int rt_grid_build_triangles(rt_ctx* ctx, const double* pos9, const double* nor9, const unsigned* id, unsigned n, const double bmin[3],
                            const double bmax[3], unsigned n_slabs, const rt_mesh_xform* xform, rt_grid* out) {
    return buildGrid(ctx, 1, pos9, nor9, id, n, bmin, bmax, n_slabs, xform, out);
}

int rt_slab_build_spheres(rt_ctx* ctx, const double* xyzr, const unsigned* id, unsigned n, double x_min, double x_max, unsigned n_slabs,
                          rt_grid* out) {
    const double bmin[3] = {x_min, 0.0, 0.0}, bmax[3] = {x_max, 0.0, 0.0};
    return buildGrid(ctx, 0, xyzr, nullptr, id, n, bmin, bmax, n_slabs, nullptr, out, 1);
}

int rt_slab_build_triangles(rt_ctx* ctx, const double* pos9, const double* nor9, const unsigned* id, unsigned n, double x_min, double x_max,
                            unsigned n_slabs, rt_grid* out) {
    const double bmin[3] = {x_min, 0.0, 0.0}, bmax[3] = {x_max, 0.0, 0.0};
    return buildGrid(ctx, 1, pos9, nor9, id, n, bmin, bmax, n_slabs, nullptr, out, 1);
}

int rt_grid_release(rt_ctx* ctx, rt_grid* g) {
    RT_CHECK_CTX(ctx);
    if (!g) return RT_ERR_INVALID;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->grids.size(); i++)
        if (ctx->grids[i].box_size == g->box_size) {
            cudaFree(ctx->grids[i].pre_ng);
            cudaFree(ctx->grids[i].pre_pe);
            cudaFree(ctx->grids[i].macro_occ);
            ctx->grids.erase(ctx->grids.begin() + i);
            break;
        }
    cudaFree(g->prim);
    cudaFree(g->normal);
    cudaFree(g->matid);
    cudaFree(g->box_size);
    cudaFree(g->occupancy);
    memset(g, 0, sizeof *g);
    return RT_OK;
}

}  // extern "C"
