// rt_internal.h -- host-side internals shared by the translation units of librt2015.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rt2015.h"

struct rt_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string last_error;
    // optional per-work-item walk statistics (device pointers, may be null)
    unsigned* st_hit = nullptr;
    unsigned* st_cells = nullptr;
    unsigned* st_tests = nullptr;
    unsigned long long* st_totals = nullptr;   // rt_set_walk_totals: 8 counters accumulated by every grid-walk launch
    unsigned long long launches = 0;   // kernels launched through this context
    // scratch allocations (grid build, rpp == 1 launcher) come from a pool of the context's own that KEEPS its memory
    // between calls: the device's default pool releases everything at each synchronisation, which made a per-pass
    // 16 MB cudaMallocAsync cost 27 ms
    cudaMemPool_t pool = nullptr;
    // The grids built through this context, keyed by the cell table the launchers receive.  The launchers keep the reference
    // kernels' argument lists (which only know the cell table); for a table of ours they find the occupancy bits here (skip the
    // two table loads of every EMPTY cell a ray crosses -- same walk, same floats) and, for the Assignment-7 traces, what the
    // queue walker needs (face vectors / edge form, coarse occupancy), built on first use.
    struct GridAux {
        const void* box_size = nullptr;
        const unsigned* occupancy = nullptr;
        const void* prim = nullptr;
        unsigned n_refs = 0, n_slabs = 0, kind = 0, dims = 3;
        float4* pre_ng = nullptr;
        float4* pre_pe = nullptr;
        unsigned* macro_occ = nullptr;
        unsigned macro_shift = 0, macro_n = 0;
        bool aux_ready = false;
    };
    std::vector<GridAux> grids;
};

static inline rt_ctx::GridAux* rt_grid_aux_of(rt_ctx* ctx, const void* box_size) {
    for (auto& g : ctx->grids)
        if (g.box_size == box_size) return &g;
    return nullptr;
}
static inline const unsigned* rt_occupancy_of(rt_ctx* ctx, const void* box_size) {
    rt_ctx::GridAux* g = rt_grid_aux_of(ctx, box_size);
    return g ? g->occupancy : nullptr;
}

// rt_frame.cu: face vectors / edge form / coarse occupancy of a registered multi-cell grid (for the queue-walker route of the
// Assignment-7 launchers); built with the grid when it is big enough to take that route.
int rt_grid_aux_build(rt_ctx* ctx, rt_ctx::GridAux& g);
static inline bool rt_grid_wants_walker(const rt_ctx::GridAux& g) {
    return g.dims == 3 && g.occupancy && g.n_slabs >= 32 && g.n_slabs <= 1023 && g.n_refs >= 65536;
}

static inline cudaError_t rt_scratch_alloc(rt_ctx* ctx, void** p, size_t bytes) {
    return ctx->pool ? cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream) : cudaMallocAsync(p, bytes, ctx->stream);
}

#define RT_CHECK_CTX(ctx)              \
    do {                               \
        if (!(ctx)) return RT_ERR_INVALID; \
    } while (0)

static inline int rt_fail(rt_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        ctx->last_error = buf;
    }
    return code;
}

#define RT_CUDA(ctx, call)                                         \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return rt_fail(ctx, RT_ERR_CUDA, #call, _e); \
    } while (0)

// Check the launch that was just issued.
#define RT_LAUNCH_CHECK(ctx, name)                                 \
    do {                                                           \
        (ctx)->launches++;                                         \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return rt_fail(ctx, RT_ERR_CUDA, name, _e); \
    } while (0)

static inline unsigned rt_blocks(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }
