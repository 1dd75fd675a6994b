// rt_comm.cu -- the ONE exchange step of a multi-GPU render: every rank renders its own
// range of each pixel's ray slots, then the per-pixel accumulation images (float4 per pixel) are summed
// into the root rank with a single ncclReduce over NVLink / NVSwitch, issued on the context's own stream
// right behind the pass that produced the image -- no torch, no second stream, no host round trip.
// The reference has no counterpart (one device, one in-order queue, A10/code.js:592).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a single-GPU host never needs it, and a process
// that already carries an NCCL (e.g. PyTorch's bundled one) gets THAT instance instead of a second copy.
#include <dlfcn.h>
#include <nccl.h>

#include "rt_frame.h"

struct rt_comm {
    rt_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
};

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.handle || !api.error.empty()) return api;
    const char* names[] = {getenv("RT2015_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        const char* why = dlerror();   // one call: dlerror() clears the message it returns
        api.error = std::string("NCCL not found (libnccl.so.2): ") + (why ? why : "dlopen failed");
        return api;
    }
    auto sym = [&](const char* s) -> void* {
        void* p = dlsym(api.handle, s);
        if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    return api;
}

int ncclFail(rt_ctx* ctx, const char* what, ncclResult_t e) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(e) : "NCCL error");
    return rt_fail(ctx, RT_ERR_CUDA, buf);
}

}  // namespace

extern "C" {

static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "rt_comm id size");

int rt_comm_unique_id(unsigned char id[RT_COMM_ID_BYTES]) {
    if (!id) return RT_ERR_INVALID;
    NcclApi& n = nccl();
    if (!n.error.empty()) return RT_ERR_STATE;
    ncclUniqueId u;
    if (n.GetUniqueId(&u) != ncclSuccess) return RT_ERR_CUDA;
    memcpy(id, &u, sizeof u);
    return RT_OK;
}

int rt_comm_create(rt_ctx* ctx, int world, int rank, const unsigned char id[RT_COMM_ID_BYTES], rt_comm** out) {
    RT_CHECK_CTX(ctx);
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return RT_ERR_INVALID;
    *out = nullptr;
    NcclApi& n = nccl();
    if (!n.error.empty()) return rt_fail(ctx, RT_ERR_STATE, n.error.c_str());
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    rt_comm* c = new rt_comm();
    c->ctx = ctx;
    c->world = world;
    c->rank = rank;
    ncclResult_t e = n.CommInitRank(&c->comm, world, u, rank);
    if (e != ncclSuccess) { delete c; return ncclFail(ctx, "ncclCommInitRank", e); }
    *out = c;
    return RT_OK;
}

int rt_comm_destroy(rt_comm* c) {
    if (!c) return RT_ERR_INVALID;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
    return RT_OK;
}

// accum (float4 per pixel) of every rank -> summed in place into rank `root`'s image; asynchronous on the
// context's stream like every launcher.
int rt_render_reduce(rt_render* r, rt_comm* c, int root) {
    if (!r || !c || c->ctx != r->ctx || root < 0 || root >= c->world) return RT_ERR_INVALID;
    rt_ctx* ctx = r->ctx;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclResult_t e = nccl().Reduce(r->accum, r->accum, 4 * r->pixels, ncclFloat32, ncclSum, root, c->comm, ctx->stream);
    if (e != ncclSuccess) return ncclFail(ctx, "ncclReduce", e);
    return RT_OK;
}

}  // extern "C"
