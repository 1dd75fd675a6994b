// rt_ctx.cu -- context, stream and buffer management behind the C ABI (include/rt2015.h).
// Replaces the WebCL context / command queue / buffer objects the reference host creates
// in createCLBasicResources (Assign10-Path_Tracing/code.js:576-608) and releases in
// releaseCLResources (:1539-1552).  There is deliberately no CPU fallback: without a CUDA
// device rt_ctx_create fails with RT_ERR_NO_DEVICE.
#include "rt_internal.h"

extern "C" {

int rt_ctx_create(int device_ordinal, rt_ctx** out) {
    if (!out) return RT_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return RT_ERR_NO_DEVICE;
    if (device_ordinal < 0 || device_ordinal >= count) return RT_ERR_INVALID;
    rt_ctx* ctx = new rt_ctx();
    ctx->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device_ordinal) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return RT_ERR_CUDA;
    }
    cudaMemPoolProps pp;
    memset(&pp, 0, sizeof pp);
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = ctx->device;
    if (cudaMemPoolCreate(&ctx->pool, &pp) == cudaSuccess) {
        unsigned long long keep = ~0ull;   // never trim at synchronisation points
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
        ctx->pool = nullptr;               // fall back to the device's default pool
        cudaGetLastError();
    }
    *out = ctx;
    return RT_OK;
}

int rt_ctx_destroy(rt_ctx* ctx) {
    RT_CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RT_OK;
}

const char* rt_last_error_string(rt_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int rt_finish(rt_ctx* ctx) {
    RT_CHECK_CTX(ctx);
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

void* rt_ctx_stream(rt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int rt_device_info(rt_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes, size_t* total_mem) {
    RT_CHECK_CTX(ctx);
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (l2_bytes) *l2_bytes = (size_t)ctx->prop.l2CacheSize;
    if (total_mem) *total_mem = ctx->prop.totalGlobalMem;
    return RT_OK;
}

int rt_buffer_create(rt_ctx* ctx, size_t bytes, void** dptr) {
    RT_CHECK_CTX(ctx);
    if (!dptr) return RT_ERR_INVALID;
    *dptr = nullptr;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) return rt_fail(ctx, RT_ERR_NOMEM, "cudaMalloc", e);
    if (e != cudaSuccess) return rt_fail(ctx, RT_ERR_CUDA, "cudaMalloc", e);
    return RT_OK;
}

int rt_buffer_release(rt_ctx* ctx, void* dptr) {
    RT_CHECK_CTX(ctx);
    if (!dptr) return RT_OK;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RT_CUDA(ctx, cudaFree(dptr));
    return RT_OK;
}

int rt_buffer_write(rt_ctx* ctx, void* dptr, size_t offset, size_t bytes, const void* host) {
    RT_CHECK_CTX(ctx);
    if (!dptr || (!host && bytes)) return RT_ERR_INVALID;
    if (!bytes) return RT_OK;
    // In-order with the kernels on the context stream; the copy is complete (for pageable
    // memory: staged) when the call returns, so the caller may reuse `host` at once.
    RT_CUDA(ctx, cudaMemcpyAsync((char*)dptr + offset, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_buffer_read(rt_ctx* ctx, const void* dptr, size_t offset, size_t bytes, void* host) {
    RT_CHECK_CTX(ctx);
    if (!dptr || (!host && bytes)) return RT_ERR_INVALID;
    if (!bytes) return RT_OK;
    RT_CUDA(ctx, cudaMemcpyAsync(host, (const char*)dptr + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_buffer_fill(rt_ctx* ctx, void* dptr, int byte_value, size_t bytes) {
    RT_CHECK_CTX(ctx);
    if (!dptr) return RT_ERR_INVALID;
    RT_CUDA(ctx, cudaMemsetAsync(dptr, byte_value, bytes, ctx->stream));
    return RT_OK;
}

unsigned rt_struct_size(const char* name, int assignment) {
    if (!name) return 0;
    if (!strcmp(name, "Ray")) return (assignment >= 3 && assignment <= 10) ? 48u : 0u;
    if (!strcmp(name, "Poi")) {
        if (assignment == 10) return 64u;
        if (assignment == 8 || assignment == 9) return 48u;
    }
    return 0;
}

int rt_set_walk_stats(rt_ctx* ctx, void* hit_id_u32, void* cells_u32, void* tests_u32) {
    RT_CHECK_CTX(ctx);
    ctx->st_hit = (unsigned*)hit_id_u32;
    ctx->st_cells = (unsigned*)cells_u32;
    ctx->st_tests = (unsigned*)tests_u32;
    return RT_OK;
}

int rt_set_walk_totals(rt_ctx* ctx, void* totals_u64x8) {
    RT_CHECK_CTX(ctx);
    ctx->st_totals = (unsigned long long*)totals_u64x8;
    return RT_OK;
}

}  // extern "C"
