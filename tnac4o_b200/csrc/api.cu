// Context, error reporting and scratch management of the tnac4o_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void tn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int tn_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    tn_set_error("%s:%d: CUDA error %d (%s) in %s", file, line, (int)e, cudaGetErrorString(e), what);
    return TN_ERR_CUDA;
}

static std::mutex g_capture_mutex;
void tn_capture_lock() { g_capture_mutex.lock(); }
void tn_capture_unlock() { g_capture_mutex.unlock(); }

cudaError_t tn_malloc_async(tn_ctx* ctx, void** p, size_t bytes, cudaStream_t st) {
    if (ctx->pool) return cudaMallocFromPoolAsync(p, bytes, ctx->pool, st);
    return cudaMallocAsync(p, bytes, st);
}

void* tn_scratch(tn_ctx* ctx, int slot, size_t bytes) {
    if (bytes <= ctx->scratch_bytes[slot]) return ctx->scratch[slot];
    if (ctx->capturing) {
        tn_set_error("scratch slot %d would have to grow (%zu bytes) during a stream capture", slot, bytes);
        return nullptr;
    }
    ctx->scratch_gen++;
    // grow-only; a grow synchronises the device so that no in-flight kernel still reads the old block
    size_t want = bytes + bytes / 4 + (1 << 20);
    std::lock_guard<std::mutex> lock(g_capture_mutex);      // no stream capture of another thread is open in here
    cudaDeviceSynchronize();
    if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
    if (e != cudaSuccess) {
        tn_cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        return nullptr;
    }
    ctx->scratch_bytes[slot] = want;
    return ctx->scratch[slot];
}

extern "C" {

int tn_version(void) { return 100; }

const char* tn_last_error(void) { return g_err; }

int tn_create(int device, tn_ctx** out) {
    TN_REQUIRE(out != nullptr, "null output pointer");
    int count = 0;
    TN_CUDA(cudaGetDeviceCount(&count));
    TN_REQUIRE(device >= 0 && device < count, "no such CUDA device");
    TN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        tn_set_error("tnac4o_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return TN_ERR_ARG;
    }
    tn_ctx* ctx = new tn_ctx();
    ctx->device = device;
    {
        // temporaries of the native drivers come from a stream-ordered pool owned by this context; freed blocks stay cached
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
            ctx->pool = nullptr;
        }
    }
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaMallocHost(&ctx->pinned, 4096);
    if (e != cudaSuccess) {
        delete ctx;
        return tn_cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
    }
    if (!tn_scratch(ctx, TN_SLOT_GEMM, (size_t)32 << 20)) {
        cudaFreeHost(ctx->pinned);
        delete ctx;
        return TN_ERR_NOMEM;
    }
    e = cudaMalloc(&ctx->counters, 4096);
    if (e == cudaSuccess) e = cudaMemset(ctx->counters, 0, 4096);
    if (e != cudaSuccess) {
        cudaFreeHost(ctx->pinned);
        delete ctx;
        return tn_cuda_fail(e, "cudaMalloc(counters)", __FILE__, __LINE__);
    }
    *out = ctx;
    return TN_OK;
}

int tn_destroy(tn_ctx* ctx) {
    if (!ctx) return TN_OK;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < tn_ctx::SLOTS; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(ctx->pool); }
    delete ctx;
    return TN_OK;
}

int64_t tn_launch_count(const tn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int tn_profile(tn_ctx* ctx, int on) {
    TN_REQUIRE(ctx != nullptr, "null context");
    for (auto& r : ctx->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    ctx->prof_recs.clear();
    for (int c = 0; c < TN_P_COUNT; ++c)
        for (int k = 0; k < 4; ++k) ctx->prof_acc[c][k] = 0.0;
    ctx->prof_on = on != 0;
    return TN_OK;
}

int tn_profile_read(tn_ctx* ctx, double* h_out, int ncat) {
    TN_REQUIRE(ctx != nullptr && h_out != nullptr && ncat >= 1, "bad arguments");
    for (auto& r : ctx->prof_recs) {
        TN_CUDA(cudaEventSynchronize(r.e1));
        float ms = 0.f;
        TN_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        ctx->prof_acc[r.cat][0] += 1e-3 * ms;
        ctx->prof_acc[r.cat][1] += r.flops;
        ctx->prof_acc[r.cat][2] += r.bytes;
        ctx->prof_acc[r.cat][3] += 1.0;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    ctx->prof_recs.clear();
    for (int c = 0; c < ncat && c < TN_P_COUNT; ++c)
        for (int k = 0; k < 4; ++k) h_out[c * 4 + k] = ctx->prof_acc[c][k];
    return TN_OK;
}

int tn_set_blocking_sync(int on) {
    // host threads waiting in a synchronising call sleep instead of spinning: many solver threads can share few cores
    // 0: driver default (spin), 1: sleep on an interrupt (frees the core, ~0.4 ms wake-up), 2: spin with sched_yield (keeps
    // the wake-up short while letting other runnable threads of an oversubscribed host take the core)
    TN_CUDA(cudaSetDeviceFlags(on == 1 ? cudaDeviceScheduleBlockingSync : (on == 2 ? cudaDeviceScheduleYield : cudaDeviceScheduleAuto)));
    return TN_OK;
}

}  // extern "C"
