// capi.cu -- error plumbing and bookkeeping behind include/pdm_ops.h.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace pdm {

thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code == 0 ? PDM_ERR_INVALID_ARG : code;
}

int ensure_dynamic_smem(const void *func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail((int)e, "cudaGetDevice: %s", cudaGetErrorString(e));
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = done[{func, dev}];
    if (bytes <= cur) return PDM_OK;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e));
    cur = bytes;
    return PDM_OK;
}

void prefer_max_smem(const void *func) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, bool> done;
    static const bool off = [] { const char *e = getenv("PDM_CARVEOUT"); return e && e[0] == 'o'; }();
    if (off) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(mu);
    bool &d = done[{func, dev}];
    if (d) return;
    d = true;
    if (cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        (void)cudaGetLastError();   // a hint only: never fail a launch over it
}

void *stream_scratch(cudaStream_t st, size_t bytes) {
    struct Buf { void *ptr = nullptr; size_t size = 0; };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { fail(PDM_ERR_INVALID_ARG, "stream_scratch: cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lock(mu);
    Buf &b = bufs[{dev, st}];
    if (bytes <= b.size) return b.ptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap != cudaStreamCaptureStatusNone) {
        fail(PDM_ERR_UNSUPPORTED, "scratch would grow to %zu B during stream capture: run the call once before capturing", bytes);
        return nullptr;
    }
    const size_t want = bytes + bytes / 4;   // head-room so that slightly larger calls do not reallocate
    void *p = nullptr;
    const cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { fail((int)e, "stream_scratch: cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return nullptr; }
    b.ptr = p;      // the outgrown buffer stays allocated: earlier work / captured graphs may still use it
    b.size = want;
    return p;
}

// cuda_utils.h:10-14 of the reference, evaluated the same way (double log ratio, truncation).
int ref_fps_block_size(int n) {
    const int pow_2 = (int)(log((double)n) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

}  // namespace pdm

extern "C" {

int pdm_abi_version(void) { return 1; }
const char *pdm_last_error(void) { return pdm::g_err; }
long long pdm_launch_count(void) { return pdm::g_launches.load(); }
void pdm_reset_launch_count(void) { pdm::g_launches.store(0); }

}  // extern "C"
