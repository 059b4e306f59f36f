// capi.cu -- error plumbing and bookkeeping behind include/pdm_ops.h.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
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

// Scratch ownership.
//   * Eager calls on a stream share one buffer per (device, stream): work on a stream is serialised.  When it has
//     to grow, the old buffer is released (cudaFree synchronises the device, so nothing can still be using it).
//   * A call made while its stream is being CAPTURED gets a buffer that belongs to that capture alone (keyed by
//     the capture id and the stream -- a capture may fork onto side streams whose work runs concurrently): the first such call adopts the stream's eager buffer -- sized by the warm-up run every
//     capture needs anyway, allocation is not possible during capture -- and the stream's eager slot is emptied,
//     so later eager calls, and later captures on the same stream, get a different buffer.  Two graphs therefore
//     never share scratch, whatever streams they are replayed on (a graph does not run concurrently with itself),
//     and eager work on the capture stream cannot collide with a replay either.  The calls inside ONE graph are
//     ordered by the captured dependencies of that stream and share the graph's buffer.
//   Buffers adopted by a capture stay allocated for the life of the process (the library cannot see a graph die);
//   that is bounded by the number of captures, not by the number of calls.
void *stream_scratch(cudaStream_t st, size_t bytes) {
    struct Buf { void *ptr = nullptr; size_t size = 0; };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Buf> eager;
    static std::map<std::tuple<int, unsigned long long, cudaStream_t>, Buf> captured;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { fail(PDM_ERR_INVALID_ARG, "stream_scratch: cudaGetDevice failed"); return nullptr; }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    unsigned long long cap_id = 0;
    if (cudaStreamGetCaptureInfo(st, &cap, &cap_id) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    std::lock_guard<std::mutex> lock(mu);
    if (cap != cudaStreamCaptureStatusNone) {
        Buf &g = captured[std::make_tuple(dev, cap_id, st)];   // per stream: a graph may fork onto side streams that run concurrently
        if (!g.ptr) {                       // first scratch user of this capture: adopt the stream's eager buffer
            Buf &e = eager[{dev, st}];
            g = e;
            e = Buf();
        }
        if (bytes <= g.size) return g.ptr;
        fail(PDM_ERR_UNSUPPORTED, "scratch of %zu B needed during stream capture but only %zu B were prepared: run the same "
             "calls once on this stream before capturing", bytes, g.size);
        return nullptr;
    }
    Buf &b = eager[{dev, st}];
    if (bytes <= b.size) return b.ptr;
    const size_t want = bytes + bytes / 4;   // head-room so that slightly larger calls do not reallocate
    if (b.ptr) {
        if (cudaFree(b.ptr) != cudaSuccess) (void)cudaGetLastError();   // implicit device synchronisation: no user left
        b = Buf();
    }
    void *p = nullptr;
    const cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { fail((int)e, "stream_scratch: cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return nullptr; }
    b.ptr = p;
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
