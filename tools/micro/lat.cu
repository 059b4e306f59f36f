// Dependent-chain latency microbenchmarks for the primitives on the FPS critical path (B200).
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define ITERS 256
#define CHAIN(NAME, BODY)                                              \
    {                                                                  \
        unsigned v = seed + lane;                                      \
        long long t0 = clock64();                                      \
        _Pragma("unroll 16") for (int i = 0; i < ITERS; ++i) { BODY; } \
        long long t1 = clock64();                                      \
        sink += v;                                                     \
        if (threadIdx.x == 0) printf("%-28s %6.1f cyc\n", NAME, double(t1 - t0) / ITERS); \
    }
__global__ void lat(unsigned seed, unsigned *out) {
    __shared__ unsigned sm[1024];
    const int lane = threadIdx.x & 31;
    unsigned sink = 0;
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (threadIdx.x < 32) {
        CHAIN("iadd (baseline)", v = v * 3 + 1)
        CHAIN("redux.max.u32", v = __reduce_max_sync(FULL, v ^ lane) + 1)
        CHAIN("redux.min.u32", v = __reduce_min_sync(FULL, v + lane) + 1)
        CHAIN("ballot", v = __ballot_sync(FULL, (v + lane) & 1) + i)
        CHAIN("shfl.idx", v = __shfl_sync(FULL, v, (v + 1) & 31) + 1)
        CHAIN("ffs (brev+flo)", v = __ffs(v | 0x100) + v)
        CHAIN("popc", v = __popc(v) + v + 7)
        CHAIN("clz", v = __clz(v | 1) + v + 1)
        CHAIN("lds", v = sm[v & 1023])
        CHAIN("sts+lds same thread", sm[lane] = v; v = sm[lane] + 1)
        CHAIN("match_any", v = __match_any_sync(FULL, v & 3) + i)
        CHAIN("redux+ballot+ffs+shfl", { unsigned mx = __reduce_max_sync(FULL, v ^ lane); unsigned b = __ballot_sync(FULL, (v ^ lane) == mx); v = __shfl_sync(FULL, v, __ffs(b) - 1) + 1; })
        CHAIN("redux+redux(payload)", { unsigned x = v ^ lane; unsigned mx = __reduce_max_sync(FULL, x); v = __reduce_max_sync(FULL, x == mx ? (v + lane) : 0u) + 1; })
        CHAIN("fmnmx/ffma chain x4", { float f = __uint_as_float(v & 0x3fffffff); f = fmaf(f, f, 1.f); f = fminf(f, 3.f); f = fmaf(f, f, 1.f); f = fmaxf(f, 0.5f); v = __float_as_uint(f); })
    }
    __syncthreads();
    // barrier round trip with all warps of the CTA: sts -> bar -> lds
    {
        unsigned v = seed + threadIdx.x;
        long long t0 = clock64();
        for (int i = 0; i < ITERS; ++i) {
            if (lane == 0) sm[(i & 1) * 32 + (threadIdx.x >> 5)] = v;
            __syncthreads();
            v = sm[(i & 1) * 32 + (lane & (blockDim.x / 32 - 1))] + 1;
        }
        long long t1 = clock64();
        sink += v;
        if (threadIdx.x == 0) printf("sts+bar(%2d warps)+lds        %6.1f cyc\n", blockDim.x / 32, double(t1 - t0) / ITERS);
    }
    {
        unsigned v = seed + threadIdx.x;
        long long t0 = clock64();
        for (int i = 0; i < ITERS; ++i) { __syncthreads(); v += i; }
        long long t1 = clock64();
        sink += v;
        if (threadIdx.x == 0) printf("bar.sync only (%2d warps)     %6.1f cyc\n", blockDim.x / 32, double(t1 - t0) / ITERS);
    }
    out[threadIdx.x] = sink;
}
int main() {
    unsigned *d;
    cudaMalloc(&d, 4096);
    for (int nt : {32, 128, 256, 512, 1024}) {
        printf("---- block %d\n", nt);
        lat<<<1, nt>>>(12345u, d);
        cudaDeviceSynchronize();
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
