// One-hop grid-wide sum through L2 integer atomics: every CTA adds its contribution (K 64-bit words, each carrying
// an arrival count in its top 16 bits) into `G` replicas of the accumulator and polls its own replica until every
// word has seen all CTAs.  How many cycles per round at 148 CTAs, against 2356 for gather + broadcast
// (tools/pingpong.cu)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/atomic_allreduce tools/atomic_allreduce.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long ld_rel(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_add(unsigned long long *p, unsigned long long v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// words[(round * G + g) * K + k] at `stride` 8-byte words apart; step-indexed so nothing is ever reset
__global__ void allreduce(unsigned long long *words, int K, int G, int stride, int iters, long long *cyc, unsigned long long *sink) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x;
    const int myg = blockIdx.x % G;
    unsigned long long acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned long long *base = words + (size_t)i * G * K * stride;
        for (int q = lane; q < G * K; q += 32) red_add(base + (size_t)q * stride, (1ull << 48) + blockIdx.x + 1);
        unsigned long long u = 0;
        for (;;) {
            if (lane < K) u = ld_rel(base + (size_t)(myg * K + lane) * stride);
            const bool pend = lane < K && (u >> 48) != (unsigned long long)nb;
            if (!__any_sync(0xffffffffu, pend)) break;
        }
        acc += u;
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
    if (acc == 12345) *sink = acc;
}
// same, with DEPTH polls of every word in flight, issued `gap` cycles apart (a fresh observation every `gap` cycles
// instead of every L2 round trip)
template <int DEPTH>
__global__ void allreduce_staggered(unsigned long long *words, int K, int gap, int iters, long long *cyc, unsigned long long *sink) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x;
    unsigned long long acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned long long *base = words + (size_t)i * K;
        if (lane < K) red_add(base + lane, (1ull << 48) + blockIdx.x + 1);
        unsigned long long u[DEPTH];
        const unsigned long long *addr = base + (lane < K ? lane : 0);
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) {
            u[q] = ld_rel(addr);
            if (q + 1 < DEPTH) __nanosleep(gap);
        }
        unsigned long long got = 0;
        bool done = false;
        while (!done) {
#pragma unroll
            for (int q = 0; q < DEPTH; ++q) {
                if (!done) {
                    const bool pend = lane < K && (u[q] >> 48) != (unsigned long long)nb;
                    if (!__any_sync(0xffffffffu, pend)) { done = true; got = u[q]; }
                    else u[q] = ld_rel(addr);
                }
            }
        }
        acc += got;
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
    if (acc == 12345) *sink = acc;
}
template <int DEPTH>
void run_staggered(unsigned long long *buf, size_t bytes, long long *cyc, int K, int gap, int nb) {
    const int iters = 1000;
    long long h;
    cudaMemset(buf, 0, bytes);
    unsigned long long *sink = buf + (bytes / 8 - 1);
    void *args[] = {&buf, (void *)&K, (void *)&gap, (void *)&iters, &cyc, &sink};
    cudaLaunchCooperativeKernel((void *)allreduce_staggered<DEPTH>, dim3(nb), dim3(32), args, 0, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("staggered polls, %3d CTAs, %d words, depth %d, gap %3d ns: %.0f cycles per round\n", nb, K, DEPTH, gap, (double)h / iters);
}
int main() {
    unsigned long long *buf; long long *cyc, h;
    const size_t bytes = 1ull << 28;
    cudaMalloc(&buf, bytes); cudaMalloc(&cyc, 8);
    const int iters = 1000;
    for (int nb : {8, 64, 148})
        for (int K : {2, 8})
            for (int G : {1, 2, 4, 8})
                for (int stride : {1, 16}) {
                    if ((size_t)iters * G * K * stride * 8 > bytes) continue;
                    cudaMemset(buf, 0, bytes);
                    unsigned long long *sink = buf + (bytes / 8 - 1);
                    void *args[] = {&buf, (void *)&K, (void *)&G, (void *)&stride, (void *)&iters, &cyc, &sink};
                    cudaLaunchCooperativeKernel((void *)allreduce, dim3(nb), dim3(32), args, 0, 0);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("err %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                    printf("atomic all-reduce, %3d CTAs, %d words, %d replica(s), word stride %3d B: %.0f cycles per round\n", nb, K, G, stride * 8, (double)h / iters);
                }
    for (int K : {6, 8})
        for (int gap : {50, 100, 200}) {
            run_staggered<1>(buf, bytes, cyc, K, gap, 148);
            run_staggered<2>(buf, bytes, cyc, K, gap, 148);
            run_staggered<4>(buf, bytes, cyc, K, gap, 148);
        }
    return 0;
}
