// Inter-SM signalling latency through L2 on B200: two CTAs bounce a step-stamped word.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void st_rel(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_rel(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__global__ void pingpong(unsigned long long *a, unsigned long long *b, int iters, long long *cyc, int other) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != other) return;
    long long t0 = clock64();
    if (blockIdx.x == 0) {
        for (int i = 1; i <= iters; ++i) { st_rel(a, i); while (ld_rel(b) != (unsigned long long)i) {} }
    } else {
        for (int i = 1; i <= iters; ++i) { while (ld_rel(a) != (unsigned long long)i) {} st_rel(b, i); }
    }
    if (blockIdx.x == 0) *cyc = clock64() - t0;
}
// one writer, many pollers on the same word (broadcast); then all pollers ack into padded slots read by the writer
__global__ void bcast_gather(unsigned long long *flag, unsigned long long *acks, int stride, int iters, long long *cyc) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (blockIdx.x == 0) {
            if (lane == 0) st_rel(flag, i);
            // gather acks of all other CTAs
            bool pend = true;
            while (__any_sync(0xffffffffu, pend)) {
                pend = false;
                for (int c = 1 + lane; c < nb; c += 32) pend |= (ld_rel(acks + (size_t)c * stride) != (unsigned long long)i);
            }
        } else {
            if (lane == 0) { while (ld_rel(flag) != (unsigned long long)i) {} st_rel(acks + (size_t)blockIdx.x * stride, i); }
            __syncwarp();
        }
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
}
// all-gather: every CTA posts to its slot, every CTA polls all slots
__global__ void allgather(unsigned long long *slots, int stride, int iters, long long *cyc) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (lane == 0) st_rel(slots + (size_t)blockIdx.x * stride, i);
        bool pend = true;
        while (__any_sync(0xffffffffu, pend)) {
            pend = false;
            for (int c = lane; c < nb; c += 32) pend |= (ld_rel(slots + (size_t)c * stride) < (unsigned long long)i);
        }
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
}
// replicated all-gather: every CTA posts its value into one copy per destination group of `gs` CTAs;
// a CTA polls only its own group's copy (so each line is polled by `gs` CTAs instead of all of them)
__global__ void allgather_repl(unsigned long long *slots, int gs, int iters, long long *cyc) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x, ng = (nb + gs - 1) / gs, myg = blockIdx.x / gs;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        for (int g = lane; g < ng; g += 32) st_rel(slots + ((size_t)g * nb + blockIdx.x) * 2, i);
        bool pend = true;
        while (__any_sync(0xffffffffu, pend)) {
            pend = false;
            for (int c = lane; c < nb; c += 32) pend |= (ld_rel(slots + ((size_t)myg * nb + c) * 2) < (unsigned long long)i);
        }
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
}
// counter variant: post slot, bump a counter (one per destination group), spin on own group's counter, then read all slots once
__global__ void allgather_counter(unsigned long long *slots, unsigned long long *ctr, int gs, int iters, long long *cyc) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x, ng = (nb + gs - 1) / gs, myg = blockIdx.x / gs;
    long long t0 = clock64();
    unsigned long long acc = 0;
    for (int i = 1; i <= iters; ++i) {
        if (lane == 0) { st_rel(slots + (size_t)blockIdx.x * 2, i); __threadfence(); }
        __syncwarp();
        for (int g = lane; g < ng; g += 32) atomicAdd(ctr + (size_t)g * 16, 1ull);
        if (lane == 0) while (ld_rel(ctr + (size_t)myg * 16) < (unsigned long long)i * nb) {}
        __syncwarp();
        for (int c = lane; c < nb; c += 32) acc += ld_rel(slots + (size_t)c * 2);
    }
    if (acc == 12345) *cyc = 0;
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
}
// all-gather with ALL loads of a polling pass in flight at once (5 per lane covers 160 CTAs)
__global__ void allgather_batched(unsigned long long *slots, int stride, int iters, long long *cyc) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (lane == 0) st_rel(slots + (size_t)blockIdx.x * stride, i);
        for (;;) {
            unsigned long long u[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) { int c = lane + 32 * q; u[q] = ld_rel(slots + (size_t)(c < nb ? c : 0) * stride); }
            bool pend = false;
#pragma unroll
            for (int q = 0; q < 5; ++q) pend |= (lane + 32 * q < nb) && (u[q] < (unsigned long long)i);
            if (!__any_sync(0xffffffffu, pend)) break;
        }
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
}
// reducer + broadcast with batched gather at the reducer
__global__ void bcast_gather_batched(unsigned long long *flag, unsigned long long *acks, int stride, int iters, long long *cyc) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, nb = gridDim.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (blockIdx.x == 0) {
            for (;;) {
                unsigned long long u[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) { int c = lane + 32 * q; u[q] = ld_rel(acks + (size_t)(c < nb ? c : 1) * stride); }
                bool pend = false;
#pragma unroll
                for (int q = 0; q < 5; ++q) { int c = lane + 32 * q; pend |= (c >= 1 && c < nb) && (u[q] < (unsigned long long)i); }
                if (!__any_sync(0xffffffffu, pend)) break;
            }
            if (lane == 0) st_rel(flag, i);
        } else {
            if (lane == 0) { st_rel(acks + (size_t)blockIdx.x * stride, i); while (ld_rel(flag) != (unsigned long long)i) {} }
            __syncwarp();
        }
    }
    if (blockIdx.x == 0 && lane == 0) *cyc = clock64() - t0;
}
int main() {
    unsigned long long *buf; long long *cyc, h;
    cudaMalloc(&buf, 1 << 22); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int other : {1, 2, 37, 74, 100, 147}) {
        cudaMemset(buf, 0, 1 << 22);
        void *args[] = {&buf, nullptr, (void *)&iters, &cyc, (void *)&other};
        unsigned long long *b = buf + 1024; args[1] = &b;
        cudaLaunchCooperativeKernel((void *)pingpong, dim3(148), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("ping-pong CTA0 <-> CTA%d: %.0f cycles round trip (%.0f one way)\n", other, (double)h / iters, (double)h / iters / 2);
    }
    for (int stride : {1, 4, 16}) for (int nb : {8, 64, 148}) {
        cudaMemset(buf, 0, 1 << 22);
        unsigned long long *acks = buf + 4096;
        void *args[] = {&buf, &acks, (void *)&stride, (void *)&iters, &cyc};
        cudaLaunchCooperativeKernel((void *)bcast_gather, dim3(nb), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("broadcast+gather, %3d CTAs, slot stride %3d B: %.0f cycles per round\n", nb, stride * 8, (double)h / iters);
    }
    for (int stride : {1, 4, 16}) for (int nb : {8, 64, 148}) {
        cudaMemset(buf, 0, 1 << 22);
        void *args[] = {&buf, (void *)&stride, (void *)&iters, &cyc};
        cudaLaunchCooperativeKernel((void *)allgather, dim3(nb), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("all-gather, %3d CTAs, slot stride %3d B: %.0f cycles per round\n", nb, stride * 8, (double)h / iters);
    }
    for (int stride : {1, 2, 16}) for (int nb : {8, 32, 64, 148}) {
        cudaMemset(buf, 0, 1 << 22);
        void *args[] = {&buf, (void *)&stride, (void *)&iters, &cyc};
        cudaLaunchCooperativeKernel((void *)allgather_batched, dim3(nb), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("BATCHED all-gather, %3d CTAs, slot stride %3d B: %.0f cycles per round\n", nb, stride * 8, (double)h / iters);
    }
    for (int stride : {1, 16}) for (int nb : {8, 64, 148}) {
        cudaMemset(buf, 0, 1 << 22);
        unsigned long long *acks = buf + 4096;
        void *args[] = {&buf, &acks, (void *)&stride, (void *)&iters, &cyc};
        cudaLaunchCooperativeKernel((void *)bcast_gather_batched, dim3(nb), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("BATCHED gather+broadcast, %3d CTAs, slot stride %3d B: %.0f cycles per round\n", nb, stride * 8, (double)h / iters);
    }
    for (int gs : {8}) for (int nb : {148}) {
        cudaMemset(buf, 0, 1 << 22);
        void *args[] = {&buf, (void *)&gs, (void *)&iters, &cyc};
        cudaLaunchCooperativeKernel((void *)allgather_repl, dim3(nb), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("replicated all-gather, %3d CTAs, groups of %2d: %.0f cycles per round\n", nb, gs, (double)h / iters);
    }
    for (int gs : {148}) for (int nb : {148}) {
        cudaMemset(buf, 0, 1 << 22);
        unsigned long long *ctr = buf + (1 << 18);
        void *args[] = {&buf, &ctr, (void *)&gs, (void *)&iters, &cyc};
        cudaLaunchCooperativeKernel((void *)allgather_counter, dim3(nb), dim3(32), args, 0, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("counter all-gather, %3d CTAs, counter per group of %3d: %.0f cycles per round\n", nb, gs, (double)h / iters);
    }
    return 0;
}
