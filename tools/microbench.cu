// B200 FP64 microbenchmarks used to pick the kernel design and the FP64 roofline denominator.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// --- dependent DFMA chain latency (1 warp) -------------------------------------------------------
__global__ void k_dfma_dep(double *out, long long *cyc, int iters, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// --- N independent DFMA chains in one warp: issue interval ---------------------------------------
template <int NCH>
__global__ void k_dfma_ind(double *out, long long *cyc, int iters, double a, double b) {
    double x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) x[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < NCH; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// --- LDS.128 / SHFL dependent latency ---------------------------------------------------------------
__global__ void k_lds_dep(double *out, long long *cyc, int iters) {
    __shared__ double2 buf[64];
    buf[threadIdx.x] = make_double2((threadIdx.x + 1) & 31, 0.0);
    buf[threadIdx.x + 32] = make_double2(0, 0);
    __syncwarp();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) idx = (int)buf[idx].x;
    }
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl_dep(double *out, long long *cyc, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x = __shfl_xor_sync(0xffffffffu, x, 1 + (j & 3));
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl_add_dep(double *out, long long *cyc, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// STS -> syncwarp -> LDS round trip
__global__ void k_sts_lds(double *out, long long *cyc, int iters) {
    __shared__ double2 buf[2][32];
    double2 v = make_double2(threadIdx.x, 1.0);
    int lane = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        buf[i & 1][lane] = v;
        __syncwarp();
        double2 w = buf[i & 1][(lane + 1) & 31];
        v.x = w.x + 1.0;
        v.y = w.y;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v.x + v.y;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// --- chip-wide DFMA throughput -----------------------------------------------------------------
__global__ void k_dfma_peak(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// --- chip-wide DMMA m8n8k4 throughput --------------------------------------------------------------
__global__ void k_dmma_peak(double *out, int iters, double a, double b) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    double fa = a + threadIdx.x, fb = b;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[2]), "+d"(c0[3]) : "d"(fa), "d"(fb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[2]), "+d"(c1[3]) : "d"(fa), "d"(fb));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
}
__global__ void k_dmma16_peak(double *out, int iters, double a, double b) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    double fa[8], fb[4];
    for (int i = 0; i < 8; ++i) fa[i] = a + threadIdx.x + i;
    for (int i = 0; i < 4; ++i) fb[i] = b + i;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c0[0]), "+d"(c0[1]), "+d"(c0[2]), "+d"(c0[3])
                         : "d"(fa[0]), "d"(fa[1]), "d"(fa[2]), "d"(fa[3]), "d"(fa[4]), "d"(fa[5]), "d"(fa[6]), "d"(fa[7]), "d"(fb[0]), "d"(fb[1]), "d"(fb[2]), "d"(fb[3]));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c1[0]), "+d"(c1[1]), "+d"(c1[2]), "+d"(c1[3])
                         : "d"(fa[0]), "d"(fa[1]), "d"(fa[2]), "d"(fa[3]), "d"(fa[4]), "d"(fa[5]), "d"(fa[6]), "d"(fa[7]), "d"(fb[0]), "d"(fb[1]), "d"(fb[2]), "d"(fb[3]));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
}

int main() {
    double *out; long long *cyc, h;
    CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 8));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int it = 20000;
    for (int rep = 0; rep < 2; ++rep) { k_dfma_dep<<<1, 32>>>(out, cyc, it, 1.0000001, 1e-9); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("DFMA dependent latency: %.2f cycles\n", (double)h / (it * 16.0));
#define IND(N) { for (int rep = 0; rep < 2; ++rep) { k_dfma_ind<N><<<1, 32>>>(out, cyc, it, 1.0000001, 1e-9); CK(cudaDeviceSynchronize()); } \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("DFMA %d indep chains, 1 warp : %.2f cycles/DFMA\n", N, (double)h / (it * 4.0 * N)); }
    IND(2) IND(4) IND(8) IND(16)
#define INDW(N, NW) { for (int rep = 0; rep < 2; ++rep) { k_dfma_ind<N><<<1, 32 * NW>>>(out, cyc, it, 1.0000001, 1e-9); CK(cudaDeviceSynchronize()); } \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("DFMA %d chains x %d warps (1 SM): %.2f cycles per warp-DFMA per SMSP\n", N, NW, (double)h / (it * 4.0 * N * (NW / 4.0))); }
    INDW(8, 4) INDW(8, 8) INDW(8, 16)
    for (int rep = 0; rep < 2; ++rep) { k_lds_dep<<<1, 32>>>(out, cyc, it); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("LDS.128 dependent (incl. F2I): %.2f cycles\n", (double)h / (it * 8.0));
    for (int rep = 0; rep < 2; ++rep) { k_shfl_dep<<<1, 32>>>(out, cyc, it); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("SHFL f64 dependent: %.2f cycles\n", (double)h / (it * 8.0));
    for (int rep = 0; rep < 2; ++rep) { k_shfl_add_dep<<<1, 32>>>(out, cyc, it); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("5-level f64 xor butterfly: %.2f cycles\n", (double)h / it);
    for (int rep = 0; rep < 2; ++rep) { k_sts_lds<<<1, 32>>>(out, cyc, it); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); printf("STS.128 -> syncwarp -> LDS.128 -> DADD loop: %.2f cycles\n", (double)h / it);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int nw : {4, 8, 16, 32}) {
        int blocks = p.multiProcessorCount * 2, iters = 4000;
        k_dfma_peak<<<blocks, nw * 32 / 2>>>(out, 10, 1.0000001, 1e-9);
        cudaEventRecord(e0); k_dfma_peak<<<blocks, nw * 32 / 2>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * blocks * (nw * 32 / 2) * (double)iters * 64;
        printf("DFMA peak, %2d warps/SM: %.2f TFLOP/s (%.3f ms)\n", nw, flops / ms / 1e9, ms);
    }
    for (int nw : {4, 8, 16, 32}) {
        int blocks = p.multiProcessorCount * 2, iters = 4000;
        k_dmma_peak<<<blocks, nw * 32 / 2>>>(out, 10, 1.0, 1e-9);
        cudaEventRecord(e0); k_dmma_peak<<<blocks, nw * 32 / 2>>>(out, iters, 1.0, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * 8 * 4 * (double)blocks * (nw / 2) * (double)iters * 32;
        printf("DMMA m8n8k4 peak, %2d warps/SM: %.2f TFLOP/s (%.3f ms)\n", nw, flops / ms / 1e9, ms);
    }
    for (int nw : {4, 8, 16, 32}) {
        int blocks = p.multiProcessorCount * 2, iters = 2000;
        k_dmma16_peak<<<blocks, nw * 32 / 2>>>(out, 10, 1.0, 1e-9);
        cudaEventRecord(e0); k_dmma16_peak<<<blocks, nw * 32 / 2>>>(out, iters, 1.0, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * 16 * (double)blocks * (nw / 2) * (double)iters * 16;
        printf("DMMA m16n8k16 peak, %2d warps/SM: %.2f TFLOP/s (%.3f ms)\n", nw, flops / ms / 1e9, ms);
    }
    return 0;
}
