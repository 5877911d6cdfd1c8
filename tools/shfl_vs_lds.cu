// Do SHFL and LDS share one datapath on B200?  Throughput of (a) 6 LDS.128, (b) 3 LDS.128 + 12 SHFL.32,
// (c) 24 SHFL.32 per loop iteration, 16 warps per SM on all SMs.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(double *out, int iters) {
    __shared__ double2 buf[16][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    buf[warp][lane] = make_double2(lane, warp);
    __syncwarp();
    double ar = 0, ai = 0;
    double2 own = make_double2(lane * 0.5, 1.0);
    for (int i = 0; i < iters; ++i) {
        const int o = i & 3;
        if (MODE == 0) {
#pragma unroll
            for (int s = 1; s <= 6; ++s) { double2 x = buf[warp][(lane + s + o) & 31]; ar += x.x; ai += x.y; }
        } else if (MODE == 1) {
#pragma unroll
            for (int s = 1; s <= 3; ++s) { double2 x = buf[warp][(lane + s + o) & 31]; ar += x.x; ai += x.y; }
#pragma unroll
            for (int s = 4; s <= 6; ++s) { ar += __shfl_sync(0xffffffffu, own.x, (lane + s + o) & 31); ai += __shfl_sync(0xffffffffu, own.y, (lane + s + o) & 31); }
        } else {
#pragma unroll
            for (int s = 1; s <= 6; ++s) { ar += __shfl_sync(0xffffffffu, own.x, (lane + s + o) & 31); ai += __shfl_sync(0xffffffffu, own.y, (lane + s + o) & 31); }
        }
        own.x = ar * 1e-9; own.y = ai * 1e-9;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ar + ai;
}
int main() {
    double *out; cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    const int iters = 20000, blocks = 148, threads = 512;
    const char *names[3] = {"6 LDS.128", "3 LDS.128 + 12 SHFL", "24 SHFL"};
    for (int m = 0; m < 3; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<blocks, threads>>>(out, iters); else if (m == 1) k<1><<<blocks, threads>>>(out, iters); else k<2><<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1);
        }
        double cyc = ms * 1e-3 * 1.965e9 / iters / 16.0;  // cycles per warp-iteration per SM
        printf("%-22s: %.3f ms, %.1f SM-cycles per warp-iteration (6 complex neighbours)\n", names[m], ms, cyc);
    }
    return 0;
}
