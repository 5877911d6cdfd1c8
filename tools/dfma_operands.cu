// Does the register-operand pattern limit DFMA issue on B200?  28 DFMA (one 7-slot complex row times vector) per loop
// iteration, all operands in registers, 1..3 warps per SM sub-partition.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/dfma_operands tools/dfma_operands.cu
// MODE 0: 8 chains acc_i = fma(acc_i, b, c), shared b and c              (the usual "peak" loop)
// MODE 1: complex row dot as the warp kernel writes it (compiler's order)
// MODE 2: same, order pinned with volatile asm: per slot  (g.x,x.x) (g.x,x.y) (g.y,x.y) (g.y,x.x)  -> operand reuse
// MODE 3: same, order pinned, per slot (g.x,x.x) (g.y,x.y) (g.x,x.y) (g.y,x.x)                    -> no reuse
#include <cuda_runtime.h>
#include <cstdio>
#define FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double *out, long long *cyc, int iters) {
    const int lane = threadIdx.x & 31;
    double gx[7], gy[7], xx[7], xy[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        gx[s] = 0.01 * (s + 1) + 1e-3 * lane;
        gy[s] = -0.02 * (s + 2) + 1e-3 * lane;
        xx[s] = 0.5 + 0.01 * s;
        xy[s] = 0.25 - 0.01 * s;
    }
    double ar = 0.0, ai = 0.0, ar1 = 0.0, ai1 = 0.0;
    double acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    const double b = 0.999 + 1e-6 * lane, c = 1e-3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 7; ++i) acc[(r * 7 + i) & 7] = fma(acc[(r * 7 + i) & 7], b, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                ar = fma(gx[s], xx[s], ar);
                ar1 = fma(-gy[s], xy[s], ar1);
                ai = fma(gx[s], xy[s], ai);
                ai1 = fma(gy[s], xx[s], ai1);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                FMA(ar, gx[s], xx[s], ar);
                FMA(ai, gx[s], xy[s], ai);
                FMA(ar1, gy[s], xy[s], ar1);
                FMA(ai1, gy[s], xx[s], ai1);
            }
        } else {
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                FMA(ar, gx[s], xx[s], ar);
                FMA(ar1, gy[s], xy[s], ar1);
                FMA(ai, gx[s], xy[s], ai);
                FMA(ai1, gy[s], xx[s], ai1);
            }
        }
        if (MODE != 0) {  // keep the vector live and changing without adding FP64 work
            const double t = xx[0];
#pragma unroll
            for (int s = 0; s < 6; ++s) xx[s] = xx[s + 1];
            xx[6] = t;
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
    double sum = ar + ai + ar1 + ai1;
    for (int i = 0; i < 8; ++i) sum += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

template <int MODE>
void run(const char *name, int wpc, double *out, long long *cyc) {
    const int iters = 20000, blocks = 148;
    for (int rep = 0; rep < 2; ++rep) {
        k<MODE><<<blocks, wpc * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
    }
    long long h[148 * 16];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks * wpc, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks * wpc; ++i) mean += h[i];
    mean /= blocks * wpc;
    printf("%-40s warps/SMSP=%d: %.2f cycles per DFMA per warp, %.2f per SMSP\n", name, wpc / 4, mean / iters / 28.0,
           mean / iters / 28.0 / (wpc / 4));
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 1 << 20);
    for (int wpc : {4, 8, 12}) {
        run<0>("0 shared operands, 8 chains", wpc, out, cyc);
        run<1>("1 row dot, compiler order", wpc, out, cyc);
        run<2>("2 row dot, reuse-friendly order", wpc, out, cyc);
        run<3>("3 row dot, no-reuse order", wpc, out, cyc);
    }
    return 0;
}
