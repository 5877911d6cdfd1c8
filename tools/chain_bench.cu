// Latency of one Chebyshev term of the warp kernel in isolation: how many cycles does
//   STS(v_{j-1}) -> __syncwarp -> 6 LDS.128 gathers -> 28 DFMA -> v_j
// take with 1 or 2 warps per SM sub-partition, and which restructurings shorten it?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I krotov.jl_b200/csrc -o tools/chain_bench tools/chain_bench.cu
// VAR 0: cheby_step of warp_kernel.cuh as shipped (4 FMA chains per row)
// VAR 1: 8 FMA chains per row (even / odd slots), joined by 4 DADD
// VAR 2: exchange through SHFL (rotation patterns only) instead of shared memory
// VAR 3: as 0 with the loop over terms fully unrolled for m = 17 (compile-time trip count)
// VAR 4: two independent recursions interleaved in one warp (shared rows)
// VAR 5: as 0, but the gathers of the next term are issued right after the store, before the bookkeeping
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "warp_kernel.cuh"

constexpr int W = 6;
__constant__ int c_off[W] = {1, 31, 5, 27, 6, 26};

template <int NCH>
__device__ __forceinline__ void row_dot_v(const double2 (&g)[W + 1], const double2 (&x)[W], const double2 own, double &ar,
                                          double &ai) {
    if (NCH == 4) {
        double ar1 = 0.0, ai1 = 0.0;
        ar = fma(g[W].x, own.x, ar);
        ar1 = fma(-g[W].y, own.y, ar1);
        ai = fma(g[W].x, own.y, ai);
        ai1 = fma(g[W].y, own.x, ai1);
#pragma unroll
        for (int s = 0; s < W; ++s) {
            ar = fma(g[s].x, x[s].x, ar);
            ar1 = fma(-g[s].y, x[s].y, ar1);
            ai = fma(g[s].x, x[s].y, ai);
            ai1 = fma(g[s].y, x[s].x, ai1);
        }
        ar += ar1;
        ai += ai1;
    } else {
        double a0 = ar, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = ai, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        a0 = fma(g[W].x, own.x, a0);
        a1 = fma(-g[W].y, own.y, a1);
        b0 = fma(g[W].x, own.y, b0);
        b1 = fma(g[W].y, own.x, b1);
#pragma unroll
        for (int s = 0; s < W; s += 2) {
            a2 = fma(g[s].x, x[s].x, a2);
            a3 = fma(-g[s].y, x[s].y, a3);
            b2 = fma(g[s].x, x[s].y, b2);
            b3 = fma(g[s].y, x[s].x, b3);
            if (s + 1 < W) {
                a0 = fma(g[s + 1].x, x[s + 1].x, a0);
                a1 = fma(-g[s + 1].y, x[s + 1].y, a1);
                b0 = fma(g[s + 1].x, x[s + 1].y, b0);
                b1 = fma(g[s + 1].y, x[s + 1].x, b1);
            }
        }
        ar = (a0 + a1) + (a2 + a3);
        ai = (b0 + b1) + (b2 + b3);
    }
}

template <int VAR>
__global__ void __launch_bounds__(512, 1) chain(double *out, long long *cyc, int steps, int m, const double *coef) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 *bufA = reinterpret_cast<double2 *>(smem_raw) + (size_t)warp * 128;
    double2 *bufB = bufA + 32;
    double2 *bufC = bufA + 64;
    double2 *bufD = bufA + 96;
    double2 g[W + 1];
    int col[W];
#pragma unroll
    for (int s = 0; s < W; ++s) {
        col[s] = (lane + c_off[s]) & 31;
        g[s] = make_double2(0.01 * (s + 1) * ((lane & 3) - 1.5), 0.02 * (lane % 5 - 2.0));
    }
    g[W] = make_double2(0.0, 0.1 * (lane - 12));
    double2 psi = make_double2(lane == (warp & 31) ? 1.0 : 0.0, 0.0);
    double2 psi2 = make_double2(lane == ((warp + 3) & 31) ? 1.0 : 0.0, 0.0);
    const double2 phase = make_double2(0.8, 0.6);
    bufA[lane] = psi;
    __syncwarp();
    const long long t0 = clock64();
    for (int n = 0; n < steps; ++n) {
        if (VAR == 0) {
            bufC[lane] = psi;
            __syncwarp();
            psi = kr::cheby_step<W, 32>(psi, g, col, bufC, bufA, bufB, coef, m, phase, lane);
        } else if (VAR == 1 || VAR == 3) {
            constexpr int NCH = (VAR == 1) ? 8 : 4;
            bufC[lane] = psi;
            __syncwarp();
            double2 x[W];
#pragma unroll
            for (int s = 0; s < W; ++s) x[s] = bufC[col[s]];
            double2 vm2 = psi;
            double outr = coef[0] * psi.x, outi = coef[0] * psi.y;
            double ar = 0.0, ai = 0.0;
            row_dot_v<NCH>(g, x, psi, ar, ai);
            double2 vm1 = make_double2(0.5 * ar, 0.5 * ai);
            outr = fma(coef[1], vm1.x, outr);
            outi = fma(coef[1], vm1.y, outi);
            bufB[lane] = vm1;
            __syncwarp();
            double2 *cur = bufB, *nxt = bufA;
            if (VAR == 3) {
#pragma unroll
                for (int j = 2; j < 17; ++j) {
                    const double aj = coef[j];
#pragma unroll
                    for (int s = 0; s < W; ++s) x[s] = cur[col[s]];
                    ar = vm2.x;
                    ai = vm2.y;
                    row_dot_v<NCH>(g, x, vm1, ar, ai);
                    outr = fma(aj, ar, outr);
                    outi = fma(aj, ai, outi);
                    vm2 = vm1;
                    vm1 = make_double2(ar, ai);
                    nxt[lane] = vm1;
                    __syncwarp();
                    double2 *t = cur;
                    cur = nxt;
                    nxt = t;
                }
            } else {
                for (int j = 2; j < m; ++j) {
                    const double aj = coef[j];
#pragma unroll
                    for (int s = 0; s < W; ++s) x[s] = cur[col[s]];
                    ar = vm2.x;
                    ai = vm2.y;
                    row_dot_v<NCH>(g, x, vm1, ar, ai);
                    outr = fma(aj, ar, outr);
                    outi = fma(aj, ai, outi);
                    vm2 = vm1;
                    vm1 = make_double2(ar, ai);
                    nxt[lane] = vm1;
                    __syncwarp();
                    double2 *t = cur;
                    cur = nxt;
                    nxt = t;
                }
            }
            psi = make_double2(phase.x * outr - phase.y * outi, phase.x * outi + phase.y * outr);
        } else if (VAR == 2) {
            double2 vm2 = psi, vm1 = psi;
            double outr = coef[0] * psi.x, outi = coef[0] * psi.y;
            for (int j = 1; j < m; ++j) {
                const double aj = coef[j];
                double2 x[W];
#pragma unroll
                for (int s = 0; s < W; ++s) {
                    x[s].x = __shfl_sync(0xffffffffu, vm1.x, col[s]);
                    x[s].y = __shfl_sync(0xffffffffu, vm1.y, col[s]);
                }
                double ar = (j == 1) ? 0.0 : vm2.x, ai = (j == 1) ? 0.0 : vm2.y;
                row_dot_v<4>(g, x, vm1, ar, ai);
                if (j == 1) {
                    ar *= 0.5;
                    ai *= 0.5;
                }
                outr = fma(aj, ar, outr);
                outi = fma(aj, ai, outi);
                vm2 = vm1;
                vm1 = make_double2(ar, ai);
            }
            psi = make_double2(phase.x * outr - phase.y * outi, phase.x * outi + phase.y * outr);
        } else if (VAR == 5) {
            bufC[lane] = psi;
            __syncwarp();
            double2 x[W];
#pragma unroll
            for (int s = 0; s < W; ++s) x[s] = bufC[col[s]];
            double2 vm2 = psi;
            double outr = coef[0] * psi.x, outi = coef[0] * psi.y;
            double ar = 0.0, ai = 0.0;
            row_dot_v<4>(g, x, psi, ar, ai);
            double2 vm1 = make_double2(0.5 * ar, 0.5 * ai);
            bufB[lane] = vm1;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < W; ++s) x[s] = bufB[col[s]];
            outr = fma(coef[1], vm1.x, outr);
            outi = fma(coef[1], vm1.y, outi);
            double2 *cur = bufB, *nxt = bufA;
            for (int j = 2; j < m; ++j) {
                const double aj = coef[j];
                ar = vm2.x;
                ai = vm2.y;
                row_dot_v<4>(g, x, vm1, ar, ai);
                nxt[lane] = make_double2(ar, ai);
                __syncwarp();
#pragma unroll
                for (int s = 0; s < W; ++s) x[s] = nxt[col[s]];
                outr = fma(aj, ar, outr);
                outi = fma(aj, ai, outi);
                vm2 = vm1;
                vm1 = make_double2(ar, ai);
                double2 *t = cur;
                cur = nxt;
                nxt = t;
            }
            psi = make_double2(phase.x * outr - phase.y * outi, phase.x * outi + phase.y * outr);
        } else if (VAR == 4) {
            // two recursions, one instruction stream
            double2 vm2a = psi, vm2b = psi2, vm1a = psi, vm1b = psi2;
            double oar = coef[0] * psi.x, oai = coef[0] * psi.y, obr = coef[0] * psi2.x, obi = coef[0] * psi2.y;
            double2 *ca = bufA, *na = bufB, *cb = bufC, *nb = bufD;
            ca[lane] = vm1a;
            cb[lane] = vm1b;
            __syncwarp();
            for (int j = 1; j < m; ++j) {
                const double aj = coef[j];
                double2 xa[W], xb[W];
#pragma unroll
                for (int s = 0; s < W; ++s) {
                    xa[s] = ca[col[s]];
                    xb[s] = cb[col[s]];
                }
                double ar = (j == 1) ? 0.0 : vm2a.x, ai = (j == 1) ? 0.0 : vm2a.y;
                double br = (j == 1) ? 0.0 : vm2b.x, bi = (j == 1) ? 0.0 : vm2b.y;
                row_dot_v<4>(g, xa, vm1a, ar, ai);
                row_dot_v<4>(g, xb, vm1b, br, bi);
                if (j == 1) {
                    ar *= 0.5; ai *= 0.5; br *= 0.5; bi *= 0.5;
                }
                oar = fma(aj, ar, oar);
                oai = fma(aj, ai, oai);
                obr = fma(aj, br, obr);
                obi = fma(aj, bi, obi);
                vm2a = vm1a;
                vm2b = vm1b;
                vm1a = make_double2(ar, ai);
                vm1b = make_double2(br, bi);
                na[lane] = vm1a;
                nb[lane] = vm1b;
                __syncwarp();
                double2 *t = ca; ca = na; na = t;
                t = cb; cb = nb; nb = t;
            }
            psi = make_double2(phase.x * oar - phase.y * oai, phase.x * oai + phase.y * oar);
            psi2 = make_double2(phase.x * obr - phase.y * obi, phase.x * obi + phase.y * obr);
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = psi.x + psi.y + psi2.x;
}

template <int VAR>
void run(const char *name, int wpc, double *out, long long *cyc, const double *coef, int trajs_per_warp = 1) {
    const int steps = 400, m = 17, blocks = 148;
    const size_t smem = (size_t)wpc * 128 * 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        chain<VAR><<<blocks, wpc * 32, smem>>>(out, cyc, steps, m, coef);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    long long h[148 * 16];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks * wpc, cudaMemcpyDeviceToHost);
    long long mx = 0;
    double mean = 0;
    for (int i = 0; i < blocks * wpc; ++i) {
        mx = h[i] > mx ? h[i] : mx;
        mean += h[i];
    }
    mean /= blocks * wpc;
    const double terms = (double)steps * (m - 1);
    printf("%-44s wpc=%2d: %7.1f cycles/term (max %7.1f), %6.1f per trajectory-term per SMSP, %.3f ms\n", name, wpc,
           mean / terms, mx / terms, mean / terms / trajs_per_warp / (wpc < 4 ? 1.0 : wpc / 4.0) , ms);
}

int main() {
    double *out, *coef;
    long long *cyc;
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 1 << 20);
    cudaMalloc(&coef, 64 * 8);
    double hc[64];
    for (int i = 0; i < 64; ++i) hc[i] = 1.0 / (1 + i * i);
    cudaMemcpy(coef, hc, sizeof(hc), cudaMemcpyHostToDevice);
    for (int wpc : {1, 4, 8, 12}) {
        run<0>("0 shipped cheby_step (LDS.128, 4 chains)", wpc, out, cyc, coef);
        run<1>("1 8 FMA chains", wpc, out, cyc, coef);
        run<2>("2 SHFL exchange", wpc, out, cyc, coef);
        run<3>("3 unrolled m=17", wpc, out, cyc, coef);
        run<4>("4 two recursions per warp", wpc, out, cyc, coef, 2);
        run<5>("5 gathers issued right after the store", wpc, out, cyc, coef);
    }
    return 0;
}
