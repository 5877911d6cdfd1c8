// Two-trajectories-per-warp variant of the persistent Krotov kernel (ensembles; sm_100a).
//
// Same algorithm, same exchange, same data layout as krotov_warp_kernel (warp_kernel.cuh).  The difference is
// instruction-level parallelism: a warp owns TWO trajectories that share a generator (two basis states of one
// ensemble sample), keeps ONE copy of the generator rows in registers and runs the two Chebyshev recursions
// interleaved in a single instruction stream.  A single recursion is a dependent chain
// (STS -> __syncwarp -> LDS ~57 cycles, then 7 DFMA levels of 8 cycles) that one warp cannot hide; with
// 1024 trajectories on 148 SMs the one-trajectory kernel has two such warps on three of the four SM
// sub-partitions and relies on the scheduler to overlap them, the pair kernel has exactly one warp per
// sub-partition with eight independent FMA chains in flight.
#pragma once
#include "warp_kernel.cuh"

namespace kr {

// acc_t += G_row . v_t  for two vectors sharing the row; four chains per trajectory
template <int W>
__device__ __forceinline__ void row_dot2(const double2 (&g)[W + 1], const int (&col)[W], const double2 *__restrict__ va,
                                         const double2 *__restrict__ vb, const double2 owna, const double2 ownb,
                                         double &ara, double &aia, double &arb, double &aib) {
    double ra1 = 0.0, ia1 = 0.0, rb1 = 0.0, ib1 = 0.0;
    ara = fma(g[W].x, owna.x, ara);
    ra1 = fma(-g[W].y, owna.y, ra1);
    aia = fma(g[W].x, owna.y, aia);
    ia1 = fma(g[W].y, owna.x, ia1);
    arb = fma(g[W].x, ownb.x, arb);
    rb1 = fma(-g[W].y, ownb.y, rb1);
    aib = fma(g[W].x, ownb.y, aib);
    ib1 = fma(g[W].y, ownb.x, ib1);
#pragma unroll
    for (int s = 0; s < W; ++s) {
        const double2 xa = va[col[s]];
        const double2 xb = vb[col[s]];
        ara = fma(g[s].x, xa.x, ara);
        ra1 = fma(-g[s].y, xa.y, ra1);
        aia = fma(g[s].x, xa.y, aia);
        ia1 = fma(g[s].y, xa.x, ia1);
        arb = fma(g[s].x, xb.x, arb);
        rb1 = fma(-g[s].y, xb.y, rb1);
        aib = fma(g[s].x, xb.y, aib);
        ib1 = fma(g[s].y, xb.x, ib1);
    }
    ara += ra1;
    aia += ia1;
    arb += rb1;
    aib += ib1;
}

// Two Chebyshev steps interleaved.  bufs: [traj][A/B][32]; v0: [traj][32] holds psi for all lanes.
template <int W>
__device__ __forceinline__ void cheby_step2(double2 &psia, double2 &psib, const double2 (&g)[W + 1], const int (&col)[W],
                                            const double2 *v0, double2 *bufs, const double *__restrict__ a, const int m,
                                            const double2 phase, const int lane) {
    double2 a2 = psia, b2 = psib;
    const double c0 = a[0];
    double oar = c0 * psia.x, oai = c0 * psia.y, obr = c0 * psib.x, obi = c0 * psib.y;
    double ara = 0.0, aia = 0.0, arb = 0.0, aib = 0.0;
    row_dot2<W>(g, col, v0, v0 + 32, psia, psib, ara, aia, arb, aib);
    double2 a1 = make_double2(0.5 * ara, 0.5 * aia), b1 = make_double2(0.5 * arb, 0.5 * aib);
    if (m > 1) {
        const double c1 = a[1];
        oar = fma(c1, a1.x, oar);
        oai = fma(c1, a1.y, oai);
        obr = fma(c1, b1.x, obr);
        obi = fma(c1, b1.y, obi);
    }
    double2 *curA = bufs + 32, *nxtA = bufs;            // trajectory a: B then A
    double2 *curB = bufs + 96, *nxtB = bufs + 64;       // trajectory b
    curA[lane] = a1;
    curB[lane] = b1;
    __syncwarp();
    for (int j = 2; j < m; ++j) {
        const double cj = a[j];
        ara = a2.x; aia = a2.y; arb = b2.x; aib = b2.y;
        row_dot2<W>(g, col, curA, curB, a1, b1, ara, aia, arb, aib);
        oar = fma(cj, ara, oar);
        oai = fma(cj, aia, oai);
        obr = fma(cj, arb, obr);
        obi = fma(cj, aib, obi);
        a2 = a1; b2 = b1;
        a1 = make_double2(ara, aia);
        b1 = make_double2(arb, aib);
        nxtA[lane] = a1;
        nxtB[lane] = b1;
        __syncwarp();
        double2 *t = curA; curA = nxtA; nxtA = t;
        t = curB; curB = nxtB; nxtB = t;
    }
    psia = make_double2(phase.x * oar - phase.y * oai, phase.x * oai + phase.y * oar);
    psib = make_double2(phase.x * obr - phase.y * obi, phase.x * obi + phase.y * obr);
}

template <int W, int LT>
__global__ void __launch_bounds__(256, 1) krotov_warp2_kernel(const __grid_constant__ WarpParams p) {
    constexpr int NT = 1 + LT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int L = LT;
    const int wpc = p.wpc;
    double2 *vbuf = reinterpret_cast<double2 *>(smem_raw);                       // [wpc][2 traj][2][32]
    double2 *psis = vbuf + (size_t)wpc * 128;                                    // [wpc][2 traj][32]
    double *red = reinterpret_cast<double *>(psis + (size_t)wpc * 64);           // [L][wpc*32]
    double *eps_s = red + (size_t)L * wpc * 32;                                  // [kMaxCtrl]
    double *gbuf = eps_s + kMaxCtrl;                                             // [kMaxCtrl*160]
    const int nthr_all = (wpc + 1) * 32;
    const int N_T = p.N_T;
    if (warp == wpc) {
        __shared__ __align__(16) WarpParams p_sh;
        comm_warp_run(p, (int)blockIdx.x, L, lane, wpc, nthr_all, red, eps_s, gbuf, &p_sh);
        return;
    }
    const int ka = (blockIdx.x * wpc + warp) * 2, kb = ka + 1;
    const bool live = kb < p.N;  // N is even and pairs share a generator (checked on the host)
    double2 *bufs = vbuf + (size_t)warp * 128;
    double2 *myps = psis + (size_t)warp * 64;
    int col[W];
#pragma unroll
    for (int s = 0; s < W; ++s) col[s] = p.cols[s * 32 + lane];
    const size_t rowstride = (size_t)(W + 1) * 32;
    double2 P[NT][W + 1];
    double2 g[W + 1];
    const int gi = live ? p.gen_of_traj[ka] : 0;
    const long long t_begin = clock64();

    // ================================================================ backward sweep
    if (p.mode == 1 && live) {
        const double2 *Pg = p.Pb + (size_t)gi * NT * rowstride;
#pragma unroll
        for (int q = 0; q < NT; ++q) load_row<W>(Pg + q * rowstride, P[q], lane);
        double2 ca, cb;
        if (p.chiT != nullptr) {
            ca = p.chiT[(size_t)ka * 32 + lane];
            cb = p.chiT[(size_t)kb * 32 + lane];
        } else {
            const double2 fa = p.chi_coef[ka], fb = p.chi_coef[kb];
            const double2 ta = p.target[(size_t)ka * 32 + lane], tb = p.target[(size_t)kb * 32 + lane];
            ca = make_double2(fa.x * ta.x - fa.y * ta.y, fa.x * ta.y + fa.y * ta.x);
            cb = make_double2(fb.x * tb.x - fb.y * tb.y, fb.x * tb.y + fb.y * tb.x);
        }
        double2 *Xa = p.X + (size_t)ka * (N_T + 1) * 32, *Xb = p.X + (size_t)kb * (N_T + 1) * 32;
        Xa[(size_t)N_T * 32 + lane] = ca;
        Xb[(size_t)N_T * 32 + lane] = cb;
        myps[lane] = ca;
        myps[32 + lane] = cb;
        __syncwarp();
        StepMeta meta = load_meta(p.dtc_b, p.m_b, p.phase_b, p.coef_b, p.ndtc_b, p.mmax_b, gi, N_T - 1);
        double e_cur[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) e_cur[l] = p.amp_old[(size_t)l * N_T + N_T - 1];
        for (int n = N_T - 1; n >= 0; --n) {
            const int nn = n > 0 ? n - 1 : 0;
            const StepMeta meta_next = load_meta(p.dtc_b, p.m_b, p.phase_b, p.coef_b, p.ndtc_b, p.mmax_b, gi, nn);
            double e_next[LT];
#pragma unroll
            for (int l = 0; l < LT; ++l) e_next[l] = p.amp_old[(size_t)l * N_T + nn];
#pragma unroll
            for (int s = 0; s <= W; ++s) g[s] = P[0][s];
#pragma unroll
            for (int l = 0; l < LT; ++l) {
#pragma unroll
                for (int s = 0; s <= W; ++s) {
                    g[s].x = fma(e_cur[l], P[l + 1][s].x, g[s].x);
                    g[s].y = fma(e_cur[l], P[l + 1][s].y, g[s].y);
                }
            }
            cheby_step2<W>(ca, cb, g, col, myps, bufs, meta.a, meta.m, meta.phase, lane);
            myps[lane] = ca;
            myps[32 + lane] = cb;
            __syncwarp();
            Xa[(size_t)n * 32 + lane] = ca;
            Xb[(size_t)n * 32 + lane] = cb;
            meta = meta_next;
#pragma unroll
            for (int l = 0; l < LT; ++l) e_cur[l] = e_next[l];
        }
    }
    const long long t_bw_end = clock64();

    // ================================================================ forward sweep
    double2 pa = make_double2(0.0, 0.0), pb = make_double2(0.0, 0.0);
    if (live) {
        pa = p.psi0[(size_t)ka * 32 + lane];
        pb = p.psi0[(size_t)kb * 32 + lane];
        if (p.store_fw) {
            p.Phi[(size_t)ka * (N_T + 1) * 32 + lane] = pa;
            p.Phi[(size_t)kb * (N_T + 1) * 32 + lane] = pb;
        }
        const double2 *Pg = p.Pf + (size_t)gi * NT * rowstride;
#pragma unroll
        for (int q = 0; q < NT; ++q) load_row<W>(Pg + q * rowstride, P[q], lane);
    }
    myps[lane] = pa;
    myps[32 + lane] = pb;
    __syncwarp();
    const double inv_s = live ? p.inv_s_f[gi] : 0.0;
    StepMeta fmeta = load_meta(p.dtc_f, p.m_f, p.phase_f, p.coef_f, p.ndtc_f, p.mmax_f, gi, 0);
    double2 chia = make_double2(0.0, 0.0), chib = make_double2(0.0, 0.0);
    if (p.mode == 1 && live) {
        chia = p.X[(size_t)ka * (N_T + 1) * 32 + lane];
        chib = p.X[(size_t)kb * (N_T + 1) * 32 + lane];
    }
    long long t_wait_b = 0, t_overlap = 0, t_step = 0;
    for (int n = 0; n < N_T; ++n) {
        double eps[LT];
        const long long ts0 = clock64();
        if (p.mode == 1) {
            double part[LT];
#pragma unroll
            for (int l = 0; l < LT; ++l) part[l] = 0.0;
            if (live) {
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    double war = 0.0, wai = 0.0, wbr = 0.0, wbi = 0.0;
                    row_dot2<W>(P[l + 1], col, myps, myps + 32, pa, pb, war, wai, wbr, wbi);
                    part[l] = fma(inv_s, fma(chia.x, war, chia.y * wai), part[l]);  // trajectory a, then b: fixed order
                    part[l] = fma(inv_s, fma(chib.x, wbr, chib.y * wbi), part[l]);
                }
            }
#pragma unroll
            for (int l = 0; l < LT; ++l) red[(size_t)l * wpc * 32 + warp * 32 + lane] = part[l];
            bar_arrive(1, nthr_all);
            const long long w0 = clock64();
            t_overlap += w0 - ts0;
            if (n + 1 < N_T && live) {
                chia = p.X[((size_t)ka * (N_T + 1) + n + 1) * 32 + lane];
                chib = p.X[((size_t)kb * (N_T + 1) + n + 1) * 32 + lane];
            }
            bar_sync(2, nthr_all);
#pragma unroll
            for (int l = 0; l < LT; ++l) eps[l] = eps_s[l];
            if (p.prof != nullptr) {
                asm volatile("" ::"d"(eps[0]) : "memory");
                t_wait_b += clock64() - w0;
            }
        } else {
#pragma unroll
            for (int l = 0; l < LT; ++l) eps[l] = p.amp_old[(size_t)l * N_T + n];
        }
        const StepMeta fmeta_next =
            load_meta(p.dtc_f, p.m_f, p.phase_f, p.coef_f, p.ndtc_f, p.mmax_f, gi, n + 1 < N_T ? n + 1 : n);
        if (live) {
#pragma unroll
            for (int s = 0; s <= W; ++s) g[s] = P[0][s];
#pragma unroll
            for (int l = 0; l < LT; ++l) {
#pragma unroll
                for (int s = 0; s <= W; ++s) {
                    g[s].x = fma(eps[l], P[l + 1][s].x, g[s].x);
                    g[s].y = fma(eps[l], P[l + 1][s].y, g[s].y);
                }
            }
            cheby_step2<W>(pa, pb, g, col, myps, bufs, fmeta.a, fmeta.m, fmeta.phase, lane);
            myps[lane] = pa;
            myps[32 + lane] = pb;
            __syncwarp();
            if (p.store_fw) {
                const int slot = (p.mode == 1) ? n : n + 1;  // sic, src/optimize.jl:367 vs :263
                p.Phi[((size_t)ka * (N_T + 1) + slot) * 32 + lane] = pa;
                p.Phi[((size_t)kb * (N_T + 1) + slot) * 32 + lane] = pb;
            }
        }
        fmeta = fmeta_next;
        if (p.prof != nullptr) {
            asm volatile("" ::"d"(pa.x) : "memory");
            t_step += clock64() - ts0;
        }
    }
    if (p.prof != nullptr && warp == 0 && lane == 0) {
        const long long t_end = clock64();
        p.prof[blockIdx.x * 8 + 0] = t_bw_end - t_begin;
        p.prof[blockIdx.x * 8 + 1] = t_end - t_bw_end;
        p.prof[blockIdx.x * 8 + 2] = t_wait_b;
        p.prof[blockIdx.x * 8 + 6] = t_overlap;
        p.prof[blockIdx.x * 8 + 7] = t_step;
    }
    // ---- final states and tau_k
    if (live) {
        p.psi_final[(size_t)ka * 32 + lane] = pa;
        p.psi_final[(size_t)kb * 32 + lane] = pb;
        double tar = 0.0, tai = 0.0, tbr = 0.0, tbi = 0.0;
        if (p.target != nullptr) {
            const double2 ta = p.target[(size_t)ka * 32 + lane], tb = p.target[(size_t)kb * 32 + lane];
            tar = ta.x * pa.x + ta.y * pa.y;
            tai = ta.x * pa.y - ta.y * pa.x;
            tbr = tb.x * pb.x + tb.y * pb.y;
            tbi = tb.x * pb.y - tb.y * pb.x;
        }
        tar = warp_sum_xor(tar);
        tai = warp_sum_xor(tai);
        tbr = warp_sum_xor(tbr);
        tbi = warp_sum_xor(tbi);
        if (lane == 0) {
            p.tau[ka] = make_double2(tar, tai);
            p.tau[kb] = make_double2(tbr, tbi);
        }
    }
}

}  // namespace kr
