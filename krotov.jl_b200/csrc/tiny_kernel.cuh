// Register-resident Krotov kernel for tiny Hilbert spaces (d <= 4) with at most 32 trajectories, sm_100a.
//
// For a 2x2 or 3x3 generator the warp-per-trajectory kernel spends a Chebyshev term almost entirely in the
// shared-memory exchange of the state (STS -> __syncwarp -> LDS, ~57 cycles) and leaves 29 of 32 lanes idle.
// Here ONE THREAD owns a whole trajectory: state, Chebyshev vectors and the dense generator G = 2c (H - beta)
// live in registers, a term is D*D complex multiply-adds with no exchange at all, and the only communication left
// is the per-time-step sum of the overlaps over the trajectories -- a warp butterfly over ceil(log2 N) levels
// (nothing for a single trajectory).  The whole iteration (src/optimize.jl:279-371) is one launch of one warp;
// same algorithm, same HBM layout (records of 32 entries) as krotov_warp_kernel, so storage read-back and the
// host side are shared.  The costates needed by the forward sweep are prefetched 2-4 time steps ahead (they
// come back from L2), the per-step scalars one step ahead.
#pragma once
#include "warp_kernel.cuh"

namespace kr {

struct TinyParams {
    WarpParams w;           // shares every array with the warp path
    const double2 *Tf, *Tb; // dense prepared terms [g][1+L][D*D] row-major: 2c (H_t - beta delta_t0), forward / adjoint
    int coef_in_smem;       // the coefficient rows of every thread's generator fit in shared memory
};


// y = M v + add   (complex; four independent FMA chains per row).  GIMAG: every entry of M is purely imaginary --
// the generator 2c (H - beta) of a REAL Hamiltonian (two-level systems, Duffing oscillators in the rotating frame
// with real drives ...) -- so half of the products vanish identically and are not issued: the kernel is bound by
// FP64 issue slots (a warp instruction costs 2 cycles however few lanes are live), this halves them.
template <int D, bool GIMAG>
__device__ __forceinline__ void tiny_matvec(double2 (&y)[D], const double2 (&M)[D][D], const double2 (&v)[D],
                                            const double2 (&add)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        if (GIMAG) {  // (i g) (vr + i vi) = -g vi + i g vr: two chains seeded with `add`, no joining add
            double br = add[i].x, bi = add[i].y;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                br = fma(-M[i][j].y, v[j].y, br);
                bi = fma(M[i][j].y, v[j].x, bi);
            }
            y[i] = make_double2(br, bi);
        } else {
            double ar = add[i].x, ai = add[i].y, br = 0.0, bi = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                ar = fma(M[i][j].x, v[j].x, ar);
                ai = fma(M[i][j].x, v[j].y, ai);
                br = fma(-M[i][j].y, v[j].y, br);
                bi = fma(M[i][j].y, v[j].x, bi);
            }
            y[i] = make_double2(ar + br, ai + bi);
        }
    }
}

// One Chebyshev step (same recursion and operation order as cheby_step of the warp kernel).
template <int D, bool GIMAG>
__device__ __forceinline__ void tiny_cheby(double2 (&psi)[D], const double2 (&G)[D][D], const double *a, const int astride,
                                           const int m, const double2 phase) {
    double2 v0[D], v1[D], v2[D], out[D], zero[D];
    const double a0 = a[0];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        v0[i] = psi[i];
        zero[i] = make_double2(0.0, 0.0);
        out[i] = make_double2(a0 * psi[i].x, a0 * psi[i].y);
    }
    tiny_matvec<D, GIMAG>(v1, G, v0, zero);
    const double a1 = m > 1 ? a[astride] : 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        v1[i] = make_double2(0.5 * v1[i].x, 0.5 * v1[i].y);
        out[i].x = fma(a1, v1[i].x, out[i].x);
        out[i].y = fma(a1, v1[i].y, out[i].y);
    }
    for (int j = 2; j < m; ++j) {
        const double aj = a[(size_t)j * astride];
        tiny_matvec<D, GIMAG>(v2, G, v1, v0);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            out[i].x = fma(aj, v2[i].x, out[i].x);
            out[i].y = fma(aj, v2[i].y, out[i].y);
            v0[i] = v1[i];
            v1[i] = v2[i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
        psi[i] = make_double2(phase.x * out[i].x - phase.y * out[i].y, phase.x * out[i].y + phase.y * out[i].x);
}

// The (1+LT) prepared terms of a thread's generator: in registers when they fit, else in shared memory laid out
// [entry][lane] (conflict-free).
template <int D, int LT, bool TREG>
struct TinyTerms {
    double2 r[TREG ? (1 + LT) * D * D : 1];
    const double2 *s;  // shared memory, entry e of this lane at s[e * 32]
    __device__ __forceinline__ void load(const double2 *__restrict__ src, double2 *smem, const int lane, const bool live) {
        s = smem + lane;
#pragma unroll
        for (int e = 0; e < (1 + LT) * D * D; ++e) {
            const double2 v = live ? src[e] : make_double2(0.0, 0.0);
            if (TREG)
                r[e] = v;
            else
                smem[e * 32 + lane] = v;
        }
        __syncwarp();
    }
    __device__ __forceinline__ double2 get(const int e) const { return TREG ? r[e] : s[e * 32]; }
    template <bool GIMAG>
    __device__ __forceinline__ void build(double2 (&G)[D][D], const double (&eps)[LT]) const {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double2 g = get(i * D + j);
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    const double2 t = get((l + 1) * D * D + i * D + j);
                    if (!GIMAG) g.x = fma(eps[l], t.x, g.x);
                    g.y = fma(eps[l], t.y, g.y);
                }
                G[i][j] = g;
            }
    }
};

struct TinyMeta {
    int m;
    double2 phase;
    const double *a;
    int astride;
};
// coefficient row of (generator, dt class): from the shared-memory copy ([class][j][lane]) or from global memory
__device__ __forceinline__ TinyMeta tiny_meta(const int *dtc, const int *m_tab, const double2 *ph_tab, const double *coef,
                                              const double *coef_s, int ndtc, int mmax, int gi, int n, int lane) {
    const int c = dtc[n];
    const int ci = gi * ndtc + c;
    TinyMeta s;
    s.m = m_tab[ci];
    s.phase = ph_tab[ci];
    if (coef_s != nullptr) {
        s.a = coef_s + (size_t)c * mmax * 32 + lane;
        s.astride = 32;
    } else {
        s.a = coef + (size_t)ci * mmax;
        s.astride = 1;
    }
    return s;
}

template <int D, int LT, bool TREG, bool GIMAG>
__global__ void __launch_bounds__(32, 1) krotov_tiny_kernel(const __grid_constant__ TinyParams tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const WarpParams &p = tp.w;
    const int k = threadIdx.x;  // one thread per trajectory
    const bool live = k < p.N;
    const int N_T = p.N_T;
    const int gi = live ? p.gen_of_traj[k] : 0;
    constexpr int TE = (1 + LT) * D * D;
    constexpr int kTinyPF = (D == 2) ? 4 : 2;  // prefetch distance of the costates (time steps; they return from L2)
    double2 *term_s = reinterpret_cast<double2 *>(smem_raw);                  // [TE][32] (unused when TREG)
    double *coef_fs = reinterpret_cast<double *>(term_s + (TREG ? 0 : TE * 32));  // [ndtc_f][mmax_f][32]
    double *coef_bs = coef_fs + (size_t)p.ndtc_f * p.mmax_f * 32;             // [ndtc_b][mmax_b][32]
    if (tp.coef_in_smem) {
        for (int c = 0; c < p.ndtc_f; ++c)
            for (int j = 0; j < p.mmax_f; ++j)
                coef_fs[((size_t)c * p.mmax_f + j) * 32 + k] = p.coef_f[((size_t)gi * p.ndtc_f + c) * p.mmax_f + j];
        if (p.mode == 1)
            for (int c = 0; c < p.ndtc_b; ++c)
                for (int j = 0; j < p.mmax_b; ++j)
                    coef_bs[((size_t)c * p.mmax_b + j) * 32 + k] = p.coef_b[((size_t)gi * p.ndtc_b + c) * p.mmax_b + j];
        __syncwarp();
    } else {
        coef_fs = nullptr;
        coef_bs = nullptr;
    }
    TinyTerms<D, LT, TREG> T;
    double2 G[D][D];
    double eps[LT];

    // ================================================================ backward sweep  (src/optimize.jl:303-317)
    if (p.mode == 1) {
        T.load(tp.Tb + (size_t)gi * TE, term_s, k, live);
        double2 chi[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            chi[i] = make_double2(0.0, 0.0);
            if (live) {
                if (p.chiT != nullptr) {
                    chi[i] = p.chiT[(size_t)k * 32 + i];
                } else {
                    const double2 c = p.chi_coef[k], tg = p.target[(size_t)k * 32 + i];
                    chi[i] = make_double2(c.x * tg.x - c.y * tg.y, c.x * tg.y + c.y * tg.x);
                }
            }
        }
        double2 *Xk = p.X + (size_t)(live ? k : 0) * (N_T + 1) * 32;
        if (live) {
#pragma unroll
            for (int i = 0; i < D; ++i) Xk[(size_t)N_T * 32 + i] = chi[i];
        }
        TinyMeta meta = tiny_meta(p.dtc_b, p.m_b, p.phase_b, p.coef_b, coef_bs, p.ndtc_b, p.mmax_b, gi, N_T - 1, k);
        double e_cur[LT];
#pragma unroll
        for (int l = 0; l < LT; ++l) e_cur[l] = p.eps_old[(size_t)l * N_T + N_T - 1];
        for (int n = N_T - 1; n >= 0; --n) {
            const int nn = n > 0 ? n - 1 : 0;  // next step's scalars, one step ahead
            const TinyMeta meta_next = tiny_meta(p.dtc_b, p.m_b, p.phase_b, p.coef_b, coef_bs, p.ndtc_b, p.mmax_b, gi, nn, k);
            double e_next[LT];
#pragma unroll
            for (int l = 0; l < LT; ++l) e_next[l] = p.eps_old[(size_t)l * N_T + nn];
            T.template build<GIMAG>(G, e_cur);
            tiny_cheby<D, GIMAG>(chi, G, meta.a, meta.astride, meta.m, meta.phase);
            if (live) {
#pragma unroll
                for (int i = 0; i < D; ++i) Xk[(size_t)n * 32 + i] = chi[i];
            }
            meta = meta_next;
#pragma unroll
            for (int l = 0; l < LT; ++l) e_cur[l] = e_next[l];
        }
        __syncwarp();
    }

    // ================================================================ forward sweep  (:321-368, or :247-265 for mode 0)
    T.load(tp.Tf + (size_t)gi * TE, term_s, k, live);
    double2 psi[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        psi[i] = live ? p.psi0[(size_t)k * 32 + i] : make_double2(0.0, 0.0);
        if (live && p.store_fw) p.Phi[(size_t)k * (N_T + 1) * 32 + i] = psi[i];
    }
    const double inv_s = live ? p.inv_s_f[gi] : 0.0;
    int levels = 0;
    while ((1 << levels) < p.N) ++levels;
    double ga[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) ga[l] = 0.0;
    const double2 *Xk = p.X + (size_t)(live ? k : 0) * (N_T + 1) * 32;
    // costate ring: ring[r] holds chi(t_n) for the step n = r (mod kTinyPF) that comes next
    double2 ring[kTinyPF][D];
#pragma unroll
    for (int r = 0; r < kTinyPF; ++r)
#pragma unroll
        for (int i = 0; i < D; ++i)
            ring[r][i] = (p.mode == 1 && live && r < N_T) ? Xk[(size_t)r * 32 + i] : make_double2(0.0, 0.0);
    TinyMeta meta = tiny_meta(p.dtc_f, p.m_f, p.phase_f, p.coef_f, coef_fs, p.ndtc_f, p.mmax_f, gi, 0, k);
    double eo_cur[LT], al_cur[LT], dt_cur = p.dt[0];
#pragma unroll
    for (int l = 0; l < LT; ++l) {
        eo_cur[l] = p.eps_old[(size_t)l * N_T];
        al_cur[l] = p.mode == 1 ? p.alpha[(size_t)l * N_T] : 0.0;
    }
    for (int n0 = 0; n0 < N_T; n0 += kTinyPF) {
#pragma unroll
        for (int r = 0; r < kTinyPF; ++r) {
            const int n = n0 + r;
            if (n >= N_T) break;
            const int nn = n + 1 < N_T ? n + 1 : n;
            const TinyMeta meta_next = tiny_meta(p.dtc_f, p.m_f, p.phase_f, p.coef_f, coef_fs, p.ndtc_f, p.mmax_f, gi, nn, k);
            double eo_next[LT], al_next[LT];
            const double dt_next = p.dt[nn];
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                eo_next[l] = p.eps_old[(size_t)l * N_T + nn];
                al_next[l] = p.mode == 1 ? p.alpha[(size_t)l * N_T + nn] : 0.0;
            }
            if (p.mode == 1) {
                double2 chi[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    chi[i] = ring[r][i];
                    if (live && n + kTinyPF < N_T) ring[r][i] = Xk[(size_t)(n + kTinyPF) * 32 + i];
                }
                // ---- overlaps Im <chi_k| mu_l |psi_k>  (:339-349), summed over the trajectories in a fixed order
#pragma unroll
                for (int l = 0; l < LT; ++l) {
                    double part = 0.0;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        double wr = 0.0, wi = 0.0, wr1 = 0.0, wi1 = 0.0;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            const double2 t = T.get((l + 1) * D * D + i * D + j);
                            if (!GIMAG) {
                                wr = fma(t.x, psi[j].x, wr);
                                wi = fma(t.x, psi[j].y, wi);
                            }
                            wr1 = fma(-t.y, psi[j].y, wr1);
                            wi1 = fma(t.y, psi[j].x, wi1);
                        }
                        part = fma(chi[i].x, wr + wr1, fma(chi[i].y, wi + wi1, part));
                    }
                    double du = inv_s * part;  // P_l = -i s mu_l  =>  Im<chi|mu_l|psi> = Re<chi|P_l psi> / s
                    for (int o = 0; o < levels; ++o) du += __shfl_xor_sync(0xffffffffu, du, 1 << o);
                    du = __shfl_sync(0xffffffffu, du, 0);
                    const double e_new = __dadd_rn(eo_cur[l], __dmul_rn(al_cur[l], du));  // :355-356
                    eps[l] = e_new;
                    if (k == 0) {
                        p.eps_new[(size_t)l * N_T + n] = e_new;
                        ga[l] = __dadd_rn(ga[l], __dmul_rn(__dmul_rn(al_cur[l], __dmul_rn(fabs(du), fabs(du))), dt_cur));  // :357
                    }
                }
            } else {
#pragma unroll
                for (int l = 0; l < LT; ++l) eps[l] = eo_cur[l];
            }
            // ---- forward step with the (updated) pulse value  (:360-368)
            T.template build<GIMAG>(G, eps);
            tiny_cheby<D, GIMAG>(psi, G, meta.a, meta.astride, meta.m, meta.phase);
            if (live && p.store_fw) {
                const int slot = (p.mode == 1) ? n : n + 1;  // sic, src/optimize.jl:367 vs :263
#pragma unroll
                for (int i = 0; i < D; ++i) p.Phi[((size_t)k * (N_T + 1) + slot) * 32 + i] = psi[i];
            }
            meta = meta_next;
            dt_cur = dt_next;
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                eo_cur[l] = eo_next[l];
                al_cur[l] = al_next[l];
            }
        }
    }
    if (p.mode == 1 && k == 0) {
#pragma unroll
        for (int l = 0; l < LT; ++l) p.g_a_int[l] = ga[l];
    }
    // ---- final states and tau_k = <tgt_k|psi_k(T)>  (:378-381)
    if (live) {
        double tr = 0.0, ti = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            p.psi_final[(size_t)k * 32 + i] = psi[i];
            if (p.target != nullptr) {
                const double2 tg = p.target[(size_t)k * 32 + i];
                tr += tg.x * psi[i].x + tg.y * psi[i].y;
                ti += tg.x * psi[i].y - tg.y * psi[i].x;
            }
        }
        p.tau[k] = make_double2(tr, ti);
    }
}

}  // namespace kr
