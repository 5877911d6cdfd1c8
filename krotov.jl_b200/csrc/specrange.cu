// Host utility of libkrotov_cuda: extreme eigenvalues of a batch of small complex Hermitian matrices, threaded.
//
// This is the arithmetic behind the spectral envelope of the Chebyshev propagator (QuantumPropagators'
// `specrange(H; method=:diag)` evaluated at the corners of the control ranges, called through `reinit_prop!` at
// src/optimize.jl:251,306,324 of the reference whenever Krotov's range hook :238-244 fires).  An ensemble
// optimisation re-derives it for every ensemble member at once -- 2 x 256 matrices of 25 x 25 for BASELINE's C4 --
// and through NumPy that costs 30 ms per event (LAPACK one matrix at a time under the GIL) against a 14 ms
// iteration.  Here: Householder reduction to a real symmetric tridiagonal matrix, then bisection on the Sturm
// count for the smallest and the largest eigenvalue only (both backward stable: errors of a few ulp of ||A||),
// matrices spread over std::threads.  Pure host code; no device work.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/krotov_cuda.h"

namespace {

using cplx = std::complex<double>;

// Householder tridiagonalisation of the Hermitian matrix (ar + i ai) (n x n, row-major, destroyed): on exit
// dg = diagonal, e2 = squared moduli of the sub-diagonal (the spectrum of a Hermitian tridiagonal matrix depends
// on the sub-diagonal through its moduli only).  Real and imaginary parts are kept in separate arrays and the
// complex products are written out: std::complex multiplication goes through __muldc3.
void tridiagonalise(int n, double *__restrict__ ar, double *__restrict__ ai, double *__restrict__ dg,
                    double *__restrict__ e2, double *__restrict__ vr, double *__restrict__ vi, double *__restrict__ pr,
                    double *__restrict__ pi) {
    for (int k = 0; k + 2 < n; ++k) {
        const int m = n - k - 1;
        double xnorm2 = 0.0;
        for (int i = 0; i < m; ++i) {
            vr[i] = ar[(size_t)(k + 1 + i) * n + k];  // x = a[k+1.., k]
            vi[i] = ai[(size_t)(k + 1 + i) * n + k];
            xnorm2 += vr[i] * vr[i] + vi[i] * vi[i];
        }
        const double x0n2 = vr[0] * vr[0] + vi[0] * vi[0];
        if (xnorm2 - x0n2 <= 0.0) continue;  // already tridiagonal in this column
        const double xnorm = std::sqrt(xnorm2), ax0 = std::sqrt(x0n2);
        const double phr = ax0 > 0.0 ? vr[0] / ax0 : 1.0, phi = ax0 > 0.0 ? vi[0] / ax0 : 0.0;
        const double alr = -phr * xnorm, ali = -phi * xnorm;  // H x = alpha e_1, no cancellation in v_0 = x_0 - alpha
        vr[0] -= alr;
        vi[0] -= ali;
        double vnorm2 = 0.0;
        for (int i = 0; i < m; ++i) vnorm2 += vr[i] * vr[i] + vi[i] * vi[i];
        const double beta = 2.0 / vnorm2;
        // trailing block B = a[k+1.., k+1..]:  B <- H B H,  H = I - beta v v^H
        //   p = beta B v;  K = beta/2 v^H p (real for Hermitian B);  w = p - K v;  B -= v w^H + w v^H
        double vhp = 0.0;
        for (int i = 0; i < m; ++i) {
            const double *__restrict__ rr = ar + (size_t)(k + 1 + i) * n + (k + 1);
            const double *__restrict__ ri = ai + (size_t)(k + 1 + i) * n + (k + 1);
            double sr = 0.0, si = 0.0;
            for (int j = 0; j < m; ++j) {
                sr += rr[j] * vr[j] - ri[j] * vi[j];
                si += rr[j] * vi[j] + ri[j] * vr[j];
            }
            pr[i] = beta * sr;
            pi[i] = beta * si;
            vhp += vr[i] * pr[i] + vi[i] * pi[i];
        }
        const double K = 0.5 * beta * vhp;
        for (int i = 0; i < m; ++i) {  // p is w now
            pr[i] -= K * vr[i];
            pi[i] -= K * vi[i];
        }
        for (int i = 0; i < m; ++i) {
            double *__restrict__ rr = ar + (size_t)(k + 1 + i) * n + (k + 1);
            double *__restrict__ ri = ai + (size_t)(k + 1 + i) * n + (k + 1);
            const double vir = vr[i], vii = vi[i], wir = pr[i], wii = pi[i];
            for (int j = 0; j < m; ++j) {
                // v_i conj(w_j) + w_i conj(v_j)
                rr[j] -= vir * pr[j] + vii * pi[j] + wir * vr[j] + wii * vi[j];
                ri[j] -= vii * pr[j] - vir * pi[j] + wii * vr[j] - wir * vi[j];
            }
        }
        ar[(size_t)(k + 1) * n + k] = alr;
        ai[(size_t)(k + 1) * n + k] = ali;
        for (int i = 1; i < m; ++i) {
            ar[(size_t)(k + 1 + i) * n + k] = 0.0;
            ai[(size_t)(k + 1 + i) * n + k] = 0.0;
        }
    }
    for (int i = 0; i < n; ++i) dg[i] = ar[(size_t)i * n + i];
    for (int i = 0; i + 1 < n; ++i) {
        const double er = ar[(size_t)(i + 1) * n + i], ei = ai[(size_t)(i + 1) * n + i];
        e2[i] = er * er + ei * ei;
    }
}

// Smallest and largest eigenvalue of the real symmetric tridiagonal matrix (dg, e2) by multi-section on the Sturm
// count (the count of eigenvalues below x is the number of negative terms of q_0 = d_0 - x,
// q_i = d_i - x - e_{i-1}^2 / q_{i-1}, as in LAPACK's dstebz).  A count is a serial chain of n divisions, so each
// round evaluates kPts interior points of BOTH brackets as independent chains (the loop over points vectorises)
// and keeps the sub-interval that still holds the eigenvalue: 3 bits per round instead of 1.
constexpr int kPts = 7;

void extreme_eigenvalues(int n, const double *dg, const double *e2, double &emin, double &emax) {
    double glo = dg[0], ghi = dg[0], e2max = 0.0;
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? std::sqrt(e2[i - 1]) : 0.0) + (i + 1 < n ? std::sqrt(e2[i]) : 0.0);
        glo = std::min(glo, dg[i] - r);
        ghi = std::max(ghi, dg[i] + r);
        if (i + 1 < n) e2max = std::max(e2max, e2[i]);
    }
    const double tnorm = std::max(std::fabs(glo), std::fabs(ghi));
    const double eps = 2.220446049250313e-16, safemin = 2.2250738585072014e-308;
    const double pivmin = std::max(safemin * std::max(1.0, e2max), safemin);
    glo -= 2.0 * eps * tnorm * n + 2.0 * pivmin;
    ghi += 2.0 * eps * tnorm * n + 2.0 * pivmin;
    // bracket 0 holds the smallest eigenvalue (count(lo) = 0 < count(hi)), bracket 1 the largest
    double lo[2] = {glo, glo}, hi[2] = {ghi, ghi};
    const int want[2] = {0, n - 1};  // the eigenvalue with this many eigenvalues below it
    const double tol = eps * tnorm;
    for (int round = 0; round < 64; ++round) {
        if (hi[0] - lo[0] <= tol && hi[1] - lo[1] <= tol) break;
        double x[2 * kPts], q[2 * kPts];
        int cnt[2 * kPts];
        for (int b = 0; b < 2; ++b)
            for (int j = 0; j < kPts; ++j) x[b * kPts + j] = lo[b] + (hi[b] - lo[b]) * ((j + 1) / (double)(kPts + 1));
        for (int j = 0; j < 2 * kPts; ++j) {
            double t = dg[0] - x[j];
            if (std::fabs(t) < pivmin) t = -pivmin;
            q[j] = t;
            cnt[j] = t < 0.0;
        }
        for (int i = 1; i < n; ++i) {
            const double di = dg[i], ei = e2[i - 1];
            for (int j = 0; j < 2 * kPts; ++j) {
                double t = di - x[j] - ei / q[j];
                if (std::fabs(t) < pivmin) t = -pivmin;
                q[j] = t;
                cnt[j] += t < 0.0;
            }
        }
        for (int b = 0; b < 2; ++b) {
            // the wanted eigenvalue lies right of every point with count <= want and left of the first with count > want
            double nlo = lo[b], nhi = hi[b];
            for (int j = 0; j < kPts; ++j) {
                const double xj = x[b * kPts + j];
                if (!(xj > lo[b] && xj < hi[b])) continue;  // bracket exhausted to neighbouring floats
                if (cnt[b * kPts + j] > want[b]) {
                    nhi = std::min(nhi, xj);
                } else {
                    nlo = std::max(nlo, xj);
                }
            }
            lo[b] = nlo;
            hi[b] = nhi;
        }
    }
    emin = 0.5 * (lo[0] + hi[0]);
    emax = 0.5 * (lo[1] + hi[1]);
}

}  // namespace

// One worker: extremes of matrix (ar + i ai), destroying it.
static void extremes_of(int d, std::vector<double> &ar, std::vector<double> &ai, std::vector<double> &vr,
                        std::vector<double> &vi, std::vector<double> &pr, std::vector<double> &pi,
                        std::vector<double> &dg, std::vector<double> &e2, double &lo, double &hi) {
    if (d == 1) {
        lo = hi = ar[0];
        return;
    }
    tridiagonalise(d, ar.data(), ai.data(), dg.data(), e2.data(), vr.data(), vi.data(), pr.data(), pi.data());
    extreme_eigenvalues(d, dg.data(), e2.data(), lo, hi);
}

template <typename F>
static void run_threads(int n_items, int n_threads, F &&work) {
    // default: all hardware threads -- shared fairly when several ranks of one job run on this host (torchrun exports
    // LOCAL_WORLD_SIZE; with the replicated forward sweep every rank solves the same envelopes at the same moment, and
    // 8 x 32 threads on 32 cores cost more than 8 x 4).  KROTOV_HOST_THREADS overrides.
    int nt = n_threads;
    if (nt <= 0) {
        const char *e = getenv("KROTOV_HOST_THREADS");
        if (e && atoi(e) > 0) {
            nt = atoi(e);
        } else {
            const char *lw = getenv("LOCAL_WORLD_SIZE");
            nt = (int)std::thread::hardware_concurrency() / std::max(1, lw ? atoi(lw) : 1);
        }
    }
    nt = std::max(1, std::min(nt, std::min(n_items, 64)));
    if (nt == 1) {
        work(0, 1);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, t, nt);
    for (auto &th : pool) th.join();
}

extern "C" int krotov_envelope_extremes(int n_gen, int d, int n_ctrl, const double *H0, const double *Hc, int n_corner,
                                        const double *amps, double *e_min, double *e_max, int n_threads) {
    if (n_gen < 0 || d < 1 || n_ctrl < 0 || n_corner < 1 || !H0 || (n_ctrl > 0 && (!Hc || !amps)) || !e_min || !e_max)
        return KROTOV_ERR_ARG;
    if (n_gen == 0) return KROTOV_OK;
    const size_t dd = (size_t)d * d;
    run_threads(n_gen, n_threads, [&](int t, int nt) {
        std::vector<double> ar(dd), ai(dd), vr(d), vi(d), pr(d), pi(d), dg(d), e2(d);
        for (int g = t; g < n_gen; g += nt) {
            const cplx *h0 = reinterpret_cast<const cplx *>(H0) + (size_t)g * dd;
            double lo_all = 0.0, hi_all = 0.0;
            for (int c = 0; c < n_corner; ++c) {
                // G = H0 + sum_l amps[c][l] Hc[l], Hermitian part
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j < d; ++j) {
                        double xr = h0[(size_t)i * d + j].real() + h0[(size_t)j * d + i].real();
                        double xi = h0[(size_t)i * d + j].imag() - h0[(size_t)j * d + i].imag();
                        for (int l = 0; l < n_ctrl; ++l) {
                            const cplx *hc = reinterpret_cast<const cplx *>(Hc) + ((size_t)l * n_gen + g) * dd;
                            const double a = amps[(size_t)c * n_ctrl + l];
                            xr += a * (hc[(size_t)i * d + j].real() + hc[(size_t)j * d + i].real());
                            xi += a * (hc[(size_t)i * d + j].imag() - hc[(size_t)j * d + i].imag());
                        }
                        ar[(size_t)i * d + j] = 0.5 * xr;
                        ai[(size_t)i * d + j] = 0.5 * xi;
                    }
                double lo, hi;
                extremes_of(d, ar, ai, vr, vi, pr, pi, dg, e2, lo, hi);
                lo_all = c == 0 ? lo : std::min(lo_all, lo);
                hi_all = c == 0 ? hi : std::max(hi_all, hi);
            }
            e_min[g] = lo_all;
            e_max[g] = hi_all;
        }
    });
    return KROTOV_OK;
}

extern "C" int krotov_hermitian_extremes(int n_mat, int d, const double *mats, double *e_min, double *e_max,
                                         int n_threads) {
    if (n_mat < 0 || d < 1 || !mats || !e_min || !e_max) return KROTOV_ERR_ARG;
    if (n_mat == 0) return KROTOV_OK;
    run_threads(n_mat, n_threads, [&](int t, int nt) {
        std::vector<double> ar((size_t)d * d), ai((size_t)d * d), vr(d), vi(d), pr(d), pi(d), dg(d), e2(d);
        for (int q = t; q < n_mat; q += nt) {
            const cplx *src = reinterpret_cast<const cplx *>(mats) + (size_t)q * d * d;
            // Hermitian part of the input, symmetrised exactly so that row- and column-major callers agree
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    const cplx x = src[(size_t)i * d + j], y = src[(size_t)j * d + i];
                    ar[(size_t)i * d + j] = 0.5 * (x.real() + y.real());
                    ai[(size_t)i * d + j] = 0.5 * (x.imag() - y.imag());
                }
            extremes_of(d, ar, ai, vr, vi, pr, pi, dg, e2, e_min[q], e_max[q]);
        }
    });
    return KROTOV_OK;
}
