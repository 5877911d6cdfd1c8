/*
 * CPU oracle for the Krotov iteration, C restatement  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libkrotov_cuda) never links or calls it.
 *
 * PARITY UNPINNED: the reference (JuliaQuantumControl/Krotov.jl) is Julia and cannot be run in
 * this image; its tests hold no numeric golden vector for this path.  This file follows
 *   src/optimize.jl:247-265  (krotov_initial_fw_prop!)        -> initial_fw_prop()
 *   src/optimize.jl:279-371  (krotov_iteration)               -> krotov_iteration()
 *   src/optimize.jl:238-244  (transform_control_ranges)       -> transform_ranges()
 *   src/optimize.jl:374-386  (update_result!: taus, J_T)      -> update_result()
 * and restates the QuantumPropagators.jl `Cheby` piecewise propagator that the reference
 * calls at src/optimize.jl:251,257,306,309,324,361 (init_prop / reinit_prop! / prop_step!)
 * from its published algorithm.  It is checked against the NumPy
 * twin (krotov_oracle.py) in tests/test_oracle.py.
 *
 * Parallel structure mirrors the reference's `@threadsif` (src/optimize.jl:182,303,321,360):
 * OpenMP over trajectories in the backward sweep, the re-arm loop and the forward step; the
 * overlap accumulation (src/optimize.jl:340-349) stays serial, l outer / k inner.
 *
 * Generators: one CSR matrix per (generator g, term t), t = 0 drift, t = 1..L controls; a
 * term with rowptr[d] == 0 nonzeros is "nothing" (src/optimize.jl:344).  The lazy operator
 * applies one term at a time, drift first, like Operator `mul!`.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

typedef struct {
    int d;
    const int *rowptr; /* d+1 */
    const int *col;
    const cplx *val;
    int nnz;
} csr_t;

typedef struct {
    int backward;
    int g;
    double *rng_min, *rng_max; /* control ranges, L */
    double E_min, Delta, dt;
    int m;
    double *coef;
    cplx *state;
    int n; /* time-grid index of the state */
} prop_t;

typedef struct {
    int d, N, L, N_T, n_gen;
    const double *tlist;
    const int *gen_of_traj;
    csr_t *fw_terms; /* n_gen*(1+L) */
    csr_t *bw_terms; /* adjoint terms */
    int *bw_rowptr, *bw_col;
    cplx *bw_val;
    int has_explicit_range;
    double ex_Emin, ex_Emax;
    double limit, buffer;
    /* non-linear amplitudes a_l(eps, n) = shape[l][n] * sum_p poly[l][p] eps^p (src/optimize.jl:268-272); NULL = linear */
    int amp_deg;
    const double *amp_poly;  /* [L][amp_deg+1] */
    const double *amp_shape; /* [L][N_T] */
    double *amp_shape_max;   /* [L] largest |shape| (spectral envelope) */
    /* spectral-envelope cache per (direction, generator): key = control ranges */
    double *cache_key; /* 2*n_gen*(2L) */
    double *cache_val; /* 2*n_gen*2 */
    int *cache_ok;
} ctx_t;

/* ---------------- small dense Hermitian eigenvalue range (cyclic Jacobi on the real embedding) */
static void herm_eig_range(const cplx *A, int d, double *emin, double *emax) {
    int n = 2 * d;
    double *M = (double *)calloc((size_t)n * n, sizeof(double));
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            double re = 0.5 * (creal(A[i * d + j]) + creal(A[j * d + i]));
            double im = 0.5 * (cimag(A[i * d + j]) - cimag(A[j * d + i]));
            M[i * n + j] = re;
            M[(i + d) * n + (j + d)] = re;
            M[(i + d) * n + j] = im;
            M[i * n + (j + d)] = -im;
        }
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++) {
            diag += M[i * n + i] * M[i * n + i];
            for (int j = i + 1; j < n; j++) off += M[i * n + j] * M[i * n + j];
        }
        if (off <= 1e-32 * (diag + 1e-300)) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = M[p * n + q];
                if (fabs(apq) < 1e-300) continue;
                double theta = (M[q * n + q] - M[p * n + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = M[k * n + p], akq = M[k * n + q];
                    M[k * n + p] = c * akp - s * akq;
                    M[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = M[p * n + k], aqk = M[q * n + k];
                    M[p * n + k] = c * apk - s * aqk;
                    M[q * n + k] = s * apk + c * aqk;
                }
            }
    }
    double lo = M[0], hi = M[0];
    for (int i = 1; i < n; i++) {
        double v = M[i * n + i];
        if (v < lo) lo = v;
        if (v > hi) hi = v;
    }
    *emin = lo;
    *emax = hi;
    free(M);
}

static double amp_poly_value(const ctx_t *c, int l, double eps) {
    if (!c->amp_poly) return eps;
    const double *q = c->amp_poly + (size_t)l * (c->amp_deg + 1);
    double v = 0.0;
    for (int p = c->amp_deg; p >= 0; p--) v = v * eps + q[p];
    return v;
}
/* coefficient of H_l on interval n for the control value eps */
static double amp_value(const ctx_t *c, int l, int n, double eps) {
    double v = amp_poly_value(c, l, eps);
    return c->amp_shape ? c->amp_shape[(size_t)l * c->N_T + n] * v : v;
}
/* d a_l / d eps: the factor of H_l in mu_l = dH / d eps_l (evaluated at the guess pulse, src/optimize.jl:337) */
static double amp_deriv(const ctx_t *c, int l, int n, double eps) {
    double v = 1.0;
    if (c->amp_poly) {
        const double *q = c->amp_poly + (size_t)l * (c->amp_deg + 1);
        v = 0.0;
        for (int p = c->amp_deg; p >= 1; p--) v = v * eps + p * q[p];
    }
    return c->amp_shape ? c->amp_shape[(size_t)l * c->N_T + n] * v : v;
}
static int amp_nonlinear(const ctx_t *c) { return c->amp_poly != NULL || c->amp_shape != NULL; }

static void evaluate_dense(const ctx_t *c, const csr_t *terms, int g, const double *vals, cplx *G) {
    int d = c->d;
    memset(G, 0, sizeof(cplx) * d * d);
    for (int t = 0; t <= c->L; t++) {
        const csr_t *T = &terms[g * (1 + c->L) + t];
        double cf = 1.0;
        if (t > 0) { /* range corner: polynomial at the corner times the largest |shape| (unpinned convention) */
            cf = amp_poly_value(c, t - 1, vals[t - 1]);
            if (c->amp_shape) cf *= c->amp_shape_max[t - 1];
        }
        for (int i = 0; i < d; i++)
            for (int q = T->rowptr[i]; q < T->rowptr[i + 1]; q++) G[i * d + T->col[q]] += cf * T->val[q];
    }
}

static int g_oracle_diverged = 0;
/* amplitudes of the NEXT oracle_krotov_optimize call (set by oracle_set_amplitudes, consumed by the call) */
static int g_amp_deg = 0;
static const double *g_amp_poly = NULL, *g_amp_shape = NULL; /* a pulse grew beyond anything a polynomial can be derived for */

static int cheby_coeffs(double Delta, double dt, double limit, double **out) {
    double alpha = fabs(0.5 * Delta * dt);
    if (!(alpha < 1e6)) { /* NaN / Inf / absurd spectral radius: a diverged optimisation, not a crash */
        double *z = (double *)malloc(sizeof(double));
        z[0] = NAN;
        *out = z;
        g_oracle_diverged = 1;
        return 1;
    }
    int cap = (int)(alpha * 1.5) + 64, m = 0;
    double *a = (double *)malloc(sizeof(double) * cap);
    a[m++] = jn(0, alpha);
    double eps = fabs(a[0]);
    int i = 1;
    while (eps > limit || i <= alpha) {
        if (m == cap) {
            cap *= 2;
            a = (double *)realloc(a, sizeof(double) * cap);
        }
        a[m] = 2.0 * jn(i, alpha);
        eps = fabs(a[m]);
        m++;
        i++;
    }
    *out = a;
    return m;
}

/* src/optimize.jl:238-244 */
static void transform_ranges(double lo, double hi, int check, double *olo, double *ohi) {
    double f = check ? 2.0 : 5.0;
    *olo = fmin(lo, f * lo);
    *ohi = fmax(hi, f * hi);
}

static void set_envelope(ctx_t *c, prop_t *p) {
    double E_min, E_max;
    if (c->has_explicit_range) {
        E_min = c->ex_Emin;
        E_max = c->ex_Emax;
    } else {
        int L = c->L;
        int slot = (p->backward ? c->n_gen : 0) + p->g;
        double *key = &c->cache_key[slot * 2 * L];
        int hit = 0;
#pragma omp critical(envcache)
        {
            if (c->cache_ok[slot]) {
                hit = 1;
                for (int l = 0; l < L; l++)
                    if (key[l] != p->rng_min[l] || key[L + l] != p->rng_max[l]) hit = 0;
            }
            if (hit) {
                E_min = c->cache_val[slot * 2];
                E_max = c->cache_val[slot * 2 + 1];
            }
        }
        if (!hit) {
            int d = c->d;
            cplx *G = (cplx *)malloc(sizeof(cplx) * d * d);
            const csr_t *terms = p->backward ? c->bw_terms : c->fw_terms;
            double lo1, hi1, lo2, hi2;
            evaluate_dense(c, terms, p->g, p->rng_max, G);
            herm_eig_range(G, d, &lo1, &hi1);
            evaluate_dense(c, terms, p->g, p->rng_min, G);
            herm_eig_range(G, d, &lo2, &hi2);
            free(G);
            E_min = fmin(lo1, lo2);
            E_max = fmax(hi1, hi2);
#pragma omp critical(envcache)
            {
                for (int l = 0; l < L; l++) {
                    key[l] = p->rng_min[l];
                    key[L + l] = p->rng_max[l];
                }
                c->cache_val[slot * 2] = E_min;
                c->cache_val[slot * 2 + 1] = E_max;
                c->cache_ok[slot] = 1;
            }
        }
    }
    double Delta = E_max - E_min;
    double delta = c->buffer * Delta;
    p->E_min = E_min - delta / 2;
    p->Delta = Delta + delta;
    p->dt = c->tlist[1] - c->tlist[0];
    if (p->backward) p->dt = -p->dt;
    free(p->coef);
    p->m = cheby_coeffs(p->Delta, p->dt, c->limit, &p->coef);
}

static void pulse_minmax(const double *e, int n, double *lo, double *hi) {
    double a = e[0], b = e[0];
    for (int i = 1; i < n; i++) {
        if (e[i] < a) a = e[i];
        if (e[i] > b) b = e[i];
    }
    *lo = a;
    *hi = b;
}

static void init_prop(ctx_t *c, prop_t *p, int g, int backward, const double *pulses) {
    p->backward = backward;
    p->g = g;
    p->rng_min = (double *)malloc(sizeof(double) * c->L);
    p->rng_max = (double *)malloc(sizeof(double) * c->L);
    p->state = (cplx *)calloc(c->d, sizeof(cplx));
    p->coef = NULL;
    for (int l = 0; l < c->L; l++) pulse_minmax(pulses + (size_t)l * c->N_T, c->N_T, &p->rng_min[l], &p->rng_max[l]);
    set_envelope(c, p);
}

/* reinit_prop!(propagator, state; transform_control_ranges) */
static void reinit_prop(ctx_t *c, prop_t *p, const cplx *state, const double *pulses) {
    memcpy(p->state, state, sizeof(cplx) * c->d);
    p->n = p->backward ? c->N_T : 0;
    int need = 0;
    for (int l = 0; l < c->L; l++) {
        double lo, hi, clo, chi;
        pulse_minmax(pulses + (size_t)l * c->N_T, c->N_T, &lo, &hi);
        transform_ranges(lo, hi, 1, &clo, &chi);
        if (clo < p->rng_min[l] || chi > p->rng_max[l]) need = 1;
    }
    if (need) {
        for (int l = 0; l < c->L; l++) {
            double lo, hi;
            pulse_minmax(pulses + (size_t)l * c->N_T, c->N_T, &lo, &hi);
            transform_ranges(lo, hi, 0, &p->rng_min[l], &p->rng_max[l]);
        }
        set_envelope(c, p);
    }
}

/* lazy Operator mul!: out = sum_t coeff_t * (A_t v), one term at a time, drift first */
static void apply_op(const ctx_t *c, const csr_t *terms, int g, const double *eps_n, const cplx *v, cplx *out,
                     cplx *tmp) {
    int d = c->d;
    memset(out, 0, sizeof(cplx) * d);
    for (int t = 0; t <= c->L; t++) {
        const csr_t *T = &terms[g * (1 + c->L) + t];
        if (T->nnz == 0) continue;
        double cf = (t == 0) ? 1.0 : eps_n[t - 1];
        for (int i = 0; i < d; i++) {
            cplx s = 0;
            for (int q = T->rowptr[i]; q < T->rowptr[i + 1]; q++) s += T->val[q] * v[T->col[q]];
            tmp[i] = s;
        }
        for (int i = 0; i < d; i++) out[i] += cf * tmp[i];
    }
}

/* prop_step!(ChebyPropagator); work = 5*d scratch */
static void prop_step(ctx_t *c, prop_t *p, const double *pulses, cplx *work) {
    int d = c->d, L = c->L;
    int ni;
    double dt;
    if (p->backward) {
        ni = p->n - 1;
        dt = -(c->tlist[p->n] - c->tlist[p->n - 1]);
    } else {
        ni = p->n;
        dt = c->tlist[p->n + 1] - c->tlist[p->n];
    }
    if (fabs(dt - p->dt) > 1e-12 * fmax(1.0, fabs(p->dt))) {
        p->dt = dt;
        free(p->coef);
        p->m = cheby_coeffs(p->Delta, dt, c->limit, &p->coef);
    }
    double eps_n[16];
    for (int l = 0; l < L; l++) eps_n[l] = amp_value(c, l, ni, pulses[(size_t)l * c->N_T + ni]);
    const csr_t *terms = p->backward ? c->bw_terms : c->fw_terms;
    cplx *v0 = work, *v1 = work + d, *v2 = work + 2 * d, *psi = work + 3 * d, *tmp = work + 4 * d;
    const double *a = p->coef;
    double Delta = p->Delta, beta = Delta / 2 + p->E_min;
    cplx cc = -2.0 * I / Delta;
    if (dt < 0) cc = -cc;
    memcpy(v0, p->state, sizeof(cplx) * d);
    for (int i = 0; i < d; i++) psi[i] = a[0] * v0[i];
    apply_op(c, terms, p->g, eps_n, v0, v1, tmp);
    for (int i = 0; i < d; i++) v1[i] = cc * (v1[i] - beta * v0[i]);
    if (p->m > 1)
        for (int i = 0; i < d; i++) psi[i] += a[1] * v1[i];
    cplx c2 = 2.0 * cc;
    for (int j = 2; j < p->m; j++) {
        apply_op(c, terms, p->g, eps_n, v1, v2, tmp);
        for (int i = 0; i < d; i++) v2[i] = c2 * (v2[i] - beta * v1[i]) + v0[i];
        for (int i = 0; i < d; i++) psi[i] += a[j] * v2[i];
        cplx *t = v0;
        v0 = v1;
        v1 = v2;
        v2 = t;
    }
    cplx ph = cexp(-I * beta * dt);
    for (int i = 0; i < d; i++) p->state[i] = ph * psi[i];
    p->n += p->backward ? -1 : 1;
}

static double J_T_value(int kind, const cplx *tau, const double *w, int N) {
    if (kind == 0) { /* sm */
        cplx F = 0;
        for (int k = 0; k < N; k++) F += w[k] * tau[k];
        F /= N;
        return 1.0 - cabs(F) * cabs(F);
    } else if (kind == 1) { /* ss */
        double F = 0;
        for (int k = 0; k < N; k++) F += w[k] * cabs(tau[k]) * cabs(tau[k]);
        return 1.0 - F / N;
    } else { /* re */
        cplx F = 0;
        for (int k = 0; k < N; k++) F += w[k] * tau[k];
        return 1.0 - creal(F / N);
    }
}

static void chi_states(int kind, const cplx *tau, const double *w, const cplx *target, int N, int d, cplx *chi) {
    if (kind == 0) {
        cplx s = 0;
        for (int k = 0; k < N; k++) s += w[k] * tau[k];
        for (int k = 0; k < N; k++)
            for (int i = 0; i < d; i++) chi[(size_t)k * d + i] = (w[k] / ((double)N * N)) * s * target[(size_t)k * d + i];
    } else if (kind == 1) {
        for (int k = 0; k < N; k++)
            for (int i = 0; i < d; i++) chi[(size_t)k * d + i] = (w[k] / N) * tau[k] * target[(size_t)k * d + i];
    } else {
        for (int k = 0; k < N; k++)
            for (int i = 0; i < d; i++) chi[(size_t)k * d + i] = (w[k] / (2.0 * N)) * target[(size_t)k * d + i];
    }
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/*
 * Run `iters` Krotov iterations.  All complex arrays are interleaved (re, im) doubles.
 *   rowptr : [n_gen*(1+L)][d+1] ints, offsets relative to each term's own col/val block
 *   term_off : [n_gen*(1+L)+1] ints, start of each term's block in col/val
 *   pulses : [L][N_T] guess pulses (in), optimised pulses (out)
 *   out_JT : [iters+1]; out_ga : [iters][L]; out_tau : [N] complex; out_states : [N][d] complex
 *   out_m : [2] Chebyshev coefficient counts (fw, bw) of trajectory 0 after the last iteration
 *   out_secs : wall seconds of the iteration loop only (no init, no initial forward sweep)
 * Returns 0, or a negative error code.
 */
int oracle_krotov_optimize(int d, int N, int L, int N_T, int n_gen, const double *tlist, const int *gen_of_traj,
                           const int *rowptr, const int *term_off, const int *col, const double *val,
                           const double *psi0, const double *target, const double *weight, double *pulses,
                           const double *S, const double *lam, int functional, double cheby_limit,
                           double specrange_buffer, int has_range, double E_min, double E_max, int iters,
                           int n_threads, double *out_JT, double *out_ga, double *out_tau, double *out_states,
                           int *out_m, double *out_secs) {
    if (L > 16) return -1;
    g_oracle_diverged = 0;
    const int amp_deg = g_amp_deg;
    const double *amp_poly = g_amp_poly, *amp_shape = g_amp_shape;
    g_amp_poly = NULL;
    g_amp_shape = NULL;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    /* `use_threads` of the reference (src/workspace.jl:66): worth it only with several trajectories */
    const int use_threads = (N >= 8);
    ctx_t c;
    memset(&c, 0, sizeof(c));
    c.d = d;
    c.N = N;
    c.L = L;
    c.N_T = N_T;
    c.n_gen = n_gen;
    c.amp_deg = amp_deg;
    c.amp_poly = amp_poly;
    c.amp_shape = amp_shape;
    c.amp_shape_max = (double *)calloc(L, sizeof(double));
    if (amp_shape)
        for (int l = 0; l < L; l++)
            for (int n = 0; n < N_T; n++) c.amp_shape_max[l] = fmax(c.amp_shape_max[l], fabs(amp_shape[(size_t)l * N_T + n]));
    c.tlist = tlist;
    c.gen_of_traj = gen_of_traj;
    c.limit = cheby_limit;
    c.buffer = specrange_buffer;
    c.has_explicit_range = has_range;
    c.ex_Emin = E_min;
    c.ex_Emax = E_max;
    int nterm = n_gen * (1 + L);
    c.fw_terms = (csr_t *)calloc(nterm, sizeof(csr_t));
    c.bw_terms = (csr_t *)calloc(nterm, sizeof(csr_t));
    int total_nnz = term_off[nterm];
    c.bw_rowptr = (int *)calloc((size_t)nterm * (d + 1), sizeof(int));
    c.bw_col = (int *)malloc(sizeof(int) * (total_nnz + 1));
    c.bw_val = (cplx *)malloc(sizeof(cplx) * (total_nnz + 1));
    const cplx *cval = (const cplx *)val;
    for (int t = 0; t < nterm; t++) {
        csr_t *F = &c.fw_terms[t];
        F->d = d;
        F->rowptr = rowptr + (size_t)t * (d + 1);
        F->col = col + term_off[t];
        F->val = cval + term_off[t];
        F->nnz = F->rowptr[d];
        /* adjoint term (src/workspace.jl:69): transpose + conjugate */
        int *rp = c.bw_rowptr + (size_t)t * (d + 1);
        int *bc = c.bw_col + term_off[t];
        cplx *bv = c.bw_val + term_off[t];
        for (int i = 0; i < d; i++)
            for (int q = F->rowptr[i]; q < F->rowptr[i + 1]; q++) rp[F->col[q] + 1]++;
        for (int i = 0; i < d; i++) rp[i + 1] += rp[i];
        int *fill = (int *)calloc(d, sizeof(int));
        for (int i = 0; i < d; i++)
            for (int q = F->rowptr[i]; q < F->rowptr[i + 1]; q++) {
                int j = F->col[q];
                int pos = rp[j] + fill[j]++;
                bc[pos] = i;
                bv[pos] = conj(F->val[q]);
            }
        free(fill);
        csr_t *B = &c.bw_terms[t];
        B->d = d;
        B->rowptr = rp;
        B->col = bc;
        B->val = bv;
        B->nnz = F->nnz;
    }
    c.cache_key = (double *)calloc((size_t)2 * n_gen * 2 * L, sizeof(double));
    c.cache_val = (double *)calloc((size_t)2 * n_gen * 2, sizeof(double));
    c.cache_ok = (int *)calloc((size_t)2 * n_gen, sizeof(int));

    const cplx *P0 = (const cplx *)psi0, *TG = (const cplx *)target;
    /* pulses0 / pulses1 double buffer (src/workspace.jl:123-125) */
    double *e0 = (double *)malloc(sizeof(double) * L * N_T);
    double *e1 = (double *)malloc(sizeof(double) * L * N_T);
    memcpy(e0, pulses, sizeof(double) * L * N_T);
    memcpy(e1, pulses, sizeof(double) * L * N_T);
    prop_t *fw = (prop_t *)calloc(N, sizeof(prop_t));
    prop_t *bw = (prop_t *)calloc(N, sizeof(prop_t));
    for (int k = 0; k < N; k++) {
        init_prop(&c, &fw[k], gen_of_traj[k], 0, e0);
        init_prop(&c, &bw[k], gen_of_traj[k], 1, e0);
    }
    size_t slab = (size_t)d * (N_T + 1);
    cplx *X = (cplx *)malloc(sizeof(cplx) * slab * N); /* bw_storage: per trajectory d x (N_T+1), column n = t_n */
    cplx *chi = (cplx *)malloc(sizeof(cplx) * (size_t)N * d);
    cplx *tau = (cplx *)out_tau;
    int nthr = 1;
#ifdef _OPENMP
    nthr = omp_get_max_threads();
#endif
    cplx *work = (cplx *)malloc(sizeof(cplx) * (size_t)nthr * 5 * d);
    cplx *mu_psi = (cplx *)malloc(sizeof(cplx) * 2 * d);

    /* initial forward sweep (src/optimize.jl:182-184, 247-265) */
#pragma omp parallel for schedule(static) if (use_threads)
    for (int k = 0; k < N; k++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        reinit_prop(&c, &fw[k], P0 + (size_t)k * d, e0);
        for (int n = 0; n < N_T; n++) prop_step(&c, &fw[k], e0, work + (size_t)tid * 5 * d);
    }
    /* update_result!(wrk, 0) */
    for (int k = 0; k < N; k++) {
        cplx s = 0;
        for (int i = 0; i < d; i++) s += conj(TG[(size_t)k * d + i]) * fw[k].state[i];
        tau[k] = s;
    }
    out_JT[0] = J_T_value(functional, tau, weight, N);

    double *eps_i = e0, *eps_ip1 = e1;
    double t_start = now_s();
    for (int it = 1; it <= iters; it++) {
        /* ---- krotov_iteration (src/optimize.jl:279-371) ---- */
        chi_states(functional, tau, weight, TG, N, d, chi); /* :297-302 */
#pragma omp parallel for schedule(static) if (use_threads)
        for (int k = 0; k < N; k++) { /* :303-317 */
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            reinit_prop(&c, &bw[k], chi + (size_t)k * d, eps_i);
            cplx *Xk = X + slab * k;
            memcpy(Xk + (size_t)N_T * d, chi + (size_t)k * d, sizeof(cplx) * d);
            for (int n = N_T - 1; n >= 0; n--) {
                prop_step(&c, &bw[k], eps_i, work + (size_t)tid * 5 * d);
                memcpy(Xk + (size_t)n * d, bw[k].state, sizeof(cplx) * d);
            }
        }
#pragma omp parallel for schedule(static) if (use_threads)
        for (int k = 0; k < N; k++) /* :321-325 */
            reinit_prop(&c, &fw[k], P0 + (size_t)k * d, eps_ip1);
        double *ga = out_ga + (size_t)(it - 1) * L;
        for (int l = 0; l < L; l++) ga[l] = 0.0; /* :327 */
        for (int n = 0; n < N_T; n++) {          /* :328 */
            double dt = tlist[n + 1] - tlist[n];
            double du[16];
            for (int l = 0; l < L; l++) { /* :340-349, serial, l outer / k inner */
                du[l] = 0.0;
                for (int k = 0; k < N; k++) {
                    const csr_t *mu = &c.fw_terms[gen_of_traj[k] * (1 + L) + 1 + l];
                    if (mu->nnz == 0) continue;
                    const cplx *psi = fw[k].state;
                    const cplx *chik = X + slab * k + (size_t)n * d;
                    cplx s = 0;
                    for (int i = 0; i < d; i++) {
                        cplx r = 0;
                        for (int q = mu->rowptr[i]; q < mu->rowptr[i + 1]; q++) r += mu->val[q] * psi[mu->col[q]];
                        s += conj(chik[i]) * r;
                    }
                    if (amp_nonlinear(&c)) s *= amp_deriv(&c, l, n, eps_i[(size_t)l * N_T + n]);
                    du[l] += cimag(s);
                }
            }
            for (int l = 0; l < L; l++) { /* :351-358 */
                double alpha = S[(size_t)l * N_T + n] / lam[l];
                eps_ip1[(size_t)l * N_T + n] = eps_i[(size_t)l * N_T + n] + alpha * du[l];
                ga[l] += alpha * fabs(du[l]) * fabs(du[l]) * dt;
            }
#pragma omp parallel for schedule(static) if (use_threads)
            for (int k = 0; k < N; k++) { /* :360-368 */
                int tid = 0;
#ifdef _OPENMP
                tid = omp_get_thread_num();
#endif
                prop_step(&c, &fw[k], eps_ip1, work + (size_t)tid * 5 * d);
            }
        }
        /* ---- update_result!(wrk, i) (:374-386) ---- */
        for (int k = 0; k < N; k++) {
            cplx s = 0;
            for (int i = 0; i < d; i++) s += conj(TG[(size_t)k * d + i]) * fw[k].state[i];
            tau[k] = s;
        }
        out_JT[it] = J_T_value(functional, tau, weight, N);
        double *t = eps_i; /* :216 swap */
        eps_i = eps_ip1;
        eps_ip1 = t;
    }
    *out_secs = now_s() - t_start;
    memcpy(pulses, eps_i, sizeof(double) * L * N_T);
    cplx *OS = (cplx *)out_states;
    for (int k = 0; k < N; k++) memcpy(OS + (size_t)k * d, fw[k].state, sizeof(cplx) * d);
    out_m[0] = fw[0].m;
    out_m[1] = bw[0].m;

    for (int k = 0; k < N; k++) {
        free(fw[k].rng_min); free(fw[k].rng_max); free(fw[k].state); free(fw[k].coef);
        free(bw[k].rng_min); free(bw[k].rng_max); free(bw[k].state); free(bw[k].coef);
    }
    free(fw); free(bw); free(X); free(chi); free(work); free(mu_psi); free(e0); free(e1);
    free(c.fw_terms); free(c.bw_terms); free(c.bw_rowptr); free(c.bw_col); free(c.bw_val);
    free(c.cache_key); free(c.cache_val); free(c.cache_ok); free(c.amp_shape_max);
    return g_oracle_diverged ? -2 : 0;
}

/* Non-linear amplitudes for the next oracle_krotov_optimize call: poly [L][deg+1] (ascending powers) or NULL,
 * shape [L][N_T] or NULL.  The arrays must stay alive until that call returns. */
void oracle_set_amplitudes(int deg, const double *poly, const double *shape) {
    g_amp_deg = deg;
    g_amp_poly = poly;
    g_amp_shape = shape;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
