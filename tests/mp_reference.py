"""An oracle-INDEPENDENT check of the Krotov loop: the two-level system of ``test/test_tls_optimization.jl:12-63``
optimised in 50-digit arithmetic (mpmath) with the EXACT propagator of every interval,

    exp(-i H dt) = cos(w dt) 1 - i sin(w dt) H / w,     H = -sz/2 + eps sx,  w = sqrt(1/4 + eps^2),

so that nothing of the recalled QuantumPropagators conventions of SURVEY.md Appendix A.1 (Chebyshev coefficients,
truncation rule, spectral range, buffer) enters.  What it does share with the oracles -- deliberately, these are the
lines of the reference under test -- is the loop of ``src/optimize.jl:279-371``: chi(T) = -dJ_T/d<psi| for J_T_sm,
backward propagation with the guess pulse storing chi(t_n), the sequential update
eps_new[n] = eps_old[n] + S[n]/lambda * Im<chi(t_n)|sx|psi(t_n)>, the running cost, the forward step with the updated
value; and the midpoint discretisation (first / last sample ON the grid ends).

Test infrastructure only.  Written without looking at oracle/krotov_oracle.py's propagators: a different number type,
a different propagator, a different code path."""
import mpmath as mp


def _blackman(t, t0, T, a=mp.mpf("0.16")):
    if t < t0 or t > T:
        return mp.mpf(0)
    x = (t - t0) / (T - t0)
    return (1 - a - mp.cos(2 * mp.pi * x) + a * mp.cos(4 * mp.pi * x)) / 2


def _flattop(t, T, t_rise):
    if t <= 0 or t >= T:
        return mp.mpf(0)
    if t <= t_rise:
        return _blackman(t, mp.mpf(0), 2 * t_rise)
    if t >= T - t_rise:
        return _blackman(t, T - 2 * t_rise, T)
    return mp.mpf(1)


def _step(psi, eps, dt):
    """psi <- exp(-i H dt) psi for H = -sz/2 + eps sx (closed form)."""
    w = mp.sqrt(mp.mpf(1) / 4 + eps * eps)
    c, s = mp.cos(w * dt), mp.sin(w * dt) / w
    a, b = psi
    # H psi = (-a/2 + eps b, eps a + b/2)
    return (c * a - 1j * s * (-a / 2 + eps * b), c * b - 1j * s * (eps * a + b / 2))


def tls_krotov_exact(iters=5, n_grid=501, dps=50, float_grid=True, amp_poly=None, amp_shape=None, lam=1, sigma=None):
    """J_T history, running costs and final pulses of the TLS optimisation in `dps`-digit arithmetic.

    float_grid: take the time grid and the guess pulse samples from their Float64 values (what both the reference
    and the oracles start from) so that the comparison isolates the arithmetic of the loop.

    sigma: None (first order) | a number | one number per time interval -- the second-order update of Reich, Ndong, Koch,
    J. Chem. Phys. 136, 104103 (2012): the overlap gains (sigma_n / 2) <psi_new(t_n) - psi_old(t_n)| mu |psi_new(t_n)>."""
    mp.mp.dps = dps
    T, t_rise, lam = mp.mpf(5), mp.mpf("0.3"), mp.mpf(lam)
    # non-linear amplitude: H = -sz/2 + a(eps, n) sx,  a = shape[n] * sum_p c_p eps^p;  mu = dH/d eps = a'(eps, n) sx is
    # evaluated at the GUESS value of the interval (src/optimize.jl:337)
    poly = None if amp_poly is None else [mp.mpf(float(c)) for c in amp_poly]
    shape = None if amp_shape is None else [mp.mpf(float(x)) for x in amp_shape]

    def amp(e, n):
        v = e if poly is None else sum(c * e**q for q, c in enumerate(poly))
        return v if shape is None else shape[n] * v

    def damp(e, n):
        v = mp.mpf(1) if poly is None else sum(q * c * e ** (q - 1) for q, c in enumerate(poly) if q > 0)
        return v if shape is None else shape[n] * v

    if float_grid:
        import numpy as np

        tl = [mp.mpf(float(x)) for x in np.linspace(0.0, 5.0, n_grid)]
    else:
        tl = [T * i / (n_grid - 1) for i in range(n_grid)]
    N_T = n_grid - 1

    def guess(t):
        if float_grid:
            import workloads as W

            return mp.mpf(float(0.2 * W.flattop(float(t), T=5.0, t_rise=0.3)))
        return mp.mpf("0.2") * _flattop(t, T, t_rise)

    # discretize_on_midpoints: first / last sample on the grid ends, interior on the interval midpoints
    def mid(i):
        if float_grid:
            return mp.mpf(float(tl[i]) + 0.5 * (float(tl[i + 1]) - float(tl[i])))
        return (tl[i] + tl[i + 1]) / 2

    eps = [guess(tl[0])] + [guess(mid(i)) for i in range(1, N_T - 1)] + [guess(tl[-1])]
    psi0 = (mp.mpc(1), mp.mpc(0))
    dts = [tl[n + 1] - tl[n] for n in range(N_T)]

    if sigma is not None:
        sig = [mp.mpf(float(x)) for x in sigma] if hasattr(sigma, "__len__") else [mp.mpf(float(sigma))] * N_T
    old_traj = []

    def forward(e):
        psi = psi0
        for n in range(N_T):
            old_traj.append(psi)
            psi = _step(psi, amp(e[n], n), dts[n])
        old_traj.append(psi)
        return psi

    psi = forward(eps)
    tau = psi[1]  # <target|psi>, target = (0, 1)
    J = [1 - abs(tau) ** 2]
    ga_hist = []
    for _ in range(iters):
        chi = (mp.mpc(0), tau)  # chi = (1/N^2)(sum tau) |tgt>, N = 1
        X = [None] * (N_T + 1)
        X[N_T] = chi
        for n in range(N_T - 1, -1, -1):  # exp(+i H^dagger dt) with the GUESS pulse
            chi = _step(chi, amp(eps[n], n), -dts[n])
            X[n] = chi
        new = list(eps)
        psi = psi0
        ga = mp.mpf(0)
        for n in range(N_T):
            c0, c1 = X[n]
            # <chi| sx |psi> = conj(c0) psi1 + conj(c1) psi0
            ov = mp.conj(c0) * psi[1] + mp.conj(c1) * psi[0]
            if sigma is not None:
                d0, d1 = psi[0] - old_traj[n][0], psi[1] - old_traj[n][1]
                ov += sig[n] / 2 * (mp.conj(d0) * psi[1] + mp.conj(d1) * psi[0])
                old_traj[n] = psi  # (slot n is not read again in this iteration)
            du = damp(eps[n], n) * mp.im(ov)
            alpha = mp.mpf(1) / lam  # S = 1
            new[n] = eps[n] + alpha * du
            ga += alpha * du * du * dts[n]
            psi = _step(psi, amp(new[n], n), dts[n])
        eps = new
        old_traj[N_T] = psi
        tau = psi[1]
        J.append(1 - abs(tau) ** 2)
        ga_hist.append(ga)
    return dict(J_T=[float(x) for x in J], J_T_mp=J, g_a_int=[float(x) for x in ga_hist], pulses=[float(x) for x in eps])
