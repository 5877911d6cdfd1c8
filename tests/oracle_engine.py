"""An ORACLE-backed stand-in for ``krotov_jl_b200.engine.KrotovCuda`` -- TEST INFRASTRUCTURE for the CPU suite only.

The product has no CPU path (``krotov_create`` fails loudly without a GPU), so the host driver of ``optimize.py`` /
``workspace.py`` (buffer swaps, ``continue_from``, callbacks, the second-order boundary condition, ...) cannot run in
the ``-m "not gpu"`` tests through the product engine.  Tests that exercise that host logic monkeypatch
``krotov_jl_b200.workspace.KrotovCuda`` with this class: same methods, the arithmetic done by ``oracle/krotov_oracle.py``
(which does its own ``reinit_prop!`` bookkeeping: what the host pushes through ``set_cheby`` is recorded, not used).
Nothing under ``krotov.jl_b200/`` imports this file.
"""
import numpy as np

from oracle import krotov_oracle as O

_FUNCTIONAL = {0: "host", 1: "sm", 2: "ss", 3: "re"}


class OracleEngine:
    created = 0  # how many engines the tests made (to prove the patch was in effect)

    def __init__(self, *, tlist, H0, Hc, gen_of_traj, psi0, target=None, weight=None, update_shape, lambda_a,
                 functional=0, n_traj_global=0, store_fw=False, replicated_forward=False, **_):
        OracleEngine.created += 1
        dense = lambda m: None if m is None else np.asarray(m.toarray() if hasattr(m, "toarray") else m, complex)  # noqa: E731
        psi0 = np.asarray(psi0, complex)
        self.N, self.d = psi0.shape
        self.L, self.N_T, self.n_gen = len(Hc[0]), len(tlist) - 1, len(H0)
        self.functional = _FUNCTIONAL[int(functional)]
        self.p = O.ProblemArrays(
            tlist=np.asarray(tlist, float), H0=[dense(h) for h in H0], Hc=[[dense(h) for h in row] for row in Hc],
            gen_of_traj=np.asarray(gen_of_traj, int), psi0=psi0,
            target=np.zeros_like(psi0) if target is None else np.asarray(target, complex),
            pulses=np.zeros((self.L, self.N_T)), S=np.asarray(update_shape, float).reshape(self.L, self.N_T),
            lam=np.asarray(lambda_a, float).reshape(self.L), weight=None if weight is None else np.asarray(weight, float),
            functional=self.functional if self.functional != "host" else "sm")
        self.store_fw = bool(store_fw)
        self.n_global = int(n_traj_global) or self.N  # > N: this engine holds one rank's shard of the trajectories
        self._reduce = None
        self.replicated = bool(replicated_forward)  # every rank holds all trajectories: no per-step exchange
        self.wrk = None
        self.cheby_pushed = []
        self._chi = None
        self._chi_coef = None

    # -- settings (recorded only) -----------------------------------------------------------
    def set_cheby(self, direction, *args):
        self.cheby_pushed.append(int(direction))

    def set_amplitudes(self, poly=None, shape=None):
        self.p.amp_poly = None if poly is None else [list(map(float, row)) for row in np.asarray(poly)]
        self.p.amp_shape = None if shape is None else np.asarray(shape, float)

    def info(self):
        return dict(path=0, grid_blocks=0, block_threads=0, launches_total=0, exchange=0, ms_last=0.0)

    # -- hot path -------------------------------------------------------------------------
    def _make(self, pulses):
        self.p.pulses = np.array(pulses, float).reshape(self.L, self.N_T)
        self.wrk = O.OracleWrk(self.p, store_fw=self.store_fw)
        self._bufs = [self.wrk.pulses0, self.wrk.pulses1]

    def forward(self, pulses):
        self._make(pulses)
        for k in range(self.N):
            O.krotov_initial_fw_prop(self._bufs[0], self.p.psi0[k], k, self.wrk)
        O.update_result(self.wrk)

    def set_chi(self, chi):
        self._chi = np.array(chi, complex).reshape(self.N, self.d)

    def set_chi_coeffs(self, coef):
        self._chi_coef = np.array(coef, complex).reshape(self.N)

    def iterate(self, guess_pulses, out_pulses=None):
        if self.wrk is None:  # skip_initial_forward_propagation: the handle starts from Psi(T) = Psi(0)
            self._make(guess_pulses)
            for k, prop in enumerate(self.wrk.fw_propagators):
                prop.state = self.p.psi0[k].copy()
            O.update_result(self.wrk)
        eps_i, eps_ip1 = self._bufs
        g = np.asarray(guess_pulses, float).reshape(self.L, self.N_T)
        for l in range(self.L):
            eps_i[l][:] = g[l]  # (callbacks may have edited the host copy)
        chi = None
        if self._chi is not None:
            chi = lambda Psi, c=self._chi: list(c)  # noqa: E731
        elif self._chi_coef is not None:
            chi = lambda Psi, c=self._chi_coef: [c[k] * self.p.target[k] for k in range(self.N)]  # noqa: E731
        elif self.functional == "host":
            raise RuntimeError("functional is KROTOV_CHI_HOST: call set_chi before iterate")
        elif self.n_global != self.N:
            # a shard forms its boundary condition with the GLOBAL trajectory count (krotov_problem.n_traj_global)
            if self.functional == "sm":
                raise RuntimeError("multi-rank J_T_sm needs krotov_set_chi_coeffs (global sum of tau)")
            w, n = self.p.weights(), self.n_global
            c = (w / n) * self.wrk.tau_vals if self.functional == "ss" else (w / (2 * n)).astype(complex)
            chi = lambda Psi, c=c: [c[k] * self.p.target[k] for k in range(self.N)]  # noqa: E731
        O.krotov_iteration(self.wrk, eps_i, eps_ip1, chi=chi, reduce_du=self._reduce)
        O.update_result(self.wrk)
        self._chi = self._chi_coef = None
        self._bufs = [eps_ip1, eps_i]
        new = np.array(eps_ip1)
        if out_pulses is not None:
            out_pulses[:] = new
            new = out_pulses
        return new, self.wrk.g_a_int.copy()

    # -- results --------------------------------------------------------------------------
    def states(self):
        if self.wrk is None:  # no sweep yet (skip_initial_forward_propagation): the handle holds Psi(0) and its tau
            return np.array(self.p.psi0)
        return np.array([prop.state for prop in self.wrk.fw_propagators])

    def tau(self):
        if self.wrk is None:
            return O.taus(list(self.p.psi0), self.p.target)
        return np.array(self.wrk.tau_vals)

    def storage(self, which, k, n0=0, n1=None):
        n1 = self.N_T + 1 if n1 is None else n1
        src = self.wrk.bw_storage if which == 1 else self.wrk.fw_storage
        return np.array(src[k][:, n0:n1].T)

    # -- multi-rank (gloo tests): the per-step overlap sums cross the ranks through torch.distributed -----------------
    def comm_export(self):
        return bytes(256)

    def comm_connect(self, rank, world, descs):
        import torch
        import torch.distributed as dist

        assert len(descs) == world
        if self.replicated:
            return

        def reduce(du):
            t = torch.from_numpy(np.array(du, float))
            dist.all_reduce(t)
            return t.numpy()

        self._reduce = reduce

    def close(self):
        pass
