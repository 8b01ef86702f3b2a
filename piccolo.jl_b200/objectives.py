"""Objectives of the knot path, mirrored from the reference's constructors and evaluated on the B200.

    reference (Julia, src/control/objectives.jl)                    here
    --------------------------------------------------------------  ------------------------------------
    KetInfidelityObjective(ψ_goal, name, traj; Q)        :56-64      KetInfidelityObjective(ψ_goal, name, traj, Q=)
    CoherentKetInfidelityObjective(goals, names, traj)   :181-216    CoherentKetInfidelityObjective(...)
    UnitaryInfidelityObjective(U_goal, name, traj; Q)    :347-356    UnitaryInfidelityObjective(..., subspace=)
    DensityMatrixInfidelityObjective(name, ρ_goal, traj) :404-411    DensityMatrixInfidelityObjective(...)
    DensityMatrixPureStateInfidelityObjective            :421-429    DensityMatrixPureStateInfidelityObjective(...)
    LeakageObjective(indices, name, traj; times, Qs)     :464-474    LeakageObjective(...)
    QuadraticRegularizer(name, traj, R)  (DirectTrajOpt; smooth_pulse_problem.jl:248-250)   QuadraticRegularizer(...)
    J1 + J2                                                          J1 + J2
    objective_value(J, traj), gradient!(∇, J, traj)                  objective_value(J, traj), gradient_(∇, J, traj)

Each constructor only rewrites its complex goal as the real coefficient vectors of the C ABI's term
(include/piccolo_b200.h, pb2_obj_term); all arithmetic on the trajectory happens in
libpiccolo_b200.so.  There is no CPU path.
"""
import ctypes

import numpy as np

from . import capi
from .integrators import NamedTrajectory, _as_ptr

ONE_MINUS = 1


class _Term:
    def __init__(self, rows, a_re=None, a_im=None, a_sq=None, a_lin=None, scale=1.0, flags=ONE_MINUS, times=None,
                 Q=100.0, label=""):
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        n = self.rows.size
        vec = lambda a: None if a is None else np.ascontiguousarray(np.broadcast_to(a, (n,)), dtype=np.float64)
        self.a_re, self.a_im, self.a_sq, self.a_lin = vec(a_re), vec(a_im), vec(a_sq), vec(a_lin)
        self.scale, self.flags, self.label = float(scale), int(flags), label
        self.times = None if times is None else np.ascontiguousarray(times, dtype=np.int32)
        nq = 1 if self.times is None else self.times.size
        self.Q = np.ascontiguousarray(np.broadcast_to(Q, (nq,)), dtype=np.float64)


class _Reg:
    def __init__(self, rows, R, baseline=None, dt_power=0, times=None, label=""):
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.R = np.ascontiguousarray(np.broadcast_to(R, (self.rows.size,)), dtype=np.float64)
        self.baseline = None if baseline is None else np.asfortranarray(baseline, dtype=np.float64)
        self.dt_power, self.label = int(dt_power), label
        self.times = None if times is None else np.ascontiguousarray(times, dtype=np.int32)


class B200Objective:
    """A sum of terms and regularizers over one trajectory layout; built lazily into one device handle
    (one kernel launch per evaluation whatever the number of terms)."""

    def __init__(self, traj, terms=(), regs=(), device=0):
        self.K, self.D = traj.N, traj.dim
        self.dt_off = traj.components[traj.timestep].start
        self.global_dim = traj.global_dim
        self.terms, self.regs, self.device = list(terms), list(regs), device
        self._h = None

    def __add__(self, other):
        if other is None:
            return self
        if (other.K, other.D, other.dt_off) != (self.K, self.D, self.dt_off):
            raise ValueError("objectives were built on different trajectory layouts")
        out = B200Objective.__new__(B200Objective)
        out.K, out.D, out.dt_off, out.global_dim, out.device = self.K, self.D, self.dt_off, self.global_dim, self.device
        out.terms, out.regs, out._h = self.terms + other.terms, self.regs + other.regs, None
        return out

    __radd__ = __add__

    # -- handle ---------------------------------------------------------------------------------
    def _handle(self):
        if self._h is not None:
            return self._h
        lib = capi.load_library()
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        ptr = lambda a, t: None if a is None else a.ctypes.data_as(t)
        T = (capi.pb2_obj_term * max(len(self.terms), 1))()
        for i, t in enumerate(self.terms):
            T[i].flags, T[i].n_rows, T[i].rows = t.flags, t.rows.size, ptr(t.rows, ip)
            T[i].a_re, T[i].a_im, T[i].a_sq, T[i].a_lin = ptr(t.a_re, dp), ptr(t.a_im, dp), ptr(t.a_sq, dp), ptr(t.a_lin, dp)
            T[i].scale = t.scale
            T[i].n_times, T[i].times, T[i].Q = (0 if t.times is None else t.times.size), ptr(t.times, ip), ptr(t.Q, dp)
        R = (capi.pb2_obj_reg * max(len(self.regs), 1))()
        for i, r in enumerate(self.regs):
            if r.baseline is not None and r.baseline.shape != (r.rows.size, self.K):
                raise ValueError("baseline must be (rows, K)")
            R[i].n_rows, R[i].rows, R[i].R, R[i].baseline = r.rows.size, ptr(r.rows, ip), ptr(r.R, dp), ptr(r.baseline, dp)
            R[i].dt_power = r.dt_power
            R[i].n_times, R[i].times = (0 if r.times is None else r.times.size), ptr(r.times, ip)
        d = capi.pb2_obj_desc()
        d.K, d.D, d.dt_off, d.n_terms, d.n_regs, d.device = self.K, self.D, self.dt_off, len(self.terms), len(self.regs), self.device
        d.terms, d.regs = T, R
        h = ctypes.c_void_p()
        capi.check(lib.pb2_obj_create(ctypes.byref(d), ctypes.byref(h)))
        self._lib, self._h = lib, h
        return h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb2_obj_destroy(self._h)
            self._h = None

    __del__ = close

    # -- evaluation -----------------------------------------------------------------------------
    def _Z(self, Z):
        Z = Z.data if isinstance(Z, NamedTrajectory) else np.asfortranarray(Z, dtype=np.float64)
        if Z.shape != (self.D, self.K):
            raise ValueError(f"trajectory is {Z.shape}, expected {(self.D, self.K)}")
        return Z

    def value(self, Z):
        Z, J = self._Z(Z), ctypes.c_double()
        h = self._handle()
        capi.check(self._lib.pb2_obj_value_gradient(h, Z.ctypes.data, ctypes.byref(J), None, capi.PB2_HOST))
        return J.value

    def value_gradient(self, Z):
        """J and the gradient over [datavec; globals] (globals get zeros)."""
        Z, J = self._Z(Z), ctypes.c_double()
        g = np.zeros(self.K * self.D + self.global_dim)
        h = self._handle()
        capi.check(self._lib.pb2_obj_value_gradient(h, Z.ctypes.data, ctypes.byref(J), g.ctypes.data, capi.PB2_HOST))
        return J.value, g

    def hessian_structure(self):
        """1-based (rows, cols) of the upper-triangle COO entries ``hessian_values`` fills (duplicates sum)."""
        h = self._handle()
        n = int(self._lib.pb2_obj_nnz_hess(h))
        rows, cols = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        ip = ctypes.POINTER(ctypes.c_int64)
        capi.check(self._lib.pb2_obj_structure_hess(h, rows.ctypes.data_as(ip), cols.ctypes.data_as(ip)))
        return rows, cols

    def hessian_values(self, Z, sigma=1.0):
        """Values of sigma * d2J/dz2 (Ipopt's eval_h objective part) in ``hessian_structure`` order."""
        Z = self._Z(Z)
        h = self._handle()
        vals = np.empty(int(self._lib.pb2_obj_nnz_hess(h)))
        capi.check(self._lib.pb2_obj_hessian(h, Z.ctypes.data, float(sigma), vals.ctypes.data, capi.PB2_HOST))
        return vals

    def hessian_device(self, dZ, sigma, dvals, stream=None):
        h = self._handle()
        capi.check(self._lib.pb2_obj_hessian_async(h, _as_ptr(dZ), float(sigma), _as_ptr(dvals), _as_ptr(stream)))

    def value_gradient_device(self, dZ, dJ, dgrad, stream=None):
        """Device pointers / torch CUDA tensors; asynchronous on ``stream``."""
        h = self._handle()
        capi.check(self._lib.pb2_obj_value_gradient_async(h, _as_ptr(dZ), _as_ptr(dJ), _as_ptr(dgrad), _as_ptr(stream)))


def objective_value(J, traj):
    return J.value(traj)


def gradient_(grad, J, traj):
    """gradient!(∇, J, traj)"""
    grad[:] = J.value_gradient(traj)[1]
    return grad


# ---------------------------------------------------------------------------------------------------
def _overlap_coeffs(goal):
    """<goal|psi> for psi~ = [Re psi; Im psi]:  Re = a_re.psi~, Im = a_im.psi~."""
    g = np.asarray(goal, dtype=complex).reshape(-1)
    return np.concatenate([g.real, g.imag]), np.concatenate([-g.imag, g.real])


def _rows(traj, name):
    return np.asarray(traj.components[name], dtype=np.int32)


def KetInfidelityObjective(psi_goal, name, traj, Q=100.0, device=0):
    a_re, a_im = _overlap_coeffs(psi_goal)
    return B200Objective(traj, [_Term(_rows(traj, name), a_re, a_im, Q=Q, label=f"KetInfidelity({name})")], device=device)


def coherent_fidelity_weights(weights, n):
    """objectives.jl:136-144: None for absent or uniform weights, else normalised to sum one."""
    if weights is None:
        return None
    w = np.asarray(weights, dtype=float)
    if w.size != n or (w < 0).any() or not w.sum() > 0:
        raise ValueError("weights must be non-negative, not all zero, one per state")
    return None if np.all(w == w[0]) else w / w.sum()


def CoherentKetInfidelityObjective(psi_goals, names, traj, Q=100.0, weights=None, device=0):
    n = len(psi_goals)
    if len(names) != n:
        raise ValueError("Number of names must match number of goals")
    w = coherent_fidelity_weights(weights, n)
    ww = np.full(n, 1.0 / n) if w is None else w
    rows, a_re, a_im = [], [], []
    for g, nm, wi in zip(psi_goals, names, ww):
        r, i = _overlap_coeffs(g)
        rows.append(_rows(traj, nm)); a_re.append(wi * r); a_im.append(wi * i)
    return B200Objective(traj, [_Term(np.concatenate(rows), np.concatenate(a_re), np.concatenate(a_im), Q=Q,
                                      label="CoherentKetInfidelity")], device=device)


def UnitaryInfidelityObjective(U_goal, name, traj, Q=100.0, subspace=None, device=0):
    """Full-operator goal, or (``subspace`` = 0-based level indices, ``U_goal`` the unembedded
    n_sub x n_sub unitary) the EmbeddedOperator form of objectives.jl:339-345."""
    U_goal = np.asarray(U_goal, dtype=complex)
    rows = _rows(traj, name)
    N = int(round(np.sqrt(rows.size // 2)))
    if 2 * N * N != rows.size:
        raise ValueError("state component is not an iso-vec operator")
    a_re, a_im, a_sq = np.zeros(rows.size), np.zeros(rows.size), None
    if subspace is None:
        if U_goal.shape != (N, N):
            raise ValueError("goal has the wrong size")
        for c in range(N):
            r, i = _overlap_coeffs(U_goal[:, c])
            a_re[2 * N * c: 2 * N * (c + 1)], a_im[2 * N * c: 2 * N * (c + 1)] = r, i
        scale = 1.0 / N ** 2
    else:
        sub = np.asarray(subspace, dtype=int)
        n = sub.size
        if U_goal.shape != (n, n):
            raise ValueError("goal must be the unembedded subspace operator")
        if np.abs(U_goal.conj().T @ U_goal - np.eye(n)).max() > 1e-10:
            raise ValueError("the subspace form needs a unitary goal")
        a_sq = np.zeros(rows.size)
        for p, c in enumerate(sub):
            for q, i in enumerate(sub):
                g = U_goal[q, p]
                re_idx, im_idx = 2 * N * c + i, 2 * N * c + N + i
                a_re[re_idx], a_re[im_idx] = g.real, g.imag
                a_im[re_idx], a_im[im_idx] = -g.imag, g.real
                a_sq[re_idx] = a_sq[im_idx] = 1.0
        scale = 1.0 / (n * (n + 1))
    return B200Objective(traj, [_Term(rows, a_re, a_im, a_sq, scale=scale, Q=Q, label=f"UnitaryInfidelity({name})")],
                         device=device)


def _compact_trace_coeffs(W):
    """Re tr(rho W) = a.x for rho = compact_iso_to_density(x) (isomorphisms.jl:176-191 ordering)."""
    W = np.asarray(W, dtype=complex)
    n = W.shape[0]
    a, idx = np.empty(n * n), 0
    for k in range(n):
        for j in range(k + 1):
            a[idx] = W[k, k].real if j == k else (W[k, j] + W[j, k]).real
            idx += 1
    for k in range(1, n):
        for j in range(k):
            a[idx] = (W[j, k] - W[k, j]).imag
            idx += 1
    return a


def DensityMatrixInfidelityObjective(name, rho_goal, traj, Q=100.0, device=0):
    return B200Objective(traj, [_Term(_rows(traj, name), a_lin=_compact_trace_coeffs(rho_goal), Q=Q,
                                      label=f"DensityMatrixInfidelity({name})")], device=device)


def DensityMatrixPureStateInfidelityObjective(name, psi_goal, traj, Q=100.0, device=0):
    psi = np.asarray(psi_goal, dtype=complex)
    return B200Objective(traj, [_Term(_rows(traj, name), a_lin=_compact_trace_coeffs(np.outer(psi, psi.conj())), Q=Q,
                                      label=f"DensityMatrixPureStateInfidelity({name})")], device=device)


def LeakageObjective(indices, name, traj, times=None, Qs=None, device=0):
    """``indices`` 0-based inside the component; ``times`` 0-based knots (default all)."""
    idx = np.asarray(indices, dtype=int)
    times = np.arange(traj.N) if times is None else np.asarray(times, dtype=int)
    Qs = np.ones(times.size) if Qs is None else np.asarray(Qs, dtype=float)
    rows = _rows(traj, name)[idx]
    return B200Objective(traj, [_Term(rows, a_sq=1.0 / idx.size, flags=0, times=times, Q=Qs,
                                      label=f"Leakage({name})")], device=device)


def QuadraticRegularizer(name, traj, R, baseline=None, times=None, dt_power=0, device=0):
    """1/2 sum_k sum_i R_i (v_ik - baseline_ik)^2 dt_k^dt_power  (parity with DirectTrajOpt unpinned:
    see include/piccolo_b200.h)."""
    return B200Objective(traj, regs=[_Reg(_rows(traj, name), R, baseline, dt_power, times,
                                          label=f"QuadraticRegularizer({name})")], device=device)
