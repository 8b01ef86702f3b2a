"""Objective terms and their gradients (oracle; test infrastructure).

A restatement, in complex arithmetic on the un-vectorised objects, of the loss functions in
/root/reference/src/control/objectives.jl:
  ket_fidelity_loss :24-27, KetInfidelityObjective :34-38 / :56-64,
  coherent_ket_fidelity :96-121 (uniform weights take the unweighted path), coherent_fidelity_weights :136-144,
  CoherentKetInfidelityObjective :181-216,
  unitary_fidelity_loss :330-337 (full operator) and :339-345 (EmbeddedOperator subspace),
  UnitaryInfidelityObjective :347-356, density_matrix_infidelity_loss :387-394,
  density_matrix_pure_state_infidelity_loss :412-419, LeakageObjective :464-474.
Every terminal objective is  J = Q * loss(x_K)  (DirectTrajOpt's TerminalObjective, as the reference's
own tests use it: "100.0 * (1 - 0.9025)", objectives.jl:592).

DirectTrajOpt's QuadraticRegularizer is NOT in /root/reference (dependency "DirectTrajOpt 0.9.5, 0.10",
Project.toml:48) and no reference test pins a value of it, so ``quadratic_regularizer`` below is
PARITY UNPINNED: it restates  J = 1/2 * sum_k sum_i R_i (v_ik - b_ik)^2 * dt_k^p  with the power p an
explicit argument (p = 0: plain knot sum; p = 1: rectangle rule in time; p = 2: the older
QuantumCollocation form (dt*v)' R (dt*v)).  The CUDA path implements the same three.

Gradients here are derived in complex form (not the real dot-product form the kernel uses) and are
checked against central differences in tests/test_oracle.py.  d|x|/dx follows ForwardDiff's rule
(+1 for x >= +0, -1 otherwise).
"""
import numpy as np

from . import isomorphisms as iso


def _dabs(x):
    return -1.0 if np.signbit(x) else 1.0


# ------------------------------------------------------------------ kets
def ket_infidelity(x, psi_goal, Q=100.0):
    """J, dJ/dx for  Q * |1 - |<goal|psi>|^2|  (objectives.jl:24-38)."""
    psi_goal = np.asarray(psi_goal, dtype=complex)
    psi = iso.iso_to_ket(x)
    ov = np.vdot(psi_goal, psi)
    F = abs(ov) ** 2
    # dF/dpsi* = goal * <goal|psi> ; dF/dRe = 2 Re(.), dF/dIm = 2 Im(.)
    w = psi_goal * ov
    dF = np.concatenate([2 * w.real, 2 * w.imag])
    return Q * abs(1 - F), -Q * _dabs(1 - F) * dF


def coherent_fidelity_weights(weights, n):
    if weights is None:
        return None
    w = np.asarray(weights, dtype=float)
    assert w.size == n and (w >= 0).all() and w.sum() > 0
    if np.all(w == w[0]):
        return None
    return w / w.sum()


def coherent_ket_fidelity(xs, goals, weights=None):
    n = len(xs)
    w = coherent_fidelity_weights(weights, n)
    if w is None:
        s = sum(np.vdot(np.asarray(goals[i], dtype=complex), iso.iso_to_ket(xs[i])) for i in range(n))
        return abs(s / n) ** 2
    s = sum(w[i] * np.vdot(np.asarray(goals[i], dtype=complex), iso.iso_to_ket(xs[i])) for i in range(n))
    return abs(s / w.sum()) ** 2


def coherent_ket_infidelity(xs, goals, Q=100.0, weights=None):
    """J and the list of dJ/dx_i for  Q * |1 - F_coherent|  (objectives.jl:181-216)."""
    n = len(xs)
    w = coherent_fidelity_weights(weights, n)
    ww = np.full(n, 1.0 / n) if w is None else w / w.sum()
    F = coherent_ket_fidelity(xs, goals, weights)
    s = sum(ww[i] * np.vdot(np.asarray(goals[i], dtype=complex), iso.iso_to_ket(xs[i])) for i in range(n))
    grads = []
    for i in range(n):
        g = ww[i] * np.asarray(goals[i], dtype=complex) * s
        grads.append(-Q * _dabs(1 - F) * np.concatenate([2 * g.real, 2 * g.imag]))
    return Q * abs(1 - F), grads


# ------------------------------------------------------------------ unitaries
def unitary_infidelity(x, U_goal, Q=100.0, subspace=None):
    """J, dJ/dx for Q * |1 - F|.  Full operator: F = |tr(Ug' U)|^2 / n^2 (objectives.jl:330-337).
    With ``subspace`` (0-based indices; U_goal is then the n_sub x n_sub unembedded goal):
    F = (|tr(M'M)| + |tr M|^2) / (n(n+1)),  M = Ug' U[sub, sub]  (objectives.jl:339-345)."""
    U_goal = np.asarray(U_goal, dtype=complex)
    U = iso.iso_vec_to_operator(x)
    N = U.shape[0]
    if subspace is None:
        n = N
        t = np.trace(U_goal.conj().T @ U)
        F = abs(t) ** 2 / n ** 2
        dU = U_goal * t / n ** 2                 # dF/dU* ; dF/dRe = 2 Re, dF/dIm = 2 Im
    else:
        sub = np.asarray(subspace)
        n = sub.size
        Us = U[np.ix_(sub, sub)]
        M = U_goal.conj().T @ Us
        t = np.trace(M)
        F = (abs(np.trace(M.conj().T @ M)) + abs(t) ** 2) / (n * (n + 1))
        dUs = (U_goal @ M + U_goal * t) / (n * (n + 1))
        dU = np.zeros_like(U)
        dU[np.ix_(sub, sub)] = dUs
    g = np.empty(2 * N * N)
    for i in range(N):
        g[i * 2 * N: i * 2 * N + N] = 2 * dU[:, i].real
        g[i * 2 * N + N: (i + 1) * 2 * N] = 2 * dU[:, i].imag
    return Q * abs(1 - F), -Q * _dabs(1 - F) * g


# ------------------------------------------------------------------ density matrices
def _compact_grad(W):
    """d Re tr(rho W) / d x  for rho = compact_iso_to_density(x): entry (j,k), j<k carries
    rho[j,k] = a + ib, rho[k,j] = a - ib."""
    n = W.shape[0]
    g = np.empty(n * n)
    idx = 0
    for k in range(n):
        for j in range(k + 1):
            g[idx] = W[k, j].real if j == k else (W[k, j] + W[j, k]).real
            idx += 1
    for k in range(1, n):
        for j in range(k):
            g[idx] = (1j * W[k, j] - 1j * W[j, k]).real
            idx += 1
    return g


def density_infidelity(x, rho_goal, Q=100.0):
    """Q * |1 - Re tr(rho rho_goal)|  (objectives.jl:387-394)."""
    rho_goal = np.asarray(rho_goal, dtype=complex)
    rho = iso.compact_iso_to_density(x)
    F = np.trace(rho @ rho_goal).real
    return Q * abs(1 - F), -Q * _dabs(1 - F) * _compact_grad(rho_goal)


def density_pure_state_infidelity(x, psi, Q=100.0):
    """Q * |1 - Re <psi|rho|psi>|  (objectives.jl:412-419)."""
    psi = np.asarray(psi, dtype=complex)
    rho = iso.compact_iso_to_density(x)
    F = np.vdot(psi, rho @ psi).real
    return Q * abs(1 - F), -Q * _dabs(1 - F) * _compact_grad(np.outer(psi, psi.conj()))


# ------------------------------------------------------------------ knot-point terms
def leakage(X, indices, Qs=None):
    """sum_t Qs[t] * sum(x_t[indices]^2) / len(indices) over the columns of X (objectives.jl:464-474).
    Returns J and dJ/dX."""
    idx = np.asarray(indices)
    Qs = np.ones(X.shape[1]) if Qs is None else np.asarray(Qs, dtype=float)
    G = np.zeros_like(X)
    J = 0.0
    for t in range(X.shape[1]):
        J += Qs[t] * np.sum(X[idx, t] ** 2) / idx.size
        G[idx, t] = Qs[t] * 2 * X[idx, t] / idx.size
    return J, G


def quadratic_regularizer(V, dt, R, baseline=None, dt_power=0):
    """PARITY UNPINNED (see the header).  V is dim x T, dt length T.  Returns J, dJ/dV, dJ/ddt."""
    R = np.broadcast_to(np.asarray(R, dtype=float), (V.shape[0],))
    dV = V if baseline is None else V - baseline
    q = 0.5 * np.sum(R[:, None] * dV * dV, axis=0)
    if dt_power == 0:
        return float(np.sum(q)), R[:, None] * dV, np.zeros_like(dt)
    return (float(np.sum(q * dt ** dt_power)), R[:, None] * dV * dt ** dt_power,
            q * dt_power * dt ** (dt_power - 1))


# ------------------------------------------------------------------ second derivatives
def hessian_from_gradient(grad_fn, x):
    """Hessian of a loss  Q*|1 - F(x)|  or  Q*F(x)  with F at most quadratic in x (every term above):
    away from the kink the gradient g(x) is affine in x, so column j of the Hessian is EXACTLY
    (g(x + h e_j) - g(x - h e_j)) / 2h for any h that keeps 1 - F on the same side; h is taken small and
    a power of two so that the difference quotient carries only rounding error.  ``grad_fn(x)`` returns
    the gradient (the functions above, closed over their goal).  This is the complex-form statement the
    kernel's real outer-product form (2 c scale (a_re a_re' + a_im a_im' + diag a_sq)) is checked against."""
    x = np.asarray(x, dtype=float)
    n, h = x.size, 2.0 ** -12
    H = np.empty((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = h
        H[:, j] = (grad_fn(x + e) - grad_fn(x - e)) / (2 * h)
    return 0.5 * (H + H.T)


def quadratic_regularizer_hessian(V, dt, R, baseline=None, dt_power=0):
    """PARITY UNPINNED like quadratic_regularizer.  Returns (d2J/dV2 diagonal [dim x T],
    d2J/dV ddt [dim x T], d2J/ddt2 [T]) -- the regularizer couples a knot's rows only with that knot's dt."""
    R = np.broadcast_to(np.asarray(R, dtype=float), (V.shape[0],))
    dV = V if baseline is None else V - baseline
    q = 0.5 * np.sum(R[:, None] * dV * dV, axis=0)
    p = dt_power
    if p == 0:
        return np.broadcast_to(R[:, None], V.shape).copy(), np.zeros_like(V), np.zeros_like(dt)
    return (R[:, None] * dt ** p * np.ones_like(V), R[:, None] * dV * p * dt ** (p - 1),
            q * p * (p - 1) * (dt ** (p - 2) if p > 1 else 0.0))
