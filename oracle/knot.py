"""Per-knot dynamics constraint, Jacobian and Lagrangian Hessian (oracle; test infrastructure).

What the reference evaluates inside Ipopt callbacks through DirectTrajOpt's
``BilinearIntegrator`` (constructed at /root/reference/src/control/integrators.jl:35-95;
``.f(x_next, x, u, dt)`` contract at :525; constraint definition
docs/src/concepts/index.md:21,62):

    delta_k = x_{k+1} - exp(dt_k * Ghat(u_k)) x_k ,   k = 1 .. K-1

  unitary : Ghat(u) = I_d (x) G(u)  (integrators.jl:48)  -> with X = reshape(x, 2d, d):  X+ - E X
  ket     : Ghat(u) = G(u)          (integrators.jl:71)
  density : Ghat(u) = compact Lindbladian  (integrators.jl:88-92)

DirectTrajOpt.jl / ExponentialAction.jl are not vendored in the reference
(Project.toml:8,10); the arithmetic here is the published math:  ``scipy.linalg.expm``
(Pade 13 scaling-and-squaring) for exp, ``expm_frechet`` for d/du, a block-triangular
exponential for the second Frechet derivative.  All float64.

Trajectory layout (NamedTrajectory.datavec, named_trajectory_conversion.jl:316-321):
Z is D x K column-major (each knot contiguous); numpy view ``Z[c, k]`` of shape (D, K),
Fortran order.  The NLP primal is [vec(Z); globals]  (ext/PiccoloMakieExt.jl:497-511).

Canonical COO order (OURS; DirectTrajOpt's is not available -> "parity unpinned"):
  Jacobian, knot-major; inside knot k (0-based), with n_x = b * n_b:
    (i)   for block c in 0..n_b-1: the b x b block  -E  column-major
            row = k n_x + c b + i ,  col = k D + x_off + c b + j
    (ii)  for drive j: rows 0..n_x-1 of  -F_j X     col = k D + u_off + j
    (iii) rows 0..n_x-1 of  -G E X                  col = k D + dt_off
    (iv)  +1 entries                                 row = k n_x + i, col = (k+1) D + x_off + i
  Hessian of sum_k mu_k . delta_k, upper triangle (row <= col), knot-major; inside knot k:
    (x,u_j) for j, for i ; (x,dt) for i ; (u_i,u_j) for j, for i<=j ; (u_j,dt) for j ; (dt,dt)
  Index arrays are 1-based int64 like Julia / Ipopt's MOI layer.
"""
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla


@dataclass
class KnotProblem:
    kind: str          # "ket" | "unitary" | "density"
    b: int             # generator block size (2d, or d^2 for density)
    n_b: int           # state columns sharing the generator (d for unitary, else 1)
    m: int             # number of drives
    K: int             # knots
    D: int             # reals per knot (traj.dim)
    x_off: int         # 0-based row offsets inside a knot column
    dt_off: int
    u_off: int
    G0: np.ndarray     # (b, b)
    Gj: np.ndarray     # (m, b, b)
    global_dim: int = 0

    @property
    def n_x(self):
        return self.b * self.n_b

    @property
    def dim(self):  # integrator.dim == x_dim * (N - 1)   (integrators.jl:309)
        return self.n_x * (self.K - 1)

    @property
    def nnz_jac_knot(self):
        return self.n_b * self.b * self.b + self.n_x * self.m + self.n_x + self.n_x

    @property
    def nnz_hess_knot(self):
        m = self.m
        return self.n_x * m + self.n_x + m * (m + 1) // 2 + m + 1

    def G(self, u):
        out = self.G0.copy()
        for j in range(self.m):
            out = out + u[j] * self.Gj[j]
        return out


def make_problem(kind, G0, Gj, K, m_derivs=2, extra_rows=0):
    """SmoothPulseProblem knot layout [state | dt | t | u | du | ddu]
    (named_trajectory_conversion.jl:321; smooth_pulse_problem.jl:196-201)."""
    G0 = np.ascontiguousarray(G0, dtype=float)
    Gj = np.ascontiguousarray(np.array(Gj, dtype=float).reshape(-1, *G0.shape))
    b = G0.shape[0]
    m = Gj.shape[0]
    n_b = b // 2 if kind == "unitary" else 1
    n_x = b * n_b
    D = n_x + 2 + m * (1 + m_derivs) + extra_rows
    return KnotProblem(kind, b, n_b, m, K, D, 0, n_x, n_x + 2, G0, Gj)


def _knot(prob, Z, k):
    X = Z[prob.x_off:prob.x_off + prob.n_x, k].reshape(prob.b, prob.n_b, order="F")
    Xn = Z[prob.x_off:prob.x_off + prob.n_x, k + 1].reshape(prob.b, prob.n_b, order="F")
    u = Z[prob.u_off:prob.u_off + prob.m, k]
    dt = Z[prob.dt_off, k]
    return X, Xn, u, dt


def residual(prob, Z):
    out = np.empty(prob.dim)
    for k in range(prob.K - 1):
        X, Xn, u, dt = _knot(prob, Z, k)
        E = sla.expm(dt * prob.G(u))
        out[k * prob.n_x:(k + 1) * prob.n_x] = (Xn - E @ X).reshape(-1, order="F")
    return out


def jacobian_structure(prob):
    b, n_b, m, n_x, D = prob.b, prob.n_b, prob.m, prob.n_x, prob.D
    rows, cols = [], []
    for k in range(prob.K - 1):
        r0 = k * n_x
        for c in range(n_b):
            for j in range(b):
                for i in range(b):
                    rows.append(r0 + c * b + i)
                    cols.append(k * D + prob.x_off + c * b + j)
        for j in range(m):
            for i in range(n_x):
                rows.append(r0 + i)
                cols.append(k * D + prob.u_off + j)
        for i in range(n_x):
            rows.append(r0 + i)
            cols.append(k * D + prob.dt_off)
        for i in range(n_x):
            rows.append(r0 + i)
            cols.append((k + 1) * D + prob.x_off + i)
    return np.array(rows, dtype=np.int64) + 1, np.array(cols, dtype=np.int64) + 1


def jacobian_values(prob, Z):
    vals = np.empty((prob.K - 1, prob.nnz_jac_knot))
    b, n_b, m, n_x = prob.b, prob.n_b, prob.m, prob.n_x
    for k in range(prob.K - 1):
        X, Xn, u, dt = _knot(prob, Z, k)
        Gu = prob.G(u)
        A = dt * Gu
        E = sla.expm(A)
        o = 0
        for c in range(n_b):
            vals[k, o:o + b * b] = (-E).reshape(-1, order="F")
            o += b * b
        for j in range(m):
            Fj = sla.expm_frechet(A, dt * prob.Gj[j], compute_expm=False)
            vals[k, o:o + n_x] = (-(Fj @ X)).reshape(-1, order="F")
            o += n_x
        vals[k, o:o + n_x] = (-(Gu @ (E @ X))).reshape(-1, order="F")
        o += n_x
        vals[k, o:o + n_x] = 1.0
    return vals.reshape(-1)


def _frechet2(A, B1, B2):
    """Second Frechet derivative L2_exp(A; B1, B2) via a 4x4 block-triangular exponential."""
    n = A.shape[0]
    Zr = np.zeros((n, n))
    big = np.block([[A, B1, B2, Zr], [Zr, A, Zr, B2], [Zr, Zr, A, B1], [Zr, Zr, Zr, A]])
    return sla.expm(big)[:n, 3 * n:]


def hessian_structure(prob):
    m, n_x, D = prob.m, prob.n_x, prob.D
    rows, cols = [], []

    def emit(a, c):
        rows.append(min(a, c))
        cols.append(max(a, c))

    for k in range(prob.K - 1):
        base = k * D
        for j in range(m):
            for i in range(n_x):
                emit(base + prob.x_off + i, base + prob.u_off + j)
        for i in range(n_x):
            emit(base + prob.x_off + i, base + prob.dt_off)
        for j in range(m):
            for i in range(j + 1):
                emit(base + prob.u_off + i, base + prob.u_off + j)
        for j in range(m):
            emit(base + prob.u_off + j, base + prob.dt_off)
        emit(base + prob.dt_off, base + prob.dt_off)
    return np.array(rows, dtype=np.int64) + 1, np.array(cols, dtype=np.int64) + 1


def hessian_values(prob, Z, mu):
    """Values of the Hessian of  sum_k mu_k . delta_k  in hessian_structure order."""
    vals = np.empty((prob.K - 1, prob.nnz_hess_knot))
    m, n_x = prob.m, prob.n_x
    for k in range(prob.K - 1):
        X, Xn, u, dt = _knot(prob, Z, k)
        M = mu[k * n_x:(k + 1) * n_x].reshape(prob.b, prob.n_b, order="F")
        Gu = prob.G(u)
        A = dt * Gu
        E = sla.expm(A)
        F = [sla.expm_frechet(A, dt * prob.Gj[j], compute_expm=False) for j in range(m)]
        o = 0
        for j in range(m):
            vals[k, o:o + n_x] = (-(F[j].T @ M)).reshape(-1, order="F")
            o += n_x
        vals[k, o:o + n_x] = (-((Gu @ E).T @ M)).reshape(-1, order="F")
        o += n_x
        for j in range(m):
            for i in range(j + 1):
                L2 = _frechet2(A, dt * prob.Gj[i], dt * prob.Gj[j])
                vals[k, o] = -np.sum(M * (L2 @ X))
                o += 1
        for j in range(m):
            vals[k, o] = -np.sum(M * ((prob.Gj[j] @ E + Gu @ F[j]) @ X))
            o += 1
        vals[k, o] = -np.sum(M * (Gu @ Gu @ E @ X))
    return vals.reshape(-1)


def dense(vals, rows, cols, shape):
    """COO -> dense with duplicates accumulating (test/test_utils.jl:17-30)."""
    out = np.zeros(shape)
    np.add.at(out, (rows - 1, cols - 1), vals)
    return out


def lagrangian(prob, Z, mu):
    return float(mu @ residual(prob, Z))
