"""Time-dependent (carrier-modulated) drives: the knot constraint and its Jacobian (oracle; test infrastructure).

The reference builds ``TimeDependentBilinearIntegrator(Ghat, x, u, :t, traj)`` when ``sys.time_dependent``
(/root/reference/src/control/integrators.jl:38-46, 63, 110, 158-192) with ``Ghat(u, t) = sys.G(u, t)``; for a system
whose drives are ``ModulatedDrive(LinearDrive(H_j, j), c_j)`` (src/quantum/systems/drives.jl:342-370,
quantum_systems.jl:575-597) that generator is

    Ghat(u, t) = G_drift + sum_j c_j(t) u_j G_j ,          drive_coeff(d, u, t) = c_j(t) u_j   (drives.jl:387-388)

and the constraint at knot k uses the knot's own time row,  delta_k = x_{k+1} - exp(dt_k Ghat(u_k, t_k)) x_k.
DirectTrajOpt's source is not part of the reference tree (Project.toml:8), so the evaluation point t_k and the
position of the extra d/dt column are "parity unpinned"; the derivative itself is pinned by finite differences of
the residual (tests/test_oracle.py).

Jacobian per knot: the entries of oracle/knot.py (with d/du_j scaled by c_j(t_k)) followed by one more column,
    d delta_k / d t_k = - sum_j c_j'(t_k) u_j L_exp(dt Ghat; dt G_j) x_k ,       col = k D + t_off.
"""
import numpy as np
import scipy.linalg as sla

from . import knot as KN


def coefficients(prob, Z, t_off, mods, dmods):
    """c[j, k] = c_j(t_k), cdot[j, k] = c_j'(t_k)  (None = unmodulated drive: 1, 0)."""
    t = Z[t_off, :]
    c = np.ones((prob.m, prob.K))
    cd = np.zeros((prob.m, prob.K))
    for j in range(prob.m):
        if mods[j] is not None:
            c[j] = [mods[j](tk) for tk in t]
            cd[j] = [dmods[j](tk) for tk in t]
    return c, cd


def residual(prob, Z, c):
    out = np.empty(prob.dim)
    for k in range(prob.K - 1):
        X, Xn, u, dt = KN._knot(prob, Z, k)
        E = sla.expm(dt * prob.G(c[:, k] * u))
        out[k * prob.n_x:(k + 1) * prob.n_x] = (Xn - E @ X).reshape(-1, order="F")
    return out


def nnz_jac_knot(prob):
    return prob.nnz_jac_knot + prob.n_x


def jacobian_structure(prob, t_off):
    r0, c0 = KN.jacobian_structure(prob)
    n, n_x, D = prob.nnz_jac_knot, prob.n_x, prob.D
    rows, cols = [], []
    for k in range(prob.K - 1):
        rows.append(r0[k * n:(k + 1) * n])
        cols.append(c0[k * n:(k + 1) * n])
        rows.append(k * n_x + np.arange(n_x, dtype=np.int64) + 1)
        cols.append(np.full(n_x, k * D + t_off + 1, dtype=np.int64))
    return np.concatenate(rows), np.concatenate(cols)


def jacobian_values(prob, Z, c, cd):
    b, n_b, m, n_x = prob.b, prob.n_b, prob.m, prob.n_x
    vals = np.empty((prob.K - 1, nnz_jac_knot(prob)))
    for k in range(prob.K - 1):
        X, Xn, u, dt = KN._knot(prob, Z, k)
        Gu = prob.G(c[:, k] * u)
        A = dt * Gu
        E = sla.expm(A)
        o = 0
        for cc in range(n_b):
            vals[k, o:o + b * b] = (-E).reshape(-1, order="F")
            o += b * b
        tcol = np.zeros(n_x)
        for j in range(m):
            FjX = (sla.expm_frechet(A, dt * prob.Gj[j], compute_expm=False) @ X).reshape(-1, order="F")
            vals[k, o:o + n_x] = -c[j, k] * FjX
            tcol -= cd[j, k] * u[j] * FjX
            o += n_x
        vals[k, o:o + n_x] = (-(Gu @ (E @ X))).reshape(-1, order="F")
        o += n_x
        vals[k, o:o + n_x] = 1.0
        o += n_x
        vals[k, o:o + n_x] = tcol
    return vals.reshape(-1)
