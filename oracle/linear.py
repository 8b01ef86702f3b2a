"""Linear knot constraints (oracle; test infrastructure): DerivativeIntegrator and time consistency.

DirectTrajOpt's ``DerivativeIntegrator(x, xdot, traj)`` enforces  x_{k+1} - x_k - dt_k * xdot_k = 0
(used for u -> du and du -> ddu: /root/reference/src/control/templates/smooth_pulse_problem.jl:267-275;
"du[k] = (u[k+1] - u[k]) / dt", spline_pulse_problem.jl:363-366) and DirectTrajOpt applies
t_{k+1} - t_k - dt_k = 0 whenever :t and :dt are present (smooth_pulse_problem.jl:277).  The
reference's converged ``two_qubit_zoh`` trajectory satisfies both to 3e-14 / 2e-15,
which pins this restatement.

``TimeStepsAllEqualConstraint`` (pushed when ``piccolo_options.timesteps_all_equal``,
src/control/templates/_problem_templates.jl:175-180; type defined in DirectTrajOpt, source absent) is restated as
dt_{k+1} - dt_k = 0, k = 1..K-1 ("parity unpinned" for the row form; every reference solution satisfies it exactly).

Row order: pair-major, knot-major inside a pair, component fastest; time rows, then equal-timestep rows last.  Jacobian values
per derivative row: d x_k (-1), d xdot_k (-dt), d dt_k (-xdot), d x_{k+1} (+1); per time row:
d t_k (-1), d dt_k (-1), d t_{k+1} (+1).  Hessian: (xdot_k[i], dt_k) = -mu per derivative row.
"""
import numpy as np


def residual(Z, pairs, dt_off, t_off=None, dt_all_equal=False):
    D, K = Z.shape
    out = []
    for x_off, xd_off, dim in pairs:
        r = Z[x_off:x_off + dim, 1:] - Z[x_off:x_off + dim, :-1] - Z[dt_off, :-1] * Z[xd_off:xd_off + dim, :-1]
        out.append(r.reshape(-1, order="F"))
    if t_off is not None:
        out.append(Z[t_off, 1:] - Z[t_off, :-1] - Z[dt_off, :-1])
    if dt_all_equal:   # TimeStepsAllEqualConstraint (_problem_templates.jl:175-180): dt_{k+1} - dt_k
        out.append(Z[dt_off, 1:] - Z[dt_off, :-1])
    return np.concatenate(out) if out else np.zeros(0)


def jacobian(Z, pairs, dt_off, t_off=None, dt_all_equal=False):
    """(rows, cols, vals), 1-based, in the documented order."""
    D, K = Z.shape
    rows, cols, vals = [], [], []
    r = 1
    for x_off, xd_off, dim in pairs:
        for k in range(K - 1):
            c0 = k * D + 1
            for i in range(dim):
                rows += [r] * 4
                cols += [c0 + x_off + i, c0 + xd_off + i, c0 + dt_off, c0 + D + x_off + i]
                vals += [-1.0, -Z[dt_off, k], -Z[xd_off + i, k], 1.0]
                r += 1
    if t_off is not None:
        for k in range(K - 1):
            c0 = k * D + 1
            rows += [r] * 3
            cols += [c0 + t_off, c0 + dt_off, c0 + D + t_off]
            vals += [-1.0, -1.0, 1.0]
            r += 1
    if dt_all_equal:
        for k in range(K - 1):
            c0 = k * D + 1
            rows += [r] * 2
            cols += [c0 + dt_off, c0 + D + dt_off]
            vals += [-1.0, 1.0]
            r += 1
    return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(vals)


def hessian(Z, mu, pairs, dt_off):
    D, K = Z.shape
    rows, cols, vals = [], [], []
    r = 0
    for x_off, xd_off, dim in pairs:
        for k in range(K - 1):
            for i in range(dim):
                a, c = k * D + 1 + xd_off + i, k * D + 1 + dt_off
                rows.append(min(a, c)); cols.append(max(a, c)); vals.append(-mu[r])
                r += 1
    return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(vals)
