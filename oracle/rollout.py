"""Rollout of a trajectory's piecewise-constant controls (oracle; test infrastructure).

What the reference does after a solve: ``sync_trajectory!`` (/root/reference/src/control/problems.jl:186-208)
extracts the pulse from the optimizer's trajectory and calls ``rollout!(qtraj, pulse)``
(src/quantum/trajectories/rollouts_extensions.jl:46-92), which integrates the Schroedinger / Lindblad ODE adaptively
(``abstol = reltol = 1e-8``) and saves the solution at the knot times; ``rollout_divergence``
(problems.jl:336-356) then compares the terminal states of the two solutions.  For a zero-order-hold pulse the exact
flow over the knot interval k is exp(dt_k Ghat(u_k)), so the restatement is the chain of SciPy matrix exponentials

    x_1 = x0 ,   x_{k+1} = expm(dt_k Ghat(u_k)) x_k .

Pinned by the reference's converged ``two_qubit_zoh`` solution: rolling out its controls from the identity
reproduces its stored states to 1e-9 (they satisfy the knot constraints to 6e-12 each) and its terminal unitary is
the CX gate to fidelity 0.9999999987.
"""
import numpy as np
import scipy.linalg as sla


def rollout(p, Z, x0=None):
    """States at every knot, shape (n_x, K) in Fortran order like the trajectory's state rows."""
    K = p.K
    X = (Z[p.x_off:p.x_off + p.n_x, 0] if x0 is None else np.asarray(x0, float)).reshape(p.b, p.n_b, order="F").copy()
    out = np.empty((p.n_x, K), order="F")
    for k in range(K):
        out[:, k] = X.reshape(-1, order="F")
        if k < K - 1:
            X = sla.expm(Z[p.dt_off, k] * p.G(Z[p.u_off:p.u_off + p.m, k])) @ X
    return out


def divergence(p, Z, states):
    """rollout_divergence (problems.jl:336-356) for one state component: (eps, ||dx||, ||x_collocation||)."""
    xc = Z[p.x_off:p.x_off + p.n_x, -1]
    nd, nc = np.linalg.norm(states[:, -1] - xc), np.linalg.norm(xc)
    return nd / max(nc, 1.0), nd, nc
