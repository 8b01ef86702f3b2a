"""Third, independent statement of the Lagrangian Hessian of one knot (oracle; test infrastructure): the
adjoint-pairing form of the second derivative of a matrix polynomial.

For the Taylor polynomial p(B) = sum_{l <= M} B^l / l!  of  B = dt A(u),  A(u) = G0 + sum_j u_j G_j, a multiplier block
L = reshape(mu_k) and the knot's state block X (what BilinearIntegrator evaluates: /root/reference/src/control/
integrators.jl:35-95; the blocks and their order are oracle/knot.py's), with

    V_k   = sum_{l >= k} c_l (B^T)^(l-k) L          adjoint Horner iterates   (V_M = c_M L, V_k = c_k L + B^T V_{k+1})
    y_n   = B^n X,    z^a_n = D(B^n)[E_a] X         forward power jets        (z^a_n = B z^a_{n-1} + E_a y_{n-1})

the derivatives of  phi = <L, p(B) X>  are single sums

    <L, Dp(B)[E_a] X>         = sum_{n=0}^{M-1} <V_{n+1}, E_a y_n>
    <L, D2p(B)[E_a, E_b] X>   = sum_{n=1}^{M-1} <V_{n+1}, E_a z^b_n + E_b z^a_n>

(no second-order jets: ten of the twenty column tiles the CUDA Hessian kernels carry for a 3-qubit system exist only to
be contracted with L at the end).  Directions: E_j = dt G_j for the controls, E_t = A for the time step, and the mixed
term d2B / du_j d dt = G_j adds <L, Dp(B)[G_j] X> to the (u_j, dt) entry.  The Hessian of  mu . delta  is MINUS these.

This is the formulation DESIGN.md section 8 proposes for the next Hessian kernel; here it is pinned against the
Pade / Frechet oracle (tests/test_oracle.py)."""
import math

import numpy as np


def knot_hessian(G0, Gj, X, L, u, dt, degree=40):
    """Hessian values of one knot in oracle/knot.py's order:
    (x,u_j) j<m | (x,dt) | (u_i,u_j) i<=j, j major | (u_j,dt) | (dt,dt)."""
    m = len(Gj)
    A = G0 + sum(u[j] * Gj[j] for j in range(m))
    B = dt * A
    M = degree
    c = [1.0 / math.factorial(l) for l in range(M + 1)]
    dirs = [dt * Gj[j] for j in range(m)] + [A]                 # E_1 .. E_m, E_t
    nd = m + 1
    # adjoint Horner iterates V_1 .. V_M (V[k]), k descending
    V = [None] * (M + 2)
    V[M] = c[M] * L
    for k in range(M - 1, -1, -1):
        V[k] = c[k] * L + B.T @ V[k + 1]
    # forward power sequences
    y = [X]
    z = [[np.zeros_like(X)] for _ in range(nd)]                 # z^a_0 = 0
    for n in range(1, M):
        for a in range(nd):
            z[a].append(B @ z[a][n - 1] + dirs[a] @ y[n - 1])
        y.append(B @ y[n - 1])

    def first(E):                                               # <L, Dp(B)[E] X>
        return sum(np.sum(V[n + 1] * (E @ y[n])) for n in range(M))

    def second(a, b):                                           # <L, D2p(B)[E_a, E_b] X>
        return sum(np.sum(V[n + 1] * (dirs[a] @ z[b][n] + dirs[b] @ z[a][n])) for n in range(1, M))

    # (x, .) blocks: (Dp(B)[E])^T L as matrices, from the adjoint first-order jets  W^a = sum_n (B^T)^.. -- equivalently
    # d/dX of <L, Dp(B)[E_a] X>:  sum_n (E_a B^n)^T V_{n+1}
    def x_block(E):
        acc = np.zeros_like(X)
        Bn = np.eye(B.shape[0])
        for n in range(M):
            acc += (E @ Bn).T @ V[n + 1]
            Bn = B @ Bn
        return acc

    out = []
    for j in range(m):
        out.append(-x_block(dirs[j]).reshape(-1, order="F"))
    out.append(-x_block(A).reshape(-1, order="F"))
    sc = []
    for j in range(m):
        for i in range(j + 1):
            sc.append(-second(i, j))
    for j in range(m):
        sc.append(-(second(j, m) + first(Gj[j])))
    sc.append(-second(m, m))
    return np.concatenate(out + [np.array(sc)])


def hessian_values(prob, Z, mu, degree=40):
    vals = []
    for k in range(prob.K - 1):
        X = Z[prob.x_off:prob.x_off + prob.n_x, k].reshape(prob.b, prob.n_b, order="F")
        L = mu[k * prob.n_x:(k + 1) * prob.n_x].reshape(prob.b, prob.n_b, order="F")
        u = Z[prob.u_off:prob.u_off + prob.m, k]
        vals.append(knot_hessian(prob.G0, list(prob.Gj), X, L, u, Z[prob.dt_off, k], degree))
    return np.concatenate(vals)
