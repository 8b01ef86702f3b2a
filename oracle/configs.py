"""Seeded synthetic workloads C1..C5 of SURVEY.md section 8(d) (test infrastructure).

C1 single-qubit X gate (smooth_pulse_problem.jl:792-793 system), C2 the two_qubit_zoh
system (docs/literate/two_qubit_gate_validation.jl:51-56), C3/C5 three 2-level transmons
with drives on transmons 1,2 (transmon_system.jl:199-263 assembled by hand),
C4 CatSystem(cat_levels=2, buffer_levels=2) Lindblad (cat_system.jl:54-125).

Trajectories: u ~ U(-bound, bound); du, ddu ~ N(0, 0.01^2); states = exact propagation of
the initial state under those controls + N(0, 1e-3^2) noise, clipped to [-1, 1]
(named_trajectory_conversion.jl:331-332 state bounds); mu ~ N(0,1).
"""
import numpy as np
import scipy.linalg as sla

from . import isomorphisms as iso
from . import knot as KN
from . import systems as S

SEED0 = 20261017


def system_c3():
    lv = [2, 2, 2]
    a = [S.lift_operator(S.annihilate(2), i, lv) for i in range(3)]
    H_drift = np.zeros((8, 8), dtype=complex)
    for i in range(3):
        for j in range(i + 1, 3):
            H_drift += 2 * np.pi * 0.1 * (a[i] @ a[j].conj().T + a[i].conj().T @ a[j])
    H_drives = []
    for i in range(2):
        H_drives += [2 * np.pi * (a[i] + a[i].conj().T), 2 * np.pi * 1j * (a[i] - a[i].conj().T)]
    return S.QuantumSystem(H_drift, H_drives, [0.1] * 4)


def problem(cfg, K=None):
    """Returns (KnotProblem, x0, bounds, duration)."""
    if cfg == 1:
        s = S.QuantumSystem(S.PAULI_Z, [S.PAULI_X], [1.0])
        G0, Gj = s.G_parts()
        p = KN.make_problem("unitary", G0, Gj, K or 50)
        return p, iso.operator_to_iso_vec(np.eye(2)), s.drive_bounds, 10.0
    if cfg == 2:
        s = S.MultiTransmonSystem([4.0, 4.1], [0.2, 0.2], [[0, 0.1], [0.1, 0]],
                                  levels_per_transmon=2, drive_bounds=0.1)
        G0, Gj = s.G_parts()
        p = KN.make_problem("unitary", G0, Gj, K or 200)
        return p, iso.operator_to_iso_vec(np.eye(4)), s.drive_bounds, 10.0
    if cfg in (3, 5):
        s = system_c3()
        G0, Gj = s.G_parts()
        p = KN.make_problem("unitary", G0, Gj, K or (1000 if cfg == 3 else 8000))
        return p, iso.operator_to_iso_vec(np.eye(8)), s.drive_bounds, 20.0 * (1 if cfg == 3 else 8)
    if cfg == 4:
        s = S.CatSystem(cat_levels=2, buffer_levels=2)
        G0, Gj = S.compact_generator_parts(s)
        p = KN.make_problem("density", G0, Gj, K or 500)
        rho0 = np.zeros((4, 4), dtype=complex)
        rho0[0, 0] = 1.0
        return p, iso.density_to_compact_iso(rho0), s.drive_bounds, 1.0
    if cfg == 6:  # ket variant of C2's system (covers the n_b = 1 closed-system path)
        s = S.MultiTransmonSystem([4.0, 4.1], [0.2, 0.2], [[0, 0.1], [0.1, 0]],
                                  levels_per_transmon=2, drive_bounds=0.1)
        G0, Gj = s.G_parts()
        p = KN.make_problem("ket", G0, Gj, K or 64)
        return p, iso.ket_to_iso(np.array([1.0, 0, 0, 0])), s.drive_bounds, 10.0
    raise ValueError(cfg)


def trajectory(cfg, K=None, noise=1e-3):
    """Returns (prob, Z (D x K Fortran), mu)."""
    p, x0, bounds, duration = problem(cfg, K)
    rng = np.random.default_rng(SEED0 + cfg)
    K = p.K
    Z = np.zeros((p.D, K), order="F")
    dt = duration / max(K - 1, 1)
    Z[p.dt_off, :] = dt
    Z[p.dt_off + 1, :] = dt * np.arange(K)
    bnd = np.asarray(bounds, dtype=float)
    Z[p.u_off:p.u_off + p.m, :] = rng.uniform(-1, 1, size=(p.m, K)) * bnd[:, None]
    Z[p.u_off + p.m:p.u_off + 3 * p.m, :] = 0.01 * rng.standard_normal((2 * p.m, K))
    X = x0.reshape(p.b, p.n_b, order="F").copy()
    for k in range(K):
        Z[p.x_off:p.x_off + p.n_x, k] = X.reshape(-1, order="F")
        if k < K - 1:
            X = sla.expm(dt * p.G(Z[p.u_off:p.u_off + p.m, k])) @ X
    Z[p.x_off:p.x_off + p.n_x, :] += noise * rng.standard_normal((p.n_x, K))
    np.clip(Z[p.x_off:p.x_off + p.n_x, :], -1.0, 1.0, out=Z[p.x_off:p.x_off + p.n_x, :])
    mu = rng.standard_normal(p.dim)
    return p, Z, mu
