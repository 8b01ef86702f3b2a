"""CPU tests: the oracle against the reference's own known answers and cached solutions."""
import numpy as np
import pytest

from oracle import configs as C
from oracle import cport as CP
from oracle import isomorphisms as iso
from oracle import knot as KN
from oracle import systems as S
from tests import golden_util as GU


# ---- isomorphism KATs (reference: src/quantum/primitives/isomorphisms.jl:471-643) ---------
def test_ket_iso_kat():
    psi = np.array([-1j, 2 + 3j])
    assert np.array_equal(iso.ket_to_iso(psi), [0, 2, -1, 3])            # :473
    assert np.allclose(iso.iso_to_ket(iso.ket_to_iso(psi)), psi)


def test_operator_iso_vec_kat():
    assert np.array_equal(iso.operator_to_iso_vec(np.eye(2)), [1, 0, 0, 0, 0, 1, 0, 0])   # :479
    XY = np.array([[0, 1 - 1j], [1 + 1j, 0]])
    assert np.array_equal(iso.operator_to_iso_vec(XY), [0, 1, 0, 1, 1, 0, -1, 0])          # :497-508
    assert np.allclose(iso.iso_vec_to_operator(iso.operator_to_iso_vec(XY)), XY)


def test_hamiltonian_iso_kat():
    Hc = np.array([[1, 2], [3, 4]]) + 1j * np.array([[0, 1], [1, 0]])
    GH = iso.G(Hc)
    assert np.allclose(GH, [[0, 1, 1, 2], [1, 0, 3, 4], [-1, -2, 0, 1], [-3, -4, 1, 0]])  # :630
    assert np.allclose(iso.iso(Hc), [[1, 2, 0, -1], [3, 4, -1, 0], [0, 1, 1, 2], [1, 0, 3, 4]])
    assert np.allclose(iso.H_of_G(GH), Hc)
    assert np.allclose(iso.ad_vec(S.PAULI_X.real),
                       [[0, 1, -1, 0], [1, 0, 0, -1], [-1, 0, 0, 1], [0, -1, 1, 0]])      # :636
    assert np.allclose(iso.ad_vec(S.PAULI_Y),
                       np.array([[0, -1j, -1j, 0], [1j, 0, 0, -1j], [1j, 0, 0, -1j], [0, 1j, 1j, 0]]))


@pytest.mark.parametrize("n", [2, 3, 4, 6])
def test_compact_density_iso(n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    rho = A + A.conj().T
    x = iso.density_to_compact_iso(rho)
    L, P = iso.density_lift_matrix(n), iso.density_projection_matrix(n)
    assert x.size == n * n
    assert np.count_nonzero(L) == n * (2 * n - 1)                       # :571
    assert np.allclose(P @ L, np.eye(n * n))                           # P L = I
    assert np.allclose(L @ x, iso.density_to_iso_vec(rho))
    assert np.allclose(P @ iso.density_to_iso_vec(rho), x)
    assert np.allclose(iso.compact_iso_to_density(x), rho)


def test_compact_density_order_2x2():
    a, b, c, d = 0.7, 0.3, 0.1, -0.2
    rho = np.array([[a, c + 1j * d], [c - 1j * d, b]])
    assert np.allclose(iso.density_to_compact_iso(rho), [a, c, b, d])   # :610-617


def test_lindbladian_is_trace_preserving_and_matches_master_equation():
    rng = np.random.default_rng(7)
    n = 3
    Hh = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Hh = Hh + Hh.conj().T
    Lop = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    rho = A @ A.conj().T
    sys_ = S.OpenQuantumSystem(Hh, [], [], [Lop])
    G0, _ = S.compact_generator_parts(sys_)
    drho = -1j * (Hh @ rho - rho @ Hh) + Lop @ rho @ Lop.conj().T \
        - 0.5 * (Lop.conj().T @ Lop @ rho + rho @ Lop.conj().T @ Lop)
    assert np.allclose(G0 @ iso.density_to_compact_iso(rho), iso.density_to_compact_iso(drho))


# ---- golden trajectories: the reference's converged solutions satisfy OUR residual ---------
@pytest.mark.parametrize("name", sorted(GU.TIGHT))
def test_golden_residual_is_zero(name):
    p, Z = GU.load(name)
    assert np.abs(KN.residual(p, Z)).max() < GU.TIGHT[name]
    assert np.abs(CP.residual(p, Z)).max() < GU.TIGHT[name]


def test_golden_layout_two_qubit():
    p, Z = GU.load("two_qubit_zoh")
    dt, t = Z[p.dt_off], Z[p.dt_off + 1]
    assert np.ptp(dt) < 1e-12                                    # timesteps_all_equal
    assert np.allclose(np.cumsum(dt[:-1]), t[1:], atol=1e-9)
    assert np.all(Z[p.u_off:p.u_off + p.m, 0] == 0)              # initial control pin
    u, du, ddu = (Z[p.u_off + i * p.m:p.u_off + (i + 1) * p.m] for i in range(3))
    assert np.abs(u[:, 1:] - u[:, :-1] - dt[:-1] * du[:, :-1]).max() < 1e-12   # DerivativeIntegrator
    U = iso.iso_vec_to_operator(Z[:32, -1])
    assert abs(np.trace(S.GATE_CX.conj().T @ U)) ** 2 / 16 > 0.999999          # solved CX gate


# ---- the two oracle implementations agree; derivatives agree with finite differences -------
@pytest.mark.parametrize("cfg,K", [(1, 12), (2, 8), (3, 5), (4, 8), (6, 8)])
def test_python_oracle_vs_c_port(cfg, K):
    p, Z, mu = C.trajectory(cfg, K)
    assert np.abs(KN.residual(p, Z) - CP.residual(p, Z)).max() < 1e-13
    assert np.abs(KN.jacobian_values(p, Z) - CP.jacobian_values(p, Z)).max() < 1e-12
    assert np.abs(KN.hessian_values(p, Z, mu) - CP.hessian_values(p, Z, mu)).max() < 1e-10


@pytest.mark.parametrize("name", ["systems_cat_density", "trajectories_density"])
def test_c_port_on_large_norm_golden(name):
    p, Z = GU.load(name)       # ||dt G||_1 up to 14.8: exercises scaling in the Taylor action
    mu = np.random.default_rng(0).standard_normal(p.dim)
    assert np.abs(KN.jacobian_values(p, Z) - CP.jacobian_values(p, Z)).max() < 1e-11
    h1, h2 = KN.hessian_values(p, Z, mu), CP.hessian_values(p, Z, mu)
    assert np.abs(h1 - h2).max() < 1e-9 * max(1.0, np.abs(h1).max())


@pytest.mark.parametrize("cfg,K", [(1, 5), (2, 4), (4, 4), (6, 4)])
def test_jacobian_and_hessian_vs_finite_differences(cfg, K):
    p, Z, mu = C.trajectory(cfg, K)
    n = p.D * p.K
    rows, cols = KN.jacobian_structure(p)
    J = KN.dense(KN.jacobian_values(p, Z), rows, cols, (p.dim, n))
    assert J.shape == (p.dim, p.D * p.K + p.global_dim)          # integrators.jl:780-782
    z0 = Z.reshape(-1, order="F")
    h = 1e-6
    Jfd = np.zeros_like(J)
    for c in range(n):
        zp, zm = z0.copy(), z0.copy()
        zp[c] += h
        zm[c] -= h
        Jfd[:, c] = (KN.residual(p, zp.reshape(p.D, p.K, order="F"))
                     - KN.residual(p, zm.reshape(p.D, p.K, order="F"))) / (2 * h)
    assert np.abs(J - Jfd).max() < 1e-7
    hr, hc = KN.hessian_structure(p)
    assert np.all(hr <= hc)
    Hu = KN.dense(KN.hessian_values(p, Z, mu), hr, hc, (n, n))
    Hs = Hu + np.triu(Hu, 1).T
    # d/dz (J^T mu) by central differences of the analytic Jacobian
    Hfd = np.zeros((n, n))
    for c in range(n):
        zp, zm = z0.copy(), z0.copy()
        zp[c] += h
        zm[c] -= h
        gp = KN.dense(KN.jacobian_values(p, zp.reshape(p.D, p.K, order="F")), rows, cols, (p.dim, n)).T @ mu
        gm = KN.dense(KN.jacobian_values(p, zm.reshape(p.D, p.K, order="F")), rows, cols, (p.dim, n)).T @ mu
        Hfd[:, c] = (gp - gm) / (2 * h)
    assert np.abs(Hs - Hfd).max() < 1e-5 * max(1.0, np.abs(Hs).max())


def test_structure_counts():
    for cfg, nj, nh in [(1, 48 + 8, None), (2, 416 + 32, 175), (3, 2688 + 128, 655), (4, 304 + 16, 54)]:
        p, _, _ = C.problem(cfg, 3)[0], None, None
        assert p.nnz_jac_knot == nj
        if nh is not None:
            assert p.nnz_hess_knot == nh
        r, c = KN.jacobian_structure(p)
        assert r.dtype == np.int64 and r.min() == 1 and r.max() == p.dim
        assert len(set(zip(r.tolist(), c.tolist()))) == r.size    # no duplicates emitted


def test_linear_knot_constraints_on_reference_golden():
    """DerivativeIntegrator (u->du, du->ddu) and time consistency: the reference's converged C2
    solution satisfies them (<= 3e-14 / 2e-15); Jacobian against finite differences."""
    from oracle import linear as LN
    from tests import golden_util as GU
    p, Z = GU.load("two_qubit_zoh")
    m, n_x = p.m, p.n_x
    u, du, ddu = n_x + 2, n_x + 2 + m, n_x + 2 + 2 * m
    pairs = [(u, du, m), (du, ddu, m)]
    r = LN.residual(Z, pairs, dt_off=n_x, t_off=n_x + 1)
    assert r.size == (2 * m + 1) * (p.K - 1)
    assert np.abs(r[:2 * m * (p.K - 1)]).max() < 1e-12 and np.abs(r[2 * m * (p.K - 1):]).max() < 1e-13
    rows, cols, vals = LN.jacobian(Z, pairs, n_x, n_x + 1)
    rng = np.random.default_rng(3)
    dZ = rng.standard_normal(Z.shape)
    h = 1e-6
    fd = (LN.residual(Z + h * dZ, pairs, n_x, n_x + 1) - LN.residual(Z - h * dZ, pairs, n_x, n_x + 1)) / (2 * h)
    jv = np.zeros(r.size)
    np.add.at(jv, rows - 1, vals * dZ.reshape(-1, order="F")[cols - 1])
    assert np.abs(jv - fd).max() < 1e-8
    mu = rng.standard_normal(r.size)
    hr, hc, hv = LN.hessian(Z, mu, pairs, n_x)
    # second directional derivative of mu . r along dZ equals dZ^T H dZ (H symmetric from its upper triangle)
    f = lambda t: mu @ LN.residual(Z + t * dZ, pairs, n_x, n_x + 1)
    d2 = (f(1e-3) - 2 * f(0.0) + f(-1e-3)) / 1e-6
    z = dZ.reshape(-1, order="F")
    assert abs(2 * np.sum(hv * z[hr - 1] * z[hc - 1]) - d2) < 1e-6 * max(1.0, abs(d2))
    # TimeStepsAllEqualConstraint (_problem_templates.jl:175-180): the reference solved this problem with
    # timesteps_all_equal on, and its solution has exactly equal steps; Jacobian rows against differences
    re = LN.residual(Z, pairs, n_x, n_x + 1, dt_all_equal=True)
    assert re.size == r.size + p.K - 1 and np.all(re[r.size:] == 0.0) and np.array_equal(re[:r.size], r)
    rows, cols, vals = LN.jacobian(Z, pairs, n_x, n_x + 1, dt_all_equal=True)
    fd = (LN.residual(Z + h * dZ, pairs, n_x, n_x + 1, True) - LN.residual(Z - h * dZ, pairs, n_x, n_x + 1, True)) / (2 * h)
    jv = np.zeros(re.size)
    np.add.at(jv, rows - 1, vals * dZ.reshape(-1, order="F")[cols - 1])
    assert np.abs(jv - fd).max() < 1e-8 and rows.max() == re.size


def test_time_dependent_oracle_jacobian_against_differences():
    """Carrier-modulated drives (TimeDependentBilinearIntegrator, integrators.jl:38-46; ModulatedDrive,
    drives.jl:342-388): the restated Jacobian, the d/d t_k column included, against central differences of the
    restated residual; with constant modulations the restatement reduces to the time-independent one."""
    from oracle import configs as C
    from oracle import knot as KN
    from oracle import knot_td as TD
    p, Z, mu = C.trajectory(2, 8)
    t_off = p.dt_off + 1
    Z[t_off, :] = np.cumsum(np.r_[0, Z[p.dt_off, :-1]]) + 0.3
    w = [1.3, 0.7, 2.1, None]
    mods = [(lambda t, w=w_: np.cos(w * t)) if w_ else None for w_ in w]
    dmods = [(lambda t, w=w_: -w * np.sin(w * t)) if w_ else None for w_ in w]
    c, cd = TD.coefficients(p, Z, t_off, mods, dmods)
    r = TD.residual(p, Z, c)
    rows, cols = TD.jacobian_structure(p, t_off)
    vals = TD.jacobian_values(p, Z, c, cd)
    assert rows.size == vals.size == (p.K - 1) * (p.nnz_jac_knot + p.n_x)
    rng = np.random.default_rng(0)
    dZ, h = rng.standard_normal(Z.shape), 1e-6

    def res(Zz):
        return TD.residual(p, Zz, TD.coefficients(p, Zz, t_off, mods, dmods)[0])
    fd = (res(Z + h * dZ) - res(Z - h * dZ)) / (2 * h)
    jv = np.zeros(r.size)
    np.add.at(jv, rows - 1, vals * dZ.reshape(-1, order="F")[cols - 1])
    assert np.abs(jv - fd).max() < 1e-7
    one = [None] * p.m
    c1, cd1 = TD.coefficients(p, Z, t_off, one, one)
    assert np.array_equal(TD.residual(p, Z, c1), KN.residual(p, Z))
    v1 = TD.jacobian_values(p, Z, c1, cd1).reshape(p.K - 1, -1)
    assert np.array_equal(v1[:, :p.nnz_jac_knot].reshape(-1), KN.jacobian_values(p, Z)) and np.all(v1[:, p.nnz_jac_knot:] == 0)


def test_rollout_oracle_on_reference_solution():
    """Rolling out the controls of the reference's converged two_qubit_zoh solution from the identity reproduces
    its stored states (each knot constraint holds to 6e-12), the divergence of problems.jl:336-356 is ~1e-10 and
    the rolled-out terminal unitary is the CX gate to the fidelity the reference reports."""
    from oracle import rollout as RO
    from oracle import objectives as OB
    from tests import golden_util as GU
    p, Z = GU.load("two_qubit_zoh")
    S = RO.rollout(p, Z)
    assert S.shape == (p.n_x, p.K) and np.array_equal(S[:, 0], Z[:p.n_x, 0])
    assert np.abs(S - Z[:p.n_x]).max() < 1e-9
    eps, nd, nc = RO.divergence(p, Z, S)
    assert eps < 1e-9 and abs(nc - 2.0) < 1e-8          # ||iso_vec(U)||_2 = sqrt(d) for a unitary, d = 4
    CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], complex)
    J, _ = OB.unitary_infidelity(S[:, -1], CX, 1.0)
    assert abs((1.0 - J) - 0.9999999985) < 5e-10
    # a different initial state: linearity of the chain
    S2 = RO.rollout(p, Z, x0=2.0 * Z[:p.n_x, 0])
    assert np.abs(S2 - 2.0 * S).max() < 1e-12


def test_multi_ket_golden_is_a_valid_input():
    """The reference's MultiKetTrajectory solution (stopped at max_iter): every state block obeys its own
    dynamics constraint to the reference's solve-level tolerance, with the shared control rows."""
    probs, Z = GU.load_multi()
    assert Z.shape == (16, 100) and len(probs) == 2
    for p in probs:
        d = KN.residual(p, Z)
        assert np.abs(d).max() < 1e-2                      # smooth_pulse_problem.jl:781-784
        assert np.abs(d - CP.residual(p, Z)).max() < 1e-12  # both oracle algorithms agree on it


# ---- objectives ---------------------------------------------------------------
def _fd(f, x, h=1e-6):
    g = np.zeros_like(x)
    for i in range(x.size):
        e = np.zeros_like(x)
        e[i] = h
        g[i] = (f(x + e) - f(x - e)) / (2 * h)
    return g


def test_objective_kats_from_the_reference_tests():
    """The numeric expectations the reference's own objective tests hold (objectives.jl:483-667)."""
    from oracle import objectives as OB
    p0, p1 = np.array([1, 0], complex), np.array([0, 1], complex)
    goals = [p1, p0]
    asym = [iso.ket_to_iso(p1), iso.ket_to_iso(0.5 * p0)]
    # :592-593  F = |0.9*1 + 0.1*1/2|^2 = 0.9025 and |0.1*1 + 0.9*1/2|^2 = 0.3025
    assert np.isclose(OB.coherent_ket_infidelity(asym, goals, 100.0, [0.9, 0.1])[0], 100.0 * (1 - 0.9025), rtol=1e-12)
    assert np.isclose(OB.coherent_ket_infidelity(asym, goals, 100.0, [0.1, 0.9])[0], 100.0 * (1 - 0.3025), rtol=1e-12)
    # :648-657 weighted mean normalised by the weight sum; only ratios matter
    assert np.isclose(OB.coherent_ket_fidelity(asym, goals, [0.9, 0.1]), abs(0.9 * 1 + 0.1 * 0.5) ** 2)
    assert np.isclose(OB.coherent_ket_fidelity(asym, goals, [9.0, 1.0]), OB.coherent_ket_fidelity(asym, goals, [0.9, 0.1]))
    # :660-666 uniform weights are bit-for-bit the unweighted value, also when 1/n is inexact
    x3, g3 = [iso.ket_to_iso(p1), iso.ket_to_iso(0.5 * p0), iso.ket_to_iso(0.25 * p1)], [p1, p0, p1]
    F = OB.coherent_ket_fidelity(x3, g3)
    assert OB.coherent_ket_fidelity(x3, g3, [1 / 3] * 3) == F and OB.coherent_ket_fidelity(x3, g3, [1.0] * 3) == F
    # :537 perfect coherent transfer, :559 opposite phases give F = 0
    assert OB.coherent_ket_infidelity([iso.ket_to_iso(p1), iso.ket_to_iso(p0)], goals)[0] < 1e-10
    assert OB.coherent_ket_infidelity([iso.ket_to_iso(p1), iso.ket_to_iso(-p0)], goals)[0] > 50.0
    # :627 identical kets
    assert np.isclose(OB.coherent_ket_fidelity([iso.ket_to_iso(p1), iso.ket_to_iso(p0)], goals), 1.0)


def test_objectives_on_the_reference_solutions():
    """The reference's converged two-qubit solution was obtained by minimising exactly this objective
    (smooth_pulse_problem.jl:534-542): its terminal unitary is CX up to a global phase, so the
    restated loss must vanish there; same for the coherent ket loss on the MultiKetTrajectory solution."""
    from oracle import objectives as OB
    p, Z = GU.load("two_qubit_zoh")
    CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], complex)
    J, g = OB.unitary_infidelity(Z[:p.n_x, -1], CX, 100.0)
    assert 0 <= J < 1e-5
    probs, Zm = GU.load_multi()
    p0, p1 = np.array([1, 0], complex), np.array([0, 1], complex)
    Jc, _ = OB.coherent_ket_infidelity([Zm[0:4, -1], Zm[4:8, -1]], [p1, p0], 100.0)
    assert 0 <= Jc < 1e-3
    for x, goal in ((Zm[0:4, -1], p1), (Zm[4:8, -1], p0)):
        assert OB.ket_infidelity(x, goal, 100.0)[0] < 1e-3


def test_objective_gradients_vs_finite_differences():
    from oracle import objectives as OB
    rng = np.random.default_rng(11)
    cplx = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    g = cplx(4)
    g /= np.linalg.norm(g)
    for x in (rng.standard_normal(8), 0.3 * rng.standard_normal(8)):      # F > 1 and F < 1 branches of |1 - F|
        assert np.abs(OB.ket_infidelity(x, g)[1] - _fd(lambda y: OB.ket_infidelity(y, g)[0], x)).max() < 1e-6
    n = 4
    x = 0.3 * rng.standard_normal(2 * n * n)
    Ug = np.linalg.qr(cplx(n, n))[0]
    assert np.abs(OB.unitary_infidelity(x, Ug)[1] - _fd(lambda y: OB.unitary_infidelity(y, Ug)[0], x)).max() < 1e-6
    sub, Us = [0, 2], np.linalg.qr(cplx(2, 2))[0]
    f = lambda y: OB.unitary_infidelity(y, Us, subspace=sub)[0]
    assert np.abs(OB.unitary_infidelity(x, Us, subspace=sub)[1] - _fd(f, x)).max() < 1e-6
    x = 0.3 * rng.standard_normal(9)
    A = cplx(3, 3)
    rg = A @ A.conj().T
    rg /= np.trace(rg).real
    assert np.abs(OB.density_infidelity(x, rg)[1] - _fd(lambda y: OB.density_infidelity(y, rg)[0], x)).max() < 1e-6
    psi = cplx(3)
    psi /= np.linalg.norm(psi)
    f = lambda y: OB.density_pure_state_infidelity(y, psi)[0]
    assert np.abs(OB.density_pure_state_infidelity(x, psi)[1] - _fd(f, x)).max() < 1e-6
    # density loss agrees with the definition on a physical state: tr(rho rho) of a pure state is 1
    rho = np.outer(psi, psi.conj())
    assert OB.density_infidelity(iso.density_to_compact_iso(rho), rho)[0] < 1e-12
    assert OB.density_pure_state_infidelity(iso.density_to_compact_iso(rho), psi)[0] < 1e-12
    # identical unitary up to a global phase
    assert OB.unitary_infidelity(iso.operator_to_iso_vec(np.exp(0.7j) * Ug), Ug)[0] < 1e-10
    V, dt = rng.standard_normal((3, 7)), 0.1 + rng.random(7)
    for pw in (0, 1, 2):
        J, gV, gdt = OB.quadratic_regularizer(V, dt, [1.0, 2.0, 0.5], dt_power=pw)
        f = lambda y: OB.quadratic_regularizer(y[:21].reshape(3, 7), y[21:], [1.0, 2.0, 0.5], dt_power=pw)[0]
        fd = _fd(f, np.concatenate([V.reshape(-1), dt]))
        assert np.abs(np.concatenate([gV.reshape(-1), gdt]) - fd).max() < 1e-6


def test_objective_hessians_against_central_differences():
    """Second derivatives of the objective oracle: hessian_from_gradient is exact for the (piecewise) quadratic
    losses; here it is checked against an independent second difference of the VALUE, and the regularizer's
    analytic blocks against differences of its gradient."""
    from oracle import objectives as OB
    rng = np.random.default_rng(12)
    cplx = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    N = 3
    Ug = np.linalg.qr(cplx(N, N))[0]
    x = 0.4 * rng.standard_normal(2 * N * N)
    for kw in ({}, {"subspace": [0, 2]}):
        goal = Ug if not kw else np.linalg.qr(cplx(2, 2))[0]
        f = lambda y: OB.unitary_infidelity(y, goal, 10.0, **kw)
        H = OB.hessian_from_gradient(lambda y: f(y)[1], x)
        assert np.abs(H - H.T).max() == 0.0
        h = 1e-4
        for (i, j) in ((0, 0), (1, 7), (5, 12), (17, 3)):
            ei, ej = np.eye(x.size)[i] * h, np.eye(x.size)[j] * h
            fd = (f(x + ei + ej)[0] - f(x + ei - ej)[0] - f(x - ei + ej)[0] + f(x - ei - ej)[0]) / (4 * h * h)
            assert abs(fd - H[i, j]) < 1e-5 * max(1.0, abs(H[i, j]))
    psi = cplx(4)
    psi /= np.linalg.norm(psi)
    xk = 0.5 * rng.standard_normal(8)
    Hk = OB.hessian_from_gradient(lambda y: OB.ket_infidelity(y, psi, 3.0)[1], xk)
    c = -3.0 * (1.0 if 1 - abs(np.vdot(psi, xk[:4] + 1j * xk[4:])) ** 2 >= 0 else -1.0)
    a_re, a_im = np.concatenate([psi.real, psi.imag]), np.concatenate([-psi.imag, psi.real])
    assert np.abs(Hk - 2 * c * (np.outer(a_re, a_re) + np.outer(a_im, a_im))).max() < 1e-9
    # regularizer: analytic blocks vs differences of its gradient
    V, dt, R, base = rng.standard_normal((2, 5)), 0.1 + rng.random(5), np.array([0.3, 0.7]), rng.standard_normal((2, 5))
    for pw in (0, 1, 2):
        dvv, dvt, dtt = OB.quadratic_regularizer_hessian(V, dt, R, base, pw)
        h = 1e-6
        for t in range(5):
            for i in range(2):
                Vp, Vm = V.copy(), V.copy()
                Vp[i, t] += h
                Vm[i, t] -= h
                gp, gm = OB.quadratic_regularizer(Vp, dt, R, base, pw), OB.quadratic_regularizer(Vm, dt, R, base, pw)
                assert abs((gp[1][i, t] - gm[1][i, t]) / (2 * h) - dvv[i, t]) < 1e-7
                assert abs((gp[2][t] - gm[2][t]) / (2 * h) - dvt[i, t]) < 1e-7
            dp, dm = dt.copy(), dt.copy()
            dp[t] += h
            dm[t] -= h
            gp, gm = OB.quadratic_regularizer(V, dp, R, base, pw), OB.quadratic_regularizer(V, dm, R, base, pw)
            assert abs((gp[2][t] - gm[2][t]) / (2 * h) - dtt[t]) < 1e-7


def test_hessian_adjoint_pairing_matches_frechet_oracle():
    """oracle/hessian_pairing.py (adjoint Horner iterates paired with forward power jets: no second-order jets) against
    the Pade / Frechet statement, on a unitary, a ket and a non-normal density generator."""
    from oracle import hessian_pairing as HP
    import dataclasses
    for cfg, K in ((1, 6), (2, 5), (4, 5), (6, 5)):
        p, Z, mu = C.trajectory(cfg)              # BASELINE time steps (a short K would stretch dt: ||dt G|| ~ 56 for C4)
        p = dataclasses.replace(p, K=K)
        Z, mu = np.asfortranarray(Z[:, :K]), mu[:p.dim]
        h = HP.hessian_values(p, Z, mu)
        ho = KN.hessian_values(p, Z, mu)
        assert np.abs(h - ho).max() < 1e-9 * max(1.0, np.abs(ho).max())
