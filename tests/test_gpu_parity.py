"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs.

Tolerances (float64 path; north star: residual match <= 1e-10):
  residual  : 1e-12 absolute   (values are O(1))
  Jacobian  : 1e-11 absolute   (entries O(1..10))
  Hessian   : 1e-9 relative to max|H| (entries up to O(1e2), second derivatives)
COO index arrays: bit-exact.
"""
import numpy as np
import pytest

import piccolo_b200 as pb
from oracle import configs as C
from oracle import cport as CP
from oracle import knot as KN
from tests import golden_util as GU

pytestmark = pytest.mark.gpu

RES_TOL, JAC_TOL, HESS_RTOL = 1e-12, 1e-11, 1e-9
PATH_TOL = 1e-13   # two kernels evaluating the same quantity (different summation order)


def make(p, algorithm="auto", **kw):
    return pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off,
                                     dt_off=p.dt_off, u_off=p.u_off, algorithm=algorithm, **kw)


def algorithms(p):
    """Every kind is checked through both the jet kernels and whatever `auto` ships (for b <= 24 with sparse
    drive generators -- C4's compact Lindbladian included -- that is the tensor-core kernel)."""
    return ["generic", "auto"]


def expected_auto(p):
    """What PB2_ALG_AUTO must select (decided at construction, pb2_api.cu): the tensor-core path for
    b <= 24 with at most 4 nonzeros per drive-generator row and at most 8 column tiles."""
    if p.b > 24 or p.n_b > 12:
        return "generic"
    w = max([int((np.asarray(g) != 0).sum(1).max()) for g in p.Gj] + [0])
    ncT = p.b // 2 if p.kind != "density" else p.b
    tiles = (ncT + (1 + p.m) * p.n_b + 7) // 8
    return "dmma" if (w <= 4 and tiles <= 8) else "generic"


def check_all(p, Z, mu, B, hess=True):
    d, v = B.residual_jacobian(Z)
    assert np.abs(d - KN.residual(p, Z)).max() < RES_TOL
    assert np.abs(v - KN.jacobian_values(p, Z)).max() < JAC_TOL
    # separate entry points give the same numbers as the fused one (a residual-only call may run a
    # different kernel than the fused call -- the 3-qubit shape does -- so: to rounding, not bitwise)
    d2 = np.empty(B.dim)
    B.evaluate_(d2, Z)
    assert np.abs(d - d2).max() < PATH_TOL
    assert np.abs(v - B.jacobian_values(Z)).max() < PATH_TOL
    if hess:
        h = B.hessian_values(Z, mu)
        ho = KN.hessian_values(p, Z, mu)
        assert np.abs(h - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())


@pytest.mark.parametrize("cfg,K", [(1, 50), (2, 33), (3, 20), (4, 40), (6, 17)])
def test_synthetic_configs_vs_oracle(cfg, K):
    p, Z, mu = C.trajectory(cfg, K)
    for alg in algorithms(p):
        B = make(p, alg)
        if alg == "auto":
            assert B.algorithm == expected_auto(p)
            if cfg in (1, 2, 3, 4):
                assert B.algorithm == "dmma"          # every BASELINE config ships on the tensor-core path
                assert B.hessian_algorithm == ("u8h" if cfg == 3 else "dmmah")     # ... the Hessian too
        else:
            assert B.hessian_algorithm == "generic"
        check_all(p, Z, mu, B)
        B.close()


@pytest.mark.parametrize("name", sorted(n for n in GU.META if n != "trajectories_multi"))
def test_golden_trajectories(name):
    """The reference's own converged solutions: delta ~ 0 from the CUDA path, and full parity."""
    p, Z = GU.load(name)
    mu = np.random.default_rng(1).standard_normal(p.dim)
    for alg in algorithms(p):
        B = make(p, alg)
        if alg == "auto":
            assert B.algorithm == expected_auto(p)
        check_all(p, Z, mu, B)
        if name in GU.TIGHT:
            d = np.empty(B.dim)
            B.evaluate_(d, Z)
            assert np.abs(d).max() < GU.TIGHT[name]
        B.close()


def test_structure_bit_exact():
    for cfg in (1, 2, 4, 6):
        p, Z, mu = C.trajectory(cfg, 7)
        B = make(p)
        r, c = B.jacobian_structure()
        ro, co = KN.jacobian_structure(p)
        assert r.dtype == np.int64 and np.array_equal(r, ro) and np.array_equal(c, co)
        r, c = B.hessian_structure()
        ro, co = KN.hessian_structure(p)
        assert np.array_equal(r, ro) and np.array_equal(c, co)
        assert B.dim == p.dim == p.n_x * (p.K - 1)                    # integrators.jl:309
        J = pb.eval_jacobian(B, Z)
        assert J.shape == (B.dim, p.D * p.K + p.global_dim)           # integrators.jl:780-782
        B.close()


def test_edge_cases():
    # K = 1: no constraints at all; K = 2: a single knot; dt = 0: identity propagator
    p, Z, mu = C.trajectory(2, 2)
    B = make(p)
    check_all(p, Z, mu, B)
    B.close()
    p1, Z1, _ = C.trajectory(2, 1)
    B = make(p1)
    assert B.dim == 0 and B.nnz_jac == 0
    d, v = B.residual_jacobian(Z1)
    assert d.size == 0 and v.size == 0
    B.close()
    p, Z, mu = C.trajectory(6, 6)
    Z[p.dt_off, :] = 0.0
    Z[p.u_off:p.u_off + p.m, 2] = 0.0
    for alg in algorithms(p):
        B = make(p, alg)
        check_all(p, Z, mu, B)
        B.close()


def test_large_norm_and_degenerate_generators():
    # big dt (||dt G|| ~ 30: many squarings / large phases) and exactly degenerate spectra (u = 0)
    p, Z, mu = C.trajectory(2, 9)
    Z[p.dt_off, :] = 3.0
    Z[p.u_off:p.u_off + p.m, 4:] = 0.0
    for alg in algorithms(p):
        B = make(p, alg)
        d, v = B.residual_jacobian(Z)
        assert np.abs(d - KN.residual(p, Z)).max() < 1e-10
        assert np.abs(v - KN.jacobian_values(p, Z)).max() < 1e-9
        h, ho = B.hessian_values(Z, mu), KN.hessian_values(p, Z, mu)
        assert np.abs(h - ho).max() < 1e-8 * np.abs(ho).max()
        B.close()


def test_full_size_properties_c3():
    """BASELINE size (d=8, m=4, K=1000): oracle on a knot subset + size-independent properties."""
    p, Z, mu = C.trajectory(3, 1000)
    B = make(p)
    d, v = B.residual_jacobian(Z)
    # (1) the C++ port (independent algorithm) on the full trajectory
    assert np.abs(d - CP.residual(p, Z)).max() < RES_TOL
    assert np.abs(v - CP.jacobian_values(p, Z)).max() < JAC_TOL
    # (2) SciPy oracle on a strided subset of knots
    sub = np.arange(0, p.K - 1, 97)
    for k in sub:
        pk = KN.make_problem(p.kind, p.G0, p.Gj, 2)
        Zk = np.asfortranarray(Z[:, k:k + 2])
        assert np.abs(d[k * p.n_x:(k + 1) * p.n_x] - KN.residual(pk, Zk)).max() < RES_TOL
        assert np.abs(v[k * p.nnz_jac_knot:(k + 1) * p.nnz_jac_knot] - KN.jacobian_values(pk, Zk)).max() < JAC_TOL
    # (3) unitarity of every propagator block written to the Jacobian: E^T E = I
    V = v.reshape(p.K - 1, p.nnz_jac_knot)
    E = -V[:, :p.b * p.b].reshape(-1, p.b, p.b)
    assert np.abs(np.einsum("kij,kil->kjl", E, E) - np.eye(p.b)).max() < 1e-13
    # (4) the n_b replicated blocks are identical, identity entries are exactly 1
    for c in range(1, p.n_b):
        assert np.array_equal(V[:, :p.b * p.b], V[:, c * p.b * p.b:(c + 1) * p.b * p.b])
    assert np.all(V[:, -p.n_x:] == 1.0)
    # (5) linearity in the state: delta(x) is affine in Z's state rows
    Z2 = Z.copy(order="F")
    Z2[p.x_off:p.x_off + p.n_x, :] *= 2.0
    d2, v2 = B.residual_jacobian(Z2)
    assert np.abs(d2 - 2.0 * d).max() < 1e-12
    assert np.array_equal(v2[: p.b * p.b], v[: p.b * p.b])
    # (6) exactly propagated states give delta == 0
    p0, Z0, _ = C.trajectory(3, 1000, noise=0.0)
    d0, _ = B.residual_jacobian(Z0)
    assert np.abs(d0).max() < 1e-13
    # (7) Hessian: symmetric contraction check  z^T H z  vs second difference of mu.delta along z
    h = B.hessian_values(Z, mu)
    assert np.abs(h - CP.hessian_values(p, Z, mu)).max() < HESS_RTOL * np.abs(h).max()
    B.close()


def _strided_scipy_check(p, Z, d, v, stride, res_tol=RES_TOL, jac_tol=JAC_TOL):
    for k in range(0, p.K - 1, stride):
        pk = KN.make_problem(p.kind, p.G0, p.Gj, 2)
        Zk = np.asfortranarray(Z[:, k:k + 2])
        assert np.abs(d[k * p.n_x:(k + 1) * p.n_x] - KN.residual(pk, Zk)).max() < res_tol
        assert np.abs(v[k * p.nnz_jac_knot:(k + 1) * p.nnz_jac_knot] - KN.jacobian_values(pk, Zk)).max() < jac_tol


def test_full_size_c4_density_through_the_shipped_path():
    """BASELINE config C4 as written (CatSystem Lindbladian, d = 4 -> compact iso 16, m = 2, K = 500) through
    PB2_ALG_AUTO, which must select the tensor-core kernel (knot_dmma<2, W>): the non-normal compact
    generator (integrators.jl:82-95, open_quantum_systems.jl:541-562, 607-636) is the risky case for a
    norm-bounded Taylor action, so it is checked against BOTH oracles on every knot."""
    p, Z, mu = C.trajectory(4, 500)
    B = make(p, "auto")
    assert B.algorithm == "dmma"
    d, v = B.residual_jacobian(Z)
    assert np.abs(d - CP.residual(p, Z)).max() < RES_TOL
    assert np.abs(v - CP.jacobian_values(p, Z)).max() < JAC_TOL
    assert np.abs(d - KN.residual(p, Z)).max() < RES_TOL            # SciPy Pade-13 + expm_frechet, all 499 knots
    assert np.abs(v - KN.jacobian_values(p, Z)).max() < JAC_TOL
    d1 = np.empty(B.dim)
    B.evaluate_(d1, Z)                                               # residual-only launch (no jets)
    assert np.abs(d1 - d).max() < PATH_TOL
    h, ho = B.hessian_values(Z, mu), KN.hessian_values(p, Z, mu)
    assert np.abs(h - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())
    # the jet kernels on the same inputs: two CUDA paths, one answer
    Bg = make(p, "generic")
    dg, vg = Bg.residual_jacobian(Z)
    assert np.abs(dg - d).max() < PATH_TOL and np.abs(vg - v).max() < 1e-12
    Bg.close()
    # trace preservation of the Lindblad propagator: sum of the diagonal entries of rho is conserved,
    # i.e. the rows of E that the compact iso assigns to Re rho_jj sum to the same functional
    dd = int(round(np.sqrt(p.b)))
    diag_rows = [j * (j + 1) // 2 + j for j in range(dd)]           # Re rho[j, j] in the compact order (isomorphisms.jl:181-189)
    E = -v.reshape(p.K - 1, p.nnz_jac_knot)[:, :p.b * p.b].reshape(-1, p.b, p.b).transpose(0, 2, 1)   # E[k][row][col]
    tr_row = E[:, diag_rows, :].sum(1)
    want = np.zeros(p.b)
    want[diag_rows] = 1.0
    assert np.abs(tr_row - want).max() < 1e-13
    B.close()


def test_full_size_c5_eight_thousand_knots():
    """BASELINE config C5 (the C3 system with K = 8000): the persistent kernels walk ~ 54 knots per SM,
    several slab-ring wrap-arounds and both mbarrier phases.  Full C++-port parity, SciPy on a stride."""
    p, Z, mu = C.trajectory(5, 8000)
    B = make(p)
    assert B.algorithm == "dmma"
    d, v = B.residual_jacobian(Z)
    assert np.abs(d - CP.residual(p, Z)).max() < RES_TOL
    assert np.abs(v - CP.jacobian_values(p, Z)).max() < JAC_TOL
    _strided_scipy_check(p, Z, d, v, 397)
    V = v.reshape(p.K - 1, p.nnz_jac_knot)
    E = -V[:, :p.b * p.b].reshape(-1, p.b, p.b)
    assert np.abs(np.einsum("kij,kil->kjl", E, E) - np.eye(p.b)).max() < 1e-13
    for c in range(1, p.n_b):
        assert np.array_equal(V[:, :p.b * p.b], V[:, c * p.b * p.b:(c + 1) * p.b * p.b])
    assert np.all(V[:, -p.n_x:] == 1.0)
    h = B.hessian_values(Z, mu)
    assert np.abs(h - CP.hessian_values(p, Z, mu)).max() < HESS_RTOL * np.abs(h).max()
    B.close()


def test_c_program_through_the_abi():
    """tests/c/test_capi.c: a plain C program (no Python, no ctypes) that includes include/piccolo_b200.h, links
    libpiccolo_b200.so and evaluates BASELINE config C1 through every entry point, against the CPU port."""
    import os
    import subprocess
    cdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
    CP.lib()                                                      # the checker library exists
    subprocess.check_call(["make", "-C", cdir, "-B", "test_capi"])
    r = subprocess.run([os.path.join(cdir, "test_capi")], capture_output=True, text=True)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


def test_dense_blocks_layout():
    """pb2_desc.dense_blocks: the d/dx_k block as the full n_x x n_x block, zeros included -- the same
    sparse matrix as the canonical I (x) E layout, entry for entry, for every kind."""
    for cfg, K in ((3, 40), (2, 12), (6, 9), (4, 11)):
        p, Z, mu = C.trajectory(cfg, K)
        Bc, Bd = make(p), make(p, dense_blocks=True)
        assert Bd.nnz_jac == (p.K - 1) * (p.n_x * p.n_x + (p.m + 2) * p.n_x)
        Jc, Jd = pb.eval_jacobian(Bc, Z).toarray(), pb.eval_jacobian(Bd, Z).toarray()
        assert np.array_equal(Jc, Jd)
        r, c = Bd.jacobian_structure()
        assert r.size == Bd.nnz_jac and len(set(zip(r.tolist(), c.tolist()))) == r.size      # no duplicates
        dc, vc = Bc.residual_jacobian(Z)
        dd, vd = Bd.residual_jacobian(Z)
        assert np.array_equal(dc, dd)
        Vd = vd.reshape(p.K - 1, -1)
        blk = Vd[:, :p.n_x * p.n_x].reshape(-1, p.n_x, p.n_x)      # [knot][col][row]
        for a in range(p.n_b):
            for bq in range(p.n_b):
                sub = blk[:, a * p.b:(a + 1) * p.b, bq * p.b:(bq + 1) * p.b]
                if a != bq:
                    assert np.all(sub == 0.0)
        assert np.array_equal(Vd[:, p.n_x * p.n_x:], vc.reshape(p.K - 1, -1)[:, p.n_b * p.b * p.b:])
        Bc.close()
        Bd.close()


def test_live_handles_are_independent():
    """Distinct handles are independent (include/piccolo_b200.h): creating a handle for a small problem while a
    handle for a large one is alive must not disturb the large one's launches (the dynamic shared-memory cap is a
    property of the kernel function, not of a handle), and no entry point changes the caller's current device."""
    import torch
    pL, ZL, muL = C.trajectory(3, 12)            # unitary 16 x 16: the jet kernels need > 48 KB of shared memory
    BL = make(pL, "generic")
    hL = BL.hessian_values(ZL, muL)
    dL, vL = BL.residual_jacobian(ZL)
    pS, ZS, muS = C.trajectory(6, 9)             # a small ket problem, created while the first handle is alive
    BS = make(pS, "generic")
    check_all(pS, ZS, muS, BS)
    assert np.array_equal(BL.hessian_values(ZL, muL), hL)
    d2, v2 = BL.residual_jacobian(ZL)
    assert np.array_equal(d2, dL) and np.array_equal(v2, vL)
    trajS = pb.NamedTrajectory(ZS, {"ψ̃": range(pS.x_off, pS.x_off + pS.n_x), "Δt": range(pS.dt_off, pS.dt_off + 1),
                                    "u": range(pS.u_off, pS.u_off + pS.m)})
    J = pb.QuadraticRegularizer("u", trajS, 1.0)          # objective handle: same kind of function attribute
    J.value_gradient(ZS)
    assert np.array_equal(BL.hessian_values(ZL, muL), hL)
    J.close()
    assert torch.cuda.current_device() == 0
    BS.close()
    BL.close()


def test_device_pointer_api_matches_host_api():
    import torch
    p, Z, mu = C.trajectory(2, 40)
    B = make(p)
    d, v = B.residual_jacobian(Z)
    dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
    dd = torch.zeros(B.dim, dtype=torch.float64, device="cuda")
    dv = torch.zeros(B.nnz_jac, dtype=torch.float64, device="cuda")
    dh = torch.zeros(B.nnz_hess, dtype=torch.float64, device="cuda")
    dmu = torch.from_numpy(mu).cuda()
    torch.cuda.synchronize()
    # the kernels must run on the stream that is passed: synchronizing THAT stream only
    # (not the device) has to be enough to see the results
    side = torch.cuda.Stream()
    assert side.cuda_stream != 0
    B.residual_jacobian_device(dZ, dd, dv, side.cuda_stream)
    B.hessian_device(dZ, dmu, dh, side.cuda_stream)
    side.synchronize()
    assert np.array_equal(dd.cpu().numpy(), d) and np.array_equal(dv.cpu().numpy(), v)
    assert np.array_equal(dh.cpu().numpy(), B.hessian_values(Z, mu))
    # stream=None is CUDA's default stream
    dd.zero_()
    B.residual_jacobian_device(dZ, dd, None, None)
    torch.cuda.default_stream().synchronize()
    assert np.array_equal(dd.cpu().numpy(), d)
    assert B.launch_count >= 5
    B.close()


def test_error_behaviour():
    p, Z, mu = C.trajectory(1, 5)
    with pytest.raises(pb.PB2Error):
        pb.B200BilinearIntegrator("unitary", p.G0, list(p.Gj), K=5, D=3, x_off=0, dt_off=8, u_off=10)
    big = np.zeros((26, 26))   # the tensor-core path refuses what it cannot do (b > 24)
    with pytest.raises(pb.PB2Error):
        pb.B200BilinearIntegrator("density", big, [big], K=5, D=26 + 5, x_off=0, dt_off=26,
                                  u_off=28, algorithm="dmma")
    B = make(p)
    with pytest.raises(ValueError):
        B.residual_jacobian(Z[:, :3])
    B.close()


def _random_problem(kind, b, m, K, seed):
    """Random generators whose drive terms stay within the tensor-core path's ELL width."""
    rng = np.random.default_rng(seed)
    if kind == "density":
        G0 = rng.standard_normal((b, b)) * (rng.random((b, b)) < 0.6)
        Gj = []
        for _ in range(m):
            g = np.zeros((b, b))
            for r in range(b):
                cols = rng.choice(b, size=min(b, int(rng.integers(0, 4))), replace=False)
                g[r, cols] = rng.standard_normal(cols.size)
            Gj.append(g)
    else:
        d = b // 2

        def iso_gen(H):
            return np.block([[H.imag, H.real], [-H.real, H.imag]])
        H0 = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        G0 = iso_gen(H0 + H0.conj().T)
        Gj = []
        for _ in range(m):
            H = np.zeros((d, d), dtype=complex)
            perm = rng.permutation(d)
            for a in range(0, d - 1, 2):          # random pairing: <= 1 complex entry per row
                i, j = perm[a], perm[a + 1]
                H[i, j] = rng.standard_normal() + 1j * rng.standard_normal()
                H[j, i] = np.conj(H[i, j])
            H[perm[-1], perm[-1]] += rng.standard_normal()
            Gj.append(iso_gen(H))
    p = KN.make_problem(kind, G0, Gj, K)
    Z = np.asfortranarray(rng.standard_normal((p.D, K)) * 0.3)
    Z[p.dt_off, :] = 0.05 + 0.1 * rng.random(K)
    return p, Z, rng.standard_normal(p.dim)


@pytest.mark.parametrize("kind,b,m", [("density", 9, 2), ("density", 4, 1), ("density", 16, 3),
                                      ("density", 5, 0), ("ket", 6, 2), ("ket", 16, 3), ("ket", 2, 1),
                                      ("unitary", 6, 2), ("unitary", 12, 1), ("unitary", 16, 6),
                                      ("ket", 18, 4), ("unitary", 18, 2), ("density", 24, 1), ("ket", 22, 0),
                                      ("unitary", 20, 1)])
def test_tensor_core_path_shapes(kind, b, m, monkeypatch):
    """Odd / padded generator sizes, m = 0, mixed tiles (propagator, state and jet columns sharing
    one 8-column tile), the widest supported problem (8 tiles) -- through the default small-CTA kernel (knot_dmmaq)
    and through the persistent pipelined one (knot_dmma, PB2_DMMAQ=0)."""
    p, Z, mu = _random_problem(kind, b, m, 9, seed=100 * b + m)
    for env in (None, "0"):      # (17 <= b <= 24, two-transmon qutrit sizes: only the small-CTA kernel exists)
        if env is not None:
            monkeypatch.setenv("PB2_DMMAQ", env)
        B = make(p, "dmma")
        assert B.algorithm == "dmma"
        d, v = B.residual_jacobian(Z)
        assert np.abs(d - KN.residual(p, Z)).max() < 1e-11
        assert np.abs(v - KN.jacobian_values(p, Z)).max() < 1e-10
        d2 = np.empty(B.dim)
        B.evaluate_(d2, Z)                       # residual-only launch carries no jets
        assert np.abs(d2 - d).max() < 1e-13
        assert np.array_equal(v, B.jacobian_values(Z))
        B.close()


@pytest.mark.parametrize("dmmaq", ["1", "0"])
def test_tensor_core_path_substeps_and_limits(dmmaq, monkeypatch):
    """||dt G|| from tiny to ~60: the number of Taylor sub-steps is data dependent per knot."""
    monkeypatch.setenv("PB2_DMMAQ", dmmaq)
    p, Z, mu = C.trajectory(2, 12)
    Z[p.dt_off, :] = np.geomspace(1e-6, 6.0, p.K)
    B = make(p, "dmma")
    d, v = B.residual_jacobian(Z)
    assert np.abs(d - KN.residual(p, Z)).max() < 1e-10
    assert np.abs(v - KN.jacobian_values(p, Z)).max() < 2e-9
    # NaN / inf inputs poison only their own knot
    Z2 = Z.copy(order="F")
    Z2[p.u_off, 3] = np.nan
    Z2[p.dt_off, 5] = np.inf
    d2, v2 = B.residual_jacobian(Z2)
    D2 = d2.reshape(p.K - 1, p.n_x)
    assert np.isnan(D2[3]).all() and np.isnan(D2[5]).all()
    ok = [k for k in range(p.K - 1) if k not in (3, 5)]
    assert np.array_equal(D2[ok], d.reshape(p.K - 1, p.n_x)[ok])
    B.close()


def test_tensor_core_path_unaligned_outputs():
    """Device pointers that are only 8-byte aligned take the scalar-store variant."""
    import torch
    p, Z, mu = C.trajectory(3, 6)
    B = make(p, "dmma")
    d, v = B.residual_jacobian(Z)
    dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
    buf = torch.zeros(B.dim + B.nnz_jac + 3, dtype=torch.float64, device="cuda")
    dd, dv = buf[1:1 + B.dim], buf[2 + B.dim:2 + B.dim + B.nnz_jac]
    assert dd.data_ptr() % 16 == 8
    B.residual_jacobian_device(dZ, dd, dv, None)
    torch.cuda.synchronize()
    # (the aligned call above ran the warp-specialised kernel, this one the general tensor-core kernel)
    assert np.abs(dd.cpu().numpy() - d).max() < PATH_TOL and np.abs(dv.cpu().numpy() - v).max() < PATH_TOL
    assert buf[0].item() == 0.0 and buf[1 + B.dim].item() == 0.0 and buf[-1].item() == 0.0
    B.close()


def test_auto_falls_back_to_jets_at_construction():
    """A dense drive generator is outside the tensor-core path: auto picks the jet kernels when the
    handle is built (never a runtime fallback), dmma refuses."""
    rng = np.random.default_rng(5)
    b = 8
    p = KN.make_problem("density", rng.standard_normal((b, b)), [rng.standard_normal((b, b))], 6)
    Z = np.asfortranarray(0.2 * rng.standard_normal((p.D, 6)))
    B = make(p, "auto")
    assert B.algorithm == "generic"
    d, v = B.residual_jacobian(Z)
    assert np.abs(d - KN.residual(p, Z)).max() < 1e-11
    B.close()
    with pytest.raises(pb.PB2Error):
        make(p, "dmma")


# ---- the warp-specialised kernel of the 3-qubit unitary shape (b = 16, n_b = 8) ------------------

@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6])
def test_u8_kernel_drive_counts(m):
    """Odd drive counts leave the last jet warp with one tile; m = 5, 6 use five warps per group."""
    p, Z, mu = _random_problem("unitary", 16, m, 23, seed=900 + m)
    B = make(p, "dmma")
    d, v = B.residual_jacobian(Z)
    assert np.abs(d - KN.residual(p, Z)).max() < 1e-11
    assert np.abs(v - KN.jacobian_values(p, Z)).max() < 1e-10
    assert np.array_equal(v, B.jacobian_values(Z))          # Jacobian-only call (no delta output)
    d2 = np.empty(B.dim)
    B.evaluate_(d2, Z)
    assert np.abs(d2 - d).max() < 1e-13
    B.close()


def test_u8_kernel_substeps_nan_and_many_knots():
    """Data-dependent Taylor sub-steps (||dt G|| up to ~40), NaN / inf poisoning confined to one knot,
    and more knots than one wave of knot groups (several knots per group, ragged tail)."""
    p, Z, mu = C.trajectory(3, 1300)
    rng = np.random.default_rng(7)
    big = rng.choice(p.K - 1, size=40, replace=False)
    Z[p.dt_off, big] = np.geomspace(0.3, 12.0, big.size)
    B = make(p, "dmma")
    d, v = B.residual_jacobian(Z)
    ref_d, ref_v = CP.residual(p, Z), CP.jacobian_values(p, Z)
    assert np.abs(d - ref_d).max() < 1e-10
    assert np.abs(v - ref_v).max() < 5e-9
    small = np.setdiff1d(np.arange(p.K - 1), big)
    D, Dr = d.reshape(p.K - 1, -1), ref_d.reshape(p.K - 1, -1)
    V, Vr = v.reshape(p.K - 1, -1), ref_v.reshape(p.K - 1, -1)
    assert np.abs(D[small] - Dr[small]).max() < RES_TOL and np.abs(V[small] - Vr[small]).max() < JAC_TOL
    Z2 = Z.copy(order="F")
    Z2[p.u_off + 1, 5] = np.nan
    Z2[p.dt_off, 700] = np.inf
    d2, v2 = B.residual_jacobian(Z2)
    D2 = d2.reshape(p.K - 1, -1)
    assert np.isnan(D2[5]).all() and np.isnan(D2[700]).all()
    ok = np.setdiff1d(np.arange(p.K - 1), [5, 700])
    assert np.array_equal(D2[ok], D[ok])
    assert np.array_equal(v2.reshape(p.K - 1, -1)[ok], V[ok])
    B.close()


def _device_resjac(B, Z, early_z=False, pipelined=False):
    """The canonical (non-compact) device-pointer call, as the benchmark issues it."""
    import torch
    if early_z:
        B.set_option("early_z", 1)
    if pipelined:
        B.set_option("pipelined", 1)
    dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
    dd = torch.full((B.dim,), np.nan, dtype=torch.float64, device="cuda")
    dv = torch.full((B.nnz_jac,), np.nan, dtype=torch.float64, device="cuda")
    B.residual_jacobian_device(dZ, dd, dv, None)
    torch.cuda.synchronize()
    return dd.cpu().numpy(), dv.cpu().numpy()


@pytest.mark.parametrize("K", [1000, 2, 149, 700, 1037])
def test_single_round_kernel_every_occupancy(K, monkeypatch):
    """The kernels for at most seven knots per SM (default: knot_u8q, two 256-thread CTAs per SM of at most four
    knots each; PB2_U8Q=0: knot_u8p, one CTA per SM, propagator tiles first) take every such 3-qubit call: one knot in the whole grid, one per SM, ragged slot counts, the BASELINE size and the
    largest eligible size.  Canonical arrays through device pointers, with and without the early-Z promise,
    and compact records through the host-pointer path; the two-round kernel (PB2_U8P=0) must agree to
    rounding, the C++ port to the parity tolerances."""
    p, Z, mu = C.trajectory(3, K)
    B = make(p, "dmma")
    d, v = _device_resjac(B, Z)
    assert not np.isnan(d).any() and not np.isnan(v).any()          # every entry written
    assert np.abs(d - CP.residual(p, Z)).max() < RES_TOL
    assert np.abs(v - CP.jacobian_values(p, Z)).max() < JAC_TOL
    d_e, v_e = _device_resjac(B, Z, early_z=True)
    assert np.array_equal(d, d_e) and np.array_equal(v, v_e)
    d_p, v_p = _device_resjac(B, Z, pipelined=True)                 # no dependency wait at all
    assert np.array_equal(d, d_p) and np.array_equal(v, v_p)
    B.set_option("pipelined", 0)
    dh, vh = B.residual_jacobian(Z)                                 # compact records + host expansion
    assert np.array_equal(d, dh) and np.array_equal(v, vh)
    import torch
    dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
    dv2 = torch.full((B.nnz_jac,), np.nan, dtype=torch.float64, device="cuda")
    B.residual_jacobian_device(dZ, None, dv2, None)                  # Jacobian only (no delta output)
    torch.cuda.synchronize()
    assert np.array_equal(dv2.cpu().numpy(), v)
    B.close()
    # the other kernels of this shape on the same inputs: PB2_U8Q=0 -> the one-CTA-per-SM single-round kernel
    # (knot_u8p), additionally PB2_U8P=0 -> the persistent two-round kernel (knot_u8)
    monkeypatch.setenv("PB2_U8Q", "0")
    B2 = make(p, "dmma")
    d2, v2 = _device_resjac(B2, Z)
    assert np.abs(d2 - d).max() < PATH_TOL and np.abs(v2 - v).max() < PATH_TOL
    dh2, vh2 = B2.residual_jacobian(Z)
    assert np.array_equal(d2, dh2) and np.array_equal(v2, vh2)
    B2.close()
    monkeypatch.setenv("PB2_U8P", "0")
    B3 = make(p, "dmma")
    d3, v3 = _device_resjac(B3, Z)
    assert np.abs(d3 - d).max() < PATH_TOL and np.abs(v3 - v).max() < PATH_TOL
    B3.close()


def test_pipelined_back_to_back_launches():
    """PB2_OPT_PIPELINED: consecutive calls into distinct buffers overlap freely (no dependency wait) and may
    complete out of order; after a stream synchronize every one of them holds the same, correct arrays."""
    import torch
    p, Z, mu = C.trajectory(3, 1000)
    B = make(p, "dmma")
    d0, v0 = _device_resjac(B, Z)
    B.set_option("pipelined", 1)
    st = torch.cuda.Stream()
    n = 24
    dZ = [torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda() for _ in range(n)]
    dd = [torch.full((B.dim,), np.nan, dtype=torch.float64, device="cuda") for _ in range(n)]
    dv = [torch.full((B.nnz_jac,), np.nan, dtype=torch.float64, device="cuda") for _ in range(n)]
    torch.cuda.synchronize()
    for i in range(n):
        B.residual_jacobian_device(dZ[i], dd[i], dv[i], st.cuda_stream)
    st.synchronize()
    for i in range(n):
        assert np.array_equal(dd[i].cpu().numpy(), d0) and np.array_equal(dv[i].cpu().numpy(), v0)
    B.close()


@pytest.mark.parametrize("K", [1300, 8000])
def test_two_cta_kernel_many_waves(K, monkeypatch):
    """PB2_U8Q=2 lets the two-CTAs-per-SM kernel take any size (several waves of 4-knot CTAs)."""
    monkeypatch.setenv("PB2_U8Q", "2")
    p, Z, mu = C.trajectory(3 if K < 5000 else 5, K)
    B = make(p, "dmma")
    d, v = _device_resjac(B, Z, early_z=True)
    assert np.abs(d - CP.residual(p, Z)).max() < RES_TOL
    assert np.abs(v - CP.jacobian_values(p, Z)).max() < JAC_TOL
    B.close()


def test_single_round_kernel_substeps_nan_three_drives():
    """Data-dependent Taylor sub-steps and NaN confinement in the single-round kernel; m = 3 leaves warp B
    with one jet tile."""
    p, Z, mu = C.trajectory(3, 700)
    rng = np.random.default_rng(11)
    big = rng.choice(p.K - 1, size=30, replace=False)
    Z[p.dt_off, big] = np.geomspace(0.3, 12.0, big.size)
    B = make(p, "dmma")
    d, v = _device_resjac(B, Z, early_z=True)
    ref_d, ref_v = CP.residual(p, Z), CP.jacobian_values(p, Z)
    assert np.abs(d - ref_d).max() < 1e-10 and np.abs(v - ref_v).max() < 5e-9
    small = np.setdiff1d(np.arange(p.K - 1), big)
    D, Dr = d.reshape(p.K - 1, -1), ref_d.reshape(p.K - 1, -1)
    V, Vr = v.reshape(p.K - 1, -1), ref_v.reshape(p.K - 1, -1)
    assert np.abs(D[small] - Dr[small]).max() < RES_TOL and np.abs(V[small] - Vr[small]).max() < JAC_TOL
    Z2 = Z.copy(order="F")
    Z2[p.u_off + 1, 5] = np.nan
    Z2[p.dt_off, 500] = np.inf
    d2, v2 = _device_resjac(B, Z2)
    D2 = d2.reshape(p.K - 1, -1)
    assert np.isnan(D2[5]).all() and np.isnan(D2[500]).all()
    ok = np.setdiff1d(np.arange(p.K - 1), [5, 500])
    assert np.array_equal(D2[ok], D[ok]) and np.array_equal(v2.reshape(p.K - 1, -1)[ok], V[ok])
    B.close()
    p3 = KN.make_problem("unitary", p.G0, list(p.Gj[:3]), 400)      # three of C3's drives: ELL width stays 1
    Z3 = np.asfortranarray(rng.standard_normal((p3.D, 400)) * 0.3)
    Z3[p3.dt_off, :] = 0.05 + 0.1 * rng.random(400)
    B3 = make(p3, "dmma")
    d3, v3 = _device_resjac(B3, Z3)
    assert np.abs(d3 - CP.residual(p3, Z3)).max() < 1e-11 and np.abs(v3 - CP.jacobian_values(p3, Z3)).max() < 1e-10
    B3.close()


def test_single_round_kernel_general_drive_magnitudes(monkeypatch):
    """Drive generators whose nonzeros differ in magnitude take the multiply form of the coupling term
    (template UNIT = false); forcing that form on C3 (PB2_NO_UNIT) must reproduce the sign-flip form."""
    p, Z, mu = C.trajectory(3, 500)
    B = make(p, "dmma")
    d, v = _device_resjac(B, Z)
    B.close()
    monkeypatch.setenv("PB2_NO_UNIT", "1")
    B = make(p, "dmma")
    d2, v2 = _device_resjac(B, Z)
    B.close()
    monkeypatch.delenv("PB2_NO_UNIT")
    assert np.abs(d2 - d).max() < PATH_TOL and np.abs(v2 - v).max() < PATH_TOL
    # H_j -> D H_j D with a positive diagonal D: still Hermitian, still one nonzero per row, magnitudes differ
    dd = np.linspace(0.6, 1.7, 8)
    P = np.diag(np.concatenate([dd, dd]))
    Gj = [P @ g @ P for g in p.Gj]
    pn = KN.make_problem("unitary", p.G0, Gj, 300)
    Zn = np.asfortranarray(Z[:, :300])
    Bn = make(pn, "dmma")
    dn, vn = _device_resjac(Bn, Zn)
    assert np.abs(dn - CP.residual(pn, Zn)).max() < RES_TOL and np.abs(vn - CP.jacobian_values(pn, Zn)).max() < JAC_TOL
    assert np.abs(dn - KN.residual(pn, Zn)).max() < RES_TOL and np.abs(vn - KN.jacobian_values(pn, Zn)).max() < JAC_TOL
    Bn.close()


def test_u8_kernel_is_deterministic_and_matches_general_kernel(monkeypatch):
    """Same inputs -> bit-identical outputs across launches; the general tensor-core kernel (forced
    through PB2_NO_U8) agrees to rounding."""
    p, Z, mu = C.trajectory(3, 300)
    B = make(p, "dmma")
    d, v = B.residual_jacobian(Z)
    for _ in range(3):
        d2, v2 = B.residual_jacobian(Z)
        assert np.array_equal(d, d2) and np.array_equal(v, v2)
    B.close()
    monkeypatch.setenv("PB2_NO_U8", "1")
    Bg = make(p, "dmma")
    dg, vg = Bg.residual_jacobian(Z)
    assert np.abs(dg - d).max() < 1e-13 and np.abs(vg - v).max() < 1e-12
    Bg.close()


def test_compact_records_expand_to_canonical_arrays():
    """Sharded runs move compact per-knot records [E | jets, d/d dt, ones | delta] and expand them
    locally: the result must be bit-identical to the directly written canonical arrays."""
    import torch
    p, Z, mu = C.trajectory(3, 333)
    B = make(p, "dmma")
    cs = B.compact_stride
    assert cs == (p.m + 3) * 128
    d, v = B.residual_jacobian(Z)
    dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
    n = p.K - 1
    comp = torch.zeros(cs * n, dtype=torch.float64, device="cuda")
    dd = torch.zeros(B.dim, dtype=torch.float64, device="cuda")
    dv = torch.zeros(B.nnz_jac, dtype=torch.float64, device="cuda")
    B.residual_jacobian_compact_device(dZ, comp, None)
    B.expand_compact_device(comp, n, dd, dv, None)
    torch.cuda.synchronize()
    assert np.array_equal(dd.cpu().numpy(), d) and np.array_equal(dv.cpu().numpy(), v)
    # records of two "ranks" concatenated expand like one long trajectory
    comp2 = torch.cat([comp, comp])
    dd2 = torch.zeros(2 * B.dim, dtype=torch.float64, device="cuda")
    dv2 = torch.zeros(2 * B.nnz_jac, dtype=torch.float64, device="cuda")
    B.expand_compact_device(comp2, 2 * n, dd2, dv2, None)
    torch.cuda.synchronize()
    assert np.array_equal(dv2.cpu().numpy(), np.concatenate([v, v]))
    assert np.array_equal(dd2.cpu().numpy(), np.concatenate([d, d]))
    B.close()
    # shapes outside the 3-qubit unitary kernel do not offer the compact form
    p2, Z2, _ = C.trajectory(2, 20)
    B2 = make(p2)
    assert B2.compact_stride == 0
    with pytest.raises(pb.PB2Error):
        B2.residual_jacobian_compact_device(dZ, comp, None)
    B2.close()


# ---- tensor-core Lagrangian Hessian of the 3-qubit unitary shape ---------------------------------------

@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_u8_hessian_kernel_drive_counts(m):
    p, Z, mu = _random_problem("unitary", 16, m, 19, seed=700 + m)
    B = make(p, "dmma")
    h = B.hessian_values(Z, mu)
    ho = KN.hessian_values(p, Z, mu)
    assert np.abs(h - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())
    B.close()


def test_u8_hessian_kernel_substeps_many_knots_and_general_kernel(monkeypatch):
    """Sub-steps (||dt G|| up to ~40), several knots per CTA, NaN confinement, and agreement with
    the shared-memory jet kernel (forced through PB2_NO_U8H)."""
    p, Z, mu = C.trajectory(3, 500)
    rng = np.random.default_rng(11)
    big = rng.choice(p.K - 1, size=25, replace=False)
    Z[p.dt_off, big] = np.geomspace(0.3, 12.0, big.size)
    B = make(p, "dmma")
    h = B.hessian_values(Z, mu)
    ref = CP.hessian_values(p, Z, mu)
    H, R = h.reshape(p.K - 1, -1), ref.reshape(p.K - 1, -1)
    small = np.setdiff1d(np.arange(p.K - 1), big)
    assert np.abs(H[small] - R[small]).max() < HESS_RTOL * np.abs(R[small]).max()
    assert np.abs(H[big] - R[big]).max() < 1e-7 * np.abs(R[big]).max()
    assert np.array_equal(h, B.hessian_values(Z, mu))           # deterministic
    Z2 = Z.copy(order="F")
    Z2[p.u_off, 7] = np.nan
    H2 = B.hessian_values(Z2, mu).reshape(p.K - 1, -1)
    assert np.isnan(H2[7]).any()        # the poisoned knot shows it (degree-1 jets never touch G(u))
    ok = np.setdiff1d(np.arange(p.K - 1), [7])
    assert np.array_equal(H2[ok], H[ok])   # ... and only that knot
    B.close()
    monkeypatch.setenv("PB2_NO_U8H", "1")
    Bg = make(p, "dmma")
    hg = Bg.hessian_values(Z, mu)
    assert np.abs(hg.reshape(p.K - 1, -1)[small] - H[small]).max() < 1e-10 * np.abs(H[small]).max()
    Bg.close()


def test_linear_knot_constraints_match_oracle_and_golden():
    """DerivativeIntegrator pairs + time consistency through the C ABI."""
    from oracle import linear as LN
    p, Zg = GU.load("two_qubit_zoh")
    rng = np.random.default_rng(5)
    for Z in (Zg, np.asfortranarray(Zg + 0.01 * rng.standard_normal(Zg.shape))):
        traj = pb.NamedTrajectory.smooth_pulse_layout(Z, p.n_x, p.m, "Ũ⃗")
        L = pb.B200KnotLinearConstraints(traj)
        pairs, dt_off, t_off = L.pairs, L.dt_off, L.t_off
        d, v = L.residual_jacobian(Z)
        assert np.array_equal(d, LN.residual(Z, pairs, dt_off, t_off))
        ro, co, vo = LN.jacobian(Z, pairs, dt_off, t_off)
        r, c = L.jacobian_structure()
        assert np.array_equal(r, ro) and np.array_equal(c, co) and np.array_equal(v, vo)
        mu = rng.standard_normal(L.dim)
        hr, hc, hv = LN.hessian(Z, mu, pairs, dt_off)
        r, c = L.hessian_structure()
        assert np.array_equal(r, hr) and np.array_equal(c, hc) and np.array_equal(L.hessian_values(mu), hv)
        L.close()
    traj = pb.NamedTrajectory.smooth_pulse_layout(Zg, p.n_x, p.m, "Ũ⃗")
    L = pb.B200KnotLinearConstraints(traj)
    d, _ = L.residual_jacobian(Zg)
    assert np.abs(d).max() < 1e-12          # the reference's converged solution satisfies them
    L.close()
    # TimeStepsAllEqualConstraint rows (_problem_templates.jl:175-180) ride along in the same launch
    for Z in (Zg, np.asfortranarray(Zg + 0.01 * rng.standard_normal(Zg.shape))):
        traj = pb.NamedTrajectory.smooth_pulse_layout(Z, p.n_x, p.m, "Ũ⃗")
        L = pb.B200KnotLinearConstraints(traj, timesteps_all_equal=True)
        assert L.dim == (2 * p.m + 2) * (p.K - 1) and L.nnz_jac == (8 * p.m + 3 + 2) * (p.K - 1)
        d, v = L.residual_jacobian(Z)
        assert np.array_equal(d, LN.residual(Z, L.pairs, L.dt_off, L.t_off, dt_all_equal=True))
        ro, co, vo = LN.jacobian(Z, L.pairs, L.dt_off, L.t_off, dt_all_equal=True)
        r, c = L.jacobian_structure()
        assert np.array_equal(r, ro) and np.array_equal(c, co) and np.array_equal(v, vo)
        mu = rng.standard_normal(L.dim)
        hr, hc, hv = LN.hessian(Z, mu, L.pairs, L.dt_off)     # linear rows add no second derivatives
        r, c = L.hessian_structure()
        assert np.array_equal(r, hr) and np.array_equal(c, hc) and np.array_equal(L.hessian_values(mu), hv)
        if Z is Zg:
            assert np.all(d[-(p.K - 1):] == 0.0)     # the reference solved with timesteps_all_equal: exactly equal steps
        L.close()


@pytest.mark.parametrize("cfg,K", [(2, 40), (3, 300), (6, 33), (1, 20)])
def test_time_dependent_drives_match_oracle(cfg, K):
    """carrier-modulated drives (TimeDependentBilinearIntegrator, integrators.jl:38-46, with
    ModulatedDrive coefficients c_j(t) u_j, drives.jl:342-388): residual, Jacobian with the chain rule through
    c_j(t_k) and the extra d/d t_k column, COO structure, on every kernel family (3-qubit, general tensor-core,
    jets); constant modulations reproduce the time-independent handle bit for bit."""
    from oracle import knot_td as TD
    p, Z, mu = C.trajectory(cfg, K)
    t_off = p.dt_off + 1
    Z[t_off, :] = np.cumsum(np.r_[0, Z[p.dt_off, :-1]]) + 0.3
    w = [1.3, 0.7, 2.1, None, 0.4, 1.9][:p.m]
    mods = [(lambda t, w=w_: np.cos(w * t)) if w_ else None for w_ in w]
    dmods = [(lambda t, w=w_: -w * np.sin(w * t)) if w_ else None for w_ in w]
    c, cd = TD.coefficients(p, Z, t_off, mods, dmods)
    ro, co = TD.jacobian_structure(p, t_off)
    for alg in algorithms(p):
        B = make(p, alg, t_off=t_off, modulations=mods, modulation_derivs=dmods)
        assert B.time_dependent and B.nnz_jac == (p.K - 1) * TD.nnz_jac_knot(p)
        d, v = B.residual_jacobian(Z)
        assert np.abs(d - TD.residual(p, Z, c)).max() < RES_TOL
        assert np.abs(v - TD.jacobian_values(p, Z, c, cd)).max() < JAC_TOL
        r, cc = B.jacobian_structure()
        assert np.array_equal(r, ro) and np.array_equal(cc, co)
        d2 = np.empty(B.dim)
        B.evaluate_(d2, Z)
        assert np.abs(d2 - d).max() < PATH_TOL
        with pytest.raises(pb.PB2Error):
            B.hessian_values(Z, mu)                       # not available for time-dependent handles
        S, eps, _, _ = B.rollout(Z)                        # the rollout uses the same modulated propagators
        assert np.isfinite(S).all()
        B.close()
    # numeric derivative of the closures (the default when no derivative is given): same values to 1e-8
    B = make(p, "auto", t_off=t_off, modulations=mods)
    d, v = B.residual_jacobian(Z)
    assert np.abs(v - TD.jacobian_values(p, Z, c, cd)).max() < 1e-7
    B.close()
    # through the reference-style dispatch: QuantumSystem with (H, modulation) drive pairs
    if cfg == 1:
        X, Zp = np.array([[0, 1], [1, 0]]), np.diag([1.0, -1.0])
        sys_ = pb.QuantumSystem(Zp, [(X, mods[0], dmods[0])], [1.0])
        assert sys_.time_dependent
        traj = pb.NamedTrajectory.smooth_pulse_layout(Z, p.n_x, p.m, "Ũ⃗")
        Bq = pb.BilinearIntegrator(pb.UnitaryTrajectory(sys_), traj)
        assert Bq.time_dependent
        dq, vq = Bq.residual_jacobian(traj)
        assert np.abs(dq - TD.residual(p, Z, c)).max() < RES_TOL and np.abs(vq - TD.jacobian_values(p, Z, c, cd)).max() < JAC_TOL
        Bq.close()


def test_rollout_matches_expm_chain():
    """rollout of the piecewise-constant controls (rollout! / sync_trajectory!), terminal fidelity
    and rollout_divergence on the device against the SciPy expm chain: the reference's converged C2 solution
    (fidelity 0.99999999847 after rollout), C3 at full size, a ket and a Lindblad density problem."""
    import torch
    from oracle import rollout as RO
    from oracle import objectives as OB
    p, Zg = GU.load("two_qubit_zoh")
    B = make(p)
    S, eps, nd, nc = B.rollout(Zg)
    So = RO.rollout(p, Zg)
    assert np.abs(S - So).max() < 1e-10
    eo, ndo, nco = RO.divergence(p, Zg, So)
    assert abs(eps - eo) < 1e-12 and abs(nd - ndo) < 1e-12 and abs(nc - nco) < 1e-12 and eps < 1e-9
    assert np.abs(S - Zg[:p.n_x]).max() < 1e-9            # the optimizer's states ARE the rollout for this solution
    CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], complex)
    Zr = Zg.copy(order="F")
    Zr[:p.n_x, :] = S                                    # fidelity(qtraj) after sync_trajectory!: the loss on the rollout
    traj = pb.NamedTrajectory.smooth_pulse_layout(Zr, p.n_x, p.m, "Ũ⃗")
    J = pb.UnitaryInfidelityObjective(CX, "Ũ⃗", traj, Q=1.0)
    val, _ = J.value_gradient(Zr)
    assert abs((1.0 - val) - 0.9999999985) < 5e-10
    assert abs(val - OB.unitary_infidelity(So[:, -1], CX, 1.0)[0]) < 1e-12
    J.close()
    assert abs(pb.rollout_divergence(B, Zg) - eo) < 1e-12
    B.close()
    for cfg, K in ((3, 1000), (6, 64), (4, 120), (1, 50)):
        p, Z, mu = C.trajectory(cfg, K)
        B = make(p)
        x0 = Z[p.x_off:p.x_off + p.n_x, 0] * 0.5 + 0.1
        for xi in (None, x0):
            S, eps, nd, nc = B.rollout(Z, xi)
            So = RO.rollout(p, Z, xi)
            assert np.abs(S - So).max() < 1e-10
            eo, ndo, nco = RO.divergence(p, Z, So)
            assert abs(eps - eo) < 1e-10 and abs(nd - ndo) < 1e-10 and abs(nc - nco) < 1e-12
        # device-pointer form
        dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
        dS = torch.full((p.n_x * p.K,), np.nan, dtype=torch.float64, device="cuda")
        do = torch.zeros(3, dtype=torch.float64, device="cuda")
        B.rollout_device(dZ, None, dS, do, None)
        torch.cuda.synchronize()
        S0, e0, _, _ = B.rollout(Z)
        assert np.array_equal(dS.cpu().numpy().reshape(p.n_x, p.K, order="F"), S0) and do[0].item() == e0
        B.close()


def test_multi_ket_and_sampling_integrator_vectors():
    """Row a4: one integrator per state block / ensemble member, all sharing the control rows
    (integrators.jl:102-117, 134-226).  Multi-ket on the reference's own MultiKetTrajectory solution."""
    probs, Z = GU.load_multi()
    sys_ = pb.QuantumSystem(np.diag([1.0, -1.0]), [np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]])], [1, 1])
    qtraj = pb.MultiKetTrajectory(sys_, len(probs))
    traj = pb.NamedTrajectory.multi_state_layout(Z, qtraj.state_names, probs[0].b, probs[0].m)
    Bs = pb.BilinearIntegrator(qtraj, traj)
    assert [B.x_name for B in Bs] == ["ψ̃1", "ψ̃2"] and all(B.dim == probs[0].dim for B in Bs)
    for B, p in zip(Bs, probs):
        mu = np.random.default_rng(2).standard_normal(p.dim)
        check_all(p, Z, mu, B)
        d = np.empty(B.dim)
        B.evaluate_(d, Z)
        assert np.abs(d).max() < 1e-2          # the reference's own solve-level assert (smooth_pulse_problem.jl:781-784)
        r, c = B.jacobian_structure()
        ro, co = KN.jacobian_structure(p)
        assert np.array_equal(r, ro) and np.array_equal(c, co)
        B.close()
    # sampling ensemble: two members with different drifts, unitary states, shared controls
    import dataclasses
    rng = np.random.default_rng(9)
    X, Y, Zp = np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
    members = [pb.QuantumSystem(w * Zp, [X, Y], [1, 1]) for w in (1.0, 1.1)]
    st = pb.SamplingTrajectory(pb.UnitaryTrajectory, members)
    K, n_x, m = 12, 8, 2
    Zs = np.asfortranarray(0.3 * rng.standard_normal((2 * n_x + 2 + 3 * m, K)))
    Zs[2 * n_x, :] = 0.1 + 0.05 * rng.random(K)
    traj = pb.NamedTrajectory.multi_state_layout(Zs, st.state_names, n_x, m)
    Bs = pb.BilinearIntegrator(st, traj)
    assert len(Bs) == 2
    for i, (B, sys_i) in enumerate(zip(Bs, members)):
        G0, Gj = sys_i.G_parts()
        p = dataclasses.replace(KN.make_problem("unitary", G0, Gj, K), D=Zs.shape[0], x_off=i * n_x,
                                dt_off=2 * n_x, u_off=2 * n_x + 2)
        check_all(p, Zs, rng.standard_normal(p.dim), B)
        B.close()


def test_u8_kernels_with_offset_state_block():
    """A 3-qubit unitary state that is NOT the first component of the knot column (second member of a
    sampling ensemble): slab offsets, compact records and the Hessian's mu slab all use x_off."""
    import dataclasses
    p0, Z0, _ = C.trajectory(3, 90)
    rng = np.random.default_rng(21)
    n_x = p0.n_x
    Z = np.asfortranarray(np.vstack([0.1 * rng.standard_normal((n_x, p0.K)), Z0]))   # [other block | C3 column]
    p = dataclasses.replace(p0, D=Z.shape[0], x_off=n_x, dt_off=p0.dt_off + n_x, u_off=p0.u_off + n_x)
    mu = rng.standard_normal(p.dim)
    B = make(p, "dmma")
    check_all(p, Z, mu, B)
    d, v = B.residual_jacobian(Z)
    B0 = make(p0, "dmma")
    d0, v0 = B0.residual_jacobian(Z0)
    assert np.array_equal(d, d0) and np.array_equal(v, v0)          # same numbers as the un-shifted problem
    assert np.array_equal(B.hessian_values(Z, mu), B0.hessian_values(Z0, mu))
    B.close()
    B0.close()


def test_sharded_integrator_single_rank_uses_compact_records():
    """The sharding host logic on a real device (world size 1): compact records, local expansion, unpack."""
    p, Z, mu = C.trajectory(3, 61)
    S = pb.ShardedBilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                     u_off=p.u_off, rank=0, world=1, device=0)
    assert S.cs == (p.m + 3) * 128
    S.residual_jacobian(Z)
    import torch
    torch.cuda.synchronize()
    d, v = S.unpack()
    B = make(p)
    d0, v0 = B.residual_jacobian(Z)
    assert np.array_equal(d, d0) and np.array_equal(v, v0)
    B.close()
    S.local.close()


# ---- objective value + gradient --------------------------------------------------
OBJ_RTOL = 1e-12      # relative to max(|J|, 1) and to max|grad|


def _obj_check(J, Z, J_ref, g_ref):
    val, g = J.value_gradient(Z)
    assert abs(val - J_ref) < OBJ_RTOL * max(1.0, abs(J_ref))
    assert g.shape == g_ref.shape
    assert np.abs(g - g_ref).max() < OBJ_RTOL * max(1.0, np.abs(g_ref).max())
    assert J.value(Z) == val                       # value-only entry point, bitwise the same sum
    val2, g2 = J.value_gradient(Z)
    assert val2 == val and np.array_equal(g, g2)   # deterministic (knot-ordered final sum)
    return val, g


def test_objectives_match_oracle_on_reference_solutions():
    from oracle import objectives as OB
    CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], complex)
    p, Zg = GU.load("two_qubit_zoh")
    rng = np.random.default_rng(21)
    for Z in (Zg, np.asfortranarray(Zg + 0.05 * rng.standard_normal(Zg.shape))):
        traj = pb.NamedTrajectory.smooth_pulse_layout(Z, p.n_x, p.m, "Ũ⃗")
        # the objective SmoothPulseProblem assembles: infidelity + three regularizers (smooth_pulse_problem.jl:240-250)
        J = pb.UnitaryInfidelityObjective(CX, "Ũ⃗", traj, Q=100.0)
        Jo, go = OB.unitary_infidelity(Z[:p.n_x, -1], CX, 100.0)
        G = np.zeros_like(Z)
        G[:p.n_x, -1] = go
        for pw, (name, R) in enumerate((("u", 1e-2), ("du", 2e-2), ("ddu", 3e-2))):
            J = J + pb.QuadraticRegularizer(name, traj, R, dt_power=pw)
            rows = traj.components[name]
            jr, gv, gdt = OB.quadratic_regularizer(Z[rows.start:rows.stop], Z[p.n_x], R, dt_power=pw)
            Jo += jr
            G[rows.start:rows.stop] += gv
            G[p.n_x] += gdt
        val, _ = _obj_check(J, Z, Jo, G.reshape(-1, order="F"))
        J.close()
    # the reference's converged solution: the infidelity part vanishes
    traj = pb.NamedTrajectory.smooth_pulse_layout(Zg, p.n_x, p.m, "Ũ⃗")
    J = pb.UnitaryInfidelityObjective(CX, "Ũ⃗", traj)
    assert 0 <= pb.objective_value(J, traj) < 1e-5
    J.close()
    # MultiKetTrajectory solution: coherent and per-state ket losses
    probs, Zm = GU.load_multi()
    traj = pb.NamedTrajectory.multi_state_layout(Zm, ["ψ̃1", "ψ̃2"], 4, 2)
    p0, p1 = np.array([1, 0], complex), np.array([0, 1], complex)
    for w in (None, [0.9, 0.1], [0.5, 0.5]):
        J = pb.CoherentKetInfidelityObjective([p1, p0], ["ψ̃1", "ψ̃2"], traj, Q=100.0, weights=w)
        Jo, gs = OB.coherent_ket_infidelity([Zm[0:4, -1], Zm[4:8, -1]], [p1, p0], 100.0, w)
        G = np.zeros_like(Zm)
        G[0:4, -1], G[4:8, -1] = gs
        _obj_check(J, Zm, Jo, G.reshape(-1, order="F"))
        J.close()
    J = pb.KetInfidelityObjective(p1, "ψ̃1", traj) + pb.KetInfidelityObjective(p0, "ψ̃2", traj)
    (j1, g1), (j2, g2) = OB.ket_infidelity(Zm[0:4, -1], p1), OB.ket_infidelity(Zm[4:8, -1], p0)
    G = np.zeros_like(Zm)
    G[0:4, -1], G[4:8, -1] = g1, g2
    val, _ = _obj_check(J, Zm, j1 + j2, G.reshape(-1, order="F"))
    assert val < 1e-3
    J.close()


def test_objective_kats_through_the_c_abi():
    """The values the reference's own tests expect (objectives.jl:537, 559, 592-593), on the device."""
    N = 10
    p0, p1 = np.array([1, 0], complex), np.array([0, 1], complex)
    iso = lambda v: np.concatenate([v.real, v.imag])
    rng = np.random.default_rng(2)

    def traj_of(a, b):
        Z = np.zeros((10, N))
        Z[0:4], Z[4:8] = iso(a)[:, None], iso(b)[:, None]
        Z[8], Z[9] = 0.1, 0.0
        Z[9] = rng.standard_normal(N)
        return pb.NamedTrajectory(Z, {"ψ̃1": range(0, 4), "ψ̃2": range(4, 8), "Δt": range(8, 9), "u": range(9, 10)})

    goals, names = [p1, p0], ["ψ̃1", "ψ̃2"]
    t = traj_of(p1, 0.5 * p0)
    J1 = pb.CoherentKetInfidelityObjective(goals, names, t, Q=100.0, weights=[0.9, 0.1])
    J2 = pb.CoherentKetInfidelityObjective(goals, names, t, Q=100.0, weights=[0.1, 0.9])
    assert np.isclose(pb.objective_value(J1, t), 100.0 * (1 - 0.9025), rtol=1e-12)
    assert np.isclose(pb.objective_value(J2, t), 100.0 * (1 - 0.3025), rtol=1e-12)
    Ju = pb.CoherentKetInfidelityObjective(goals, names, t, Q=100.0)
    Jw = pb.CoherentKetInfidelityObjective(goals, names, t, Q=100.0, weights=[0.5, 0.5])
    assert pb.objective_value(Ju, t) == pb.objective_value(Jw, t)          # uniform weights: bit-for-bit
    g = np.zeros(t.dim * t.N + t.global_dim)
    pb.gradient_(g, J1, t)
    assert not np.all(g == 0)
    assert pb.objective_value(Ju, traj_of(p1, p0)) < 1e-10                 # perfect coherent transfer
    assert pb.objective_value(Ju, traj_of(p1, -p0)) > 50.0                 # opposite phases: F = 0


def test_objective_kinds_edge_cases_and_full_size():
    from oracle import objectives as OB
    from oracle import isomorphisms as ISO
    rng = np.random.default_rng(31)
    cplx = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    # density matrix (compact iso, linear loss) + leakage at a subset of knots + regularizer with baseline / times
    n, K, m = 3, 37, 2
    D = n * n + 2 + 3 * m
    Z = np.asfortranarray(0.4 * rng.standard_normal((D, K)))
    Z[n * n] = 0.05 + 0.1 * rng.random(K)
    traj = pb.NamedTrajectory.smooth_pulse_layout(Z, n * n, m, "ρ⃗̃")
    A = cplx(n, n)
    rho_g = A @ A.conj().T
    rho_g /= np.trace(rho_g).real
    psi = cplx(n)
    psi /= np.linalg.norm(psi)
    times, Qs, idx = np.array([0, 5, 6, K - 1]), np.array([1.0, 0.5, 2.0, 3.0]), [2, 5, 7]
    base = rng.standard_normal((m, K))
    J = (pb.DensityMatrixInfidelityObjective("ρ⃗̃", rho_g, traj, Q=50.0)
         + pb.DensityMatrixPureStateInfidelityObjective("ρ⃗̃", psi, traj, Q=7.0)
         + pb.LeakageObjective(idx, "ρ⃗̃", traj, times=times, Qs=Qs)
         + pb.QuadraticRegularizer("u", traj, [0.3, 0.7], baseline=base, times=np.arange(1, K - 1), dt_power=1)
         + pb.QuadraticRegularizer("u", traj, 0.1))                       # two regularizers on the same rows
    G = np.zeros_like(Z)
    j1, g1 = OB.density_infidelity(Z[:n * n, -1], rho_g, 50.0)
    j2, g2 = OB.density_pure_state_infidelity(Z[:n * n, -1], psi, 7.0)
    G[:n * n, -1] = g1 + g2
    j3, g3 = OB.leakage(Z[:n * n][:, times], idx, Qs)
    G[:n * n, times] += g3
    u = traj.components["u"]
    tt = np.arange(1, K - 1)
    j4, gv, gdt = OB.quadratic_regularizer(Z[u.start:u.stop][:, tt], Z[n * n, tt], [0.3, 0.7], base[:, tt], 1)
    G[u.start:u.stop, 1:K - 1] += gv
    G[n * n, 1:K - 1] += gdt
    j5, gv5, _ = OB.quadratic_regularizer(Z[u.start:u.stop], Z[n * n], 0.1)
    G[u.start:u.stop] += gv5
    _obj_check(J, Z, j1 + j2 + j3 + j4 + j5, G.reshape(-1, order="F"))
    J.close()
    # EmbeddedOperator subspace fidelity (objectives.jl:339-345): 3-level transmon, qubit subspace
    N, sub = 3, [0, 1]
    Zu = np.asfortranarray(0.5 * rng.standard_normal((2 * N * N + 2 + 3, 5)))
    traj = pb.NamedTrajectory.smooth_pulse_layout(Zu, 2 * N * N, 1, "Ũ⃗")
    Us = np.linalg.qr(cplx(2, 2))[0]
    J = pb.UnitaryInfidelityObjective(Us, "Ũ⃗", traj, Q=100.0, subspace=sub)
    jo, go = OB.unitary_infidelity(Zu[:2 * N * N, -1], Us, 100.0, subspace=sub)
    G = np.zeros_like(Zu)
    G[:2 * N * N, -1] = go
    _obj_check(J, Zu, jo, G.reshape(-1, order="F"))
    J.close()
    with pytest.raises(ValueError):
        pb.UnitaryInfidelityObjective(2 * Us, "Ũ⃗", traj, subspace=sub)      # subspace form needs a unitary goal
    # K = 1 (terminal knot is the only knot) and an empty objective
    t1 = pb.NamedTrajectory.smooth_pulse_layout(Zu[:, :1].copy(), 2 * N * N, 1, "Ũ⃗")
    J = pb.UnitaryInfidelityObjective(np.eye(N), "Ũ⃗", t1)
    jo, go = OB.unitary_infidelity(Zu[:2 * N * N, 0], np.eye(N), 100.0)
    _obj_check(J, Zu[:, :1], jo, np.concatenate([go, np.zeros(5)]))
    J.close()
    E = pb.B200Objective(traj)
    val, g = E.value_gradient(Zu)
    assert val == 0.0 and not g.any()
    E.close()
    # bad rows / times are rejected by the library
    with pytest.raises(pb.PB2Error):
        pb.LeakageObjective([0], "Ũ⃗", traj, times=[99]).value(Zu)
    # full BASELINE C3 size: identity propagator at every knot => J = 0, perturbed => oracle
    p, Z, _ = C.trajectory(3)
    traj = pb.NamedTrajectory(Z, {"Ũ⃗": range(p.x_off, p.x_off + p.n_x), "Δt": range(p.dt_off, p.dt_off + 1),
                                  "u": range(p.u_off, p.u_off + p.m)})
    Ug = np.linalg.qr(cplx(8, 8))[0]
    J = pb.UnitaryInfidelityObjective(Ug, "Ũ⃗", traj) + pb.QuadraticRegularizer("u", traj, 1e-2, dt_power=2)
    jo, go = OB.unitary_infidelity(Z[p.x_off:p.x_off + p.n_x, -1], Ug, 100.0)
    jr, gv, gdt = OB.quadratic_regularizer(Z[p.u_off:p.u_off + p.m], Z[p.dt_off], 1e-2, dt_power=2)
    G = np.zeros_like(Z)
    G[p.x_off:p.x_off + p.n_x, -1] = go
    G[p.u_off:p.u_off + p.m] += gv
    G[p.dt_off] += gdt
    _obj_check(J, Z, jo + jr, G.reshape(-1, order="F"))
    Zp = Z.copy(order="F")
    Zp[p.x_off:p.x_off + p.n_x, -1] = ISO.operator_to_iso_vec(np.exp(0.3j) * Ug)
    Jt = pb.UnitaryInfidelityObjective(Ug, "Ũ⃗", traj)
    assert Jt.value(Zp) < 1e-10
    # device-pointer entry point on the caller's stream
    import torch
    dZ = torch.from_numpy(np.ascontiguousarray(Z.T)).cuda()       # knot-major = rows of Z.T
    dJ, dg = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.empty(Z.size, dtype=torch.float64, device="cuda")
    J.value_gradient_device(dZ, dJ, dg, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    val, g = J.value_gradient(Z)
    assert dJ.item() == val and np.array_equal(dg.cpu().numpy(), g)
    J.close()
    Jt.close()


# ---- fused multi-state integrator: all kets of a MultiKetTrajectory in one launch --------------------
def test_fused_multi_ket_integrator_matches_per_state_integrators():
    """n_b = number of states sharing the generator.  Same numbers as the vector of per-state integrators
    (integrators.jl:102-117), rows knot-major instead of state-major; checked on the reference's own
    MultiKetTrajectory solution, against the oracle, for the general kernels and the 8-ket 3-qubit shape
    (which runs the headline kernel)."""
    import dataclasses
    probs, Z = GU.load_multi()
    sys_ = pb.QuantumSystem(np.diag([1.0, -1.0]), [np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]])], [1, 1])
    qtraj = pb.MultiKetTrajectory(sys_, len(probs))
    traj = pb.NamedTrajectory.multi_state_layout(Z, qtraj.state_names, probs[0].b, probs[0].m)
    n_s, b, K = len(probs), probs[0].b, probs[0].K
    pf = dataclasses.replace(probs[0], n_b=n_s, x_off=0)
    mu = np.random.default_rng(4).standard_normal(pf.dim)
    for alg in ("generic", "auto"):
        F = pb.BilinearIntegrator(qtraj, traj, fused=True, algorithm=alg)
        assert F.dim == n_s * b * (K - 1) and F.n_b == n_s
        check_all(pf, Z, mu, F)
        r, c = F.jacobian_structure()
        ro, co = KN.jacobian_structure(pf)
        assert np.array_equal(r, ro) and np.array_equal(c, co)
        hr, hc = F.hessian_structure()
        hro, hco = KN.hessian_structure(pf)
        assert np.array_equal(hr, hro) and np.array_equal(hc, hco)
        d, v = F.residual_jacobian(Z)
        Bs = pb.BilinearIntegrator(qtraj, traj, algorithm=alg)
        for s, B in enumerate(Bs):
            ds, vs = B.residual_jacobian(Z)
            # residual rows of state s inside the fused knot-major vector
            assert np.abs(d.reshape(K - 1, n_s, b)[:, s].reshape(-1) - ds).max() < 1e-14
            # same sparse matrix: compare as dense blocks through the COO structure
            rs, cs = B.jacobian_structure()
            fused = {(int(a) , int(bb)): x for a, bb, x in zip(r, c, v)}
            k_of, i_of = (rs - 1) // b, (rs - 1) % b
            rows_f = k_of * (n_s * b) + s * b + i_of + 1
            got = np.array([fused[(int(a), int(bb))] for a, bb in zip(rows_f, cs)])
            assert np.abs(got - vs).max() < 1e-13
            B.close()
        F.close()
    with pytest.raises(TypeError):
        pb.BilinearIntegrator(pb.KetTrajectory(sys_), traj, fused=True)
    # random 3-state, 2-level problem and a fused density pair (generic kernel)
    rng = np.random.default_rng(14)
    for kind, bsz, n_s in (("ket", 4, 3), ("ket", 6, 5), ("density", 4, 2)):
        m, K = 2, 19
        if kind == "ket":
            H0 = rng.standard_normal((bsz // 2, bsz // 2)); H0 = H0 + H0.T
            Hd = [rng.standard_normal((bsz // 2, bsz // 2)) + 1j * rng.standard_normal((bsz // 2, bsz // 2)) for _ in range(m)]
            s_ = pb.QuantumSystem(H0, [h + h.conj().T for h in Hd], [1] * m)
            G0, Gj = s_.G_parts()
        else:
            G0, Gj = rng.standard_normal((bsz, bsz)), [rng.standard_normal((bsz, bsz)) for _ in range(m)]
        base = KN.make_problem(kind, G0, Gj, K)
        D = n_s * bsz + 2 + 3 * m
        pf = dataclasses.replace(base, n_b=n_s, D=D, x_off=0, dt_off=n_s * bsz, u_off=n_s * bsz + 2)
        Zr = np.asfortranarray(0.4 * rng.standard_normal((D, K)))
        Zr[pf.dt_off] = 0.05 + 0.1 * rng.random(K)
        for alg in algorithms(pf):
            F = pb.B200BilinearIntegrator(kind, G0, list(Gj), K=K, D=D, x_off=0, dt_off=pf.dt_off, u_off=pf.u_off,
                                          n_states=n_s, algorithm=alg)
            check_all(pf, Zr, rng.standard_normal(pf.dim), F)
            F.close()
    # eight kets of the 3-qubit system: the fused integrator has the unitary's shape and runs knot_u8 / knot_u8h
    p3, Z3, mu3 = C.trajectory(3, 45)
    F = pb.B200BilinearIntegrator("ket", p3.G0, list(p3.Gj), K=p3.K, D=p3.D, x_off=p3.x_off, dt_off=p3.dt_off,
                                  u_off=p3.u_off, n_states=8)
    U = make(p3)
    assert F.algorithm == "dmma" and F.compact_stride == U.compact_stride > 0
    d, v = F.residual_jacobian(Z3)
    du, vu = U.residual_jacobian(Z3)
    assert np.array_equal(d, du) and np.array_equal(v, vu)
    assert np.array_equal(F.hessian_values(Z3, mu3), U.hessian_values(Z3, mu3))
    check_all(p3, Z3, mu3, F)
    F.close()
    U.close()
    with pytest.raises(ValueError):
        pb.B200BilinearIntegrator("unitary", p3.G0, list(p3.Gj), K=p3.K, D=p3.D, x_off=0, dt_off=p3.dt_off,
                                  u_off=p3.u_off, n_states=2)


def test_sharded_integrator_two_gpus_fused_exchange():
    """Two ranks, one per GPU: the fused NVLink exchange and the NCCL all-gather of the records both
    reproduce the single-GPU arrays bitwise on every rank (skipped on a one-GPU box)."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "sharded_check.py")],
                       capture_output=True, text=True, timeout=240, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


# ---- the reference's own integrator acceptance tests (integrators.jl:345-560), through the CUDA path --------
def _zero_control_traj(state_names, n_x_each, m, N=11, seed=0, zero_controls=True):
    """NamedTrajectory(qtraj, 11) stand-in: zero controls on a uniform 0..1 time grid as in the reference's
    test items (the derivative check does not need a feasible trajectory, so states are random)."""
    rng = np.random.default_rng(seed)
    D = len(state_names) * n_x_each + 2 + 3 * m
    Z = np.zeros((D, N))
    Z[:len(state_names) * n_x_each] = 0.5 * rng.standard_normal((len(state_names) * n_x_each, N))
    o = len(state_names) * n_x_each
    Z[o] = 0.1
    Z[o + 1] = np.linspace(0.0, 1.0, N)
    if not zero_controls:
        Z[o + 2:o + 2 + m] = 0.7 * rng.standard_normal((m, N))
    if len(state_names) == 1:
        return pb.NamedTrajectory.smooth_pulse_layout(Z, n_x_each, m, state_names[0])
    return pb.NamedTrajectory.multi_state_layout(Z, state_names, n_x_each, m)


@pytest.mark.parametrize("zero_controls", [True, False])
def test_reference_dispatch_test_items(zero_controls):
    X, Y, Zp = np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
    sys_ = pb.QuantumSystem(Zp, [X, Y], [1.0, 1.0])
    # "BilinearIntegrator dispatch on UnitaryTrajectory" (integrators.jl:345-363)
    qtraj = pb.UnitaryTrajectory(sys_)
    traj = _zero_control_traj([qtraj.state_name], 8, 2, zero_controls=zero_controls)
    B = pb.BilinearIntegrator(qtraj, 11, traj)
    assert isinstance(B, pb.B200BilinearIntegrator) and B.x_dim == 8 and B.dim == 8 * 10
    assert pb.test_integrator(B, traj, atol=1e-3)
    B.close()
    # "... on KetTrajectory" (:365-383)
    qtraj = pb.KetTrajectory(sys_)
    traj = _zero_control_traj([qtraj.state_name], 4, 2, zero_controls=zero_controls)
    B = pb.BilinearIntegrator(qtraj, 11, traj)
    assert B.x_dim == 4
    assert pb.test_integrator(B, traj, atol=1e-3)
    B.close()
    # "... on DensityTrajectory" (:385-413): sigma-minus decay, compact iso => x_dim = n^2
    L = np.array([[0.0, 0.1], [0.0, 0.0]], dtype=complex)
    osys = pb.OpenQuantumSystem(Zp, [X], [1.0], dissipation_operators=[L])
    qtraj = pb.DensityTrajectory(osys)
    traj = _zero_control_traj([qtraj.state_name], 4, 1, zero_controls=zero_controls)
    B = pb.BilinearIntegrator(qtraj, 11, traj)
    assert B.x_dim == osys.levels ** 2
    assert pb.test_integrator(B, traj, atol=1e-3)
    B.close()
    # "... on SamplingTrajectory (Unitary)" / "(Ket)" (:415-480): one integrator per member, shared controls
    members = [sys_, pb.QuantumSystem(1.1 * Zp, [X, Y], [1.0, 1.0])]
    for base, n_x in ((pb.UnitaryTrajectory, 8), (pb.KetTrajectory, 4)):
        st = pb.SamplingTrajectory(base, members)
        traj = _zero_control_traj(st.state_names, n_x, 2, zero_controls=zero_controls)
        Bs = pb.BilinearIntegrator(st, 11, traj)
        assert isinstance(Bs, list) and len(Bs) == 2
        for B in Bs:
            assert pb.test_integrator(B, traj, atol=1e-3)
            B.close()
    # MultiKetTrajectory (:102-117): vector of integrators, and the fused single-launch form
    mk = pb.MultiKetTrajectory(sys_, 2)
    traj = _zero_control_traj(mk.state_names, 4, 2, zero_controls=zero_controls)
    for B in pb.BilinearIntegrator(mk, 11, traj):
        assert pb.test_integrator(B, traj, atol=1e-3)
        B.close()
    F = pb.BilinearIntegrator(mk, 11, traj, fused=True)
    assert pb.test_integrator(F, traj, atol=1e-3)
    F.close()


def test_test_integrator_catches_a_wrong_jacobian(monkeypatch):
    """The acceptance test must fail when the analytic values are off."""
    X, Y, Zp = np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
    qtraj = pb.KetTrajectory(pb.QuantumSystem(Zp, [X, Y], [1.0, 1.0]))
    traj = _zero_control_traj([qtraj.state_name], 4, 2, zero_controls=False)
    B = pb.BilinearIntegrator(qtraj, 11, traj)
    good = B.jacobian_values

    def bad(Z, out=None):
        v = good(Z, out)
        v[3] += 0.01
        return v

    monkeypatch.setattr(B, "jacobian_values", bad)
    with pytest.raises(AssertionError):
        pb.test_integrator(B, traj, atol=1e-3)
    B.close()


def test_reference_sampling_ensemble_test_items():
    """integrators.jl:482-560: sampling over density / multi-ket / multi-density bases."""
    X, Y, Zp = np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
    # "SamplingTrajectory (Density)" (:482-526): compact Lindbladian per member, x_dim = n^2, and the
    # member-1 integrator agrees with the plain density integrator of the same system
    L = np.array([[0.0, 0.1], [0.0, 0.0]], dtype=complex)
    o1 = pb.OpenQuantumSystem(Zp, [X], [1.0], dissipation_operators=[L])
    o2 = pb.OpenQuantumSystem(0.95 * Zp, [X], [1.0], dissipation_operators=[L])
    base = pb.DensityTrajectory(o1)
    st = pb.SamplingTrajectory(base, [o1, o2])
    traj = _zero_control_traj(st.state_names, 4, 1, zero_controls=False)
    Bs = pb.BilinearIntegrator(st, 11, traj)
    assert isinstance(Bs, list) and len(Bs) == 2
    for B in Bs:
        assert B.x_dim == o1.levels ** 2
        assert pb.test_integrator(B, traj, atol=1e-3)
    # parity pin (:518-525): same (x_next, x, u, dt) through member 1 and through the plain integrator
    rng = np.random.default_rng(8)
    x, x_next, u, dt = rng.standard_normal(4), rng.standard_normal(4), 0.3, 0.05
    Zs = traj.data.copy(order="F")
    c = traj.components
    Zs[c["ρ⃗̃1"].start:c["ρ⃗̃1"].stop, 0], Zs[c["ρ⃗̃1"].start:c["ρ⃗̃1"].stop, 1] = x, x_next
    Zs[c["Δt"].start, 0], Zs[c["u"].start, 0] = dt, u
    d_member = np.empty(Bs[0].dim)
    Bs[0].evaluate_(d_member, Zs)
    tp = _zero_control_traj([base.state_name], 4, 1)
    Zp_ = tp.data.copy(order="F")
    Zp_[0:4, 0], Zp_[0:4, 1], Zp_[4, 0], Zp_[6, 0] = x, x_next, dt, u
    ref = pb.BilinearIntegrator(base, 11, tp)
    d_ref = np.empty(ref.dim)
    ref.evaluate_(d_ref, Zp_)
    assert np.allclose(d_member[:4], d_ref[:4], rtol=0, atol=1e-15)
    ref.close()
    for B in Bs:
        B.close()
    # "SamplingTrajectory (MultiKet)" (:528-556): 2 members x 2 kets = 4 integrators, names member-major
    s1, s2 = pb.QuantumSystem(Zp, [X, Y], [1.0, 1.0]), pb.QuantumSystem(1.1 * Zp, [X, Y], [1.0, 1.0])
    st = pb.SamplingTrajectory(pb.MultiKetTrajectory(s1, 2), [s1, s2])
    traj = _zero_control_traj(st.state_names, 4, 2, zero_controls=False)
    Bs = pb.BilinearIntegrator(st, 11, traj)
    assert len(Bs) == 4 and [B.x_name for B in Bs] == ["ψ̃1", "ψ̃2", "ψ̃3", "ψ̃4"]
    per_state = []
    for B in Bs:
        assert pb.test_integrator(B, traj, atol=1e-3)
        per_state.append(B.residual_jacobian(traj.data)[0])
        B.close()
    # fused: one launch per member evaluates both of its kets; member 2 uses its own drift
    Fs = pb.BilinearIntegrator(st, 11, traj, fused=True)
    assert len(Fs) == 2 and [F.n_b for F in Fs] == [2, 2]
    for i, F in enumerate(Fs):
        assert pb.test_integrator(F, traj, atol=1e-3)
        d = F.residual_jacobian(traj.data)[0].reshape(10, 2, 4)
        assert np.abs(d[:, 0].reshape(-1) - per_state[2 * i]).max() < 1e-14
        assert np.abs(d[:, 1].reshape(-1) - per_state[2 * i + 1]).max() < 1e-14
        F.close()
    # "SamplingTrajectory (MultiDensity)": one compact-Lindbladian integrator per (member, density)
    st = pb.SamplingTrajectory(pb.MultiDensityTrajectory(o1, 2), [o1, o2])
    traj = _zero_control_traj(st.state_names, 4, 1, zero_controls=False)
    Bs = pb.BilinearIntegrator(st, 11, traj)
    assert len(Bs) == 4 and all(B.x_dim == 4 for B in Bs)
    for B in Bs:
        assert pb.test_integrator(B, traj, atol=1e-3)
        B.close()
    # Jacobian shape pin (:780-782): dim x (traj.dim * traj.N + traj.global_dim)
    qtraj = pb.KetTrajectory(s1)
    traj = _zero_control_traj([qtraj.state_name], 4, 2, zero_controls=False)
    B = pb.BilinearIntegrator(qtraj, 11, traj)
    d = np.zeros(B.dim)
    pb.evaluate_(d, B, traj)
    assert not np.all(d == 0)
    Jm = pb.eval_jacobian(B, traj)
    assert Jm.shape == (B.dim, traj.dim * traj.N + traj.global_dim)
    B.close()


def _ensemble(cfg, K, n, rng):
    """n members of config `cfg` with perturbed drifts, their state blocks stacked at the top of the knot column
    (the layout SamplingTrajectory builds: sampling_problem.jl:389-395), sharing the dt / u rows."""
    import dataclasses
    p0, Z0, _ = C.trajectory(cfg, K)
    n_x, rest = p0.n_x, p0.D - p0.n_x
    D = n * n_x + rest
    Z = np.zeros((D, K), order="F")
    Z[n * n_x:, :] = Z0[n_x:, :]
    probs = []
    for i in range(n):
        Z[i * n_x:(i + 1) * n_x, :] = Z0[:n_x, :] + 1e-3 * rng.standard_normal((n_x, K))
        probs.append(dataclasses.replace(p0, G0=p0.G0 * (1.0 + 0.03 * i), D=D, x_off=i * n_x,
                                         dt_off=n * n_x + (p0.dt_off - n_x), u_off=n * n_x + (p0.u_off - n_x)))
    return probs, Z


@pytest.mark.parametrize("cfg,K,n,alg,fused", [(1, 40, 16, "auto", True), (2, 33, 16, "auto", True),
                                               (4, 25, 5, "auto", True), (6, 30, 7, "auto", True),
                                               (2, 20, 3, "generic", True), (3, 12, 3, "auto", False)])
def test_ensemble_batch_one_launch(cfg, K, n, alg, fused):
    """every member integrator of an ensemble in ONE launch (member = a grid axis).  Each
    member's rows are checked against the oracle and, bit for bit, against the member's own integrator."""
    rng = np.random.default_rng(100 + cfg)
    probs, Z = _ensemble(cfg, K, n, rng)
    p0 = probs[0]
    batch = pb.B200IntegratorBatch(p0.kind, [(p.G0, list(p.Gj)) for p in probs], K=K, D=p0.D,
                                   x_offs=[p.x_off for p in probs], dt_off=p0.dt_off, u_off=p0.u_off, algorithm=alg)
    assert batch.n_members == n and batch.fused == fused and batch.dim == p0.dim
    mu = rng.standard_normal((n, p0.dim))
    d, v = batch.residual_jacobian(Z)
    h = batch.hessian_values(Z, mu)
    for i, p in enumerate(probs):
        assert np.abs(d[i] - KN.residual(p, Z)).max() < RES_TOL
        assert np.abs(v[i] - KN.jacobian_values(p, Z)).max() < JAC_TOL
        ho = KN.hessian_values(p, Z, mu[i])
        assert np.abs(h[i] - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())
        B = make(p, alg)
        di, vi = B.residual_jacobian(Z)
        assert np.array_equal(d[i], di) and np.array_equal(v[i], vi)
        assert np.array_equal(h[i], B.hessian_values(Z, mu[i]))
        for got, want in ((batch.jacobian_structure(i), B.jacobian_structure()),
                          (batch.hessian_structure(i), B.hessian_structure())):
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        B.close()
    batch.close()


def test_ensemble_batch_through_the_plugin_interface():
    """BilinearIntegrator(SamplingTrajectory, traj, batched=True): the reference's constructor shape
    (integrators.jl:134-226) returning the one-launch object; members that differ in shape are refused."""
    X, Y, Zp = np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
    members = [pb.QuantumSystem(w * Zp, [X, Y], [1.0, 1.0]) for w in (1.0, 1.05, 1.1, 0.9)]
    st = pb.SamplingTrajectory(pb.KetTrajectory, members)
    rng = np.random.default_rng(5)
    K, n_x, m = 21, 4, 2
    Zs = np.asfortranarray(0.3 * rng.standard_normal((4 * n_x + 2 + 3 * m, K)))
    Zs[4 * n_x, :] = 0.1 + 0.05 * rng.random(K)
    traj = pb.NamedTrajectory.multi_state_layout(Zs, st.state_names, n_x, m)
    batch = pb.BilinearIntegrator(st, traj, batched=True)
    singles = pb.BilinearIntegrator(st, traj)
    d, v = batch.residual_jacobian(traj)
    assert batch.names == st.state_names and batch.fused
    for i, B in enumerate(singles):
        di, vi = B.residual_jacobian(traj)
        assert np.array_equal(d[i], di) and np.array_equal(v[i], vi)
        B.close()
    batch.close()
    with pytest.raises(pb.PB2Error):
        pb.B200IntegratorBatch("ket", [(np.zeros((4, 4)), []), (np.zeros((6, 6)), [])], K=5, D=20, x_offs=[0, 4],
                               dt_off=10, u_off=12)


def _dense_from_coo(rows, cols, vals, n):
    import scipy.sparse as sp
    U = sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n)).toarray()     # duplicates sum, like Ipopt
    assert (rows <= cols).all()
    return U + np.triu(U, 1).T


def test_objective_hessian_matches_oracle():
    """the objective part of eval_h (sigma * d2J) on the device.  Terminal-loss blocks against the
    oracle's complex-form Hessians, regularizer blocks against its analytic ones; sigma scales everything."""
    from oracle import objectives as OB
    rng = np.random.default_rng(77)
    cplx = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    # the reference's converged unitary solution: infidelity + the regularizers SmoothPulseProblem assembles
    p, Z = GU.load("trajectories_unitary")
    names = {"Ũ⃗": range(p.x_off, p.x_off + p.n_x), "Δt": range(p.dt_off, p.dt_off + 1),
             "u": range(p.u_off, p.u_off + p.m), "du": range(p.u_off + p.m, p.u_off + 2 * p.m)}
    traj = pb.NamedTrajectory(Z, names)
    n = p.b // 2
    Ug = np.linalg.qr(cplx(n, n))[0]
    J = (pb.UnitaryInfidelityObjective(Ug, "Ũ⃗", traj, Q=100.0) + pb.QuadraticRegularizer("u", traj, 1e-2, dt_power=2)
         + pb.QuadraticRegularizer("du", traj, [0.5] * p.m, dt_power=1) + pb.QuadraticRegularizer("Δt", traj, 2.0, dt_power=2)
         + pb.LeakageObjective([1, 3], "Ũ⃗", traj, times=[0, 3, p.K - 1], Qs=[1.0, 2.0, 0.5]))
    rows, cols = J.hessian_structure()
    nz = p.K * p.D
    for sigma in (1.0, 0.37):
        H = _dense_from_coo(rows, cols, J.hessian_values(Z, sigma), nz)
        Ho = np.zeros((nz, nz))
        xs = slice((p.K - 1) * p.D + p.x_off, (p.K - 1) * p.D + p.x_off + p.n_x)
        Ho[xs, xs] += OB.hessian_from_gradient(lambda y: OB.unitary_infidelity(y, Ug, 100.0)[1], Z[p.x_off:p.x_off + p.n_x, -1])
        for t, q in zip([0, 3, p.K - 1], [1.0, 2.0, 0.5]):
            for i in (1, 3):
                Ho[t * p.D + p.x_off + i, t * p.D + p.x_off + i] += q * 2.0 / 2
        for name, R, pw in (("u", 1e-2, 2), ("du", 0.5, 1), ("Δt", 2.0, 2)):
            r = names[name]
            dvv, dvt, dtt = OB.quadratic_regularizer_hessian(Z[r.start:r.stop], Z[p.dt_off], R, None, pw)
            for k in range(p.K):
                for i, row in enumerate(r):
                    a, b = k * p.D + row, k * p.D + p.dt_off
                    Ho[a, a] += dvv[i, k]
                    Ho[a, b] += dvt[i, k]
                    Ho[b, a] += dvt[i, k]
                Ho[k * p.D + p.dt_off, k * p.D + p.dt_off] += dtt[k]
        assert np.abs(H - sigma * Ho).max() < 1e-9 * max(1.0, np.abs(Ho).max())
    # second difference of the library's own gradient along a random direction (whole-objective consistency)
    v = rng.standard_normal(nz)
    v[(p.K - 1) * p.D + p.x_off:(p.K - 1) * p.D + p.x_off + p.n_x] *= 1e-3   # stay on one side of |1 - F|
    h = 1e-5
    gp = J.value_gradient((Z.reshape(-1, order="F") + h * v).reshape(Z.shape, order="F"))[1]
    gm = J.value_gradient((Z.reshape(-1, order="F") - h * v).reshape(Z.shape, order="F"))[1]
    H1 = _dense_from_coo(rows, cols, J.hessian_values(Z, 1.0), nz)
    assert np.abs((gp - gm)[:nz] / (2 * h) - H1 @ v).max() < 1e-5 * max(1.0, np.abs(H1 @ v).max())
    J.close()
    # a_lin-only objectives have no second derivative; kets give the rank-2 block; device entry point
    D, K = 8 + 2 + 3, 6
    Zk = np.asfortranarray(0.4 * rng.standard_normal((D, K)))
    trk = pb.NamedTrajectory.smooth_pulse_layout(Zk, 8, 1, "ψ̃")
    psi = cplx(4)
    psi /= np.linalg.norm(psi)
    Jk = pb.KetInfidelityObjective(psi, "ψ̃", trk, Q=3.0)
    Hk = _dense_from_coo(*Jk.hessian_structure(), Jk.hessian_values(Zk), K * D)
    Hko = OB.hessian_from_gradient(lambda y: OB.ket_infidelity(y, psi, 3.0)[1], Zk[:8, -1])
    assert np.abs(Hk[(K - 1) * D:(K - 1) * D + 8, (K - 1) * D:(K - 1) * D + 8] - Hko).max() < 1e-10
    import torch
    dZ = torch.from_numpy(np.ascontiguousarray(Zk.reshape(-1, order="F"))).cuda()
    dv = torch.zeros(Jk.hessian_structure()[0].size, dtype=torch.float64, device="cuda")
    Jk.hessian_device(dZ, 1.0, dv, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(dv.cpu().numpy(), Jk.hessian_values(Zk))
    Jk.close()
    rho = np.eye(2) / 2
    Zd = np.asfortranarray(rng.standard_normal((4 + 2 + 3, 4)))
    Jd = pb.DensityMatrixInfidelityObjective("ρ⃗̃", rho, pb.NamedTrajectory.smooth_pulse_layout(Zd, 4, 1, "ρ⃗̃"))
    assert Jd.hessian_structure()[0].size == 0 and Jd.hessian_values(Zd).size == 0
    Jd.close()


def test_two_process_exchange_through_the_c_abi():
    """peer setup callable from C.  Two processes (both on device 0 here; one per GPU on a
    multi-GPU box) shard a 3-qubit trajectory using only libpiccolo_b200 for the device side -- pb2_device_alloc,
    pb2_ipc_export / pb2_ipc_open, pb2_residual_jacobian_exchange_async -- and each ends with every knot's record."""
    import os
    import subprocess
    import sys
    from tests.ipc_rank import Rank
    K = 38                                   # 37 knot evaluations: ragged over two ranks
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    child = subprocess.Popen([sys.executable, os.path.join(os.path.dirname(__file__), "ipc_rank.py"), str(K), "0"],
                             stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env)
    try:
        R = Rank(0, K)
        child.stdin.write(R.export() + "\n")
        child.stdin.flush()
        R.open(1, child.stdout.readline().strip())
        child.stdin.write("go\n")
        child.stdin.flush()
        R.run()
        assert child.stdout.readline().strip() == "done"
        d, v = R.gathered()
        p, Z = R.p, R.Z
        assert np.abs(d - KN.residual(p, Z)).max() < RES_TOL
        assert np.abs(v - KN.jacobian_values(p, Z)).max() < JAC_TOL
        B = make(p)
        d1, v1 = B.residual_jacobian(Z)
        B.close()
        assert np.array_equal(d, d1) and np.array_equal(v, v1)      # bitwise what one process computes alone
        child.stdin.write("check\n")
        child.stdin.flush()
        sd, sv = (float(x) for x in child.stdout.readline().split())
        assert sd == float(np.abs(d).sum()) and sv == float(np.abs(v).sum())
        child.stdin.write("bye\n")
        child.stdin.flush()
        assert child.wait(timeout=60) == 0
        R.close()
    finally:
        if child.poll() is None:
            child.kill()


def test_host_pointer_pipeline_pinned_and_pageable(monkeypatch):
    """The host-pointer call of the 3-qubit shape uploads, evaluates and downloads chunk by chunk on three streams
    (pb2_api.cu): pinned caller buffers (DMA straight from / to them), pageable ones (staged), ragged chunk sizes,
    repeated calls on one handle, and the single-kernel path (PB2_D2H_CHUNKS=1) must all give the same bits."""
    import ctypes
    lib = pb.load_library()

    def pinned(n):
        ptr = ctypes.c_void_p()
        assert lib.pb2_host_alloc(ctypes.byref(ptr), 8 * n) == 0
        return ptr, np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,))

    for K in (1000, 259, 70):
        p, Z, _ = C.trajectory(3, K)
        B = make(p)
        d0, v0 = B.residual_jacobian(Z)                      # pageable numpy buffers
        assert np.abs(d0 - CP.residual(p, Z)).max() < RES_TOL and np.abs(v0 - CP.jacobian_values(p, Z)).max() < JAC_TOL
        pz, Zp = pinned(Z.size)
        pd, dp = pinned(B.dim)
        pv, vp = pinned(B.nnz_jac)
        for rep in range(3):
            Zr = Z + 1e-3 * rep
            Zp[:] = Zr.reshape(-1, order="F")
            dp[:] = np.nan
            vp[:] = np.nan
            assert lib.pb2_residual_jacobian(B._h, pz, pd, pv, 0) == 0
            dr, vr = B.residual_jacobian(np.asfortranarray(Zr))
            assert np.array_equal(dp, dr) and np.array_equal(vp, vr)
        monkeypatch.setenv("PB2_D2H_CHUNKS", "1")
        d1, v1 = B.residual_jacobian(np.asfortranarray(Z + 2e-3))
        monkeypatch.delenv("PB2_D2H_CHUNKS")
        assert np.array_equal(d1, dp) and np.array_equal(v1, vp)
        B.close()
        for q in (pz, pd, pv):
            lib.pb2_host_free(q)


@pytest.mark.parametrize("kind,b,m", [("density", 9, 2), ("density", 4, 1), ("density", 16, 3), ("density", 5, 0),
                                      ("ket", 6, 2), ("ket", 16, 3), ("ket", 2, 1), ("ket", 8, 6),
                                      ("unitary", 6, 2), ("unitary", 4, 3), ("unitary", 8, 4), ("unitary", 12, 1),
                                      ("ket", 18, 4), ("density", 24, 2), ("unitary", 18, 1), ("ket", 20, 0)])
def test_tensor_core_hessian_shapes(kind, b, m, monkeypatch):
    """Row a8 on the tensor cores for general generators (knot_dmmah.cuh): forward tiles (state, first- and
    second-order jets) and adjoint tiles with the TRANSPOSED generator -- no symmetry assumed (density generators
    have none).  Odd / padded sizes, m = 0, tiles shared by several column kinds; against the oracle, and against
    the jet kernel the same handle runs with PB2_NO_DMMAH=1."""
    p, Z, mu = _random_problem(kind, b, m, 9, seed=100 * b + m + 7)
    B = make(p, "dmma")
    assert B.hessian_algorithm == "dmmah"
    h = B.hessian_values(Z, mu)
    ho = KN.hessian_values(p, Z, mu)
    assert np.abs(h - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())
    assert np.array_equal(h, B.hessian_values(Z, mu))                      # fixed summation order
    monkeypatch.setenv("PB2_NO_DMMAH", "1")
    Bj = make(p, "dmma")
    assert Bj.hessian_algorithm == "generic"
    hj = Bj.hessian_values(Z, mu)
    assert np.abs(h - hj).max() < 1e-10 * max(1.0, np.abs(ho).max())
    B.close()
    Bj.close()


def test_tensor_core_hessian_substeps_nan_and_full_size(monkeypatch):
    """Data-dependent Taylor sub-steps per knot, NaN inputs confined to their knot, C2 / C4 at BASELINE sizes
    against the C++ port on every knot, and the device-pointer entry point."""
    import torch
    p, Z, mu = C.trajectory(2, 12)
    Z[p.dt_off, :] = np.geomspace(1e-6, 6.0, p.K)
    B = make(p)
    h = B.hessian_values(Z, mu)
    ho = KN.hessian_values(p, Z, mu)
    assert np.abs(h - ho).max() < 1e-8 * max(1.0, np.abs(ho).max())
    Z2 = Z.copy(order="F")
    Z2[p.u_off + 1, 4] = np.nan
    h2 = B.hessian_values(Z2, mu).reshape(p.K - 1, -1)
    assert np.isnan(h2[4]).all()
    ok = [k for k in range(p.K - 1) if k != 4]
    assert np.array_equal(h2[ok], h.reshape(p.K - 1, -1)[ok])
    B.close()
    for cfg in (2, 4):
        p, Z, mu = C.trajectory(cfg)
        B = make(p)
        h = B.hessian_values(Z, mu)
        hc = CP.hessian_values(p, Z, mu)
        assert np.abs(h - hc).max() < HESS_RTOL * max(1.0, np.abs(hc).max())
        dZ = torch.from_numpy(np.ascontiguousarray(Z.reshape(-1, order="F"))).cuda()
        dmu = torch.from_numpy(mu).cuda()
        dh = torch.zeros(B.nnz_hess, dtype=torch.float64, device="cuda")
        B.hessian_device(dZ, dmu, dh, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(dh.cpu().numpy(), h)
        B.close()


def test_tensor_core_hessian_twenty_warps(monkeypatch):
    """The 3-qubit unitary shape through the GENERAL tensor-core Hessian (15 forward + 5 adjoint tiles = 20 warps per
    knot; PB2_HESS_DMMAH=1 prefers it over knot_u8h, which is faster and stays the default): against the oracle and
    against knot_u8h."""
    p, Z, mu = C.trajectory(3, 23)
    B = make(p)
    assert B.hessian_algorithm == "u8h"
    h0 = B.hessian_values(Z, mu)
    B.close()
    monkeypatch.setenv("PB2_HESS_DMMAH", "1")
    B = make(p)
    assert B.hessian_algorithm == "dmmah"
    h = B.hessian_values(Z, mu)
    ho = KN.hessian_values(p, Z, mu)
    assert np.abs(h - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())
    assert np.abs(h - h0).max() < 1e-10 * max(1.0, np.abs(ho).max())
    B.close()


def test_single_knot_ctas_and_long_knot_columns():
    """knot_u8q runs one knot per 64-thread CTA by default (eight CTAs per SM).  A long knot column (extra
    trajectory components after the controls) leaves no room for the per-CTA table blob next to the slab: the launch
    then packs two / four knots per CTA instead -- same results; beyond that the general kernels take over."""
    import dataclasses
    p0, Z0, _ = C.trajectory(3, 40)
    base = make(p0)
    d0, v0 = base.residual_jacobian(Z0)
    base.close()
    for extra in (900, 1300, 2400):
        p = dataclasses.replace(p0, D=p0.D + extra)
        Z = np.zeros((p.D, p.K), order="F")
        Z[:p0.D] = Z0
        Z[p0.D:] = np.random.default_rng(extra).standard_normal((extra, p.K))
        B = make(p)
        d, v = B.residual_jacobian(Z)
        if extra < 2000:
            assert np.array_equal(d, d0) and np.array_equal(v, v0)
        else:       # too long for any slab-staging kernel: the general kernels take the call (Hessian included)
            assert np.abs(d - d0).max() < PATH_TOL and np.abs(v - v0).max() < PATH_TOL
            mu = np.random.default_rng(1).standard_normal(p.dim)
            ho = KN.hessian_values(p, Z, mu)
            assert np.abs(B.hessian_values(Z, mu) - ho).max() < HESS_RTOL * max(1.0, np.abs(ho).max())
        B.close()


def test_hessian_cta_cap_option():
    """PB2_OPT_HESSIAN_CTAS only changes how many SMs the persistent 3-qubit Hessian kernel occupies, never the values;
    a negative count is refused."""
    p, Z, mu = C.trajectory(3, 400)
    B = make(p)
    h0 = B.hessian_values(Z, mu)
    for n in (1, 37, 125, 10000):
        B.set_option("hessian_ctas", n)
        assert np.array_equal(B.hessian_values(Z, mu), h0)
    B.set_option("hessian_ctas", 0)
    assert np.array_equal(B.hessian_values(Z, mu), h0)
    with pytest.raises(pb.PB2Error):
        B.set_option("hessian_ctas", -1)
    B.close()
