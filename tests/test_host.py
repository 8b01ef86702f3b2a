"""CPU tests of the host side: the C-ABI library loads, exports every declared symbol and refuses
to run without a GPU; generator factors; the sharding plumbing under gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import piccolo_b200 as pb
from oracle import configs as C
from oracle import isomorphisms as oiso
from oracle import knot as KN
from oracle import systems as S

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    from piccolo_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "piccolo_b200.h")).read()
    declared = set(re.findall(r"\b(pb2_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    L = ctypes.CDLL(pb.lib_path())
    for s in declared:
        assert hasattr(L, s), s
    assert pb.load_library().pb2_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.PB2Error) as e:
        pb.B200BilinearIntegrator("ket", np.zeros((4, 4)), [np.zeros((4, 4))], K=5, D=10,
                                  x_off=0, dt_off=4, u_off=6)
    assert e.value.code == 3
    # the other handles of the library behave the same: no device, no result
    Z = np.zeros((10, 5))
    Z[4] = 0.1
    traj = pb.NamedTrajectory(Z, {"ψ̃": range(0, 4), "Δt": range(4, 5), "t": range(5, 6), "u": range(6, 7),
                                  "du": range(7, 8), "ddu": range(8, 9)})
    with pytest.raises(pb.PB2Error) as e:
        pb.B200KnotLinearConstraints(traj)
    assert e.value.code == 3
    J = pb.KetInfidelityObjective(np.array([0, 1], complex), "ψ̃", traj) + pb.QuadraticRegularizer("u", traj, 1e-2)
    with pytest.raises(pb.PB2Error) as e:
        J.value(Z)                       # the handle is built on first use
    assert e.value.code == 3


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "piccolo.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f


def test_generators_match_oracle_restatement():
    rng = np.random.default_rng(3)
    n = 3
    H = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = H + H.conj().T
    Lop = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    assert np.array_equal(pb.G(H), oiso.G(H))
    assert np.array_equal(pb.iso(H), oiso.iso(H))
    assert np.allclose(pb.iso_D(Lop), oiso.iso_D(Lop))
    assert np.array_equal(pb.density_lift_matrix(n), oiso.density_lift_matrix(n))
    assert np.array_equal(pb.density_projection_matrix(n), oiso.density_projection_matrix(n))
    s = S.OpenQuantumSystem(H, [H * 0.5], [1.0], [Lop])
    G0, Gj = S.compact_generator_parts(s)
    G0p, Gjp = pb.OpenQuantumSystem(H, [H * 0.5], [1.0], [Lop]).G_parts()
    assert np.allclose(G0, G0p) and np.allclose(Gj[0], Gjp[0])
    with pytest.raises(ValueError):
        pb.QuantumSystem(np.array([[0, 1], [0, 0]]), [])


def test_knot_partition():
    per, rng_ = pb.knot_partition(999, 8)
    assert per == 125 and rng_[0] == (0, 125) and rng_[-1] == (875, 999)
    per, rng_ = pb.knot_partition(3, 8)
    assert per == 1 and rng_[2] == (2, 3) and rng_[3] == (3, 3)
    assert pb.knot_partition(0, 4)[0] == 0


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import piccolo_b200 as pb
from oracle import configs as C, knot as KN
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
p, Z, mu = C.trajectory(2, {K})

class CpuStandIn:                     # test-only injection; the product default is the CUDA handle
    def __init__(self, K_local, knot0):
        self.p = KN.make_problem(p.kind, p.G0, p.Gj, K_local)
    def residual_jacobian(self, Zl):
        Zl = np.asfortranarray(Zl)
        return KN.residual(self.p, Zl), KN.jacobian_values(self.p, Zl)

S = pb.ShardedBilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
        u_off=p.u_off, rank=rank, world=world, tensor_device="cpu", make_local=CpuStandIn)
S.residual_jacobian(Z)
d, v = S.unpack()
assert np.array_equal(d, KN.residual(p, Z)), "delta mismatch"
assert np.array_equal(v, KN.jacobian_values(p, Z)), "jac mismatch"
dist.barrier()
if rank == 0: print("SHARD_OK", S.per, S.ranges)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("K", [10, 4, 2])
def test_sharded_gather_world2_gloo(tmp_path, K):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, K=K))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + K), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [pr.communicate(timeout=240)[0].decode() for pr in procs]
    assert all(pr.returncode == 0 for pr in procs), outs
    assert "SHARD_OK" in outs[0]


# ---- objective host logic: complex goals -> real coefficient vectors of the C ABI's term --------------
def _term_value(term, z):
    """The formula the CUDA kernel evaluates for one term (include/piccolo_b200.h), in NumPy."""
    zz = z[term.rows]
    f = lambda a: 0.0 if a is None else float(a @ zz)
    F = term.scale * (f(term.a_re) ** 2 + f(term.a_im) ** 2 + (0.0 if term.a_sq is None else float(term.a_sq @ (zz * zz)))) \
        + f(term.a_lin)
    return term.Q[0] * (abs(1 - F) if term.flags & 1 else F)


def test_objective_terms_reproduce_the_reference_losses():
    """No GPU: the coefficient vectors the constructors hand to pb2_obj_create, evaluated with the
    documented real-form formula, equal the complex-arithmetic restatement of objectives.jl."""
    from oracle import objectives as OB
    rng = np.random.default_rng(17)
    cplx = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    # unitary (full operator and EmbeddedOperator subspace form), 3 levels
    N, m, K = 3, 1, 4
    Z = 0.5 * rng.standard_normal((2 * N * N + 2 + 3 * m, K))
    traj = pb.NamedTrajectory.smooth_pulse_layout(Z, 2 * N * N, m, "Ũ⃗")
    Ug = np.linalg.qr(cplx(N, N))[0]
    J = pb.UnitaryInfidelityObjective(Ug, "Ũ⃗", traj, Q=100.0)
    assert np.isclose(_term_value(J.terms[0], Z[:, -1]), OB.unitary_infidelity(Z[:2 * N * N, -1], Ug, 100.0)[0], rtol=1e-13)
    Us = np.linalg.qr(cplx(2, 2))[0]
    J = pb.UnitaryInfidelityObjective(Us, "Ũ⃗", traj, Q=30.0, subspace=[0, 2])
    ref = OB.unitary_infidelity(Z[:2 * N * N, -1], Us, 30.0, subspace=[0, 2])[0]
    assert np.isclose(_term_value(J.terms[0], Z[:, -1]), ref, rtol=1e-13)
    # kets: single, coherent (weighted / uniform)
    Zk = 0.5 * rng.standard_normal((2 * 4 + 2 + 3, K))
    tk = pb.NamedTrajectory.multi_state_layout(Zk, ["ψ̃1", "ψ̃2"], 4, 1)
    g1, g2 = cplx(2), cplx(2)
    g1, g2 = g1 / np.linalg.norm(g1), g2 / np.linalg.norm(g2)
    J = pb.KetInfidelityObjective(g2, "ψ̃2", tk, Q=7.0)
    assert np.isclose(_term_value(J.terms[0], Zk[:, -1]), OB.ket_infidelity(Zk[4:8, -1], g2, 7.0)[0], rtol=1e-13)
    for w in (None, [0.9, 0.1], [2.0, 2.0]):
        J = pb.CoherentKetInfidelityObjective([g1, g2], ["ψ̃1", "ψ̃2"], tk, Q=100.0, weights=w)
        ref = OB.coherent_ket_infidelity([Zk[0:4, -1], Zk[4:8, -1]], [g1, g2], 100.0, w)[0]
        assert np.isclose(_term_value(J.terms[0], Zk[:, -1]), ref, rtol=1e-13)
    # the reference's own expectation (objectives.jl:592): F = |0.9*1 + 0.1*1/2|^2
    p0, p1 = np.array([1, 0], complex), np.array([0, 1], complex)
    z = np.zeros(Zk.shape[0])
    z[0:4], z[4:8] = np.concatenate([p1.real, p1.imag]), 0.5 * np.concatenate([p0.real, p0.imag])
    J = pb.CoherentKetInfidelityObjective([p1, p0], ["ψ̃1", "ψ̃2"], tk, Q=100.0, weights=[0.9, 0.1])
    assert np.isclose(_term_value(J.terms[0], z), 100.0 * (1 - 0.9025), rtol=1e-13)
    with pytest.raises(ValueError):
        pb.CoherentKetInfidelityObjective([p1, p0], ["ψ̃1"], tk)
    # density matrices in the compact iso, leakage, composition
    n = 3
    Zd = 0.4 * rng.standard_normal((n * n + 2 + 3, K))
    td = pb.NamedTrajectory.smooth_pulse_layout(Zd, n * n, 1, "ρ⃗̃")
    A = cplx(n, n)
    rho_g = A @ A.conj().T
    rho_g /= np.trace(rho_g).real
    psi = cplx(n)
    psi /= np.linalg.norm(psi)
    J1 = pb.DensityMatrixInfidelityObjective("ρ⃗̃", rho_g, td, Q=50.0)
    J2 = pb.DensityMatrixPureStateInfidelityObjective("ρ⃗̃", psi, td, Q=5.0)
    assert np.isclose(_term_value(J1.terms[0], Zd[:, -1]), OB.density_infidelity(Zd[:n * n, -1], rho_g, 50.0)[0], rtol=1e-13)
    assert np.isclose(_term_value(J2.terms[0], Zd[:, -1]), OB.density_pure_state_infidelity(Zd[:n * n, -1], psi, 5.0)[0], rtol=1e-13)
    JL = pb.LeakageObjective([1, 4], "ρ⃗̃", td, times=[0, 2], Qs=[1.0, 3.0])
    t = JL.terms[0]
    got = sum(q * t.scale * 0 + q * float(t.a_sq @ (Zd[t.rows, k] ** 2)) for k, q in zip(t.times, t.Q))
    assert np.isclose(got, OB.leakage(Zd[:n * n][:, [0, 2]], [1, 4], [1.0, 3.0])[0], rtol=1e-13)
    Jsum = J1 + J2 + JL + pb.QuadraticRegularizer("u", td, 0.1)
    assert len(Jsum.terms) == 3 and len(Jsum.regs) == 1 and Jsum._h is None     # nothing touched the library yet
    with pytest.raises(ValueError):
        J1 + pb.KetInfidelityObjective(g1, "ψ̃1", tk)                           # different trajectory layouts


def test_c_program_builds_and_fails_loudly_without_a_gpu():
    """tests/c/test_capi.c links against libpiccolo_b200.so through the public header alone; on a box without a
    CUDA device pb2_create refuses (PB2_ENODEVICE) and the program reports it -- there is no CPU path to fall into."""
    import os
    import subprocess
    import torch
    from oracle import cport as CP
    cdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
    CP.lib()
    subprocess.check_call(["make", "-C", cdir, "-B", "test_capi"], stdout=subprocess.DEVNULL)
    if torch.cuda.is_available():
        return                                   # the GPU suite runs the program (tests/test_gpu_parity.py)
    r = subprocess.run([os.path.join(cdir, "test_capi")], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stdout, r.stdout + r.stderr
