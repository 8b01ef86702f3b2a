"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import piccolo_b200 as pb
from oracle import configs as C
from oracle import knot as KN

for cfg, K in ((3, 700), (2, 60), (4, 50), (1, 20)):
    p, Z, mu = C.trajectory(cfg, K)
    B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
    d, v = B.residual_jacobian(Z)                 # host path (compact D2H + host un-pack for C3)
    h = B.hessian_values(Z, mu)
    d2 = np.empty(B.dim)
    B.evaluate_(d2, Z)
    dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
    dd = torch.zeros(B.dim, dtype=torch.float64, device="cuda")
    dv = torch.zeros(B.nnz_jac, dtype=torch.float64, device="cuda")
    B.residual_jacobian_device(dZ, dd, dv, None)  # device path (canonical arrays, bulk stores)
    if B.compact_stride:
        comp = torch.zeros(B.compact_stride * (p.K - 1), dtype=torch.float64, device="cuda")
        B.residual_jacobian_compact_device(dZ, comp, None)
        B.expand_compact_device(comp, p.K - 1, dd, dv, None)
    torch.cuda.synchronize()
    ok = np.array_equal(dd.cpu().numpy(), d) and np.array_equal(dv.cpu().numpy(), v)
    err = float(np.abs(d - KN.residual(p, Z)).max()) if K < 100 else float("nan")
    print(cfg, B.algorithm, "device==host:", ok, "residual err vs oracle:", err)
    B.close()
    traj = pb.NamedTrajectory.smooth_pulse_layout(Z, p.n_x, p.m, "x")
    L = pb.B200KnotLinearConstraints(traj)
    L.residual_jacobian(Z)
    L.hessian_values(np.ones(L.dim))
    L.close()
    if p.kind == "unitary":
        n = int(round((p.n_x // 2) ** 0.5))
        J = (pb.UnitaryInfidelityObjective(np.eye(n), "x", traj) + pb.QuadraticRegularizer("u", traj, 1e-2, dt_power=1)
             + pb.LeakageObjective([0, 1], "x", traj, times=[0, K // 2, K - 1]))
        J.value_gradient(Z)
        J.value(Z)
        J.hessian_values(Z, 0.7)
        J.close()
    # round 2: equal-timestep rows, rollout, dense d/dx_k block, an ensemble in one launch, time-dependent drives,
    # the older kernels behind their switches
    L = pb.B200KnotLinearConstraints(traj, timesteps_all_equal=True)
    L.residual_jacobian(Z)
    L.close()
    B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off,
                                  dense_blocks=True)
    B.residual_jacobian(Z)
    B.rollout(Z)
    B.close()
    if cfg in (1, 2, 4):
        import dataclasses
        n, n_x = 3, p.n_x
        Zs = np.zeros((n * n_x + p.D - n_x, p.K), order="F")
        Zs[n * n_x:] = Z[n_x:]
        for i in range(n):
            Zs[i * n_x:(i + 1) * n_x] = Z[:n_x]
        batch = pb.B200IntegratorBatch(p.kind, [(p.G0 * (1 + 0.1 * i), list(p.Gj)) for i in range(n)], K=p.K, D=Zs.shape[0],
                                       x_offs=[i * n_x for i in range(n)], dt_off=p.dt_off + (n - 1) * n_x,
                                       u_off=p.u_off + (n - 1) * n_x)
        batch.residual_jacobian(Zs)
        batch.hessian_values(Zs, np.ones((n, batch.dim)))
        batch.close()
    if p.kind != "density":
        mods = [(lambda t, w=0.3 * (j + 1): np.cos(w * t)) for j in range(p.m)]
        Bt = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                       u_off=p.u_off, t_off=p.dt_off + 1, modulations=mods)
        Bt.residual_jacobian(Z)
        Bt.close()
for env in ({"PB2_DMMAQ": "0"}, {"PB2_U8Q_NS": "2"}, {"PB2_U8Q": "0"}, {"PB2_NO_DMMAH": "1"}, {"PB2_HESS_DMMAH": "1"}):
    os.environ.update(env)
    for cfg, K in ((3, 150), (2, 40), (4, 30)):
        p, Z, mu = C.trajectory(cfg, K)
        B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
        B.residual_jacobian(Z)
        B.hessian_values(Z, mu)
        print(env, cfg, B.algorithm, B.hessian_algorithm)
        B.close()
    for k in env:
        del os.environ[k]
