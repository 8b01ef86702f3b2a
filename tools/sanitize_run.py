import numpy as np, sys
sys.path.insert(0, "/root/repo")
import piccolo_b200 as pb
from oracle import configs as C
from oracle import knot as KN
for cfg, K in ((3, 700), (2, 60), (4, 50)):
    p, Z, mu = C.trajectory(cfg, K)
    B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
    d, v = B.residual_jacobian(Z)
    print(cfg, B.algorithm, float(np.abs(d - KN.residual(p, Z)).max()) if K < 100 else "-")
    B.close()
