"""Smallest exercise of the warp-specialised kernels for `compute-sanitizer --tool racecheck`."""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import piccolo_b200 as pb
from oracle import configs as C
p, Z, mu = C.trajectory(3, 75)
B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
d, v = B.residual_jacobian(Z)
h = B.hessian_values(Z, mu)
print("done", d.shape, h.shape)
B.close()
