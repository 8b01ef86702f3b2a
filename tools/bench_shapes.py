"""Device-resident callback times of arbitrary shapes, tensor-core path against the jet kernels.
    python tools/bench_shapes.py            # two-transmon qutrit sizes (d = 9: b = 18) by default"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piccolo_b200 as pb                     # noqa: E402
from oracle import knot as KN                 # noqa: E402


def iso_gen(H):
    return np.block([[H.imag, H.real], [-H.real, H.imag]])


def problem(kind, d, m, K, rng):
    H0 = np.diag(rng.standard_normal(d)).astype(complex)
    Gj = []
    for _ in range(m):
        H = np.zeros((d, d), dtype=complex)
        for a in range(d - 1):                       # ladder-type drive: <= 2 entries per row
            H[a, a + 1] = rng.standard_normal() + 1j * rng.standard_normal()
            H[a + 1, a] = np.conj(H[a, a + 1])
        Gj.append(iso_gen(H))
    p = KN.make_problem(kind, iso_gen(H0), Gj, K)
    Z = np.asfortranarray(0.3 * rng.standard_normal((p.D, K)))
    Z[p.dt_off] = 0.05 + 0.05 * rng.random(K)
    return p, Z


def timed(fn, reps=20):
    """microseconds per call, 20 calls captured in one CUDA graph (no Python launch overhead in the figure)"""
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn(st.cuda_stream)
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn(torch.cuda.current_stream().cuda_stream)
        g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.synchronize()
        a.record(st)
        g.replay()
        b.record(st)
        st.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


rng = np.random.default_rng(3)
for kind, d, m, K in (("ket", 9, 4, 1000), ("unitary", 9, 2, 200), ("ket", 12, 2, 1000)):
    p, Z = problem(kind, d, m, K, rng)
    dZ = torch.from_numpy(np.ascontiguousarray(Z.reshape(-1, order="F"))).cuda()
    row = [f"{kind} d={d} (b={p.b}, n_b={p.n_b}) m={m} K={K}:"]
    for alg in ("auto", "generic"):
        B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                      u_off=p.u_off, algorithm=alg)
        dd = torch.empty(B.dim, dtype=torch.float64, device="cuda")
        dv = torch.empty(B.nnz_jac, dtype=torch.float64, device="cuda")
        dmu = torch.randn(B.dim, dtype=torch.float64, device="cuda")
        dh = torch.empty(B.nnz_hess, dtype=torch.float64, device="cuda")
        t1 = timed(lambda s: B.residual_jacobian_device(dZ, dd, dv, s))
        t2 = timed(lambda s: B.hessian_device(dZ, dmu, dh, s))
        row.append(f"{alg}: {B.algorithm}/{B.hessian_algorithm} resjac {t1:.1f} us, hessian {t2:.1f} us;")
        B.close()
    print(" ".join(row), flush=True)
