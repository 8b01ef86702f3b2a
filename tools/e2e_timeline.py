"""Host-pointer callback (pb2_residual_jacobian with pinned caller buffers) on C3: median call time and, with
PB2_E2E_TIMELINE=1, the library's own host-clock stamps (enqueue done | each record chunk landed | replication done).
    PB2_E2E_TIMELINE=1 [PB2_E2E_H2D=0|1|2] [PB2_D2H_CHUNKS=n] [PB2_E2E_PIPE=0] python tools/e2e_timeline.py [K]"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piccolo_b200 as pb                     # noqa: E402
from oracle import configs as C               # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
lib = pb.load_library()
p, Z, _ = C.trajectory(3, K)
B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)


def pinned(n):
    ptr = ctypes.c_void_p()
    assert lib.pb2_host_alloc(ctypes.byref(ptr), 8 * n) == 0
    return ptr, np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,))


pz, Zp = pinned(Z.size)
pd, dp = pinned(B.dim)
pv, vp = pinned(B.nnz_jac)
Zp[:] = Z.reshape(-1, order="F")
tl = os.environ.pop("PB2_E2E_TIMELINE", None)      # the library reads it on the first call: keep warm-up quiet
ts = []
for i in range(60):
    if i == 57 and tl:
        os.environ["PB2_E2E_TIMELINE"] = "1"
    t0 = time.perf_counter()
    assert lib.pb2_residual_jacobian(B._h, pz, pd, pv, 0) == 0
    ts.append(time.perf_counter() - t0)
ts = np.array(ts[10:]) * 1e6
print(f"K={K} mode H2D={os.environ.get('PB2_E2E_H2D', '0')} chunks={os.environ.get('PB2_D2H_CHUNKS', '8')} "
      f"pipe={os.environ.get('PB2_E2E_PIPE', '1')}: median {np.median(ts):.1f} us  min {ts.min():.1f}  "
      f"-> {(K - 1) / np.median(ts) * 1e6:.3g} evals/s", flush=True)
