#!/usr/bin/env python
"""Debug: per-phase clock64 stamps of block 0 of the tensor-core knot kernel (needs a library built
with -DPB2_TRACE:  make -C piccolo.jl_b200 TRACE=1 libpiccolo_b200_trace.so)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PB2_LIB", "libpiccolo_b200_trace.so")
import piccolo_b200 as pb
from oracle import configs as C

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p, Z, _ = C.trajectory(cfg, int(sys.argv[2]) if len(sys.argv) > 2 else None)
B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
mode = os.environ.get("TRACE_MODE", "resjac")
for _ in range(3):
    if mode == "hess":
        B.hessian_values(Z, _)  if False else B.hessian_values(Z, np.random.default_rng(0).standard_normal(B.dim))
    elif os.environ.get("TRACE_DEVICE"):
        # the benchmarked configuration: device-resident canonical outputs (not the compact host path)
        import torch
        dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda()
        dd = torch.empty(B.dim, dtype=torch.float64, device="cuda")
        dv = torch.empty(B.nnz_jac, dtype=torch.float64, device="cuda")
        B.residual_jacobian_device(dZ, dd, dv, None)
        torch.cuda.synchronize()
    else:
        B.residual_jacobian(Z)
out = np.zeros(8 * 16 * 8, dtype=np.int64)
lib = pb.load_library()
lib.pb2_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.pb2_debug_trace(B._h, out.ctypes.data) == 0
T = out.reshape(8, 16, 8)
t0 = T[T > 0].min() if (T > 0).any() else 0
names = ["start", "slab", "prebar", "postbar", "loaded", "horner", "staged", "postbar2"]
for it in range(8):
    for w in range(16):
        if T[it, w].max() > 0:
            print(f"knot {it} warp {w:2d}: " + " ".join(f"{n}={int(v - t0):6d}" for n, v in zip(names, T[it, w])))

if B.algorithm == "dmma" and p.b == 16 and p.n_b == 8 and not os.environ.get("PB2_NO_U8"):
    out2 = np.zeros(16 * 20 * 4, dtype=np.int64)
    lib.pb2_debug_trace2.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert lib.pb2_debug_trace2(B._h, out2.ctypes.data) == 0
    T2 = out2[:16 * 4 * 8].reshape(16, 4, 8)
    t0 = T2[T2 > 0].min()
    print("u8 kernel, block 0: warp, knot, stamps (producer: wait-z, got-z, ready | staged-wait, staged, stored, freed;"
          " compute: start, ready, horner-start, horner-end, pre-free, free, staged)")
    for w in range(16):
        for it in range(4):
            if T2[w, it].max() > 0:
                print(f"warp {w:2d} knot {it}: " + " ".join(f"{int(v - t0) if v else -1:6d}" for v in T2[w, it, :8]))
    sys.exit(0)
out2 = np.zeros(16 * 20 * 4, dtype=np.int64)
lib.pb2_debug_trace2.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert lib.pb2_debug_trace2(B._h, out2.ctypes.data) == 0
T2 = out2.reshape(16, 20, 4)
t0 = T2[T2 > 0].min()
print("per-step stamps of the last knot of block 0 (entry, before barrier, end):")
for w in range(16):
    for st in range(19, -1, -1):
        if T2[w, st].max() > 0:
            print(f"warp {w:2d} kq {st:2d}: " + " ".join(f"{int(v - t0) if v else -1:6d}" for v in T2[w, st, :3]))
