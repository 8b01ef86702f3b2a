#!/usr/bin/env python
"""Debug: per-CTA clock64 stamps of the small-CTA residual+Jacobian kernel (knot_u8q) over a CUDA graph of
back-to-back launches, grouped by SM: which CTAs share an SM, when each enters / starts its products / ends.
Needs the debug build: make -C piccolo.jl_b200 libpiccolo_b200_trace.so"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PB2_LIB", "libpiccolo_b200_trace.so")
import torch
import piccolo_b200 as pb
from oracle import configs as C

NL = int(os.environ.get("TRACE_LAUNCHES", "12"))
p, Z, _ = C.trajectory(int(os.environ.get("TRACE_CONFIG", "3")))
B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
B.set_option("early_z", 1)
nsets = 12
dZ = [torch.from_numpy(Z.reshape(-1, order="F").copy()).cuda() for _ in range(nsets)]
dd = [torch.empty(B.dim, dtype=torch.float64, device="cuda") for _ in range(nsets)]
dv = [torch.empty(B.nnz_jac, dtype=torch.float64, device="cuda") for _ in range(nsets)]
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for i in range(3):
    B.residual_jacobian_device(dZ[i], dd[i], dv[i], st.cuda_stream)
torch.cuda.synchronize()
lib = pb.load_library()
lib.pb2_debug_trace3.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
out = np.zeros(64 * 148 * 16 * 40, dtype=np.int64)
assert lib.pb2_debug_trace3(B._h, out.ctypes.data) == 0
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st):
    cs = torch.cuda.current_stream().cuda_stream
    for i in range(NL):
        B.residual_jacobian_device(dZ[i % nsets], dd[i % nsets], dv[i % nsets], cs)
torch.cuda.synchronize()
assert lib.pb2_debug_trace3(B._h, out.ctypes.data) == 0
g.replay()
torch.cuda.synchronize()
assert lib.pb2_debug_trace3(B._h, out.ctypes.data) == 0
NC = 608
T = out[:16 * NC * 8 * 8].reshape(16, NC, 8, 8)[:NL]
smid = T[:, :, 1, 6]
valid = T[:, :, 0, 0] > 0
big = np.iinfo(np.int64).max
ent = np.where(T[..., 0] > 0, T[..., 0], big).min(axis=2)
prep = T[..., 3].max(axis=2)
hend = T[..., 5].max(axis=2)
end = T[..., 7].max(axis=2)
for sm in (0, 1, 77):
    rows = []
    for l in range(NL):
        for c in range(NC):
            if valid[l, c] and smid[l, c] == sm:
                rows.append((ent[l, c], l, c, prep[l, c], hend[l, c], end[l, c]))
    rows.sort()
    t0 = rows[0][0]
    print(f"SM {sm}: CTAs in entry order: (launch, cta) entry | products start | products end | exit   [cycles]")
    for e, l, c, pr, he, en in rows[: 6 * 8]:
        print(f"  L{l:2d} c{c:3d}: {e - t0:7d} | {pr - t0:7d} | {he - t0:7d} | {en - t0:7d}   dur {en - e}")
# per-launch span over the GPU is not comparable across SMs (clock64 is per SM); report CTA durations instead
armed = T[..., 1].max(axis=2)
landed = T[..., 2].max(axis=2)
dur = np.where(valid, end - ent, 0)
for l in range(NL):
    v = dur[l][valid[l]]
    print(f"launch {l}: {valid[l].sum()} CTAs, duration median {int(np.median(v))} max {int(v.max())}; "
          f"entry->armed {int(np.median((armed - ent)[l][valid[l]]))} ->landed {int(np.median((landed - ent)[l][valid[l]]))} "
          f"->products {int(np.median((prep - ent)[l][valid[l]]))}; products median {int(np.median((hend - prep)[l][valid[l]]))}; "
          f"tail {int(np.median((end - hend)[l][valid[l]]))}")
