#!/usr/bin/env python
"""Debug: clock64 stamps of every CTA of the one-CTA-per-SM single-round kernel (knot_u8p; run with PB2_U8Q=0) over a CUDA graph of
back-to-back launches (the benchmarked configuration), keyed by %smid so that the gap between one launch's
CTA exit and the next launch's CTA entry ON THE SAME SM can be read off (clock64 is a per-SM free-running
counter).  Needs the debug build: make -C piccolo.jl_b200 libpiccolo_b200_trace.so"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PB2_LIB", "libpiccolo_b200_trace.so")
import torch
import piccolo_b200 as pb
from oracle import configs as C

NL = int(os.environ.get("TRACE_LAUNCHES", "12"))
p, Z, _ = C.trajectory(3)
B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off)
if os.environ.get("TRACE_EARLY_Z", "1") == "1" and hasattr(B, "set_option"):
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
assert lib.pb2_debug_trace3(B._h, out.ctypes.data) == 0      # also resets the launch counter
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st):
    cs = torch.cuda.current_stream().cuda_stream
    for i in range(NL):
        B.residual_jacobian_device(dZ[i % nsets], dd[i % nsets], dv[i % nsets], cs)
torch.cuda.synchronize()
assert lib.pb2_debug_trace3(B._h, out.ctypes.data) == 0      # capture only records: nothing ran; reset again
g.replay()
torch.cuda.synchronize()
# the captured launches carry trace ids 0..NL-1 (ids are baked at capture time)
assert lib.pb2_debug_trace3(B._h, out.ctypes.data) == 0
T = out[:64 * 148 * 16 * 8].reshape(64, 148, 16, 8)[:NL]
S = out[64 * 148 * 16 * 8:].reshape(64, 148, 16, 32)[:NL]
names = ["entry", "armed", "landed", "prepared", "E_done", "horner_end", "slab_issue", "end"]
smid = T[:, :, 12, 6]
END = 7
print("launch, block -> smid constant across launches:", bool((smid == smid[0]).all()))
# per launch, per block: entry (min over warps), end (max over warps)
ent = np.where(T[..., 0] > 0, T[..., 0], np.iinfo(np.int64).max).min(axis=2)
end = T[..., END].max(axis=2)
dur = end - ent
print("CTA duration (cycles) per launch: median / max over blocks")
for l in range(NL):
    print(f"  launch {l:2d}: {int(np.median(dur[l]))} / {int(dur[l].max())}")
# gap on the same SM between consecutive launches
for l in range(1, NL):
    gaps = []
    for b in range(148):
        prev = np.where(smid[l - 1] == smid[l, b])[0]
        if prev.size == 1:
            gaps.append(ent[l, b] - end[l - 1, prev[0]])
    gaps = np.array(gaps)
    print(f"  launch {l:2d}: same-SM gap exit->entry: median {int(np.median(gaps))} min {int(gaps.min())} max {int(gaps.max())} (n={gaps.size})")
# phase breakdown for blocks with 7 knots (blocks < nk - 6*148) in a late launch
l = NL - 2
for b in (0, 1, 50, 110, 111, 147):
    rows = []
    for w in range(16):
        t = T[l, b, w]
        if t[0] > 0:
            base = ent[l, b]
            rows.append(f"w{w:2d}: " + " ".join(f"{n}={int(t[i] - base) if t[i] > 0 else -1:6d}" for i, n in enumerate(names)))
    print(f"launch {l} block {b} smid {int(smid[l, b])} duration {int(dur[l, b])}")
    print("\n".join(rows))

if True:
    print("per-step stamps (start of step, after the exchange barrier), cycles since CTA entry; launch", l)
    for b in (0, 50):
        for w in range(16):
            st = S[l, b, w]
            if st.max() > 0:
                base = ent[l, b]
                print(f"block {b} w{w:2d}: " + " ".join(f"{int(v - base) if v > 0 else -1}" for v in st[:24]) +
                      f" | prep " + " ".join(f"{int(v - base) if v > 0 else -1}" for v in st[24:32]))
