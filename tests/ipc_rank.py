"""One rank of a two-process sharded evaluation that uses NOTHING but libpiccolo_b200's C ABI for the device
side (no torch, no CUDA binding): pb2_device_alloc / pb2_ipc_export / pb2_ipc_open set up the gather buffers,
pb2_residual_jacobian_exchange_async fills every rank's buffer.  This is the sequence a Julia host (one process
per GPU) would follow; tests/test_gpu_parity.py::test_two_process_exchange_through_the_c_abi drives it.

    rank 1 (child):  python tests/ipc_rank.py K device    handles travel as hex lines over stdin / stdout
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import capi               # noqa: E402
from oracle import configs as C             # noqa: E402


class Rank:
    def __init__(self, rank, K, device=0, world=2):
        from piccolo_b200 import B200BilinearIntegrator, knot_partition
        self.lib = capi.load_library()
        self.rank, self.world, self.device = rank, world, device
        p, Z, _ = C.trajectory(3, K)
        self.p, self.Z = p, Z
        self.per, self.ranges = knot_partition(K - 1, world)
        k0, k1 = self.ranges[rank]
        self.n_local = k1 - k0
        self.B = B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=self.n_local + 1, D=p.D, x_off=p.x_off,
                                        dt_off=p.dt_off, u_off=p.u_off, device=device, knot0=k0)
        self.cs = self.B.compact_stride
        self.buf = self.alloc(8 * self.cs * self.per * world)
        slab = np.ascontiguousarray(Z[:, k0:k0 + self.n_local + 1].T).reshape(-1)
        self.dZ = self.alloc(slab.nbytes)
        capi.check(self.lib.pb2_device_copy(self.dZ, slab.ctypes.data, slab.nbytes))
        self.peers = {}

    def alloc(self, nbytes):
        ptr = ctypes.c_void_p()
        capi.check(self.lib.pb2_device_alloc(ctypes.byref(ptr), int(nbytes), self.device))
        return ptr

    def export(self):
        h = ctypes.create_string_buffer(64)
        capi.check(self.lib.pb2_ipc_export(self.buf, h))
        return h.raw.hex()

    def open(self, r, hexhandle):
        ptr = ctypes.c_void_p()
        capi.check(self.lib.pb2_ipc_open(bytes.fromhex(hexhandle), self.device, ctypes.byref(ptr)))
        self.peers[r] = ptr

    def run(self):
        bufs = [self.buf.value if r == self.rank else self.peers[r].value for r in range(self.world)]
        self.B.residual_jacobian_exchange_device(self.dZ.value, self.rank, bufs, self.rank * self.cs * self.per,
                                                 self.lib.pb2_stream(self.B._h))
        self.B.sync()

    def gathered(self):
        """Expand every rank's records from THIS rank's buffer into canonical (delta, vals)."""
        n_x, nnz = self.p.n_x, self.B.nnz_jac // self.n_local
        dd, dv = self.alloc(8 * self.per * n_x), self.alloc(8 * self.per * nnz)
        deltas, vals = [], []
        for r, (a, c) in enumerate(self.ranges):
            src = self.buf.value + 8 * r * self.cs * self.per
            self.B.expand_compact_device(src, self.per, dd.value, dv.value, self.lib.pb2_stream(self.B._h))
            self.B.sync()
            d, v = np.empty(self.per * n_x), np.empty(self.per * nnz)
            capi.check(self.lib.pb2_device_copy(d.ctypes.data, dd, d.nbytes))
            capi.check(self.lib.pb2_device_copy(v.ctypes.data, dv, v.nbytes))
            deltas.append(d[:(c - a) * n_x])
            vals.append(v[:(c - a) * nnz])
        for q in (dd, dv):
            self.lib.pb2_device_free(q)
        return np.concatenate(deltas), np.concatenate(vals)

    def close(self):
        for q in self.peers.values():
            self.lib.pb2_ipc_close(q)
        self.B.close()
        self.lib.pb2_device_free(self.buf)
        self.lib.pb2_device_free(self.dZ)


if __name__ == "__main__":
    R = Rank(1, int(sys.argv[1]), int(sys.argv[2]))
    R.open(0, sys.stdin.readline().strip())
    print(R.export(), flush=True)
    assert sys.stdin.readline().strip() == "go"
    R.run()
    print("done", flush=True)
    assert sys.stdin.readline().strip() == "check"
    d, v = R.gathered()                 # rank 1's own buffer must hold the whole trajectory too
    print(f"{float(np.abs(d).sum()):.17g} {float(np.abs(v).sum()):.17g}", flush=True)
    sys.stdin.readline()
    R.close()
