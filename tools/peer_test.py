"""2-GPU diagnostic for the fused exchange: peer mapping via CUDA IPC, then the kernel writing records
into (a) two local buffers, (b) local + peer."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piccolo_b200 as pb
from oracle import configs as C

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
lib = pb.load_library()
for r in range(world):
    if r != rank:
        rc = lib.pb2_enable_peer_access(local, r)
        print(rank, "enable peer", r, rc, lib.pb2_last_error() if rc else "", flush=True)
p, Z, _ = C.trajectory(3, 200)
n = p.K - 1
B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off, u_off=p.u_off, device=local)
cs = B.compact_stride
import torch.distributed._symmetric_memory as symm_mem
buf = symm_mem.empty(cs * n * world, dtype=torch.float64, device=dev)
buf.zero_()
buf2 = torch.zeros(cs * n * world, dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
ptrs = [int(x) for x in hdl.buffer_ptrs]
print(rank, "symmetric memory ptrs", [hex(x) for x in ptrs], "mine", hex(buf.data_ptr()), flush=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
dZ = torch.from_numpy(Z.reshape(-1, order="F").copy()).to(dev)
ref = torch.zeros(cs * n, dtype=torch.float64, device=dev)
B.residual_jacobian_compact_device(dZ, ref, None)
torch.cuda.synchronize()
# (2) both destinations local
B.residual_jacobian_exchange_device(dZ, 0, [buf2.data_ptr(), buf2.data_ptr() + 8 * cs * n], 0, None)
torch.cuda.synchronize()
print(rank, "local x2:", torch.equal(buf2[:cs * n], ref), torch.equal(buf2[cs * n:2 * cs * n], ref), flush=True)
dist.barrier()
# (3) local + peer
B.residual_jacobian_exchange_device(dZ, rank, ptrs, rank * cs * n, None)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ok = all(torch.equal(buf[r * cs * n:(r + 1) * cs * n][:-1], ref[:-1]) for r in range(world))
print(rank, "local + peer:", ok, flush=True)
os._exit(0)
