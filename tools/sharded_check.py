"""Multi-GPU check of ShardedBilinearIntegrator (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_check.py
Every rank must end up with the whole trajectory's [delta | values], bitwise equal to what one GPU
computes alone, through the fused NVLink exchange and through the NCCL all-gather of the records."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import piccolo_b200 as pb
from oracle import configs as C

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
ok = True
for K in (102, 1000):                       # 101 knot evals: ragged over 2, 4 and 8 ranks
    p, Z, _ = C.trajectory(3, K)
    B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                  u_off=p.u_off, device=local)
    d0, v0 = B.residual_jacobian(Z)
    B.close()
    for fused in ("auto", False):
        S = pb.ShardedBilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                         u_off=p.u_off, rank=rank, world=world, device=local, fused=fused)
        for rep in range(3):                # repeated calls reuse the record buffer: the entry barrier orders them
            Zr = np.asfortranarray(Z + rep * 1e-3)
            S.residual_jacobian(Zr)
            torch.cuda.synchronize()
        d, v = S.unpack()
        Bf = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                       u_off=p.u_off, device=local)
        d1, v1 = Bf.residual_jacobian(np.asfortranarray(Z + 2e-3))
        Bf.close()
        good = np.array_equal(d, d1) and np.array_equal(v, v1)
        print(f"rank {rank} K={K} fused={S.fused} (asked {fused}): {'ok' if good else 'MISMATCH'}", flush=True)
        ok = ok and good and (S.fused or fused is False or world == 1)
        S.local.close()
        del S
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0 if flag.item() == 1 else 1)
