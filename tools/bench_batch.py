"""Ensemble throughput: n members in one launch (B200IntegratorBatch) against one launch per member.

    python tools/bench_batch.py [--members 16] [--configs 2,4] [--iters 200]

Device-resident inputs, CUDA events on the launching stream.  These systems are tiny (the whole working set
sits in L2), which is the point: the per-member path is launch-latency bound.  Both arms are also replayed from
CUDA graphs, so the comparison is not an artefact of CPU launch overhead.  Prints one JSON line per configuration."""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piccolo_b200 as pb                     # noqa: E402
from oracle import configs as C               # noqa: E402


def ensemble(cfg, K, n):
    p0, Z0, _ = C.trajectory(cfg, K)
    rng = np.random.default_rng(cfg)
    n_x, rest = p0.n_x, p0.D - p0.n_x
    D = n * n_x + rest
    Z = np.zeros((D, K), order="F")
    Z[n * n_x:, :] = Z0[n_x:, :]
    probs = []
    for i in range(n):
        Z[i * n_x:(i + 1) * n_x, :] = Z0[:n_x, :] + 1e-3 * rng.standard_normal((n_x, K))
        probs.append(dataclasses.replace(p0, G0=p0.G0 * (1.0 + 0.03 * i), D=D, x_off=i * n_x,
                                         dt_off=n * n_x + (p0.dt_off - n_x), u_off=n * n_x + (p0.u_off - n_x)))
    return probs, Z


def timed(fn, iters, stream):
    with torch.cuda.stream(stream):
        for _ in range(10):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        a.record(stream)
        for _ in range(iters):
            fn()
        b.record(stream)
        stream.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def graphed(fn, stream, reps=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(reps):
                fn()
    return g, reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=16)
    ap.add_argument("--configs", default="2,4")
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    st = torch.cuda.Stream()
    for cfg in [int(c) for c in a.configs.split(",")]:
        K = {1: 50, 2: 200, 4: 500, 6: 64}.get(cfg, 200)
        n = a.members
        probs, Z = ensemble(cfg, K, n)
        p0 = probs[0]
        batch = pb.B200IntegratorBatch(p0.kind, [(p.G0, list(p.Gj)) for p in probs], K=K, D=p0.D,
                                       x_offs=[p.x_off for p in probs], dt_off=p0.dt_off, u_off=p0.u_off)
        singles = [pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=K, D=p.D, x_off=p.x_off, dt_off=p.dt_off,
                                             u_off=p.u_off) for p in probs]
        dZ = torch.from_numpy(np.ascontiguousarray(Z.reshape(-1, order="F"))).cuda()
        dd = torch.empty(n, batch.dim, dtype=torch.float64, device="cuda")
        dv = torch.empty(n, batch.nnz_jac, dtype=torch.float64, device="cuda")
        dmu = torch.randn(n, batch.dim, dtype=torch.float64, device="cuda")
        dh = torch.empty(n, batch.nnz_hess, dtype=torch.float64, device="cuda")
        s = st.cuda_stream

        def one():
            batch.residual_jacobian_device(dZ.data_ptr(), dd.data_ptr(), dv.data_ptr(), s)

        def per_member():
            for i, B in enumerate(singles):
                B.residual_jacobian_device(dZ.data_ptr(), dd[i].data_ptr(), dv[i].data_ptr(), s)

        def one_h():
            batch.hessian_device(dZ.data_ptr(), dmu.data_ptr(), dh.data_ptr(), s)

        def per_member_h():
            for i, B in enumerate(singles):
                B.hessian_device(dZ.data_ptr(), dmu[i].data_ptr(), dh[i].data_ptr(), s)

        out = {"config": f"C{cfg}", "K": K, "members": n, "fused": batch.fused, "algorithm": singles[0].algorithm}
        for name, f1, fn_ in (("resjac", one, per_member), ("hess", one_h, per_member_h)):
            t1, tn = timed(f1, a.iters, st), timed(fn_, max(a.iters // 4, 10), st)
            g1, r1 = graphed(f1, st)
            gn, rn = graphed(fn_, st)
            tg1 = timed(g1.replay, 20, st) / r1
            tgn = timed(gn.replay, 20, st) / rn
            evals = n * (K - 1)
            out[name] = {"batch_us": round(t1, 2), "per_member_us": round(tn, 2), "speedup": round(tn / t1, 2),
                         "graph_batch_us": round(tg1, 2), "graph_per_member_us": round(tgn, 2),
                         "graph_speedup": round(tgn / tg1, 2),
                         "batch_knot_evals_per_s": round(evals / (tg1 * 1e-6), 0),
                         "per_member_knot_evals_per_s": round(evals / (tgn * 1e-6), 0)}
        print(json.dumps(out), flush=True)
        batch.close()
        for B in singles:
            B.close()


if __name__ == "__main__":
    main()
