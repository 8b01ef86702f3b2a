#!/usr/bin/env python
"""Benchmark of the knot path: knot-constraint + Jacobian evals / second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3]

A "step" is one Ipopt-callback's worth of work: the fused residual + Jacobian evaluation of
every knot of one trajectory (BASELINE.json config C3: 3-transmon unitary, d=8 (iso dim 128),
4 drives, K=1000 knots -> 999 knot evals per step per GPU).  N > 1: weak scaling, every rank owns
999 knot evals of a (999 N + 1)-knot trajectory and each step ends with ONE all-gather of the
[delta | Jacobian values] shards (NCCL).  Prints one JSON line on rank 0.

`--impl reference`: the reference is 100% Julia and no Julia toolchain exists here, so this arm
times the C++ port of the reference's algorithm (oracle/c/knot_ref.cpp) on all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "knot_constraint_jacobian_evals_per_sec"
UNIT = "evals/s"
L2_BYTES = 126 * 2 ** 20


def algorithmic_bytes_per_eval(p):
    """8 * [(n_x + m + 1) read + (n_x + n_b b^2 + n_x m + n_x) written]
    (the constant identity block is excluded)."""
    return 8 * ((p.n_x + p.m + 1) + (p.n_x + p.n_b * p.b * p.b + p.n_x * p.m + p.n_x))


def fp64_tensor_work(p, Z):
    """FLOPs the tensor-core kernels issue for one callback: per knot (48 M - 24) DMMA.8x8x4
    (E, X tiles M steps; 4 jet tiles M - 1 steps; one more product for d/d dt) of 512 FLOP each,
    with the per-knot Taylor degree M the kernel derives from |dt| (||G_0||_1 + sum |u_j| ||G_j||_1).
    Only meaningful for the 3-qubit unitary shape (b = 16, n_b = 8, m = 4)."""
    import math
    if not (p.b == 16 and p.n_b == 8 and p.m == 4):
        return None

    def theta(q, tol=2.0 ** -53):
        lo, hi = 0.0, q + 1.0
        for _ in range(200):
            th = 0.5 * (lo + hi)
            lg = (q + 1) * math.log(th) - math.lgamma(q + 2) - math.log1p(-th / (q + 2))
            lo, hi = (th, hi) if lg <= math.log(tol) else (lo, th)
        return lo
    th = np.array([theta(q) for q in range(1, 18)])
    n0 = np.abs(p.G0).sum(0).max()
    nj = np.array([np.abs(g).sum(0).max() for g in p.Gj])
    u = Z[p.u_off:p.u_off + p.m, :-1]
    nrm = np.abs(Z[p.dt_off, :-1]) * (n0 + (np.abs(u) * nj[:, None]).sum(0))
    M = 1 + (th[None, :] < nrm[:, None]).sum(1)
    return float(((48 * M - 24) * 512).sum()), float(M.mean())


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        v = json.load(open(path)).get(workload)
        return v if isinstance(v, (int, float)) else None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag = index, period, False
        self.sm, self.reasons, self.max_mhz, self.err = [], 0, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while True:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.reasons |= nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    self.reasons |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.stop_flag:
                    break
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report, never fake
            self.err = repr(e)

    def summary(self):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
                 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
                 0x100: "display_clock_setting"}
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
               "sm_max_mhz": self.max_mhz, "samples": len(self.sm),
               "reasons": [n for b, n in names.items() if self.reasons & b and n != "gpu_idle"]}
        if self.err:
            out["error"] = self.err
        return out


def run_reference(args):
    """CPU arm: the C++ port of the reference algorithm, all host threads, bounded steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import configs as C
    from oracle import cport as CP
    p, Z, _ = C.trajectory(args.config)
    threads = CP.max_threads()
    evals = p.K - 1

    def one_pass():
        CP.residual(p, Z, threads)
        CP.jacobian_values(p, Z, threads)

    # a reference "step" is a bounded sample of `passes` whole passes over the trajectory, sized from a
    # calibration pass so that the K timed steps last >= ~3 s (20 single passes were 0.15 s: +-30 % noise)
    # (the pool's threads take a few passes to spin up: calibrate on warm passes, and on the FASTEST of them)
    for _ in range(5):
        one_pass()
    t_pass = float("inf")
    for _ in range(10):
        t0 = time.perf_counter()
        one_pass()
        t_pass = min(t_pass, time.perf_counter() - t0)
    passes = int(min(2000, max(1, np.ceil(3.0 / max(args.steps, 1) / t_pass))))

    def step():
        for _ in range(passes):
            one_pass()

    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = evals * passes * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(p, args.config, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {passes} passes over the {evals}-eval C{args.config} trajectory "
                                   f"({dt:.1f} s timed; residual + Jacobian), C++ port of the reference's expv+dual algorithm; "
                                   "the Julia reference itself cannot run here (no Julia toolchain)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(p, cfg, n_gpus):
    return {"workload": f"C{cfg}: {p.kind} d={p.b // 2 if p.kind != 'density' else int(round(p.b ** 0.5))} "
                        f"(iso n_x={p.n_x}) m={p.m} drives, K={p.K} knots per GPU "
                        f"({p.K - 1} knot evals/step/GPU), fused residual+Jacobian",
            "knots_per_gpu": p.K, "n_x": p.n_x, "drives": p.m, "D": p.D,
            "parallelism": f"knot-sharded x{n_gpus}" if n_gpus > 1 else "single GPU"}


def cpu_baseline(p, Z, budget_s=12.0):
    from oracle import cport as CP
    threads = CP.max_threads()
    CP.residual(p, Z, threads)
    CP.jacobian_values(p, Z, threads)
    n, t0 = 0, time.perf_counter()
    while True:
        CP.residual(p, Z, threads)
        CP.jacobian_values(p, Z, threads)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 2000:
            break
    return {"value": (p.K - 1) * n / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} passes over the {p.K - 1}-eval trajectory in {el:.1f} s "
                      "(oracle/c/knot_ref.cpp: Taylor expv + forward jets, std::thread over knots)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import piccolo_b200 as pb
    from oracle import configs as C   # input generator + cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and "PB2_HOST_THREADS" not in os.environ:
        # one process per GPU on one host: divide the cores between the ranks' un-packing pools (e2e leg)
        os.environ["PB2_HOST_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
        os.environ.setdefault("PB2_HOST_NT", "1")   # world x 23.5 MB of host arrays per step exceed the last-level cache
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this benchmark has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner / debug lines on stdout; stdout must carry the JSON line only,
        # so file descriptor 1 points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    if args.strong and world > 1:
        # strong scaling (BASELINE config 5 as written: ONE trajectory of K knots sharded over the ranks): every rank
        # owns (K - 1) / world knot evaluations of the same problem
        full, _, _ = C.problem(args.config)[0], None, None
        per = (full.K - 1 + world - 1) // world
        p, Z, _ = C.trajectory(args.config, per + 1)
    else:
        p, Z, _ = C.trajectory(args.config)
    n_eval = p.K - 1
    B = pb.B200BilinearIntegrator(p.kind, p.G0, list(p.Gj), K=p.K, D=p.D, x_off=p.x_off,
                                  dt_off=p.dt_off, u_off=p.u_off, device=local,
                                  knot0=rank * n_eval, algorithm=args.algorithm)
    # the trajectory buffers of this benchmark are written once, long before the timed region (in an Ipopt
    # loop Z arrives by a copy): the promise PB2_OPT_EARLY_Z asks for (include/piccolo_b200.h)
    # ... and each step writes its own rotating output set, which no other kernel of the stream touches:
    # PB2_OPT_PIPELINED (no dependency wait at all between consecutive callbacks' grids)
    early_z = os.environ.get("PB2_BENCH_EARLY_Z", "1") == "1"
    # (N > 1: consecutive steps use different gather buffers as well, and every step ends inside its kernel with the
    # exchange of arrival words, so the same promise holds)
    # (N > 1: steps are not overlapped -- measured: overlapping them hides the barrier at N = 2 (22.4 -> 19.9 us) but
    # a run at N = 8 did not finish; PB2_BENCH_PIPELINED=2 forces it for experiments)
    pipelined = early_z and ((world == 1 and os.environ.get("PB2_BENCH_PIPELINED", "1") == "1") or
                             os.environ.get("PB2_BENCH_PIPELINED") == "2")
    if early_z:
        B.set_option("early_z", 1)
    if pipelined:
        B.set_option("pipelined", 1)
    chunk = B.dim + B.nnz_jac            # doubles of the canonical [delta | values] arrays per rank
    # N > 1: what crosses NVLink is the compact record per knot (the d/dx_k block is n_b copies of one
    # b x b block); one all-gather of the records, then a local expansion into the canonical arrays
    cs = B.compact_stride if world > 1 else 0
    # rotating buffer sets so that consecutive steps never find their lines in L2
    set_bytes = 8 * (chunk * world + p.D * p.K + (cs * n_eval * (world + 1) if cs else 0))
    nsets = max(2, int(np.ceil(2.2 * L2_BYTES / set_bytes)))
    rng = np.random.default_rng(rank)
    zflat = Z.reshape(-1, order="F")
    Zs, outs, comp_loc, comp_all = [], [], [], []
    for s in range(nsets):
        zz = zflat.copy()
        if s:  # distinct data per set (tiny perturbation of the controls' low bits is enough)
            zz += 1e-9 * rng.standard_normal(zz.size)
        Zs.append(torch.from_numpy(zz).to(dev))
        outs.append(torch.empty(chunk * world, dtype=torch.float64, device=dev))
        if cs:
            comp_all.append(torch.empty(cs * n_eval * world, dtype=torch.float64, device=dev))
            comp_loc.append(comp_all[-1][rank * cs * n_eval:(rank + 1) * cs * n_eval])
    # N > 1, fused exchange: every rank maps every other rank's gather buffers (symmetric memory over NVLink) and the
    # kernel writes each finished knot's record into all of them; the only collective left is a barrier
    peer_ptrs, symm_handles, fused = [], [], False
    symm_barrier = os.environ.get("PB2_SYMM_BARRIER", "1") == "1"
    kernel_barrier = os.environ.get("PB2_KERNEL_BARRIER", "1") == "1"
    if cs and os.environ.get("PB2_FUSED_XCHG", "1") != "0":
        try:
            lib0 = pb.load_library()
            for r in range(world):
                if r != rank:
                    assert lib0.pb2_enable_peer_access(local, r) == 0, lib0.pb2_last_error().decode()
            import torch.distributed._symmetric_memory as symm_mem
            for s in range(nsets):
                # re-home this set's gather buffer in symmetric memory (cuMem-mapped into every rank)
                t = symm_mem.empty(comp_all[s].numel() + 16, dtype=torch.float64, device=dev)   # + the arrival words
                t.zero_()
                hdl = symm_mem.rendezvous(t, dist.group.WORLD)
                comp_all[s] = t[:cs * n_eval * world]
                comp_loc[s] = t[rank * cs * n_eval:(rank + 1) * cs * n_eval]
                peer_ptrs.append([int(x) for x in hdl.buffer_ptrs])
                symm_handles.append(hdl)
            fused = True
        except Exception as e:   # no peer mapping available: the NCCL all-gather path below is used
            print(f"[bench] fused exchange unavailable ({e!r}); using NCCL all-gather", file=sys.stderr)
            fused = False
        flags = torch.tensor([1 if fused else 0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        fused = bool(flags.item())
    sync_token = torch.zeros(1, device=dev)
    # a dedicated non-default stream: the kernels, the timing events and (N > 1) the collective
    # all go on it, so the CUDA events bracket exactly the work they claim to
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    def step(i, cuda_stream, collective=True):
        s = i % nsets
        if cs:
            # the step's result is the gathered set of per-knot records on every rank (what the one consumer --
            # the rank that hosts Ipopt -- turns into its COO arrays with pb2_expand_compact_async, outside the step)
            if collective and fused:
                if kernel_barrier and collective != "nobarrier":
                    # the step's barrier is inside the kernel: the knot whose record leaves last exchanges
                    # arrival words with every rank (pb2_residual_jacobian_exchange_sync_async)
                    B.residual_jacobian_exchange_sync_device(Zs[s], rank, peer_ptrs[s], rank * cs * n_eval,
                                                             cs * n_eval * world, cuda_stream)
                    return
                B.residual_jacobian_exchange_device(Zs[s], rank, peer_ptrs[s], rank * cs * n_eval, cuda_stream)
                if collective == "nobarrier":        # diagnostic graph only: the exchange kernel without the barrier
                    return
                if symm_barrier:
                    symm_handles[0].barrier(channel=0)   # signal-pad barrier of the symmetric-memory handle
                else:
                    dist.all_reduce(sync_token)      # barrier: every rank's records have landed everywhere
                return
            B.residual_jacobian_compact_device(Zs[s], comp_loc[s], cuda_stream)
            if collective:
                dist.all_gather_into_tensor(comp_all[s], comp_loc[s])
            return
        slot = outs[s][rank * chunk:(rank + 1) * chunk]
        B.residual_jacobian_device(Zs[s], slot[:B.dim], slot[B.dim:], cuda_stream)
        if world > 1 and collective:
            dist.all_gather_into_tensor(outs[s], slot)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def capture(collective):
        """The K timed steps as ONE CUDA graph: a callback's kernel lasts a few microseconds, less
        than the host needs to issue it, so replaying a captured launch sequence is the only way to
        time the device work rather than the Python interpreter."""
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            cs = torch.cuda.current_stream().cuda_stream
            for i in range(args.steps):
                step(i, cs, collective)
        return g

    warmup = max(args.warmup, 3)
    for i in range(warmup):
        step(i, stream.cuda_stream)
    barrier()
    if world > 1:
        # self-check of the exchange (outside the timed region): buffer set 0 holds the same trajectory on
        # every rank, so after a step each rank's shard of the gathered arrays must equal the local result
        step(0, stream.cuda_stream)
        barrier()
        if cs:   # the consumer's expansion of the gathered records (not part of the timed step)
            B.expand_compact_device(comp_all[0], n_eval * world, outs[0][:B.dim * world], outs[0][B.dim * world:],
                                    stream.cuda_stream)
            torch.cuda.synchronize()
        ref = torch.empty(chunk, dtype=torch.float64, device=dev)
        B.residual_jacobian_device(Zs[0], ref[:B.dim], ref[B.dim:], stream.cuda_stream)
        torch.cuda.synchronize()
        got = outs[0]
        for r in range(world):
            if cs:
                ok = torch.equal(got[r * B.dim:(r + 1) * B.dim], ref[:B.dim]) and \
                     torch.equal(got[B.dim * world + r * B.nnz_jac:B.dim * world + (r + 1) * B.nnz_jac], ref[B.dim:])
            else:
                ok = torch.equal(got[r * chunk:(r + 1) * chunk], ref)
            if not ok:
                raise SystemExit(f"bench.py: rank {rank}: gathered shard of rank {r} differs from the local result")
        barrier()
    l0 = B.launch_count
    g_step = capture(True)
    g_kern = capture(False) if world > 1 else g_step
    g_xk = capture("nobarrier") if (world > 1 and fused and os.environ.get("PB2_BENCH_DIAG") == "1") else None
    launches = (B.launch_count - l0) - (args.steps if world > 1 else 0)   # launches recorded per replay of the step graph
    g_step.replay()          # untimed: graph upload, first-touch of every rotating set
    g_kern.replay()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    reps = 5                 # the same K-step graph, timed several times; the median is reported
    step_ms, kern_ms_l = [], []
    for _ in range(reps):
        barrier()
        ev[0].record()
        g_step.replay()
        ev[1].record()
        barrier()
        step_ms.append(ev[0].elapsed_time(ev[1]))
        ev[2].record()
        g_kern.replay()
        ev[3].record()
        barrier()
        kern_ms_l.append(ev[2].elapsed_time(ev[3]))
    total_ms = float(np.median(step_ms))
    kern_ms = float(np.median(kern_ms_l)) / args.steps
    xk_ms = None
    if g_xk is not None:
        g_xk.replay()
        barrier()
        xs = []
        for _ in range(reps):
            barrier()
            ev[0].record()
            g_xk.replay()
            ev[1].record()
            barrier()
            xs.append(ev[0].elapsed_time(ev[1]))
        xk_ms = float(np.median(xs)) / args.steps
        del g_xk

    # ---- ONE isolated launch (what a serial Ipopt callback sees: nothing to overlap with, caches cold) ----
    # Each sample is a one-launch CUDA graph replayed after an L2 flush: the events bracket the device work, not the
    # interpreter's launch path.
    iso_us = None
    if world == 1:
        flush = torch.empty(L2_BYTES * 2 // 8, dtype=torch.float64, device=dev)
        singles = []
        for i in range(4):
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, stream=stream):
                step(i, torch.cuda.current_stream().cuda_stream)
            singles.append(g1)
        iso = []
        for i in range(16):
            flush.zero_()                                  # evict the outputs / inputs of earlier launches
            torch.cuda.synchronize()
            ev[0].record()
            singles[i % 4].replay()
            ev[1].record()
            torch.cuda.synchronize()
            iso.append(ev[0].elapsed_time(ev[1]) * 1e3)
        iso_us = float(np.median(iso[4:]))
        del flush, singles

    # ---- the Lagrangian-Hessian callback, reported separately (device-resident, same graph scheme) ----
    hess = None
    if world == 1:
        dmu = torch.randn(B.dim, dtype=torch.float64, device=dev)
        dH = [torch.empty(B.nnz_hess, dtype=torch.float64, device=dev) for _ in range(2)]
        hsteps = max(3, min(args.steps, 50))
        for i in range(3):
            B.hessian_device(Zs[i % nsets], dmu, dH[i & 1], stream.cuda_stream)
        torch.cuda.synchronize()
        gh = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gh, stream=stream):
            hs = torch.cuda.current_stream().cuda_stream
            for i in range(hsteps):
                B.hessian_device(Zs[i % nsets], dmu, dH[i & 1], hs)
        gh.replay()
        torch.cuda.synchronize()
        ev[0].record()
        gh.replay()
        ev[1].record()
        torch.cuda.synchronize()
        h_ms = ev[0].elapsed_time(ev[1]) / hsteps
        hess = {"ms_per_callback": h_ms, "evals_per_sec": n_eval / (h_ms * 1e-3), "nnz_per_knot": B.nnz_hess // max(1, n_eval),
                "kernel": {"u8h": "knot_u8h_kernel (DMMA forward + adjoint jets, 3-qubit shape)",
                           "dmmah": "knot_dmmah_kernel (DMMA forward tiles: state, first- and second-order jets; adjoint "
                                    "tiles with the transposed generator)",
                           "generic": "knot_generic_kernel<2> (second-order Taylor jets in shared memory)"}[B.hessian_algorithm]}
        del gh

    # ---- objective value + gradient, device-resident, same graph scheme ----
    objective = None
    if world == 1 and p.kind == "unitary":
        d = int(round(np.sqrt(p.n_x // 2)))
        rng_o = np.random.default_rng(5)
        Ug = np.linalg.qr(rng_o.standard_normal((d, d)) + 1j * rng_o.standard_normal((d, d)))[0]
        comps = {"U": range(p.x_off, p.x_off + p.n_x), "Δt": range(p.dt_off, p.dt_off + 1),
                 "u": range(p.u_off, p.u_off + p.m)}
        traj = pb.NamedTrajectory(Z, comps)
        Jobj = pb.UnitaryInfidelityObjective(Ug, "U", traj, Q=100.0) + pb.QuadraticRegularizer("u", traj, 1e-2)
        dJ = torch.zeros(1, dtype=torch.float64, device=dev)
        dG = [torch.empty(p.D * p.K, dtype=torch.float64, device=dev) for _ in range(2)]
        osteps = max(3, min(args.steps, 200))
        for i in range(3):
            Jobj.value_gradient_device(Zs[i % nsets], dJ, dG[i & 1], stream.cuda_stream)
        torch.cuda.synchronize()
        go = torch.cuda.CUDAGraph()
        with torch.cuda.graph(go, stream=stream):
            os_ = torch.cuda.current_stream().cuda_stream
            for i in range(osteps):
                Jobj.value_gradient_device(Zs[i % nsets], dJ, dG[i & 1], os_)
        go.replay()
        torch.cuda.synchronize()
        ev[0].record()
        go.replay()
        ev[1].record()
        torch.cuda.synchronize()
        o_ms = ev[0].elapsed_time(ev[1]) / osteps
        obytes = 16 * p.D * p.K
        objective = {"ms_per_callback": o_ms, "terms": "UnitaryInfidelityObjective + QuadraticRegularizer(u)",
                     "algorithmic_bytes": obytes, "achieved_GBps": obytes / (o_ms * 1e-3) / 1e9,
                     "kernel": "knot_objective_kernel (one CTA per knot, one launch for value and gradient)",
                     "note": "latency-bound at this size: the whole trajectory is %d KB" % (8 * p.D * p.K // 1024)}
        del go
        # ---- one whole NLP iterate on the device: eval_g + eval_jac_g (dynamics and linear rows),
        #      eval_f + eval_grad_f, eval_h, all reading the same resident trajectory ----
        iterate = None
        if hess is not None and p.D == p.n_x + 2 + 3 * p.m and p.x_off == 0:
            traj_s = pb.NamedTrajectory.smooth_pulse_layout(Z, p.n_x, p.m, "U")
            Lc = pb.B200KnotLinearConstraints(traj_s)
            dLd = torch.empty(Lc.dim, dtype=torch.float64, device=dev)
            dLv = torch.empty(Lc.nnz_jac, dtype=torch.float64, device=dev)
            dOh = torch.empty(max(Jobj.hessian_structure()[0].size, 1), dtype=torch.float64, device=dev)
            isteps = max(3, min(args.steps, 50))

            def one_iterate(i, st_):
                s_ = i % nsets
                B.residual_jacobian_device(Zs[s_], outs[s_][:B.dim], outs[s_][B.dim:], st_)
                Lc.residual_jacobian_device(Zs[s_], dLd, dLv, st_)
                Jobj.value_gradient_device(Zs[s_], dJ, dG[i & 1], st_)
                B.hessian_device(Zs[s_], dmu, dH[i & 1], st_)
                Jobj.hessian_device(Zs[s_], 1.0, dOh, st_)          # sigma * d2J: the objective block of eval_h

            for i in range(3):
                one_iterate(i, stream.cuda_stream)
            torch.cuda.synchronize()
            gi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gi, stream=stream):
                is_ = torch.cuda.current_stream().cuda_stream
                for i in range(isteps):
                    one_iterate(i, is_)
            gi.replay()
            torch.cuda.synchronize()
            ev[0].record()
            gi.replay()
            ev[1].record()
            torch.cuda.synchronize()
            it_ms = ev[0].elapsed_time(ev[1]) / isteps
            del gi
            # the five callbacks of an iterate are independent of one another (same trajectory in, disjoint outputs):
            # as parallel branches of the graph the iterate costs about its longest member, the Hessian
            # (the 3-qubit Hessian kernel needs a whole SM per CTA: it runs on a high-priority branch and leaves
            # PB2_BENCH_HESS_RESERVE SMs to the other four callbacks, PB2_OPT_HESSIAN_CTAS)
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            # one knot more per Hessian CTA than the full machine would give: C3, 999 knots -> 8 instead of 7 per CTA
            # -> 125 CTAs, 23 SMs left to the other callbacks
            kp = -(-n_eval // n_sm) + 1
            reserve = int(os.environ.get("PB2_BENCH_HESS_RESERVE", str(max(0, n_sm - (-(-n_eval // kp))))))
            if B.hessian_algorithm == "u8h" and 0 < reserve < n_sm:
                B.set_option("hessian_ctas", n_sm - reserve)
            sides = [torch.cuda.Stream(device=dev) for _ in range(2)] + [torch.cuda.Stream(device=dev, priority=-1)] + \
                    [torch.cuda.Stream(device=dev)]

            def one_iterate_par(i, main):
                s_ = i % nsets
                for sd in sides:
                    sd.wait_stream(main)
                B.hessian_device(Zs[s_], dmu, dH[i & 1], sides[2].cuda_stream)
                B.residual_jacobian_device(Zs[s_], outs[s_][:B.dim], outs[s_][B.dim:], main.cuda_stream)
                Lc.residual_jacobian_device(Zs[s_], dLd, dLv, sides[0].cuda_stream)
                Jobj.value_gradient_device(Zs[s_], dJ, dG[i & 1], sides[1].cuda_stream)
                Jobj.hessian_device(Zs[s_], 1.0, dOh, sides[3].cuda_stream)
                for sd in sides:
                    main.wait_stream(sd)

            gp = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gp, stream=stream):
                for i in range(isteps):
                    one_iterate_par(i, torch.cuda.current_stream())
            gp.replay()
            torch.cuda.synchronize()
            ev[0].record()
            gp.replay()
            ev[1].record()
            torch.cuda.synchronize()
            itp_ms = ev[0].elapsed_time(ev[1]) / isteps
            del gp
            B.set_option("hessian_ctas", 0)
            iterate = {"ms_per_iterate": it_ms, "ms_per_iterate_concurrent": itp_ms, "launches_per_iterate": 5,
                       "calls": "residual+Jacobian (dynamics), residual+Jacobian (derivative pairs, time consistency), "
                                "objective value+gradient, Lagrangian Hessian (dynamics), objective Hessian; "
                                "one resident trajectory; ms_per_iterate: one stream, ms_per_iterate_concurrent: the five "
                                "as parallel branches of the graph (the 3-qubit Hessian kernel on a high-priority branch "
                                "and on all but %d SMs)" % reserve}
            Lc.close()
        objective["nlp_iterate"] = iterate
        Jobj.close()

    # ---- end to end through the public host-pointer API (pinned host buffers, H2D + D2H) ----
    import ctypes
    lib = pb.load_library()

    def pinned(n):
        ptr = ctypes.c_void_p()
        assert lib.pb2_host_alloc(ctypes.byref(ptr), 8 * n) == 0
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(n,)), ptr

    hZ, pZ = pinned(p.D * p.K)
    hD, pD = pinned(max(1, B.dim))
    hV, pV = pinned(max(1, B.nnz_jac))
    hZ[:] = zflat
    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(2):
        B.residual_jacobian(hZ, hD, hV)
    barrier()
    e2e_runs = []
    for _ in range(3):                       # median of three batches: host threads take part in this path
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            B.residual_jacobian(hZ, hD, hV)     # H2D(Z) + kernel + D2H(delta, vals), synchronous
        torch.cuda.synchronize()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_runs))
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    times = torch.tensor([total_ms, kern_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_ms = (float(x) for x in times.cpu())

    if rank == 0:
        value = n_eval * world * args.steps / (total_ms * 1e-3)
        bytes_launch = (8 * (p.n_x + p.m + 1 + cs) if cs else algorithmic_bytes_per_eval(p)) * n_eval
        peak, peak_src = measured_peak()
        achieved = bytes_launch / (kern_ms * 1e-3) / 1e9
        workload = f"C{args.config}"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(p, args.config, world),
            "detail": dict(algorithm=B.algorithm, early_z=bool(early_z), pipelined=bool(pipelined),
                           exchange_kernel_ms_without_barrier=xk_ms,
                           l2=f"rotating {nsets} buffer sets ({nsets * set_bytes / 2**20:.0f} MiB > 126 MiB L2); "
                              "inputs and outputs resident in HBM",
                           timing=f"the {args.steps} steps are captured once as a CUDA graph and replayed; CUDA events "
                                  f"around the replay, median of {reps} replays, max over ranks",
                           collective=("fused exchange: the kernel stores each knot's compact record "
                                       f"({8 * cs} B instead of {8 * (p.n_x + p.nnz_jac_knot)} B) into every rank's gather buffer "
                                       "over NVLink (torch symmetric memory) as the knot finishes, then one "
                                       + ("exchange of arrival words between the ranks at the end of the same kernel" if kernel_barrier else
                                          "signal-pad barrier of the symmetric-memory handle" if symm_barrier else "1-element all_reduce as barrier")
                                       + "; result = the gathered records on every rank (NVLink ingress per rank and step: "
                                       f"{8 * cs * n_eval * (world - 1) / 1e6:.1f} MB = {8 * cs * n_eval * (world - 1) / 770e3:.1f} us at the "
                                       "770 GB/s measured for peer copies)" if (cs and fused) else
                                       "one NCCL all_gather_into_tensor of the compact per-knot records "
                                       f"({8 * cs} B/knot instead of {8 * (p.n_x + p.nnz_jac_knot)} B)" if cs
                                       else "one NCCL all_gather_into_tensor of [delta|vals] per step") if world > 1 else "none"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(workload),
                         "kernel": f"knot resjac ({B.algorithm})", "kernel_ms": kern_ms,
                         "isolated_launch_us": iso_us,
                         "isolated_frac": (bytes_launch / (iso_us * 1e-6) / 1e9 / peak) if iso_us else None,
                         "algorithmic_bytes_per_launch": bytes_launch, "peak_source": peak_src},
            "fp64_tensor": None,
            "e2e": {"value": n_eval * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": 8 * p.D * p.K,
                    "d2h_bytes_per_step": 8 * B.compact_stride * n_eval if B.compact_stride else 8 * (B.dim + B.nnz_jac),
                    "host_array_bytes_per_step": 8 * (B.dim + B.nnz_jac),
                    "steps": e2e_steps,
                    "note": "pb2_residual_jacobian with pinned host buffers; per rank its own shard. The d/dx_k block is "
                            "n_b mirrored copies of one half block and the identity entries are constant, so the library "
                            "moves the non-redundant record per knot over PCIe in chunks and host threads replicate each "
                            "chunk into the caller's COO-ordered arrays while the next chunk is in flight"},
            "hessian": hess,
            "objective": objective,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        work = fp64_tensor_work(p, Z)
        if work is not None:
            # the kernel sits at the FP64 ridge: co-report the tensor pipe.  Peak = DMMA rate
            # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak_b200.json).
            tf = work[0] / (kern_ms * 1e-3) / 1e12
            line["fp64_tensor"] = {"achieved": tf, "peak": 37.13, "unit": "TFLOP/s", "frac": tf / 37.13,
                                   "mean_taylor_degree": work[1],
                                   "flops_per_launch": work[0], "peak_source": "measured (profiles/r01_fp64_peak_b200.json)"}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(p, Z)
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    for ptr in (pZ, pD, pV):
        lib.pb2_host_free(ptr)
    # teardown order matters with captured NCCL work: drop the graphs, drain the device, then the
    # handle; the process group is left to process exit (destroying a communicator that captured
    # graphs still reference can block)
    del g_step, g_kern
    torch.cuda.synchronize()
    B.close()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--algorithm", default="auto")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: shard ONE trajectory of the configuration's K knots over the ranks (default: K per rank)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
