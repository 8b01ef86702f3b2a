"""Knot sharding across GPUs: one process per GPU, one all-gather per callback.

The constraint couples only knots (k, k+1), so rank r owns a contiguous range of knots and
needs the Z columns of that range plus ONE halo column.  Each rank's
kernel writes its [delta shard | Jacobian-value shard] straight into its slot of the gather
buffer (no pack kernel), then a single in-place ``all_gather_into_tensor`` (NCCL over
NVLink/NVSwitch on the GPU box; gloo in the CPU plumbing tests) assembles the full arrays on
every rank.  COO structure is computed redundantly per rank (no communication).

For the 3-qubit unitary shape the exchange is fused into the kernel: the record buffer lives in
symmetric memory (mapped into every rank over NVLink), each finished knot's compact record is
stored by the producing kernel straight into every rank's buffer, and the only collective left is
the symmetric-memory barrier; every rank then expands all records locally.  ``fused="auto"`` uses
it whenever the mapping can be set up on all ranks and falls back to the NCCL all-gather of the
records otherwise (a transport choice; both run the same CUDA kernels).

The reference has no distributed code at all; this is the new exchange
step the north star asks for.
"""
import numpy as np


def knot_partition(n_constraints, world):
    """Equal-count partition of constraint indices 0..n-1: rank r owns [r*per, min((r+1)*per, n))."""
    per = -(-n_constraints // world) if n_constraints > 0 else 0
    return per, [(min(r * per, n_constraints), min((r + 1) * per, n_constraints)) for r in range(world)]


class ShardedBilinearIntegrator:
    """Evaluates the residual + Jacobian of one integrator with the knots split over
    ``world`` ranks and gathers the result on every rank.

    ``make_local(K_local, knot0)`` builds the rank-local evaluator; the default is the CUDA
    handle (B200BilinearIntegrator).  Tests inject a CPU stand-in to exercise the partition /
    gather / unpack logic under gloo -- the product path never does.
    """

    def __init__(self, kind, G_drift, G_drives, *, K, D, x_off, dt_off, u_off, rank, world,
                 device=0, group=None, tensor_device=None, make_local=None, algorithm="auto", fused="auto"):
        import torch
        self.torch = torch
        self.K, self.D = K, D
        self.rank, self.world, self.group = rank, world, group
        b = np.asarray(G_drift).shape[0]
        n_b = b // 2 if kind == "unitary" else 1
        m = len(G_drives)
        self.n_x = b * n_b
        self.nnz_knot = n_b * b * b + self.n_x * m + 2 * self.n_x
        self.n_con = K - 1
        self.per, self.ranges = knot_partition(self.n_con, world)
        self.k0, self.k1 = self.ranges[rank]
        self.n_local = self.k1 - self.k0
        if make_local is None:
            from .integrators import B200BilinearIntegrator

            def make_local(K_local, knot0):
                return B200BilinearIntegrator(kind, G_drift, G_drives, K=K_local, D=D, x_off=x_off,
                                              dt_off=dt_off, u_off=u_off, device=device,
                                              knot0=knot0, algorithm=algorithm)
        self.local = make_local(self.n_local + 1, self.k0)
        self.tensor_device = tensor_device or (f"cuda:{device}" if torch.cuda.is_available() else "cpu")
        self.chunk = self.per * (self.n_x + self.nnz_knot)       # doubles per rank slot
        self.gather = torch.zeros(max(1, self.chunk * world), dtype=torch.float64,
                                  device=self.tensor_device)
        self.zslab = torch.zeros(D * (self.n_local + 1), dtype=torch.float64, device=self.tensor_device)
        # compact records (include/piccolo_b200.h): when the local evaluator offers them, the records
        # are what crosses NVLink and the canonical arrays are rebuilt locally after the gather
        self.cs = int(getattr(self.local, "compact_stride", 0)) if str(self.tensor_device).startswith("cuda") else 0
        self.fused, self._symm, self._peer_ptrs = False, None, None
        if self.cs:
            self.comp = torch.zeros(max(1, self.cs * self.per * world), dtype=torch.float64, device=self.tensor_device)
            if world > 1 and fused in ("auto", True):
                self._setup_fused(device, required=fused is True)

    def _setup_fused(self, device, required):
        """Re-home the record buffer in symmetric memory and map every rank's copy (NVLink peer memory)."""
        torch = self.torch
        dist = torch.distributed
        ok = True
        try:
            from . import capi
            lib = capi.load_library()
            for r in range(torch.cuda.device_count()):
                if r != device:
                    lib.pb2_enable_peer_access(device, r)      # best effort: symmetric memory maps via cuMem anyway
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(self.comp.numel(), dtype=torch.float64, device=self.tensor_device)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            self.comp, self._symm = t, hdl
            self._peer_ptrs = [int(x) for x in hdl.buffer_ptrs]
        except Exception as e:                                  # no peer mapping on this rank
            if required:
                raise
            ok, self._why_not_fused = False, repr(e)
        flag = torch.tensor([1 if ok else 0], device=self.tensor_device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.fused = bool(flag.item())
        if required and not self.fused:
            raise RuntimeError("fused exchange could not be set up on every rank")

    # views into the gather buffer ---------------------------------------------------------
    def _slot(self, r):
        return self.gather[r * self.chunk:(r + 1) * self.chunk]

    def my_delta(self):
        return self._slot(self.rank)[:self.per * self.n_x]

    def my_vals(self):
        return self._slot(self.rank)[self.per * self.n_x:]

    def local_columns(self, Z):
        """Columns [k0, k0 + n_local] of the D x K trajectory (owned knots + one halo)."""
        return Z[:, self.k0:self.k0 + self.n_local + 1]

    def evaluate_local(self, Z_host=None, stream=None):
        """Stage this rank's slab and launch the fused kernel into the gather slot."""
        torch = self.torch
        if Z_host is not None:
            slab = np.ascontiguousarray(self.local_columns(np.asarray(Z_host)).T).reshape(-1)
            self.zslab.copy_(torch.from_numpy(slab), non_blocking=True)
        if self.fused:
            # nobody may still be expanding the previous call's records when new ones arrive
            # (every rank takes part, also one that owns no knots)
            self._symm.barrier(channel=1)
        if self.n_local > 0:
            if self.zslab.is_cuda:
                st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
                if self.fused:
                    self.local.residual_jacobian_exchange_device(self.zslab, self.rank, self._peer_ptrs,
                                                                 self.rank * self.cs * self.per, st)
                elif self.cs:
                    mine = self.comp[self.rank * self.cs * self.per:(self.rank + 1) * self.cs * self.per]
                    self.local.residual_jacobian_compact_device(self.zslab, mine, st)
                else:
                    self.local.residual_jacobian_device(self.zslab, self.my_delta(), self.my_vals(), st)
            else:  # injected CPU stand-in (tests only)
                d, v = self.local.residual_jacobian(self.zslab.numpy().reshape(self.n_local + 1, self.D).T)
                self.my_delta()[:d.size] = torch.from_numpy(d)
                self.my_vals()[:v.size] = torch.from_numpy(v)

    def all_gather(self):
        """One collective per callback: every rank ends with every rank's [delta | vals] slot."""
        if self.cs:
            torch = self.torch
            mine = self.comp[self.rank * self.cs * self.per:(self.rank + 1) * self.cs * self.per]
            if self.fused:
                self._symm.barrier(channel=0)     # every rank's records have landed everywhere
            elif self.world > 1:
                torch.distributed.all_gather_into_tensor(self.comp, mine, group=self.group)
            st = torch.cuda.current_stream().cuda_stream
            for r in range(self.world):       # every rank slot holds `per` records (the last may be ragged)
                slot = self._slot(r)
                self.local.expand_compact_device(self.comp[r * self.cs * self.per:], self.per,
                                                 slot[:self.per * self.n_x], slot[self.per * self.n_x:], st)
            return self.gather
        if self.world > 1:
            self.torch.distributed.all_gather_into_tensor(self.gather, self._slot(self.rank).clone()
                                                          if not self.gather.is_cuda else self._slot(self.rank),
                                                          group=self.group)
        return self.gather

    def residual_jacobian(self, Z_host=None):
        self.evaluate_local(Z_host)
        self.all_gather()
        return self.gather

    def unpack(self):
        """(delta, vals) of the whole trajectory in canonical order, from the gather buffer."""
        g = self.gather.detach().cpu().numpy()
        deltas, vals = [], []
        for r, (a, c) in enumerate(self.ranges):
            n = c - a
            slot = g[r * self.chunk:(r + 1) * self.chunk]
            deltas.append(slot[:n * self.n_x])
            vals.append(slot[self.per * self.n_x:self.per * self.n_x + n * self.nnz_knot])
        return np.concatenate(deltas), np.concatenate(vals)
