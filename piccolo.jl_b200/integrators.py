"""Host-side mirror of the reference's integrator interface for the knot path.

Reference: /root/reference/src/control/integrators.jl:35-95 (dispatch on trajectory type),
DirectTrajOpt's evaluate!/eval_jacobian/jacobian_structure/hessian_structure contract as seen
from the reference's call sites (integrators.jl:307-311,780-782; display/inspect.jl:611-623).
Every numeric call goes through the C ABI (capi.py); nothing here computes the path on the CPU.
"""
import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi, generators as gen


# ----------------------------------------------------------------------------- #
# minimal containers mirroring the reference's (only what the path reads)
# ----------------------------------------------------------------------------- #

@dataclass
class QuantumSystem:
    """QuantumSystem(H_drift, H_drives, drive_bounds)  (quantum_systems.jl:192-249)."""
    H_drift: np.ndarray
    H_drives: List[np.ndarray]
    drive_bounds: Sequence[float] = ()

    def __post_init__(self):
        self.H_drift = np.asarray(self.H_drift, dtype=complex)
        # a drive may be given as (H, modulation) -- the reference's `H => t -> c(t)` pair, which becomes
        # ModulatedDrive(LinearDrive(H, j), c) and makes the system time dependent (quantum_systems.jl:540-597)
        drives, self.modulations, self.modulation_derivs = [], [], []
        for d in self.H_drives:
            if isinstance(d, tuple):
                drives.append(d[0])
                self.modulations.append(d[1])
                self.modulation_derivs.append(d[2] if len(d) > 2 else None)
            else:
                drives.append(d)
                self.modulations.append(None)
                self.modulation_derivs.append(None)
        self.H_drives = [np.asarray(H, dtype=complex) for H in drives]
        if not np.allclose(self.H_drift, self.H_drift.conj().T):
            raise ValueError("Drift Hamiltonian H_drift is not Hermitian")
        for i, H in enumerate(self.H_drives):
            if not np.allclose(H, H.conj().T):
                raise ValueError(f"Drive Hamiltonian H_drives[{i + 1}] is not Hermitian")

    levels = property(lambda self: self.H_drift.shape[0])
    n_drives = property(lambda self: len(self.H_drives))
    time_dependent = property(lambda self: any(f is not None for f in self.modulations))

    def td_kwargs(self, traj):
        """Extra constructor arguments of a time-dependent system's integrators (integrators.jl:38-46: the
        reference passes the :t component to TimeDependentBilinearIntegrator)."""
        if not self.time_dependent:
            return {}
        return dict(t_off=traj.components["t"].start, modulations=self.modulations,
                    modulation_derivs=self.modulation_derivs)

    def G_parts(self):
        return gen.G(self.H_drift), [gen.G(H) for H in self.H_drives]


@dataclass
class OpenQuantumSystem(QuantumSystem):
    """OpenQuantumSystem(H_drift, H_drives, bounds; dissipation_operators)  (open_quantum_systems.jl:30-41)."""
    dissipation_operators: List[np.ndarray] = field(default_factory=list)
    rates: Optional[List[float]] = None

    def G_parts(self):
        return gen.compact_generator_parts(self.H_drift, self.H_drives,
                                           self.dissipation_operators, self.rates)


class NamedTrajectory:
    """The slice of NamedTrajectories.NamedTrajectory the path touches: ``datavec`` as a
    D x K column-major matrix plus named row ranges (named_trajectory_conversion.jl:316-321)."""

    def __init__(self, data: np.ndarray, components: Dict[str, range], timestep="Δt", global_dim=0):
        self.data = np.asfortranarray(data, dtype=np.float64)
        self.components = dict(components)
        self.timestep = timestep
        self.global_dim = global_dim

    dim = property(lambda self: self.data.shape[0])
    N = property(lambda self: self.data.shape[1])

    @property
    def datavec(self):
        return self.data.reshape(-1, order="F")

    @classmethod
    def multi_state_layout(cls, data, state_names, n_x_each, m, derivs=2):
        """[state_1 | state_2 | ... | Δt | t | u | du | ddu]  (multi-ket / sampling trajectories)."""
        comps, o = {}, 0
        for name in state_names:
            comps[name] = range(o, o + n_x_each)
            o += n_x_each
        comps["Δt"], comps["t"], comps["u"] = range(o, o + 1), range(o + 1, o + 2), range(o + 2, o + 2 + m)
        o += 2 + m
        for name in ["du", "ddu", "dddu"][:derivs]:
            comps[name] = range(o, o + m)
            o += m
        return cls(data, comps)

    @classmethod
    def smooth_pulse_layout(cls, data, n_x, m, state_name, derivs=2):
        """[state | Δt | t | u | du | ddu]   (smooth_pulse_problem.jl:196-201)."""
        comps = {state_name: range(0, n_x), "Δt": range(n_x, n_x + 1), "t": range(n_x + 1, n_x + 2),
                 "u": range(n_x + 2, n_x + 2 + m)}
        o = n_x + 2 + m
        for name in ["du", "ddu", "dddu"][:derivs]:
            comps[name] = range(o, o + m)
            o += m
        return cls(data, comps)


@dataclass
class _QTraj:
    system: QuantumSystem


class UnitaryTrajectory(_QTraj):
    kind, state_name = "unitary", "Ũ⃗"


class KetTrajectory(_QTraj):
    kind, state_name = "ket", "ψ̃"


class DensityTrajectory(_QTraj):
    kind, state_name = "density", "ρ⃗̃"


class MultiKetTrajectory(_QTraj):
    """Several kets driven by one system and one control row: one integrator per state
    (integrators.jl:102-117); state components ψ̃1, ψ̃2, ... precede the shared [Δt | t | u ...]
    tail (named_trajectory_conversion.jl:465-)."""
    kind = "ket"

    def __init__(self, system, n_states):
        super().__init__(system)
        self.state_names = [f"ψ̃{i + 1}" for i in range(n_states)]


class MultiDensityTrajectory(_QTraj):
    """Several density matrices driven by one open system (compact iso each): one integrator per state."""
    kind = "density"

    def __init__(self, system, n_states):
        super().__init__(system)
        self.state_names = [f"ρ⃗̃{i + 1}" for i in range(n_states)]


class SamplingTrajectory:
    """An ensemble of systems sharing the controls: one integrator per member and state, each member
    with its own generator (integrators.jl:134-226).  State components are named ``<base>1 .. <base>n``,
    member-major for multi-state bases: member i owns names[(i-1)K .. iK) with K the number of states of
    the base trajectory (sampling_trajectory.jl:97-126, 306-349).  ``base`` is a trajectory type
    (single state) or an instance (``SamplingTrajectory(base_qtraj, systems)`` as in the reference)."""

    def __init__(self, base, systems):
        self.kind, self.systems = base.kind, list(systems)
        multi = isinstance(base, (MultiKetTrajectory, MultiDensityTrajectory))
        self.n_substates = len(base.state_names) if multi else 1
        stem = (base.state_names[0][:-1] if multi else base.state_name)
        n = len(self.systems) * self.n_substates
        self.state_names = [f"{stem}{i + 1}" for i in range(n)]

    def member_states(self):
        """sampling_member_states: the state names of each member (lists of K names)."""
        K = self.n_substates
        return [self.state_names[i * K:(i + 1) * K] for i in range(len(self.systems))]


# ----------------------------------------------------------------------------- #
# the integrator
# ----------------------------------------------------------------------------- #

def _as_ptr(a):
    """numpy array -> address; torch CUDA tensor / int -> address."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


class B200BilinearIntegrator:
    """One dynamics integrator (one state component) evaluated on a B200.

    Fields mirror DirectTrajOpt's BilinearIntegrator as used by the reference:
    ``dim == x_dim * (N - 1)`` (integrators.jl:309), ``x_dim``, ``x_name``.
    """

    def __init__(self, kind, G_drift, G_drives, *, K, D, x_off, dt_off, u_off, x_name="x",
                 u_name="u", device=0, algorithm="auto", knot0=0, global_dim=0, n_states=1,
                 t_off=None, modulations=None, modulation_derivs=None, dense_blocks=False):
        self._lib = capi.load_library()
        G0 = np.asfortranarray(G_drift, dtype=np.float64)
        b = G0.shape[0]
        m = len(G_drives)
        Gj = (np.concatenate([np.asfortranarray(g, dtype=np.float64).reshape(-1, order="F")
                              for g in G_drives]) if m else np.zeros(1))
        self.kind, self.b, self.m = kind, b, m
        if kind == "unitary" and n_states != 1:
            raise ValueError("n_states applies to ket / density blocks sharing one generator")
        self.n_b = b // 2 if kind == "unitary" else int(n_states)
        self.x_dim = b * self.n_b
        self.K, self.D = int(K), int(D)
        self.x_off, self.dt_off, self.u_off = int(x_off), int(dt_off), int(u_off)
        self.x_name, self.u_name = x_name, u_name
        self.knot0, self.global_dim, self.device = int(knot0), int(global_dim), int(device)
        self._G0, self._Gj = G0, Gj  # keep alive during create
        # time-dependent systems (TimeDependentBilinearIntegrator, integrators.jl:38-46): drive j is
        # ModulatedDrive(LinearDrive(H_j, j), c_j) (drives.jl:342-388); c_j = None means unmodulated
        self.modulations = list(modulations) if modulations is not None else None
        self.time_dependent = self.modulations is not None and any(f is not None for f in self.modulations)
        if self.time_dependent:
            if t_off is None or len(self.modulations) != m:
                raise ValueError("a time-dependent integrator needs t_off and one modulation (or None) per drive")
            derivs = list(modulation_derivs) if modulation_derivs is not None else [None] * m

            def numeric(f):     # ForwardDiff in the reference; here a central difference of the closure
                return lambda t: (f(t + 1e-6) - f(t - 1e-6)) / 2e-6
            self.modulation_derivs = [None if f is None else (g if g is not None else numeric(f))
                                      for f, g in zip(self.modulations, derivs)]
        self.t_off = int(t_off) if t_off is not None else 0
        d = capi.pb2_desc(capi.KIND[kind], b, self.n_b, m, self.K, self.D, self.x_off, self.dt_off,
                          self.u_off, self.global_dim, self.knot0, self.device, capi.ALG[algorithm],
                          G0.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                          Gj.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                          self.t_off, 1 if self.time_dependent else 0, 1 if dense_blocks else 0)
        self.dense_blocks = bool(dense_blocks)
        h = ctypes.c_void_p()
        capi.check(self._lib.pb2_create(ctypes.byref(d), ctypes.byref(h)))
        self._h = h
        self.dim = int(self._lib.pb2_dim(h))
        self.nnz_jac = int(self._lib.pb2_nnz_jac(h))
        self.nnz_hess = int(self._lib.pb2_nnz_hess(h))
        self.algorithm = {1: "generic", 2: "dmma"}[self._lib.pb2_algorithm(h)]
        self.hessian_algorithm = {0: None, 1: "generic", 2: "dmmah", 3: "u8h"}[self._lib.pb2_hessian_algorithm(h)]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb2_destroy(self._h)
            self._h = None

    __del__ = close

    # -- structure -------------------------------------------------------------------------
    def jacobian_structure(self):
        rows = np.empty(self.nnz_jac, dtype=np.int64)
        cols = np.empty(self.nnz_jac, dtype=np.int64)
        ip = ctypes.POINTER(ctypes.c_int64)
        capi.check(self._lib.pb2_structure_jac(self._h, rows.ctypes.data_as(ip), cols.ctypes.data_as(ip)))
        return rows, cols

    def hessian_structure(self):
        rows = np.empty(self.nnz_hess, dtype=np.int64)
        cols = np.empty(self.nnz_hess, dtype=np.int64)
        ip = ctypes.POINTER(ctypes.c_int64)
        capi.check(self._lib.pb2_structure_hess(self._h, rows.ctypes.data_as(ip), cols.ctypes.data_as(ip)))
        return rows, cols

    # -- host-pointer calls (what the Ipopt callbacks use) ---------------------------------
    def _Z(self, Z):
        Z = Z.data if isinstance(Z, NamedTrajectory) else Z
        Z = np.asarray(Z, dtype=np.float64)
        if Z.ndim == 2:
            if Z.shape != (self.D, self.K):
                raise ValueError(f"trajectory is {Z.shape}, integrator expects {(self.D, self.K)}")
            Z = np.asfortranarray(Z)
        elif Z.size != self.D * self.K:
            raise ValueError("datavec has the wrong length")
        if self.time_dependent:
            self.set_time_coefficients(*self.time_coefficients(Z))
        return Z

    def time_coefficients(self, Z):
        """c[j, k] = c_j(t_k), c'[j, k]: the modulation closures evaluated at the trajectory's current time row
        (what the Julia shim does before each callback; the closures themselves cannot cross the C ABI)."""
        t = np.asarray(Z).reshape(-1, order="F")[self.t_off::self.D][:self.K] if np.ndim(Z) == 1 else np.asarray(Z)[self.t_off, :]
        c, cd = np.ones((self.m, self.K), order="F"), np.zeros((self.m, self.K), order="F")
        for j, (f, g) in enumerate(zip(self.modulations, self.modulation_derivs)):
            if f is not None:
                c[j, :] = [f(tk) for tk in t]
                cd[j, :] = [g(tk) for tk in t]
        return c, cd

    def set_time_coefficients(self, c, cdot):
        c = np.asfortranarray(c, dtype=np.float64)
        cdot = np.asfortranarray(cdot, dtype=np.float64)
        if c.shape != (self.m, self.K) or cdot.shape != (self.m, self.K):
            raise ValueError("coefficient tables are m x K")
        capi.check(self._lib.pb2_set_time_coefficients(self._h, c.ctypes.data, cdot.ctypes.data, capi.PB2_HOST))

    @staticmethod
    def _out(a, n, what):
        """A caller-supplied output buffer goes straight to the C ABI: it must be exactly what the library
        writes (contiguous float64 of the exact length), or the call would overrun it silently."""
        if a is None:
            return np.empty(n)
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.size != n or not a.flags.c_contiguous \
                or not a.flags.writeable:
            raise ValueError(f"{what} must be a writeable contiguous float64 array of length {n}")
        return a

    def evaluate_(self, delta, Z):
        Z = self._Z(Z)
        delta = self._out(delta, self.dim, "delta")
        capi.check(self._lib.pb2_residual(self._h, Z.ctypes.data, delta.ctypes.data, capi.PB2_HOST))
        return delta

    def jacobian_values(self, Z, out=None):
        Z = self._Z(Z)
        out = self._out(out, self.nnz_jac, "out")
        capi.check(self._lib.pb2_jacobian(self._h, Z.ctypes.data, out.ctypes.data, capi.PB2_HOST))
        return out

    def residual_jacobian(self, Z, delta=None, vals=None):
        Z = self._Z(Z)
        delta = self._out(delta, self.dim, "delta")
        vals = self._out(vals, self.nnz_jac, "vals")
        capi.check(self._lib.pb2_residual_jacobian(self._h, Z.ctypes.data, delta.ctypes.data,
                                                   vals.ctypes.data, capi.PB2_HOST))
        return delta, vals

    def hessian_values(self, Z, mu, out=None):
        Z = self._Z(Z)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        if mu.size != self.dim:
            raise ValueError("mu must have length B.dim")
        out = self._out(out, self.nnz_hess, "out")
        capi.check(self._lib.pb2_hess_lagrangian(self._h, Z.ctypes.data, mu.ctypes.data,
                                                 out.ctypes.data, capi.PB2_HOST))
        return out

    # -- device-pointer, asynchronous calls (torch CUDA tensors or raw addresses) -----------
    def residual_jacobian_device(self, dZ, ddelta, dvals, stream=None):
        capi.check(self._lib.pb2_residual_jacobian_async(self._h, _as_ptr(dZ), _as_ptr(ddelta),
                                                         _as_ptr(dvals), _as_ptr(stream)))

    def hessian_device(self, dZ, dmu, dvals, stream=None):
        capi.check(self._lib.pb2_hess_lagrangian_async(self._h, _as_ptr(dZ), _as_ptr(dmu),
                                                       _as_ptr(dvals), _as_ptr(stream)))

    # -- compact records for sharded runs (see include/piccolo_b200.h) ------------------------
    @property
    def compact_stride(self):
        return int(self._lib.pb2_compact_stride(self._h))

    def residual_jacobian_compact_device(self, dZ, dcompact, stream=None):
        capi.check(self._lib.pb2_residual_jacobian_compact_async(self._h, _as_ptr(dZ), _as_ptr(dcompact),
                                                                 _as_ptr(stream)))

    def residual_jacobian_exchange_device(self, dZ, rank, gather_ptrs, slot_offset, stream=None):
        """Fused compute + exchange: records go straight into every rank's gather buffer (peer memory)."""
        arr = (ctypes.c_void_p * len(gather_ptrs))(*[int(x) for x in gather_ptrs])
        capi.check(self._lib.pb2_residual_jacobian_exchange_async(self._h, _as_ptr(dZ), len(gather_ptrs), int(rank), arr,
                                                                  int(slot_offset), _as_ptr(stream)))

    def residual_jacobian_exchange_sync_device(self, dZ, rank, gather_ptrs, slot_offset, flag_offset, stream=None):
        """Fused compute + exchange + step barrier in one kernel (pb2_residual_jacobian_exchange_sync_async)."""
        arr = (ctypes.c_void_p * len(gather_ptrs))(*[int(x) for x in gather_ptrs])
        capi.check(self._lib.pb2_residual_jacobian_exchange_sync_async(
            self._h, _as_ptr(dZ), len(gather_ptrs), int(rank), arr, int(slot_offset), int(flag_offset), _as_ptr(stream)))

    def expand_compact_device(self, dcompact, n_knots, ddelta, dvals, stream=None):
        capi.check(self._lib.pb2_expand_compact_async(self._h, _as_ptr(dcompact), int(n_knots), _as_ptr(ddelta),
                                                      _as_ptr(dvals), _as_ptr(stream)))

    # -- rollout of the piecewise-constant controls (rollout!, sync_trajectory!, rollout_divergence) ---------
    def rollout(self, Z, x0=None):
        """States at every knot from x_1 = x0 (default: Z's first state column), x_{k+1} = exp(Δt_k Ĝ(u_k)) x_k
        (rollouts_extensions.jl:46-92 for the zero-order-hold pulse of sync_trajectory!, problems.jl:186-208).
        Returns (states [x_dim x K, Fortran order], rollout_divergence, ||Δx_K||, ||x_K^collocation||)."""
        Z = self._Z(Z)
        states = np.empty((self.x_dim, self.K), order="F")
        out3 = np.empty(3)
        x0p = None
        if x0 is not None:
            x0 = np.ascontiguousarray(x0, dtype=np.float64)
            if x0.size != self.x_dim:
                raise ValueError("x0 must have length B.x_dim")
            x0p = x0.ctypes.data
        capi.check(self._lib.pb2_rollout(self._h, Z.ctypes.data, x0p, states.ctypes.data, out3.ctypes.data, capi.PB2_HOST))
        return states, float(out3[0]), float(out3[1]), float(out3[2])

    def rollout_divergence(self, Z, x0=None):
        """rollout_divergence(qcp)  (problems.jl:336-356) for this integrator's state component."""
        return self.rollout(Z, x0)[1]

    def rollout_device(self, dZ, dx0, dstates, dout3, stream=None):
        capi.check(self._lib.pb2_rollout_async(self._h, _as_ptr(dZ), _as_ptr(dx0), _as_ptr(dstates), _as_ptr(dout3),
                                               _as_ptr(stream)))

    def sync(self):
        capi.check(self._lib.pb2_sync(self._h))

    def set_option(self, name, value):
        """pb2_set_option (include/piccolo_b200.h): e.g. ``set_option("early_z", 1)``."""
        capi.check(self._lib.pb2_set_option(self._h, capi.OPT[name], int(value)))

    @property
    def launch_count(self):
        return int(self._lib.pb2_launch_count(self._h))


class B200IntegratorBatch:
    """All integrators of an ensemble (SamplingTrajectory members, or the states of a Multi*Trajectory) in ONE launch.

    The reference attaches one BilinearIntegrator per member and state (integrators.jl:102-117, 134-226); Ipopt's
    callbacks then walk the vector.  For the small systems those problems use, that is one launch latency per
    member; here the member is a grid axis.  Outputs are member-major: ``delta[i]``, ``vals[i]`` are what member
    ``i``'s own integrator would return, bit for bit."""

    def __init__(self, kind, generators, *, K, D, x_offs, dt_off, u_off, device=0, algorithm="auto",
                 global_dim=0, n_states=1, names=None):
        self._lib = capi.load_library()
        n = len(generators)
        if n < 1 or len(x_offs) != n:
            raise ValueError("one (G_drift, G_drives) pair and one x_off per member")
        keep, descs = [], (capi.pb2_desc * n)()
        dp = ctypes.POINTER(ctypes.c_double)
        for i, ((G_drift, G_drives), xo) in enumerate(zip(generators, x_offs)):
            G0 = np.asfortranarray(G_drift, dtype=np.float64)
            b, m = G0.shape[0], len(G_drives)
            Gj = (np.concatenate([np.asfortranarray(g, dtype=np.float64).reshape(-1, order="F")
                                  for g in G_drives]) if m else np.zeros(1))
            keep += [G0, Gj]
            n_b = b // 2 if kind == "unitary" else int(n_states)
            descs[i] = capi.pb2_desc(capi.KIND[kind], b, n_b, m, int(K), int(D), int(xo), int(dt_off), int(u_off),
                                     int(global_dim), 0, int(device), capi.ALG[algorithm],
                                     G0.ctypes.data_as(dp), Gj.ctypes.data_as(dp), 0, 0, 0)
        h = ctypes.c_void_p()
        capi.check(self._lib.pb2_batch_create(descs, n, ctypes.byref(h)))
        self._h = h
        self.n_members = n
        self.kind, self.K, self.D = kind, int(K), int(D)
        self.x_offs = [int(x) for x in x_offs]
        self.names = list(names) if names is not None else [f"x{i}" for i in range(n)]
        self.dim = int(self._lib.pb2_batch_dim(h))
        self.nnz_jac = int(self._lib.pb2_batch_nnz_jac(h))
        self.nnz_hess = int(self._lib.pb2_batch_nnz_hess(h))
        self.fused = bool(self._lib.pb2_batch_fused(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb2_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def _structure(self, fn, i, n):
        rows, cols = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        ip = ctypes.POINTER(ctypes.c_int64)
        capi.check(fn(self._h, int(i), rows.ctypes.data_as(ip), cols.ctypes.data_as(ip)))
        return rows, cols

    def jacobian_structure(self, member):
        return self._structure(self._lib.pb2_batch_structure_jac, member, self.nnz_jac)

    def hessian_structure(self, member):
        return self._structure(self._lib.pb2_batch_structure_hess, member, self.nnz_hess)

    def _Z(self, Z):
        Z = Z.data if isinstance(Z, NamedTrajectory) else Z
        Z = np.asarray(Z, dtype=np.float64)
        if Z.size != self.D * self.K:
            raise ValueError("trajectory size does not match the batch")
        return np.ascontiguousarray(Z.reshape(-1, order="F") if Z.ndim == 2 else Z)

    def residual_jacobian(self, Z):
        """-> (delta [n_members, dim], vals [n_members, nnz_jac])"""
        Z = self._Z(Z)
        delta = np.empty((self.n_members, self.dim))
        vals = np.empty((self.n_members, self.nnz_jac))
        capi.check(self._lib.pb2_batch_residual_jacobian(self._h, _as_ptr(Z), _as_ptr(delta), _as_ptr(vals), capi.PB2_HOST))
        return delta, vals

    def hessian_values(self, Z, mu):
        """mu: [n_members, dim] multipliers -> vals [n_members, nnz_hess]"""
        Z = self._Z(Z)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        if mu.size != self.n_members * self.dim:
            raise ValueError("one multiplier per constraint row of every member")
        vals = np.empty((self.n_members, self.nnz_hess))
        capi.check(self._lib.pb2_batch_hess_lagrangian(self._h, _as_ptr(Z), _as_ptr(mu), _as_ptr(vals), capi.PB2_HOST))
        return vals

    def residual_jacobian_device(self, dZ, ddelta, dvals, stream=None):
        capi.check(self._lib.pb2_batch_residual_jacobian_async(self._h, dZ, ddelta, dvals, stream))

    def hessian_device(self, dZ, dmu, dvals, stream=None):
        capi.check(self._lib.pb2_batch_hess_lagrangian_async(self._h, dZ, dmu, dvals, stream))


class B200KnotLinearConstraints:
    """Every ``DerivativeIntegrator(x, xdot, traj)`` of a problem plus the time-consistency constraint and
    ``TimeStepsAllEqualConstraint`` (``timesteps_all_equal=True``, _problem_templates.jl:175-180),
    evaluated by one launch (smooth_pulse_problem.jl:267-277; include/piccolo_b200.h for the orders)."""

    def __init__(self, traj, pairs=(("u", "du"), ("du", "ddu")), time_consistency=True, device=0,
                 timesteps_all_equal=False):
        self._lib = capi.load_library()
        comps = traj.components
        d = capi.pb2_aux_desc()
        d.K, d.D, d.dt_off = traj.N, traj.dim, comps[traj.timestep].start
        d.t_off = comps["t"].start if (time_consistency and "t" in comps) else -1
        d.global_dim, d.n_pairs, d.device = traj.global_dim, len(pairs), device
        d.timesteps_all_equal = 1 if timesteps_all_equal else 0
        self.timesteps_all_equal = bool(timesteps_all_equal)
        if len(pairs) > capi.PB2_AUX_MAX_PAIRS:
            raise ValueError("too many derivative pairs")
        self.pairs = []
        for i, (x, xd) in enumerate(pairs):
            if len(comps[x]) != len(comps[xd]):
                raise ValueError(f"components {x} and {xd} differ in size")
            d.x_off[i], d.xdot_off[i], d.dim[i] = comps[x].start, comps[xd].start, len(comps[x])
            self.pairs.append((comps[x].start, comps[xd].start, len(comps[x])))
        self.dt_off, self.t_off = d.dt_off, (d.t_off if d.t_off >= 0 else None)
        self.K, self.D = traj.N, traj.dim
        h = ctypes.c_void_p()
        capi.check(self._lib.pb2_aux_create(ctypes.byref(d), ctypes.byref(h)))
        self._h = h
        self.dim = int(self._lib.pb2_aux_dim(h))
        self.nnz_jac = int(self._lib.pb2_aux_nnz_jac(h))
        self.nnz_hess = int(self._lib.pb2_aux_nnz_hess(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb2_aux_destroy(self._h)
            self._h = None

    __del__ = close

    def _structure(self, fn, n):
        rows, cols = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
        ip = ctypes.POINTER(ctypes.c_int64)
        capi.check(fn(self._h, rows.ctypes.data_as(ip), cols.ctypes.data_as(ip)))
        return rows, cols

    def jacobian_structure(self):
        return self._structure(self._lib.pb2_aux_structure_jac, self.nnz_jac)

    def hessian_structure(self):
        return self._structure(self._lib.pb2_aux_structure_hess, self.nnz_hess)

    def residual_jacobian(self, Z):
        Z = Z.data if isinstance(Z, NamedTrajectory) else np.asfortranarray(Z, dtype=np.float64)
        if Z.shape != (self.D, self.K):
            raise ValueError(f"trajectory is {Z.shape}, expected {(self.D, self.K)}")
        delta, vals = np.empty(self.dim), np.empty(self.nnz_jac)
        capi.check(self._lib.pb2_aux_residual_jacobian(self._h, Z.ctypes.data, delta.ctypes.data, vals.ctypes.data,
                                                       capi.PB2_HOST))
        return delta, vals

    def hessian_values(self, mu):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        if mu.size < self.nnz_hess:
            raise ValueError("mu must cover the derivative rows")
        out = np.empty(self.nnz_hess)
        capi.check(self._lib.pb2_aux_hess_lagrangian(self._h, mu.ctypes.data, out.ctypes.data, capi.PB2_HOST))
        return out

    def residual_jacobian_device(self, dZ, ddelta, dvals, stream=None):
        capi.check(self._lib.pb2_aux_residual_jacobian_async(self._h, _as_ptr(dZ), _as_ptr(ddelta), _as_ptr(dvals),
                                                             _as_ptr(stream)))


def BilinearIntegrator(qtraj, traj_or_N, traj=None, **kw):
    """BilinearIntegrator(qtraj, N) -- dispatch on the trajectory type (integrators.jl:35-95).

    ``traj`` (a NamedTrajectory with the state / Δt / u components) supplies the knot layout;
    the reference rebuilds it from ``qtraj`` and ``N``, here it is passed in."""
    if isinstance(traj_or_N, NamedTrajectory):
        traj = traj_or_N
    if kw.pop("batched", False):
        # one launch for every member / state integrator of the problem (member = a grid axis)
        if not isinstance(qtraj, (MultiKetTrajectory, MultiDensityTrajectory, SamplingTrajectory)) or traj is None:
            raise TypeError("batched=True needs a multi-state trajectory or an ensemble and its NamedTrajectory")
        if isinstance(qtraj, SamplingTrajectory):
            systems = [s_ for s_ in qtraj.systems for _ in range(qtraj.n_substates)]
        else:
            systems = [qtraj.system] * len(qtraj.state_names)
        if any(s_.time_dependent for s_ in systems):
            raise TypeError("batched=True does not take time-dependent systems")
        comps = traj.components
        return B200IntegratorBatch(
            qtraj.kind, [s_.G_parts() for s_ in systems], K=traj.N, D=traj.dim,
            x_offs=[comps[n].start for n in qtraj.state_names], dt_off=comps[traj.timestep].start,
            u_off=comps["u"].start, global_dim=traj.global_dim, names=qtraj.state_names, **kw)
    if kw.pop("fused", False):
        # every state of a MultiKetTrajectory obeys the same generator and the blocks are contiguous in
        # the knot column, so one integrator (one launch) can evaluate them all: rows are knot-major,
        # state-major inside a knot (the per-state vector below is state-major, knot-major inside)
        multi = isinstance(qtraj, (MultiKetTrajectory, MultiDensityTrajectory))
        ensemble = isinstance(qtraj, SamplingTrajectory) and qtraj.n_substates > 1
        if not (multi or ensemble) or traj is None:
            raise TypeError("fused=True needs a multi-state trajectory (or an ensemble of them) and its NamedTrajectory")
        comps = traj.components

        def fuse(names, sys_):
            blocks = [comps[n] for n in names]
            if any(b.start != a.stop for a, b in zip(blocks, blocks[1:])) or len({len(b) for b in blocks}) != 1:
                raise ValueError("state blocks must be contiguous and equally sized")
            G0, Gj = sys_.G_parts()
            return B200BilinearIntegrator(
                qtraj.kind, G0, Gj, K=traj.N, D=traj.dim, x_off=blocks[0].start, dt_off=comps[traj.timestep].start,
                u_off=comps["u"].start, x_name="+".join(names), global_dim=traj.global_dim,
                n_states=len(blocks), **sys_.td_kwargs(traj), **kw)

        if multi:
            return fuse(qtraj.state_names, qtraj.system)
        return [fuse(names, sys_) for names, sys_ in zip(qtraj.member_states(), qtraj.systems)]   # one per member
    if isinstance(qtraj, (MultiKetTrajectory, MultiDensityTrajectory, SamplingTrajectory)) and traj is not None:
        # a vector of integrators, one per state block (member-major for ensembles of multi-state
        # trajectories), all reading the same Δt / u rows
        if isinstance(qtraj, SamplingTrajectory):
            systems = [s_ for s_ in qtraj.systems for _ in range(qtraj.n_substates)]
        else:
            systems = [qtraj.system] * len(qtraj.state_names)
        comps = traj.components
        out = []
        for name, sys_ in zip(qtraj.state_names, systems):
            G0, Gj = sys_.G_parts()
            out.append(B200BilinearIntegrator(
                qtraj.kind, G0, Gj, K=traj.N, D=traj.dim, x_off=comps[name].start,
                dt_off=comps[traj.timestep].start, u_off=comps["u"].start, x_name=name,
                global_dim=traj.global_dim, **sys_.td_kwargs(traj), **kw))
        return out
    if traj is None:
        raise TypeError("pass the NamedTrajectory that defines the knot layout")
    if not isinstance(traj_or_N, NamedTrajectory) and int(traj_or_N) != traj.N:
        raise ValueError("N does not match the trajectory")
    G0, Gj = qtraj.system.G_parts()
    comps = traj.components
    x = comps[qtraj.state_name]
    return B200BilinearIntegrator(
        qtraj.kind, G0, Gj, K=traj.N, D=traj.dim, x_off=x.start, dt_off=comps[traj.timestep].start,
        u_off=comps["u"].start, x_name=qtraj.state_name, global_dim=traj.global_dim,
        **qtraj.system.td_kwargs(traj), **kw)


# free functions with the reference's names ------------------------------------------------

def evaluate_(delta, B, traj):
    """evaluate!(δ, B, traj)"""
    return B.evaluate_(delta, traj)


def jacobian_structure(B):
    return B.jacobian_structure()


def hessian_structure(B):
    return B.hessian_structure()


def eval_jacobian(B, traj):
    """Sparse ``dim x (D*N + global_dim)`` Jacobian (integrators.jl:780-782) as scipy COO."""
    import scipy.sparse as sp
    rows, cols = B.jacobian_structure()
    vals = B.jacobian_values(traj)
    n = B.D * B.K + B.global_dim
    total_rows = int(rows.max()) if B.knot0 else B.dim
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(total_rows, n))


def hessian_of_lagrangian(B, traj, mu):
    import scipy.sparse as sp
    rows, cols = B.hessian_structure()
    vals = B.hessian_values(traj, mu)
    n = B.D * B.K + B.global_dim
    return sp.coo_matrix((vals, (rows - 1, cols - 1)), shape=(n, n))


def rollout_divergence(integrators, traj):
    """rollout_divergence(qcp) over a vector of state integrators (multi-state trajectories: the 2-norm of the
    stacked difference over the 2-norm of the stacked collocation state, problems.jl:341-356)."""
    Bs = integrators if isinstance(integrators, (list, tuple)) else [integrators]
    sq_d = sq_c = 0.0
    for B in Bs:
        _, _, nd, nc = B.rollout(traj)
        sq_d += nd * nd
        sq_c += nc * nc
    return np.sqrt(sq_d) / max(np.sqrt(sq_c), 1.0)


def test_integrator(B, traj, atol=1e-3, h=1e-5, seed=0):
    """DirectTrajOpt's acceptance test for an integrator, as the reference calls it
    (``test_integrator(integrator, traj; atol = 1e-3)``, integrators.jl:359 and the dispatch test items
    after it): the analytic Jacobian and the Hessian of the Lagrangian must agree with derivatives of
    the integrator's own residual.  DirectTrajOpt differentiates with ForwardDiff; here every
    evaluation is the CUDA path itself and the derivatives are central differences of it (step ``h``),
    so the check needs no second implementation.  Raises AssertionError with the worst entry."""
    Z0 = np.asfortranarray(traj.data if isinstance(traj, NamedTrajectory) else traj, dtype=np.float64)
    D, K = Z0.shape
    n = D * K
    rows, cols = B.jacobian_structure()
    vals = B.jacobian_values(Z0)
    J = np.zeros((B.dim, n))
    np.add.at(J, (rows - 1, cols - 1), vals)

    def residual(z):
        out = np.empty(B.dim)
        B.evaluate_(out, z.reshape(D, K, order="F"))
        return out

    z0 = Z0.reshape(-1, order="F")
    touched = np.unique(cols - 1)
    Jfd = np.zeros_like(J)
    for c in touched:
        e = np.zeros(n)
        e[c] = h
        Jfd[:, c] = (residual(z0 + e) - residual(z0 - e)) / (2 * h)
    err = np.abs(J - Jfd)
    assert err.max() <= atol, f"Jacobian differs from finite differences by {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    # columns the structure does not mention must not influence the residual
    rest = np.setdiff1d(np.arange(n), touched)
    if rest.size:
        e = np.zeros(n)
        e[rest] = h
        assert np.abs(residual(z0 + e) - residual(z0)).max() <= atol * h

    mu = np.random.default_rng(seed).standard_normal(B.dim)
    hr, hc = B.hessian_structure()
    hv = B.hessian_values(Z0, mu)
    H = np.zeros((n, n))
    np.add.at(H, (hr - 1, hc - 1), hv)
    H = H + np.triu(H, 1).T                      # upper triangle -> symmetric

    def grad(z):                                 # J(z)^T mu through the analytic Jacobian values
        g = np.zeros(n)
        np.add.at(g, cols - 1, B.jacobian_values(z.reshape(D, K, order="F")) * mu[rows - 1])
        return g

    Hfd = np.zeros((n, n))
    for c in touched:
        e = np.zeros(n)
        e[c] = h
        Hfd[:, c] = (grad(z0 + e) - grad(z0 - e)) / (2 * h)
    err = np.abs(H - Hfd)
    assert err.max() <= atol * max(1.0, np.abs(mu).max()), \
        f"Hessian of the Lagrangian differs from finite differences by {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    return True


test_integrator.__test__ = False   # not a pytest item
