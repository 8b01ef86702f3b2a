/* piccolo_b200.h -- C ABI of the B200-native knot evaluator (libpiccolo_b200.so).
 *
 * Drop-in boundary.  The reference (harmoniqs/Piccolo.jl, 100 % Julia) reaches this path
 * through DirectTrajOpt's AbstractIntegrator interface: a template receives an
 * `integrator=` object (src/control/templates/smooth_pulse_problem.jl:123,213-233), built
 * by `BilinearIntegrator(qtraj, N)` (src/control/integrators.jl:35-95), and Ipopt's
 * eval_g / eval_jac_g / eval_h callbacks call, per integrator,
 *     evaluate!(delta, B, traj)                     integrators.jl:311, display/inspect.jl:623
 *     eval_jacobian(B, traj) / jacobian_structure   integrators.jl:780-782, test/aqua.jl:6-9
 *     hessian_of_lagrangian / hessian_structure     test/aqua.jl:6-9, spline_pulse_problem.jl:96
 * There is no FFI in the reference; these entry points are what a Julia `ccall` shim
 * (julia/PiccoloB200.jl) binds, one per callback above.
 *
 * Conventions
 *  - every function returns 0 on success, a PB2_E* code otherwise; the message is in
 *    pb2_last_error() (thread-local).  Nothing throws or aborts across the boundary.
 *  - all reals are IEEE float64; all matrices dense column-major (Julia layout);
 *    index arrays are 1-based int64 (Julia / MOI layout).
 *  - Z is NamedTrajectory.datavec: D x K column-major, one knot per column
 *    (src/quantum/trajectories/named_trajectory_conversion.jl:316-321).  Read-only.
 *  - caller owns every array; the handle owns device workspaces and pinned staging only.
 *  - a handle is not re-entrant (Ipopt is serial); distinct handles are independent.
 *  - there is NO CPU fallback: without a usable CUDA device pb2_create fails.
 */
#ifndef PICCOLO_B200_H
#define PICCOLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_VERSION 100

enum {
  PB2_OK = 0,
  PB2_EINVAL = 1,      /* bad descriptor / null pointer / unsupported size */
  PB2_ECUDA = 2,       /* CUDA runtime error (text in pb2_last_error) */
  PB2_ENODEVICE = 3,   /* no CUDA device: the library refuses to run */
  PB2_ENOMEM = 4
};

/* state kinds: which BilinearIntegrator dispatch the handle replaces */
enum {
  PB2_KET = 0,      /* integrators.jl:58-74    Ghat = G(u),        x = [Re psi; Im psi]     */
  PB2_UNITARY = 1,  /* integrators.jl:35-51    Ghat = I_d (x) G(u), x = operator_to_iso_vec  */
  PB2_DENSITY = 2   /* integrators.jl:82-95    Ghat = compact Lindbladian (non-normal)       */
};

/* where Z / output pointers live */
enum {
  PB2_HOST = 0,   /* host memory: library stages H2D / D2H on its stream, call is synchronous */
  PB2_DEVICE = 1  /* device memory on desc.device (e.g. CUDA.jl CuArray): no copies          */
};

enum {
  PB2_ALG_AUTO = 0,     /* tensor-core path when the generator qualifies, else generic          */
  PB2_ALG_GENERIC = 1,  /* scaling-and-squaring Taylor jets in shared memory; any real generator */
  PB2_ALG_DMMA = 2      /* FP64 tensor-core Taylor action, one warp per knot: b <= 16, drive
                           generators with <= 4 nonzeros per row                                */
};

typedef struct pb2_desc {
  int32_t kind;        /* PB2_KET | PB2_UNITARY | PB2_DENSITY */
  int32_t b;           /* generator block size: 2d (ket, unitary) or d^2 (density) */
  int32_t n_b;         /* contiguous state blocks sharing the generator: d for unitary; 1 for one ket /
                          density, or the number of states of a MultiKetTrajectory / MultiDensityTrajectory
                          (one system, src/control/integrators.jl:102-117) fused into one integrator */
  int32_t m;           /* number of (linear) drives */
  int32_t K;           /* knot columns in the Z passed to this handle -> K-1 constraints */
  int32_t D;           /* reals per knot (traj.dim) */
  int32_t x_off;       /* 0-based row of the state block inside a knot column */
  int32_t dt_off;      /* 0-based row of the timestep */
  int32_t u_off;       /* 0-based row of the first drive amplitude */
  int32_t global_dim;  /* trailing global variables in the NLP primal (structure only) */
  int64_t knot0;       /* global index of this handle's first knot (sharded use; else 0) */
  int32_t device;      /* CUDA device ordinal */
  int32_t algorithm;   /* PB2_ALG_* */
  const double* G0;    /* host, b*b column-major: drift generator (dissipators folded in) */
  const double* Gj;    /* host, m blocks of b*b column-major: drive generators */
  int32_t t_off;       /* time-dependent handles: 0-based row of the knot time t (else ignored) */
  int32_t time_dependent; /* != 0: Ghat(u, t) = G0 + sum_j c_j(t) u_j G_j  (ModulatedDrive over LinearDrive,
                          src/quantum/systems/drives.jl:342-388; TimeDependentBilinearIntegrator,
                          src/control/integrators.jl:38-46).  The Jacobian gains one column per knot, d/d t_k
                          (the LAST n_x values of the knot's segment, col = k D + t_off); the coefficients
                          c_j(t_k), c_j'(t_k) are supplied by pb2_set_time_coefficients before each evaluation.
                          The Lagrangian Hessian and the compact / exchange forms are not available. */
  int32_t dense_blocks;   /* != 0: the d/dx_k block of every knot is emitted as the FULL n_x x n_x block, column-major,
                          structural zeros included (n_x^2 instead of n_b b^2 entries per knot): the layout of an
                          integrator that does not know the I (x) E structure of the unitary case (SURVEY.md 8b and
                          Appendix B: what DirectTrajOpt's BilinearIntegrator is believed to scatter).  A pure layout
                          option: same kernels, one more streaming pass; compact / exchange forms unavailable. */
} pb2_desc;

typedef struct pb2_handle pb2_handle;

int pb2_version(void);
const char* pb2_last_error(void);
int pb2_device_count(void);

int pb2_create(const pb2_desc* desc, pb2_handle** out);
void pb2_destroy(pb2_handle* h);

/* sizes: dim = n_x (K-1)  (integrator.dim, integrators.jl:309) */
int64_t pb2_dim(const pb2_handle* h);
int64_t pb2_nnz_jac(const pb2_handle* h);
int64_t pb2_nnz_hess(const pb2_handle* h);
int32_t pb2_algorithm(const pb2_handle* h); /* the path actually selected */
/* kernel behind pb2_hess_lagrangian: 0 none (time-dependent handle), 1 shared-memory jet kernel, 2 general tensor-core
 * kernel (b <= 16, <= 12 column tiles), 3 the 3-qubit tensor-core kernel (16-byte aligned device pointers) */
int32_t pb2_hessian_algorithm(const pb2_handle* h);

/* COO structure, 1-based, deterministic, computed on the host (no device work).
 * Order documented in DESIGN.md ("canonical COO order"); rows of the Jacobian are
 * (knot0 + k) n_x + i, columns index the NLP primal [vec(Z); globals]. */
int pb2_structure_jac(const pb2_handle* h, int64_t* rows, int64_t* cols);
int pb2_structure_hess(const pb2_handle* h, int64_t* rows, int64_t* cols);

/* evaluate!(delta, B, traj) */
int pb2_residual(pb2_handle* h, const double* Z, double* delta, int space);
/* eval_jacobian values in pb2_structure_jac order */
int pb2_jacobian(pb2_handle* h, const double* Z, double* vals, int space);
/* fused residual + Jacobian (one kernel): the benchmarked entry */
int pb2_residual_jacobian(pb2_handle* h, const double* Z, double* delta, double* vals, int space);
/* Hessian of sum_k mu_k . delta_k, upper triangle, in pb2_structure_hess order */
int pb2_hess_lagrangian(pb2_handle* h, const double* Z, const double* mu, double* vals, int space);

/* Device-pointer, asynchronous forms: enqueue on `stream` (a cudaStream_t, used exactly as
 * given: NULL is CUDA's default stream, as everywhere in the runtime API) and return without
 * synchronizing.  pb2_stream() returns the handle's own non-blocking stream (the one the
 * host-pointer calls use); pb2_sync() waits for it.  delta / vals may be NULL to skip that
 * output. */
int pb2_residual_jacobian_async(pb2_handle* h, const double* dZ, double* ddelta, double* dvals,
                                void* stream);
int pb2_hess_lagrangian_async(pb2_handle* h, const double* dZ, const double* dmu, double* dvals,
                              void* stream);
/* Sharded (multi-GPU) use: the d/dx_k block of a unitary knot is I (x) E -- n_b copies of one b x b
 * block whose second half of columns mirrors the first (E = [[P,-Q],[Q,P]]) -- and the d/dx_{k+1}
 * entries are constant, so what has to cross NVLink (or PCIe) per knot is a COMPACT record
 * [E columns 0..b/2-1 (b*b/2) | jets, d/d dt ((m+1) n_x) | delta (n_x)] of pb2_compact_stride()
 * doubles (0: this handle cannot produce it).  Each rank writes its records, one all-gather moves
 * them, and pb2_expand_compact_async turns any number of gathered records into the canonical delta /
 * value arrays (pb2_structure_jac order). */
int64_t pb2_compact_stride(const pb2_handle* h);
int pb2_residual_jacobian_compact_async(pb2_handle* h, const double* dZ, double* dcompact, void* stream);
int pb2_expand_compact_async(pb2_handle* h, const double* dcompact, int64_t n_knots, double* ddelta,
                             double* dvals, void* stream);
/* Fused compute + exchange for sharded runs on one NVSwitch domain: gather_bufs[r] is rank r's gather
 * buffer (of every rank's compact records) mapped into THIS process (CUDA IPC / symmetric memory;
 * gather_bufs[rank] is the local one).  The kernel writes each finished knot's record into all
 * n_ranks buffers at slot_offset (+ knot * stride) doubles -- bulk stores over NVLink that overlap the
 * math of the following knots -- so no separate all-gather runs; the caller only needs a barrier
 * across ranks before expanding / reading. */
int pb2_residual_jacobian_exchange_async(pb2_handle* h, const double* dZ, int32_t n_ranks, int32_t rank,
                                         double* const* gather_bufs, int64_t slot_offset, void* stream);
/* The same, closing the step inside the kernel: flag_offset (doubles, 16-byte aligned, beyond the records) locates
 * 16 64-bit words inside EVERY gather buffer (8 arrival words and two counters; zero-initialised by the caller once).  The knot whose record
 * leaves last writes this rank's word on every peer (release, system scope) and waits for every peer's word here,
 * so when the call's kernel completes all ranks' records of this step are in gather_bufs[rank]: no barrier kernel
 * follows.  Every rank must use its gather buffers in the same order (the words carry a per-buffer use counter);
 * consecutive steps must use different gather buffers (a rank may run a step ahead of a slower one), and two calls
 * on the same buffer must not overlap. */
int pb2_residual_jacobian_exchange_sync_async(pb2_handle* h, const double* dZ, int32_t n_ranks, int32_t rank,
                                              double* const* gather_bufs, int64_t slot_offset, int64_t flag_offset,
                                              void* stream);
/* cudaDeviceEnablePeerAccess(device -> peer) (idempotent); needed once before kernels on `device`
 * write into memory that lives on `peer` */
int pb2_enable_peer_access(int32_t device, int32_t peer);
/* Gather buffers without any other CUDA binding, so a host in the reference's language (Julia, one process per
 * GPU under MPI.jl or Distributed) can shard a trajectory with this library alone:
 *   every rank:  pb2_device_alloc(&buf, bytes, device)  (zero-filled);  pb2_ipc_export(buf, handle)
 *   exchange the 64-byte handles through the host's own channel (MPI_Allgather, a socket, a file)
 *   every rank:  pb2_ipc_open(handle_of_rank_r, device, &peer_buf[r])  for r != rank  (peer access is enabled
 *                lazily by the driver);  gather_bufs = { peer_buf[0], .., buf, .., peer_buf[n-1] }
 * and then pb2_residual_jacobian_exchange_sync_async as above.  A single process driving several GPUs needs only
 * pb2_device_alloc per device and pb2_enable_peer_access per ordered pair.  pb2_device_copy is cudaMemcpy with
 * cudaMemcpyDefault (host<->device, device<->device), for reading the gathered records back. */
#define PB2_IPC_HANDLE_BYTES 64
int pb2_device_alloc(void** ptr, int64_t bytes, int32_t device);
int pb2_device_free(void* ptr);
int pb2_device_copy(void* dst, const void* src, int64_t bytes);
int pb2_ipc_export(const void* dptr, void* handle64);
int pb2_ipc_open(const void* handle64, int32_t device, void** dptr);
int pb2_ipc_close(void* dptr);
void* pb2_stream(const pb2_handle* h);
int pb2_sync(pb2_handle* h);

/* Per-handle options (all default 0): promises of the caller about the *_async entry points.  The kernels are
 * launched with programmatic stream serialization, i.e. their CTAs may be scheduled while the kernel enqueued
 * immediately before on the same stream is still draining; by default they then wait for that kernel's completion
 * (`griddepcontrol.wait`) before touching the trajectory or the outputs, which makes a call behave like any
 * stream-ordered launch.
 *  PB2_OPT_EARLY_Z    dZ is complete before the kernel enqueued immediately before this call STARTS (true whenever Z
 *                     arrives by a copy -- the Ipopt callback case -- and for back-to-back evaluator calls).  The
 *                     kernels then load Z, build G(u_k) and run their products before the dependency wait; outputs
 *                     are still written only after it.
 *  PB2_OPT_PIPELINED  in addition, the kernel enqueued immediately before does not ACCESS this call's output buffers
 *                     (e.g. consecutive evaluations into different buffers, or a preceding copy): no dependency wait
 *                     at all, so consecutive calls overlap freely (one grid's tail under the next one's products) and
 *                     need not complete in order.  Later stream operations are ordered after both as usual.
 * The host-pointer entry points always run in the second mode: the library itself copies Z to, and the results from,
 * its own buffers on the handle's stream.
 *  PB2_OPT_HESSIAN_CTAS  value = how many SMs the persistent 3-qubit Hessian kernel may occupy (0: all).  That kernel
 *                     needs a whole SM per CTA, so nothing else runs beside it; a caller that enqueues the other
 *                     callbacks of the same NLP iterate on other streams (they are independent) leaves them some SMs
 *                     this way: C3, 116 of 148 SMs for the Hessian -> the iterate costs about the Hessian alone. */
enum { PB2_OPT_EARLY_Z = 1, PB2_OPT_PIPELINED = 2, PB2_OPT_HESSIAN_CTAS = 3 };
int pb2_set_option(pb2_handle* h, int32_t option, int64_t value);

/* Time-dependent handles: the modulation values at the trajectory's CURRENT time row, c[j + m k] = c_j(t_k) and
 * cdot[j + m k] = c_j'(t_k), j < m, k < K (the host evaluates the closures: m x K numbers per callback).  They
 * stay in force until the next call.  space = PB2_HOST or PB2_DEVICE. */
int pb2_set_time_coefficients(pb2_handle* h, const double* c, const double* cdot, int space);

/* ---- ensembles: all members in one launch --------------------------------------------
 * SamplingTrajectory / MultiKet / MultiDensity problems attach one integrator per ensemble member and state
 * (src/control/integrators.jl:102-117, 134-226; sampling_problem.jl:389-395): the same knot mathematics with the
 * member's own generator on the member's own state block, all reading the same dt / u rows.  For the small systems
 * these problems are built from, a launch per member is pure launch latency; a batch evaluates them together
 * (member = blockIdx.y of one grid, per-member generator tables).
 * descs[i] must agree in kind, b, n_b, m, K, D, dt_off, u_off, global_dim, device and algorithm; x_off, G0, Gj differ.
 * Outputs are member-major: delta [n_members][dim], vals [n_members][nnz_jac], Hessian values likewise; member i's
 * COO structure is pb2_batch_structure_*(i) (rows local to the member, like a single handle's).
 * Members whose kernels cannot share a launch (e.g. different sparsity widths, or the 3-qubit shape with its own
 * kernels) are evaluated concurrently on the batch's own streams instead; results are identical. */
typedef struct pb2_batch pb2_batch;
int pb2_batch_create(const pb2_desc* descs, int32_t n_members, pb2_batch** out);
void pb2_batch_destroy(pb2_batch* b);
int32_t pb2_batch_size(const pb2_batch* b);
int32_t pb2_batch_fused(const pb2_batch* b);      /* 1: one launch for all members */
int64_t pb2_batch_dim(const pb2_batch* b);        /* per member */
int64_t pb2_batch_nnz_jac(const pb2_batch* b);
int64_t pb2_batch_nnz_hess(const pb2_batch* b);
int pb2_batch_structure_jac(const pb2_batch* b, int32_t member, int64_t* rows, int64_t* cols);
int pb2_batch_structure_hess(const pb2_batch* b, int32_t member, int64_t* rows, int64_t* cols);
int pb2_batch_residual_jacobian(pb2_batch* b, const double* Z, double* delta, double* vals, int space);
int pb2_batch_hess_lagrangian(pb2_batch* b, const double* Z, const double* mu, double* vals, int space);
int pb2_batch_residual_jacobian_async(pb2_batch* b, const double* dZ, double* ddelta, double* dvals, void* stream);
int pb2_batch_hess_lagrangian_async(pb2_batch* b, const double* dZ, const double* dmu, double* dvals, void* stream);

/* ---- rollout of the trajectory's piecewise-constant controls -----------------------
 * x_1 = x0 (NULL: the first state column of Z), x_{k+1} = exp(dt_k Ghat(u_k)) x_k: what `rollout!(qtraj, pulse)`
 * (src/quantum/trajectories/rollouts_extensions.jl:46-92) produces at the knot times for the zero-order-hold pulse
 * that `sync_trajectory!` (src/control/problems.jl:186-208) extracts from the optimizer's trajectory -- the reference
 * integrates the ODE adaptively (abstol = reltol = 1e-8), here every interval is propagated exactly by the E_k the
 * knot kernels compute.  states: n_x * K doubles, column k = state at knot k (may be NULL).
 * out3: [0] rollout_divergence (problems.jl:336-356) = ||x_K^rollout - x_K^collocation||_2 / max(||x_K^collocation||_2, 1),
 * [1] the numerator, [2] ||x_K^collocation||_2 (so a caller can stack several state components as the reference
 * does for multi-state trajectories); may be NULL.  Terminal fidelities of the rolled-out state are the objective
 * terms of pb2_obj_* evaluated on `states`. */
int pb2_rollout(pb2_handle* h, const double* Z, const double* x0, double* states, double* out3, int space);
int pb2_rollout_async(pb2_handle* h, const double* dZ, const double* dx0, double* dstates, double* dout3, void* stream);

/* ---- linear knot constraints evaluated in the same callbacks -----------------
 * DerivativeIntegrator(x, xdot, traj): r_k = x_{k+1} - x_k - dt_k xdot_k   (u -> du, du -> ddu;
 *   src/control/templates/smooth_pulse_problem.jl:267-275, spline_pulse_problem.jl:363-366)
 * time consistency: t_{k+1} - t_k - dt_k   (applied by DirectTrajOpt when :t and :dt exist,
 *   smooth_pulse_problem.jl:277).
 * TimeStepsAllEqualConstraint: dt_{k+1} - dt_k, k = 1 .. K-1   (pushed by the templates when
 *   piccolo_options.timesteps_all_equal is set: src/control/templates/_problem_templates.jl:175-180; the
 *   constraint type lives in DirectTrajOpt, source absent -- row form "parity unpinned", the reference's
 *   solutions satisfy it exactly).
 * One handle evaluates every pair of a trajectory in one launch.  Row order: pair-major (the order
 * given), knot-major inside a pair, component fastest; time rows last.  Jacobian values per
 * derivative row: d x_k[i] (-1), d xdot_k[i] (-dt_k), d dt_k (-xdot_k[i]), d x_{k+1}[i] (+1); per
 * time row: d t_k (-1), d dt_k (-1), d t_{k+1} (+1); per equal-timestep row (they come last): d dt_k (-1),
 * d dt_{k+1} (+1).  Hessian of sum mu.r: one entry per derivative
 * row, (xdot_k[i], dt_k) = -mu (upper triangle).  Rows / columns 1-based like pb2_structure_*;
 * rows are local to this handle (the caller offsets them into the NLP's constraint vector). */
#define PB2_AUX_MAX_PAIRS 8
typedef struct pb2_aux_desc {
  int32_t K, D, dt_off;
  int32_t t_off;                       /* < 0: no time-consistency rows */
  int32_t global_dim;
  int32_t n_pairs;
  int32_t x_off[PB2_AUX_MAX_PAIRS], xdot_off[PB2_AUX_MAX_PAIRS], dim[PB2_AUX_MAX_PAIRS];
  int32_t device;
  int32_t timesteps_all_equal;         /* != 0: K-1 rows dt_{k+1} - dt_k after the time rows */
} pb2_aux_desc;
typedef struct pb2_aux pb2_aux;
int pb2_aux_create(const pb2_aux_desc* desc, pb2_aux** out);
void pb2_aux_destroy(pb2_aux* h);
int64_t pb2_aux_dim(const pb2_aux* h);
int64_t pb2_aux_nnz_jac(const pb2_aux* h);
int64_t pb2_aux_nnz_hess(const pb2_aux* h);
int pb2_aux_structure_jac(const pb2_aux* h, int64_t* rows, int64_t* cols);
int pb2_aux_structure_hess(const pb2_aux* h, int64_t* rows, int64_t* cols);
/* delta / vals may be NULL to skip that output */
int pb2_aux_residual_jacobian(pb2_aux* h, const double* Z, double* delta, double* vals, int space);
int pb2_aux_hess_lagrangian(pb2_aux* h, const double* mu, double* vals, int space);
int pb2_aux_residual_jacobian_async(pb2_aux* h, const double* dZ, double* ddelta, double* dvals, void* stream);

/* ---- objective value + gradient on the device-resident trajectory -------------
 * J(Z) = sum of terms + sum of quadratic regularizers, with its dense gradient over the K*D
 * trajectory entries (the caller appends zeros for global variables).  Replaces the eval_f /
 * eval_grad_f callbacks DirectTrajOpt builds from the objectives Piccolo's templates assemble
 * (src/control/templates/smooth_pulse_problem.jl:240-250).
 *
 * A term is a loss of one knot column z (restricted to `rows`), in real form
 *     F = scale * ((a_re.z)^2 + (a_im.z)^2 + sum_i a_sq[i] z_i^2) + a_lin.z,
 *     loss = Q*|1 - F|  (PB2_OBJ_ONE_MINUS)  or  Q*F,
 * evaluated at the terminal knot (n_times = 0) or at `times` with weights Q[t].  It covers
 *   KetInfidelityObjective              src/control/objectives.jl:24-38, 56-64   (a_re, a_im from the goal ket)
 *   CoherentKetInfidelityObjective      :96-121, 181-216   (rows = all state blocks, weights folded into a_*)
 *   UnitaryInfidelityObjective          :330-337, 347-356  (scale = 1/n^2)
 *     ... with an EmbeddedOperator goal :339-345           (a_sq = subspace mask, scale = 1/(n(n+1)); unitary goal)
 *   DensityMatrix[PureState]InfidelityObjective :387-394, 412-419  (a_lin; F is linear in the compact iso)
 *   LeakageObjective                    :464-474           (a_sq = 1/len on the leakage rows, Q*F at `times`)
 * The free-phase objectives (:283-324, :358-383) take a Julia closure over global variables and the
 * quartic UnitarySensitivityObjective (:447-458) are not expressible here and stay in the reference.
 * d|x|/dx is +1 at x = +0 (ForwardDiff's rule).
 *
 * A regularizer is  1/2 sum_{k in times} sum_i R[i] (z_k[rows[i]] - baseline[i,k])^2 dt_k^dt_power,
 * dt_power in {0,1,2}.  DirectTrajOpt's QuadraticRegularizer source is not part of the reference tree
 * and no reference test pins its value, so the power is the caller's to choose (parity unpinned). */
#define PB2_OBJ_ONE_MINUS 1
typedef struct pb2_obj_term {
  int32_t flags;
  int32_t n_rows;
  const int32_t* rows;          /* 0-based rows of the knot column, distinct */
  const double *a_re, *a_im, *a_sq, *a_lin;   /* n_rows each; NULL = all zero */
  double scale;
  int32_t n_times;              /* 0: the terminal knot, weight Q[0] */
  const int32_t* times;         /* 0-based knots */
  const double* Q;              /* max(n_times, 1) weights */
} pb2_obj_term;
typedef struct pb2_obj_reg {
  int32_t n_rows;
  const int32_t* rows;          /* distinct */
  const double* R;              /* n_rows */
  const double* baseline;       /* n_rows x K column-major, or NULL */
  int32_t dt_power;
  int32_t n_times;              /* 0: every knot */
  const int32_t* times;
} pb2_obj_reg;
typedef struct pb2_obj_desc {
  int32_t K, D, dt_off;
  int32_t n_terms, n_regs;
  const pb2_obj_term* terms;
  const pb2_obj_reg* regs;
  int32_t device;
} pb2_obj_desc;
typedef struct pb2_obj pb2_obj;
int pb2_obj_create(const pb2_obj_desc* desc, pb2_obj** out);   /* copies everything it needs */
void pb2_obj_destroy(pb2_obj* h);
/* J -> *J; gradient (K*D doubles) -> grad unless NULL.  space = PB2_HOST or PB2_DEVICE for Z, J, grad.
 * J is summed in knot order (bitwise reproducible).  A handle owns per-launch scratch: calls on one
 * handle must be ordered on one stream (use one handle per stream for concurrent evaluations). */
int pb2_obj_value_gradient(pb2_obj* h, const double* Z, double* J, double* grad, int space);
int pb2_obj_value_gradient_async(pb2_obj* h, const double* dZ, double* dJ, double* dgrad, void* stream);
/* Hessian of sigma * J (Ipopt's obj_factor; eval_h adds it to the constraint Hessians of pb2_hess_lagrangian and
 * pb2_aux_hess_lagrangian), upper triangle over the K*D trajectory entries, rows / columns 1-based, COO with
 * duplicates allowed (Ipopt sums them).  Order: knot-major; inside a knot the term instances (dense term: the
 * upper triangle of its rows x rows block, column-major; a_sq-only term: its diagonal; a_lin-only: nothing), then
 * the active regularizers (diagonal, and for dt_power > 0 the (z_i, dt) entries and one (dt, dt)).
 * |1 - F| is differentiated away from its kink: d2 = -sign(1 - F) Q d2F.  DirectTrajOpt's own layout of the
 * objective Hessian is not in the reference tree (parity unpinned); the values are pinned by the oracle's complex
 * form and central differences of its gradient (tests/test_oracle.py). */
int64_t pb2_obj_nnz_hess(const pb2_obj* h);
int pb2_obj_structure_hess(const pb2_obj* h, int64_t* rows, int64_t* cols);
int pb2_obj_hessian(pb2_obj* h, const double* Z, double sigma, double* vals, int space);
int pb2_obj_hessian_async(pb2_obj* h, const double* dZ, double sigma, double* dvals, void* stream);

/* pinned host memory for callers that want DMA without the staging copy */
int pb2_host_alloc(void** ptr, int64_t bytes);
int pb2_host_free(void* ptr);

/* launch accounting for benchmarks: kernels launched by this handle since creation */
int64_t pb2_launch_count(const pb2_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* PICCOLO_B200_H */
