// C ABI of libpiccolo_b200.so (declared in include/piccolo_b200.h).
//
// Host side of the drop-in boundary: owns the handle (device copies of the generator
// factors, device/pinned staging for host-pointer calls, launch configuration) and turns
// each DirectTrajOpt callback -- evaluate! / eval_jacobian / hessian_of_lagrangian, see the
// header for the reference call sites -- into one kernel launch.  No CPU compute path exists
// in this library: if CUDA is unavailable pb2_create fails with PB2_ENODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <functional>
#include <vector>

#include "../../include/piccolo_b200.h"
#include "knot_generic.cuh"
#include "knot_dmma.cuh"
#include "knot_u8.cuh"
#include "knot_u8h.cuh"
#include "knot_dmmah.cuh"
#include "knot_dmmaq.cuh"
#include "knot_aux.cuh"
#include "knot_objective.cuh"
#include "knot_rollout.cuh"
#include "knot_td.cuh"
#include "knot_u8p.cuh"
#include "knot_u8q.cuh"
#include "host_pool.h"
#include <emmintrin.h>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// Every entry point works on its handle's device and leaves the caller's current device as it found it
// (a host framework -- torch, CUDA.jl -- driving several GPUs from one process relies on it).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define PB2_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(PB2_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

constexpr int kNT = 256;
constexpr size_t kSmemLimit = 227 * 1024;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION (per device), not of a handle: it is
// always raised to everything the device allows next to the kernel's static shared memory, so that a second live
// handle with a smaller problem can never lower the cap under an earlier handle's launches.
template <class K>
cudaError_t raise_dynamic_smem(K kernel) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  const size_t room = kSmemLimit > fa.sharedSizeBytes ? kSmemLimit - fa.sharedSizeBytes : 0;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room);
}

struct LaunchCfg {
  int GS = 0, KPC = 0, gj_in_smem = 0;
  size_t smem = 0;
};

bool plan_generic(int order, int b, int n_b, int m, LaunchCfg& cfg) {
  const int bb = b * b;
  const int npair = order == 2 ? m * (m + 1) / 2 : 0;
  const int work = (1 + m + npair) * bb;
  int GS = std::min(kNT, std::max(32, (work + 31) / 32 * 32));
  // prefer more knots per block for tiny generators, bounded by shared memory
  for (int gj = 1; gj >= 0; --gj) {
    for (int KPC = kNT / GS; KPC >= 1; KPC /= 2) {
      size_t smem = pb2::generic_smem_bytes(order, b, n_b, m, GS, KPC, gj != 0);
      if (smem + 64 <= kSmemLimit) {   // (the kernels hold a few bytes of static shared memory as well)
        cfg.GS = GS;
        cfg.KPC = KPC;
        cfg.gj_in_smem = gj;
        cfg.smem = smem;
        return true;
      }
    }
  }
  return false;
}

double theta_bound(int q) {
  double lo = 0.0, hi = q + 1.0;
  const double target = std::log(std::ldexp(1.0, -53));
  for (int it = 0; it < 200; ++it) {
    double th = 0.5 * (lo + hi);
    double lg = (q + 1) * std::log(th) - std::lgamma(q + 2.0) - std::log1p(-th / (q + 2.0));
    if (lg <= target) lo = th; else hi = th;
  }
  return lo;
}

}  // namespace

// dense_blocks: canonical values [n_b x (b x b) | rest] -> [n_x x n_x dense block, zeros included | rest]; one CTA
// per knot (grid-stride).  Pure data movement (HBM-bound).
__global__ void __launch_bounds__(256) dense_blocks_kernel(const double* __restrict__ canon, double* __restrict__ out,
                                                           long long n_knots, int b, int n_b, int nnz_canon, int nnz_dense) {
  const int n_x = b * n_b, bb = b * b, rest = nnz_canon - n_b * bb;
  for (long long k = blockIdx.x; k < n_knots; k += gridDim.x) {
    const double* src = canon + k * nnz_canon;
    double* dst = out + k * nnz_dense;
    for (int e = threadIdx.x; e < n_x * n_x; e += blockDim.x) {
      const int col = e / n_x, row = e - col * n_x, cb = col / b, rb = row / b;
      dst[e] = cb == rb ? src[cb * bb + (col - cb * b) * b + (row - rb * b)] : 0.0;
    }
    for (int e = threadIdx.x; e < rest; e += blockDim.x) dst[n_x * n_x + e] = src[n_b * bb + e];
  }
}

// One launch for all members of an ensemble (pb2_batch_*): member-dependent arrays and strides (blockIdx.y = member)
struct BatchLaunch {
  int n = 0, max_xoff = 0;
  const int* x_offs = nullptr;
  const double *G0 = nullptr, *Gj = nullptr, *gfrag = nullptr, *norms = nullptr;
  const pb2::EllEntry* ell = nullptr;
  long long mem_G0 = 0, mem_Gj = 0, mem_gfrag = 0, mem_ell = 0, mem_norms = 0, mem_delta = 0, mem_jac = 0, mem_hess = 0;
  // tensor-core Hessian tables (knot_dmmah.cuh); null: the members run the jet kernel
  const double *hgfrag = nullptr, *hgfragT = nullptr, *hnorms = nullptr;
  const pb2::EllEntry *hell = nullptr, *hellT = nullptr;
  long long mem_hgfrag = 0, mem_hell = 0, mem_hnorms = 0;
};

struct pb2_handle {
  pb2_desc d{};
  const BatchLaunch* batch = nullptr;   // set around a batched launch on the ensemble's first member
  std::vector<double> G0, Gj;
  int alg = PB2_ALG_GENERIC;
  cudaStream_t stream = nullptr;
  double *dG0 = nullptr, *dGj = nullptr;
  pb2::DmmaPlan plan;          // tensor-core path tables (host copy) and their device mirrors
  double* dGfrag = nullptr;
  pb2::DmmahPlan hplan;        // tensor-core Hessian (general b <= 24): tables and their device mirrors
  double *dHGfrag = nullptr, *dHGfragT = nullptr, *dHNorms = nullptr;
  pb2::EllEntry *dHEll = nullptr, *dHEllT = nullptr;
  int dmmah = 1;               // PB2_NO_DMMAH=1: Hessian through the jet kernel
  int hess_prefer_dmmah = 0;   // PB2_HESS_DMMAH=1: the general tensor-core Hessian also for the 3-qubit shape
  int dmmaq = 1;               // residual + Jacobian of the general shapes as one small CTA per knot (PB2_DMMAQ=0:
                               // the persistent pipelined kernel knot_dmma)
  double* dNorms = nullptr;
  double* dTab = nullptr;
  double *dComp = nullptr, *hComp = nullptr;   // compact records: device buffer and pinned landing zone
  double *dCoef = nullptr, *dCoefDot = nullptr, *dZs = nullptr;   // time-dependent handles: c, c', scaled trajectory
  bool coef_set = false;
  double *dRoJac = nullptr, *dRoStates = nullptr, *dRoX0 = nullptr, *dRoOut = nullptr;   // rollout scratch (lazy)
  cudaEvent_t chunk_ev[16] = {};
  // host-pointer pipeline: upload, kernel and download of a callback run chunk by chunk on three streams
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t in_ev[16] = {}, k_ev[16] = {};
  int64_t nk_sub = -1;         // >= 0: this launch evaluates a sub-range of the knots (pointers already offset)
  double* dTables = nullptr;   // [gfrag | norms (even) | theta | 1/k!] contiguous, the u8 kernels' smem order
  long long* dTrace = nullptr;
  long long* dTrace2 = nullptr;
  long long* dTrace3 = nullptr;   // debug build: [launch 64][block 148][warp 16][8] stamps of the single-round kernels
  int trace_launch = 0;
  int n_sm = 148, gpc_default = 3, gpc_override = 0;
  bool u8_ok = false;
  bool u8h_ok = false;
  int stagger = 5000;   // cycles; PB2_STAGGER overrides (0 = off)
  int pdl = 1;
  int stagger_g = 0;
  int direct_last = 1;
  int u8p = 1;             // single-round kernel for <= 7 knots per SM (knot_u8p.cuh); PB2_U8P=0 disables it
  int u8q = 2;             // small-CTA kernel (knot_u8q.cuh): 2 = every size, 1 = at most 7 knots per SM, 0 = off (PB2_U8Q)
  int u8q_ns = 1;          // knots per CTA: 1 (eight 64-thread CTAs per SM, the default), 2 or 4; PB2_U8Q_NS
  unsigned long long* dSyncWords = nullptr;   // [0] ticket, [1] exchange epoch (knot_u8q.cuh, in-kernel step barrier)
  double* dTablesP = nullptr;   // knot_u8p's table blob (u8p_tables)
  double* dTablesQ = nullptr;   // knot_u8q's table blob (u8q_tables)
  bool u8p_ok = false, u8p_unit = false;
  double u8p_cj[4] = {1.0, 1.0, 1.0, 1.0};
  long long xchg_flag_off = -1;   // set around an exchange_sync launch
  int early_z = 0;         // PB2_OPT_EARLY_Z: device-pointer calls may read Z before the programmatic dependency wait
  int hess_ctas = 0;       // PB2_OPT_HESSIAN_CTAS: SMs the persistent Hessian kernel may take (0: all)
  int pipelined = 0;       // PB2_OPT_PIPELINED: no dependency wait at all (outputs not shared with the preceding kernel)
  pb2::EllEntry* dEll = nullptr;
  // staging for host-pointer calls
  double *dZ = nullptr, *dDelta = nullptr, *dJac = nullptr, *dMu = nullptr, *dHess = nullptr;
  double *hZ = nullptr, *hDelta = nullptr, *hJac = nullptr, *hMu = nullptr, *hHess = nullptr;
  LaunchCfg cfg1, cfg2;
  int64_t launches = 0;

  int n_x() const { return d.b * d.n_b; }
  int64_t nk() const { return nk_sub >= 0 ? nk_sub : (int64_t)d.K - 1; }
  // canonical values per knot (what the kernels write) and what the caller sees (dense_blocks: full n_x x n_x block)
  int nnz_jac_knot() const { return d.n_b * d.b * d.b + n_x() * d.m + 2 * n_x() + (d.time_dependent ? n_x() : 0); }
  int nnz_jac_user_knot() const { return d.dense_blocks ? nnz_jac_knot() - d.n_b * d.b * d.b + n_x() * n_x() : nnz_jac_knot(); }
  double* dCanon = nullptr;    // dense_blocks: the kernels' canonical values before re-layout
  int nnz_hess_knot() const { return n_x() * d.m + n_x() + d.m * (d.m + 1) / 2 + d.m + 1; }
};

namespace {

pb2::KnotParams make_params(const pb2_handle* h) {
  pb2::KnotParams p{};
  p.b = h->d.b; p.n_b = h->d.n_b; p.m = h->d.m; p.K = h->d.K; p.D = h->d.D;
  p.x_off = h->d.x_off; p.dt_off = h->d.dt_off; p.u_off = h->d.u_off;
  p.nnz_jac = h->nnz_jac_knot(); p.nnz_hess = h->nnz_hess_knot();
  p.G0 = h->dG0; p.Gj = h->dGj;
  return p;
}

int launch_resjac_core(pb2_handle* h, const double* dZ, double* ddelta, double* djac, cudaStream_t st, int compact,
                       int n_peers, double* const* peers, int self, int z_stable);

int launch_resjac_td(pb2_handle* h, const double* dZ, double* ddelta, double* djac, cudaStream_t st, int compact,
                     int n_peers, double* const* peers, int self, int z_stable);

// Every residual / Jacobian launch of the entry points goes through here.  dense_blocks handles: the kernels write
// their canonical values into the handle's scratch, one streaming pass re-lays them out for the caller.
int launch_resjac(pb2_handle* h, const double* dZ, double* ddelta, double* djac, cudaStream_t st, int compact = 0,
                  int n_peers = 0, double* const* peers = nullptr, int self = 0, int z_stable = 0) {
  if (!h->d.dense_blocks || !djac) return launch_resjac_td(h, dZ, ddelta, djac, st, compact, n_peers, peers, self, z_stable);
  if (h->nk() <= 0) return PB2_OK;
  if (compact || n_peers) return fail(PB2_EINVAL, "dense_blocks handles do not produce compact records");
  if (!h->dCanon) PB2_CUDA(cudaMalloc(&h->dCanon, (size_t)h->nnz_jac_knot() * h->nk() * sizeof(double)));
  int rc = launch_resjac_td(h, dZ, ddelta, h->dCanon, st, 0, 0, nullptr, 0, z_stable);
  if (rc) return rc;
  dense_blocks_kernel<<<(unsigned)std::min<int64_t>(h->nk(), 148 * 8), 256, 0, st>>>(
      h->dCanon, djac, h->nk(), h->d.b, h->d.n_b, h->nnz_jac_knot(), h->nnz_jac_user_knot());
  PB2_CUDA(cudaGetLastError());
  h->launches++;
  return PB2_OK;
}

// Time-dependent handles: scale the drive rows by c_j(t_k), run the knot kernels on the scaled copy, then apply the
// chain rule to the finished values (knot_td.cuh).
int launch_resjac_td(pb2_handle* h, const double* dZ, double* ddelta, double* djac, cudaStream_t st, int compact,
                     int n_peers, double* const* peers, int self, int z_stable) {
  if (!h->d.time_dependent) return launch_resjac_core(h, dZ, ddelta, djac, st, compact, n_peers, peers, self, z_stable);
  if (h->nk() <= 0) return PB2_OK;
  if (compact || n_peers) return fail(PB2_EINVAL, "time-dependent handles do not produce compact records");
  if (!h->coef_set) return fail(PB2_EINVAL, "time-dependent handle: call pb2_set_time_coefficients before evaluating");
  pb2::TdParams t{};
  t.K = h->d.K; t.D = h->d.D; t.m = h->d.m; t.n_x = h->n_x(); t.u_off = h->d.u_off;
  t.nnz_jac = h->nnz_jac_knot(); t.o_jets = (long long)h->d.n_b * h->d.b * h->d.b;
  t.c = h->dCoef; t.cdot = h->dCoefDot; t.Z = dZ; t.Zs = h->dZs; t.jac = djac;
  const long long n = (long long)t.K * t.D;
  pb2::td_scale_controls_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>(t);
  PB2_CUDA(cudaGetLastError());
  // (the scaled copy is produced by the kernel just enqueued: the knot kernels must not read it early)
  int rc = launch_resjac_core(h, h->dZs, ddelta, djac, st, 0, 0, nullptr, 0, 0);
  if (rc) return rc;
  if (djac) {
    pb2::td_finish_kernel<<<(unsigned)std::min<int64_t>(h->nk(), 148 * 8), 128, 0, st>>>(t);
    PB2_CUDA(cudaGetLastError());
  }
  h->launches += djac ? 2 : 1;
  return PB2_OK;
}

int launch_resjac_core(pb2_handle* h, const double* dZ, double* ddelta, double* djac, cudaStream_t st, int compact,
                       int n_peers, double* const* peers, int self, int z_stable) {
  if (h->nk() <= 0) return PB2_OK;
  pb2::KnotParams p = make_params(h);
  p.Z = dZ; p.delta = ddelta; p.jac = djac;
  const BatchLaunch* bl = h->batch;
  const bool aligned16 = !bl && ((uintptr_t)dZ % 16 == 0) && ((uintptr_t)ddelta % 16 == 0) && ((uintptr_t)djac % 16 == 0);
  if (bl) {
    p.mem_n = bl->n; p.x_offs = bl->x_offs; p.G0 = bl->G0; p.Gj = bl->Gj;
    p.mem_G0 = bl->mem_G0; p.mem_Gj = bl->mem_Gj; p.mem_delta = bl->mem_delta; p.mem_jac = bl->mem_jac;
  }
  // knots per CTA of the small-CTA 3-qubit kernel: the configured count, or the next larger one whose CTA still holds the
  // table blob and its knot slots (the blob is per CTA: long knot columns leave room for fewer, larger CTAs); 0 = the
  // column does not fit at all and the older kernels below take the call
  int u8q_ns = 0;
  if (h->u8p_ok && h->u8q) {
    pb2::U8qParams probe{};
    probe.m = p.m; probe.zlen = p.D + p.x_off + 128; probe.n_peers = n_peers ? n_peers : (compact ? 8 : 0);
    probe.cstride = (p.m + 3) * 128;
    for (int ns = h->u8q_ns; ns <= 4 && !u8q_ns; ns *= 2)
      if (pb2::u8q_layout(probe, ns) <= kSmemLimit / (size_t)(8 / ns)) u8q_ns = ns;
  }
  // (when even four knots per CTA do not fit, the column is too long for any of the slab-staging 3-qubit kernels: the
  // general kernel, which reads the trajectory straight from global memory, takes the call)
  const bool slab_fits = !(h->u8p_ok && h->u8q) || u8q_ns != 0;
  if (h->alg == PB2_ALG_DMMA && h->u8_ok && h->u8p_ok && h->u8q && u8q_ns && djac && aligned16 && (p.D % 2 == 0) &&
      (p.x_off % 2 == 0) && (n_peers == 0 || !std::getenv("PB2_U8Q_NO_PEERS")) &&
      (h->u8q >= 2 || h->nk() <= (int64_t)7 * h->n_sm)) {
    // two 256-thread CTAs per SM, at most four knots each: one CTA's prologue and tail run underneath the
    // other one's products, and a freed half-SM goes to the next grid of the stream at once (knot_u8q.cuh)
    pb2::U8qParams q{};
    q.m = p.m; q.D = p.D; q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off;
    q.nnz_jac = p.nnz_jac; q.max_sub = 4096; q.nk = (int)h->nk();
    q.zlen = p.D + p.x_off + 128;
    // (time-dependent handles read a scaled copy written by the kernel enqueued just before: never early)
    const bool td = h->d.time_dependent != 0;
    q.early_z = (!td && (z_stable || h->early_z || h->pipelined)) ? 1 : 0;
    q.nowait = (!td && (z_stable || h->pipelined)) ? 1 : 0;   // z_stable: the library's own host-pointer path
    q.compact = compact; q.cstride = (p.m + 3) * 128;
    q.n_peers = n_peers; q.self = self;
    for (int r = 0; r < n_peers && r < 8; ++r) q.peers[r] = peers[r];
    if (n_peers > 0) q.slot_off = (long long)(djac - peers[self]);
    q.flag_off = n_peers > 1 ? h->xchg_flag_off : -1;
    q.tables = h->dTablesQ; q.ell = h->dEll;
    for (int j = 0; j < 4; ++j) q.cj[j] = h->u8p_cj[j];
    q.Z = dZ; q.delta = ddelta; q.jac = djac;
#ifdef PB2_TRACE
    q.trace = h->dTrace3; q.trace_id = (h->trace_launch++) % 64;
#endif
    const int ns = u8q_ns, cps = 8 / ns;   // knots per CTA, CTAs per SM
    if (const char* env = std::getenv("PB2_NOWAIT")) q.nowait = std::atoi(env);   // (measurement knob)
    unsigned grid;
    if (q.nk <= 7 * h->n_sm) {
      // cps CTAs per SM; the last n_sm CTAs own one knot less: blocks b, b + n_sm, ... share an SM on an idle
      // device (7 knots per SM, as evenly as 6.75 allows); under back-to-back launches the hardware balances
      grid = (unsigned)std::min(cps * h->n_sm, q.nk);
      q.split = (cps - 1) * h->n_sm;
    } else {
      grid = (unsigned)((q.nk + ns - 1) / ns);
      q.split = 0;
    }
    const size_t smem = pb2::u8q_layout(q, ns);
    if (smem > kSmemLimit / cps) return fail(PB2_EINVAL, "u8q resjac: knot column too large for the shared-memory staging");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(64 * ns);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pb2::u8q_kernel(h->u8p_unit, ns), q);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PB2_ECUDA, std::string("u8q resjac launch: ") + cudaGetErrorString(e));
    h->launches++;
    return PB2_OK;
  }
  if (h->alg == PB2_ALG_DMMA && h->u8_ok && h->u8p_ok && h->u8p && djac && aligned16 && slab_fits && (p.D % 2 == 0) &&
      (p.x_off % 2 == 0) && n_peers == 0 && h->nk() <= (int64_t)pb2::kU8pSlots * h->n_sm) {
    // at most seven knots per SM: every knot of an SM in flight at once, propagator tiles first (knot_u8p.cuh)
    pb2::U8pParams q{};
    q.m = p.m; q.D = p.D; q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off;
    q.nnz_jac = p.nnz_jac; q.max_sub = 4096; q.nk = (int)h->nk();
    q.zlen = p.D + p.x_off + 128;
    q.early_z = (z_stable || h->early_z) ? 1 : 0;
    q.compact = compact; q.cstride = (p.m + 3) * 128;
    q.tables = h->dTablesP; q.ell = h->dEll;
    for (int j = 0; j < 4; ++j) q.cj[j] = h->u8p_cj[j];
    q.Z = dZ; q.delta = ddelta; q.jac = djac;
#ifdef PB2_TRACE
    q.trace = h->dTrace3; q.trace_id = (h->trace_launch++) % 64;
#endif
    const size_t smem = pb2::u8p_layout(q);
    if (smem > kSmemLimit) return fail(PB2_EINVAL, "u8p resjac: knot column too large for the shared-memory staging");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)std::min<int64_t>(h->n_sm, q.nk));
    cfg.blockDim = dim3(pb2::kU8pThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pb2::u8p_kernel(h->u8p_unit), q);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PB2_ECUDA, std::string("u8p resjac launch: ") + cudaGetErrorString(e));
    h->launches++;
    return PB2_OK;
  }
  if (h->alg == PB2_ALG_DMMA && h->u8_ok && djac && aligned16 && slab_fits && (p.D % 2 == 0) && (p.x_off % 2 == 0)) {
    // the 3-qubit unitary shape: warp-specialised kernel (producer warp + (E,X) warp + jet warps)
    const pb2::DmmaPlan& pl = h->plan;
    pb2::U8Params q{};
    q.m = p.m; q.D = p.D; q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off;
    q.nnz_jac = p.nnz_jac; q.max_sub = 4096; q.nk = (int)h->nk();
    q.zlen = p.D + p.x_off + 128;
    q.gw = 2 + (p.m + 1) / 2;
    q.stagger = h->stagger;
    q.stagger_g = h->stagger_g;
    q.compact = compact;
    q.n_peers = n_peers;
    q.peer_tma = std::getenv("PB2_PEER_TMA") ? std::atoi(std::getenv("PB2_PEER_TMA")) : 1;   // 0: 16-byte st.global instead
    q.self = self;
    for (int r = 0; r < n_peers && r < 8; ++r) q.peers[r] = peers[r];
    q.direct_last = h->direct_last;
    q.cstride = (p.m + 3) * 128;
#ifdef PB2_TRACE
    if (const char* env = std::getenv("PB2_DRY")) q.dry = std::atoi(env);   // launch-floor measurement (debug build only)
#endif
    q.tables = h->dTables; q.ell = h->dEll;
    q.Z = dZ; q.delta = ddelta; q.jac = djac; q.trace = h->dTrace2;
    int maxg = std::min(pb2::kU8MaxGroups, pb2::kU8MaxThreads / (32 * q.gw));
    while (maxg > 1 && pb2::u8_layout(q, maxg) > kSmemLimit) --maxg;
    const int per_sm = (q.nk + h->n_sm - 1) / h->n_sm;
    q.gpc = std::max(1, std::min(maxg, h->gpc_override > 0 ? h->gpc_override : per_sm));
    const int blocks = std::min(h->n_sm, (q.nk + q.gpc - 1) / q.gpc);
    const size_t smem = pb2::u8_layout(q, q.gpc);
    if (smem > kSmemLimit) return fail(PB2_EINVAL, "u8 resjac: knot column too large for the shared-memory staging");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(32 * q.gw * q.gpc);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pb2::u8_kernel(pl.W), q);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PB2_ECUDA, std::string("u8 resjac launch: ") + cudaGetErrorString(e));
  } else if (compact) {
    return fail(PB2_EINVAL, "compact records are only produced by the 3-qubit unitary kernel with aligned pointers");
  } else if (h->alg == PB2_ALG_DMMA && h->dmmaq) {
    // one small CTA per knot, several resident per SM (knot_dmmaq.cuh)
    const bool jets = djac != nullptr;
    const pb2::DmmaPlan& pl = h->plan;
    pb2::DmmaqParams q{};
    q.b = p.b; q.n_b = p.n_b; q.m = p.m; q.D = p.D; q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off;
    q.nnz_jac = p.nnz_jac; q.max_sub = 4096; q.nk = (int)h->nk();
    q.ncT = jets ? pl.ncT : 0; q.iso = pl.iso; q.m_jets = jets ? p.m : 0;
    q.tiles = jets ? pl.tiles_full : pl.tiles_res;
    q.Gfrag = h->dGfrag; q.ell = h->dEll; q.norms = h->dNorms;
    q.Z = dZ; q.delta = ddelta; q.jac = djac;
    if (bl) {
      q.mem_n = bl->n; q.x_offs = bl->x_offs; q.Gfrag = bl->gfrag; q.ell = bl->ell; q.norms = bl->norms;
      q.mem_gfrag = bl->mem_gfrag; q.mem_ell = bl->mem_ell; q.mem_norms = bl->mem_norms;
      q.mem_delta = bl->mem_delta; q.mem_jac = bl->mem_jac;
    }
    const size_t smem = pb2::dmmaq_layout(q, pl.NT);
    const int threads = 32 * q.tiles;
    pb2::DmmaqKernel kern = pb2::dmmaq_kernel(pl.NT, pl.W);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      occ = 1;
    }
    const int n_mem = bl ? bl->n : 1;
    const int blocks = (int)std::min<int64_t>(q.nk, std::max(1, h->n_sm * occ / n_mem));
    kern<<<dim3(blocks, n_mem), threads, smem, st>>>(q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PB2_ECUDA, std::string("dmmaq resjac launch: ") + cudaGetErrorString(e));
  } else if (h->alg == PB2_ALG_DMMA) {
    // residual-only calls carry just the state columns; anything with a Jacobian carries the
    // propagator columns and one jet per drive as well
    const bool jets = djac != nullptr;
    const pb2::DmmaPlan& pl = h->plan;
    pb2::DmmaParams q{};
    q.b = p.b; q.n_b = p.n_b; q.m = p.m; q.D = p.D;
    q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off; q.nnz_jac = p.nnz_jac;
    q.ncT = jets ? pl.ncT : 0;
    q.m_jets = jets ? p.m : 0;
    q.iso = pl.iso; q.max_sub = 4096;
    q.tiles = jets ? pl.tiles_full : pl.tiles_res;
    q.nk = (int)h->nk();
    q.zlen = p.D + p.x_off + p.b * p.n_b;
    q.Gfrag = h->dGfrag; q.ell = h->dEll; q.norms = h->dNorms;
    q.Z = dZ; q.delta = ddelta; q.jac = djac; q.trace = h->dTrace;
    const int bb = p.b * p.b, n_x = p.b * p.n_b;
    if (bl) {
      // all members of the ensemble in this launch: own tables and state block per member (blockIdx.y)
      q.mem_n = bl->n; q.x_offs = bl->x_offs; q.Gfrag = bl->gfrag; q.ell = bl->ell; q.norms = bl->norms;
      q.mem_gfrag = bl->mem_gfrag; q.mem_ell = bl->mem_ell; q.mem_norms = bl->mem_norms;
      q.mem_delta = bl->mem_delta; q.mem_jac = bl->mem_jac;
      q.zlen = p.D + bl->max_xoff + n_x;        // the slab covers the furthest member's next-knot state
    }
    q.bulk_in = (p.D % 2 == 0) && (q.zlen % 2 == 0) && ((uintptr_t)dZ % 16 == 0);
    q.bulk_out = (bb % 2 == 0) && (n_x % 2 == 0) && ((uintptr_t)ddelta % 16 == 0) && ((uintptr_t)djac % 16 == 0) &&
                 (!bl || (bl->mem_delta % 2 == 0 && bl->mem_jac % 2 == 0));
    // persistent grid: one CTA per SM, gpc knot groups per CTA (bounded by threads, barriers, smem)
    int maxg = std::min(pb2::kDmmaMaxGroups, pb2::dmma_max_threads(pl.NT, pl.W) / (32 * q.tiles));
    while (maxg > 1 && pb2::dmma_layout(q, pl.NT, maxg) > kSmemLimit) --maxg;
    // (an ensemble shares the SMs: one wave of persistent CTAs over all members, each CTA walking its member's knots
    // with as many groups as fit -- a CTA per two knots would pay the prologue once per wave instead of once)
    const int n_mem = bl ? bl->n : 1;
    const int per_sm = (int)(((long long)q.nk * n_mem + h->n_sm - 1) / h->n_sm);
    q.gpc = std::max(1, std::min(maxg, h->gpc_override > 0 ? h->gpc_override : std::min(per_sm, bl ? maxg : h->gpc_default)));
    const int blocks = std::min(std::max(1, h->n_sm / n_mem), (q.nk + q.gpc - 1) / q.gpc);
    const size_t smem = pb2::dmma_layout(q, pl.NT, q.gpc);
    if (smem > kSmemLimit) return fail(PB2_EINVAL, "dmma resjac: knot column too large for the shared-memory staging");
    pb2::dmma_kernel(pl.NT, pl.W)<<<dim3(blocks, bl ? bl->n : 1), 32 * q.tiles * q.gpc, smem, st>>>(q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PB2_ECUDA, std::string("dmma resjac launch: ") + cudaGetErrorString(e));
  } else {
    const LaunchCfg& c = h->cfg1;
    const int blocks = (int)((h->nk() + c.KPC - 1) / c.KPC);
    pb2::knot_generic_kernel<1, kNT><<<dim3(blocks, bl ? bl->n : 1), kNT, c.smem, st>>>(p, c.GS, c.KPC, c.gj_in_smem);
    PB2_CUDA(cudaGetLastError());
  }
  h->launches++;
  return PB2_OK;
}

int launch_hess(pb2_handle* h, const double* dZ, const double* dmu, double* dhess, cudaStream_t st) {
  if (h->nk() <= 0) return PB2_OK;
  if (h->d.time_dependent)
    return fail(PB2_EINVAL, "the Lagrangian Hessian of a time-dependent handle is not available (use the reference's "
                            "integrator or a quasi-Newton Hessian)");
  pb2::KnotParams p = make_params(h);
  p.Z = dZ; p.mu = dmu; p.hess = dhess;
  const BatchLaunch* bl = h->batch;
  if (bl) {
    p.mem_n = bl->n; p.x_offs = bl->x_offs; p.G0 = bl->G0; p.Gj = bl->Gj;
    p.mem_G0 = bl->mem_G0; p.mem_Gj = bl->mem_Gj; p.mem_delta = bl->mem_delta; p.mem_hess = bl->mem_hess;
  }
  bool u8h_fits = false;
  if (h->u8h_ok) {
    pb2::U8hParams probe{};
    probe.m = p.m; probe.zlen = p.D + p.x_off + 128; probe.nnz_hess = p.nnz_hess;
    probe.ntiles = 2 + 2 * p.m + p.m * (p.m + 1) / 2; probe.ncw = (probe.ntiles + 1) / 2;   // (layout does not depend on it)
    u8h_fits = pb2::u8h_layout(probe) <= kSmemLimit;
  }
  if (!bl && h->u8h_ok && u8h_fits && !(h->hplan.ok && h->hess_prefer_dmmah) && ((uintptr_t)dZ % 16 == 0) &&
      ((uintptr_t)dmu % 16 == 0) && (p.D % 2 == 0) && (p.x_off % 2 == 0)) {
    // 3-qubit unitary shape, anti-symmetric generators: tensor-core Hessian (forward + adjoint jets)
    pb2::U8hParams q{};
    q.m = p.m; q.D = p.D; q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off;
    q.nnz_hess = p.nnz_hess; q.max_sub = 4096; q.nk = (int)h->nk();
    q.zlen = p.D + p.x_off + 128;
    q.ntiles = 2 + 2 * p.m + p.m * (p.m + 1) / 2;
    // forward and adjoint tiles in separate warps, two tiles per warp
    q.ncw = (1 + p.m + p.m * (p.m + 1) / 2 + 1) / 2 + (1 + p.m + 1) / 2;
    q.tables = h->dTables; q.ell = h->dEll;
    q.Z = dZ; q.mu = dmu; q.hess = dhess; q.trace = h->dTrace2;
    const size_t smem = pb2::u8h_layout(q);
    if (smem > kSmemLimit) return fail(PB2_EINVAL, "u8h hessian: knot column too large for the shared-memory staging");
    const int blocks = std::min(h->hess_ctas > 0 ? std::min(h->hess_ctas, h->n_sm) : h->n_sm, q.nk);
    pb2::u8h_kernel(h->plan.W)<<<blocks, 32 * (1 + q.ncw), smem, st>>>(q);
    PB2_CUDA(cudaGetLastError());
    h->launches++;
    return PB2_OK;
  }
  if (h->hplan.ok && (!bl || bl->hgfrag)) {
    // general tensor-core Hessian: one CTA per knot, forward + adjoint tiles (knot_dmmah.cuh)
    const pb2::DmmahPlan& hp = h->hplan;
    pb2::DmmahParams q{};
    q.b = p.b; q.n_b = p.n_b; q.m = p.m; q.D = p.D; q.x_off = p.x_off; q.dt_off = p.dt_off; q.u_off = p.u_off;
    q.nnz_hess = p.nnz_hess; q.max_sub = 4096; q.nk = (int)h->nk();
    q.tiles_f = hp.tiles_f; q.tiles_a = hp.tiles_a;
    q.Gfrag = h->dHGfrag; q.GfragT = h->dHGfragT; q.ell = h->dHEll; q.ellT = h->dHEllT; q.norms = h->dHNorms;
    q.Z = dZ; q.mu = dmu; q.hess = dhess;
    if (bl) {
      q.mem_n = bl->n; q.x_offs = bl->x_offs;
      q.Gfrag = bl->hgfrag; q.GfragT = bl->hgfragT; q.ell = bl->hell; q.ellT = bl->hellT; q.norms = bl->hnorms;
      q.mem_gfrag = bl->mem_hgfrag; q.mem_ell = bl->mem_hell; q.mem_norms = bl->mem_hnorms;
      q.mem_mu = bl->mem_delta; q.mem_hess = bl->mem_hess;
    }
    const size_t smem = pb2::dmmah_layout(q, hp.NT);
    const int threads = 32 * (hp.tiles_f + hp.tiles_a);
    pb2::DmmahKernel kern = pb2::dmmah_kernel(hp.NT, hp.W, hp.tiles_f + hp.tiles_a);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      occ = 1;
    }
    const int n_mem = bl ? bl->n : 1;
    const int blocks = (int)std::min<int64_t>(q.nk, std::max(1, h->n_sm * occ / n_mem));
    kern<<<dim3(blocks, n_mem), threads, smem, st>>>(q);
    PB2_CUDA(cudaGetLastError());
  } else {
    // every other shape: the jet kernel (second-order jets in shared memory)
    const LaunchCfg& c = h->cfg2;
    const int blocks = (int)((h->nk() + c.KPC - 1) / c.KPC);
    pb2::knot_generic_kernel<2, kNT><<<dim3(blocks, bl ? bl->n : 1), kNT, c.smem, st>>>(p, c.GS, c.KPC, c.gj_in_smem);
    PB2_CUDA(cudaGetLastError());
  }
  h->launches++;
  return PB2_OK;
}

// Host-side replication of a landed chunk into the caller's arrays.  With PB2_HOST_NT=1 the copies use streaming
// (non-temporal) stores: several processes un-packing 23.5 MB each per callback on one host exceed the last-level
// cache, and a plain store first READS every destination line it is about to overwrite.
inline void host_copy(double* dst, const double* src, size_t n, bool nt) {
  if (!nt || ((uintptr_t)dst & 15u) != 0) {
    std::memcpy(dst, src, n * sizeof(double));
    return;
  }
  size_t i = 0;
  for (; i + 2 <= n; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
  if (i < n) dst[i] = src[i];
}

bool is_pinned_or_device(const void* ptr) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// host -> device (through pinned staging unless the caller's buffer is already pinned)
int stage_in(pb2_handle* h, const double* src, double* pinned, double* dst, size_t n) {
  if (n == 0) return PB2_OK;
  const double* from = src;
  if (!is_pinned_or_device(src)) {
    std::memcpy(pinned, src, n * sizeof(double));
    from = pinned;
  }
  PB2_CUDA(cudaMemcpyAsync(dst, from, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return PB2_OK;
}

struct PendingOut {
  double* user;
  double* pinned;
  size_t n;
  bool direct;
};

int stage_out_begin(pb2_handle* h, double* user, double* pinned, const double* dsrc, size_t n,
                    PendingOut& po) {
  po = {user, pinned, n, false};
  if (n == 0) return PB2_OK;
  po.direct = is_pinned_or_device(user);
  PB2_CUDA(cudaMemcpyAsync(po.direct ? user : pinned, dsrc, n * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
  return PB2_OK;
}

void stage_out_finish(const PendingOut& po) {
  if (po.n && !po.direct) std::memcpy(po.user, po.pinned, po.n * sizeof(double));
}

int check(const pb2_handle* h) {
  if (!h) return fail(PB2_EINVAL, "null handle");
  return PB2_OK;
}

}  // namespace

extern "C" {

int pb2_version(void) { return PB2_VERSION; }

const char* pb2_last_error(void) { return g_err.c_str(); }

int pb2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pb2_create(const pb2_desc* desc, pb2_handle** out) {
  if (!desc || !out) return fail(PB2_EINVAL, "pb2_create: null argument");
  *out = nullptr;
  const pb2_desc& d = *desc;
  if (d.kind < PB2_KET || d.kind > PB2_DENSITY) return fail(PB2_EINVAL, "pb2_create: bad kind");
  if (d.b < 1 || d.n_b < 1 || d.m < 0 || d.K < 1 || d.D < 1)
    return fail(PB2_EINVAL, "pb2_create: sizes must be positive");
  if ((d.kind == PB2_KET || d.kind == PB2_UNITARY) && (d.b % 2))
    return fail(PB2_EINVAL, "pb2_create: ket/unitary generators have even size 2d");
  if (d.kind == PB2_UNITARY && d.n_b * 2 != d.b)
    return fail(PB2_EINVAL, "pb2_create: unitary needs n_b = b/2");
  const int n_x = d.b * d.n_b;
  auto inside = [&](int off, int len) { return off >= 0 && off + len <= d.D; };
  if (!inside(d.x_off, n_x) || !inside(d.dt_off, 1) || !inside(d.u_off, d.m))
    return fail(PB2_EINVAL, "pb2_create: component offsets outside the knot column");
  if (!d.G0 || (d.m > 0 && !d.Gj)) return fail(PB2_EINVAL, "pb2_create: null generator");
  if (d.time_dependent && !inside(d.t_off, 1))
    return fail(PB2_EINVAL, "pb2_create: time-dependent handle needs the row of the knot time inside the knot column");
  if (d.algorithm < PB2_ALG_AUTO || d.algorithm > PB2_ALG_DMMA)
    return fail(PB2_EINVAL, "pb2_create: bad algorithm");

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(PB2_ENODEVICE, "pb2_create: no CUDA device (this library has no CPU path)");
  }
  if (d.device < 0 || d.device >= ndev) return fail(PB2_EINVAL, "pb2_create: bad device ordinal");
  DeviceGuard guard_1(d.device);

  pb2_handle* h = new (std::nothrow) pb2_handle();
  if (!h) return fail(PB2_ENOMEM, "pb2_create: out of memory");
  h->d = d;
  const size_t bb = (size_t)d.b * d.b;
  h->G0.assign(d.G0, d.G0 + bb);
  if (d.m) h->Gj.assign(d.Gj, d.Gj + (size_t)d.m * bb);
  h->d.G0 = h->G0.data();
  h->d.Gj = h->Gj.data();

  // algorithm selection (decided once, at construction -- never a runtime fallback)
  // the tensor-core path covers generators up to 16 x 16 with sparse drive terms; ket / unitary
  // generators additionally use the half (real-isomorphism) layout of the propagator
  if (d.algorithm != PB2_ALG_GENERIC)
    h->plan = pb2::dmma_plan(d.b, d.n_b, d.m, d.kind != PB2_DENSITY, h->G0.data(), h->Gj.data());
  if (d.algorithm == PB2_ALG_DMMA && !h->plan.ok) {
    delete h;
    return fail(PB2_EINVAL, "pb2_create: tensor-core path unsupported for this generator "
                            "(needs b <= 24, <= 4 nonzeros per drive-generator row, <= 8 column tiles)");
  }
  h->alg = h->plan.ok ? PB2_ALG_DMMA : PB2_ALG_GENERIC;

  // the jet kernels: residual/Jacobian when the tensor-core path is off, the Hessian always
  const bool need1 = h->alg == PB2_ALG_GENERIC;
  if ((need1 && !plan_generic(1, d.b, d.n_b, d.m, h->cfg1)) || !plan_generic(2, d.b, d.n_b, d.m, h->cfg2)) {
    delete h;
    return fail(PB2_EINVAL, "pb2_create: generator too large for the shared-memory kernels");
  }

  auto cleanup_fail = [&](int code, const std::string& msg) {
    pb2_destroy(h);
    return fail(code, msg);
  };
#define PB2_CUDA_H(call)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return cleanup_fail(PB2_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

  PB2_CUDA_H(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (d.time_dependent) {
    const size_t nc = std::max<size_t>(1, (size_t)d.m * d.K);
    PB2_CUDA_H(cudaMalloc(&h->dCoef, nc * sizeof(double)));
    PB2_CUDA_H(cudaMalloc(&h->dCoefDot, nc * sizeof(double)));
    PB2_CUDA_H(cudaMalloc(&h->dZs, (size_t)d.D * d.K * sizeof(double)));
  }
  PB2_CUDA_H(cudaMalloc(&h->dG0, bb * sizeof(double)));
  PB2_CUDA_H(cudaMalloc(&h->dGj, std::max<size_t>(1, (size_t)d.m * bb) * sizeof(double)));
  PB2_CUDA_H(cudaMemcpy(h->dG0, h->G0.data(), bb * sizeof(double), cudaMemcpyHostToDevice));
  if (d.m)
    PB2_CUDA_H(cudaMemcpy(h->dGj, h->Gj.data(), (size_t)d.m * bb * sizeof(double), cudaMemcpyHostToDevice));

  double theta[pb2::kMaxDeg + 1];
  theta[0] = 0.0;
  for (int q = 1; q <= pb2::kMaxDeg; ++q) theta[q] = theta_bound(q);
  PB2_CUDA_H(cudaMemcpyToSymbol(pb2::c_theta, theta, sizeof(theta)));

    PB2_CUDA_H(raise_dynamic_smem(pb2::knot_generic_kernel<1, kNT>));
  PB2_CUDA_H(raise_dynamic_smem(pb2::knot_generic_kernel<2, kNT>));
  if (h->alg == PB2_ALG_DMMA) {
    const size_t ng = h->plan.gfrag.size() * sizeof(double), ne = h->plan.ell.size() * sizeof(pb2::EllEntry);
    const size_t nn = h->plan.norms.size() * sizeof(double);
    PB2_CUDA_H(cudaMalloc(&h->dGfrag, ng));
    PB2_CUDA_H(cudaMalloc(&h->dEll, ne));
    PB2_CUDA_H(cudaMalloc(&h->dNorms, nn));
    PB2_CUDA_H(cudaMemcpy(h->dGfrag, h->plan.gfrag.data(), ng, cudaMemcpyHostToDevice));
    PB2_CUDA_H(cudaMemcpy(h->dEll, h->plan.ell.data(), ne, cudaMemcpyHostToDevice));
    PB2_CUDA_H(cudaMemcpy(h->dNorms, h->plan.norms.data(), nn, cudaMemcpyHostToDevice));
    // the Lagrangian Hessian on the tensor cores too (the 3-qubit shape has its own kernel; time-dependent handles
    // have no Hessian)
    h->dmmah = std::getenv("PB2_NO_DMMAH") ? 0 : 1;
    h->hess_prefer_dmmah = std::getenv("PB2_HESS_DMMAH") ? std::atoi(std::getenv("PB2_HESS_DMMAH")) : 0;
    h->dmmaq = std::getenv("PB2_DMMAQ") ? std::atoi(std::getenv("PB2_DMMAQ")) : 1;
    if (!h->plan.persistent_ok) h->dmmaq = 1;   // 17 <= b <= 24 or more than 8 state columns: only the small-CTA kernel
    if (h->dmmaq) PB2_CUDA_H(raise_dynamic_smem(pb2::dmmaq_kernel(h->plan.NT, h->plan.W)));
    if (h->dmmah && !d.time_dependent) h->hplan = pb2::dmmah_plan(d.b, d.n_b, d.m, h->G0.data(), h->Gj.data());
    if (h->hplan.ok) {
      const pb2::DmmahPlan& hp = h->hplan;
      const size_t hg = hp.gfrag.size() * sizeof(double), he = hp.ell.size() * sizeof(pb2::EllEntry);
      PB2_CUDA_H(cudaMalloc(&h->dHGfrag, hg));
      PB2_CUDA_H(cudaMalloc(&h->dHGfragT, hg));
      PB2_CUDA_H(cudaMalloc(&h->dHEll, he));
      PB2_CUDA_H(cudaMalloc(&h->dHEllT, he));
      PB2_CUDA_H(cudaMalloc(&h->dHNorms, hp.norms.size() * sizeof(double)));
      PB2_CUDA_H(cudaMemcpy(h->dHGfrag, hp.gfrag.data(), hg, cudaMemcpyHostToDevice));
      PB2_CUDA_H(cudaMemcpy(h->dHGfragT, hp.gfragT.data(), hg, cudaMemcpyHostToDevice));
      PB2_CUDA_H(cudaMemcpy(h->dHEll, hp.ell.data(), he, cudaMemcpyHostToDevice));
      PB2_CUDA_H(cudaMemcpy(h->dHEllT, hp.ellT.data(), he, cudaMemcpyHostToDevice));
      PB2_CUDA_H(cudaMemcpy(h->dHNorms, hp.norms.data(), hp.norms.size() * sizeof(double), cudaMemcpyHostToDevice));
      PB2_CUDA_H(raise_dynamic_smem(pb2::dmmah_kernel(hp.NT, hp.W, hp.tiles_f + hp.tiles_a)));
    }
    double invfact[pb2::kMaxDeg + 1];
    invfact[0] = 1.0;
    for (int q = 1; q <= pb2::kMaxDeg; ++q) invfact[q] = invfact[q - 1] / (double)q;
    PB2_CUDA_H(cudaMemcpyToSymbol(pb2::c_invfact, invfact, sizeof(invfact)));
    {
      double tab[40] = {0};
      for (int q = 0; q <= pb2::kMaxDeg; ++q) { tab[q] = theta[q]; tab[20 + q] = invfact[q]; }
      PB2_CUDA_H(cudaMalloc(&h->dTab, sizeof(tab)));
      PB2_CUDA_H(cudaMemcpy(h->dTab, tab, sizeof(tab), cudaMemcpyHostToDevice));
      std::vector<double> tb(h->plan.gfrag);
      tb.insert(tb.end(), h->plan.norms.begin(), h->plan.norms.end());
      if (tb.size() % 2) tb.push_back(0.0);
      tb.insert(tb.end(), tab, tab + 40);
      PB2_CUDA_H(cudaMalloc(&h->dTables, tb.size() * sizeof(double)));
      PB2_CUDA_H(cudaMemcpy(h->dTables, tb.data(), tb.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (h->plan.persistent_ok)
      PB2_CUDA_H(cudaFuncSetAttribute(pb2::dmma_kernel(h->plan.NT, h->plan.W),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    PB2_CUDA_H(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, d.device));
    if (const char* env = std::getenv("PB2_GPC")) h->gpc_override = std::atoi(env);
    if (const char* env = std::getenv("PB2_STAGGER")) h->stagger = std::atoi(env);
    if (const char* env = std::getenv("PB2_PDL")) h->pdl = std::atoi(env);
    if (const char* env = std::getenv("PB2_STAGGER_G")) h->stagger_g = std::atoi(env);
    if (const char* env = std::getenv("PB2_DIRECT_LAST")) h->direct_last = std::atoi(env);
    if (const char* env = std::getenv("PB2_U8P")) h->u8p = std::atoi(env);
    if (const char* env = std::getenv("PB2_U8Q")) h->u8q = std::atoi(env);
    if (const char* env = std::getenv("PB2_U8Q_NS")) {
      const int v = std::atoi(env);
      h->u8q_ns = (v == 1 || v == 2) ? v : 4;
    }
    if (const char* env = std::getenv("PB2_EARLY_Z")) h->early_z = std::atoi(env);
    if (const char* env = std::getenv("PB2_PIPELINED")) h->pipelined = std::atoi(env);
    h->u8_ok = h->plan.iso && d.b == 16 && d.n_b == 8 && d.m >= 1 && d.m <= 6 && !std::getenv("PB2_NO_U8");
    if (h->u8_ok) {
      PB2_CUDA_H(cudaFuncSetAttribute(pb2::u8_kernel(h->plan.W), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemLimit));
      // the single-round kernel: m = 3 or 4 drives whose generators have one nonzero per row (ELL width 1)
      std::vector<double> blob;
      if ((d.m == 3 || d.m == 4) && pb2::u8p_tables(h->plan, d.m, blob, h->u8p_unit, h->u8p_cj)) {
        double* tabp = blob.data() + blob.size() - 40;
        for (int q = 0; q <= pb2::kMaxDeg; ++q) { tabp[q] = theta[q]; tabp[20 + q] = invfact[q]; }
        if (std::getenv("PB2_NO_UNIT")) h->u8p_unit = false;
        PB2_CUDA_H(cudaMalloc(&h->dTablesP, blob.size() * sizeof(double)));
        PB2_CUDA_H(cudaMemcpy(h->dTablesP, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice));
        PB2_CUDA_H(cudaFuncSetAttribute(pb2::u8p_kernel(h->u8p_unit), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kSmemLimit));
        for (int ns : {1, 2, 4}) {
          PB2_CUDA_H(cudaFuncSetAttribute(pb2::u8q_kernel(h->u8p_unit, ns), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(kSmemLimit * ns / 8)));
          PB2_CUDA_H(cudaFuncSetAttribute(pb2::u8q_kernel(h->u8p_unit, ns), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
        {
          double tab40[40] = {0};
          for (int q = 0; q <= pb2::kMaxDeg; ++q) { tab40[q] = theta[q]; tab40[20 + q] = invfact[q]; }
          const std::vector<double> bq = pb2::u8q_tables(h->plan, d.m, tab40);
          PB2_CUDA_H(cudaMalloc(&h->dTablesQ, bq.size() * sizeof(double)));
          PB2_CUDA_H(cudaMemcpy(h->dTablesQ, bq.data(), bq.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        PB2_CUDA_H(cudaMalloc(&h->dSyncWords, 2 * sizeof(unsigned long long)));
        PB2_CUDA_H(cudaMemset(h->dSyncWords, 0, 2 * sizeof(unsigned long long)));
        h->u8p_ok = true;
      }
    }
    // the tensor-core Hessian additionally uses E^T = exp(-dt G): every generator anti-symmetric
    // (true for the isomorphism of any Hermitian Hamiltonian, isomorphisms.jl:350,359)
    bool antisym = true;
    for (int mat = 0; mat <= d.m && antisym; ++mat) {
      const double* A = mat == 0 ? h->G0.data() : h->Gj.data() + (size_t)(mat - 1) * bb;
      for (int i = 0; i < d.b && antisym; ++i)
        for (int j = 0; j <= i; ++j)
          if (A[i + (size_t)j * d.b] != -A[j + (size_t)i * d.b]) { antisym = false; break; }
    }
    h->u8h_ok = h->u8_ok && antisym && d.m <= 4 && !std::getenv("PB2_NO_U8H");
    if (h->u8h_ok)
      PB2_CUDA_H(cudaFuncSetAttribute(pb2::u8h_kernel(h->plan.W), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemLimit));
#ifdef PB2_TRACE
    PB2_CUDA_H(cudaMalloc(&h->dTrace, 8 * 16 * 8 * sizeof(long long)));
    PB2_CUDA_H(cudaMemset(h->dTrace, 0, 8 * 16 * 8 * sizeof(long long)));
    {
      long long* t2 = nullptr;
      PB2_CUDA_H(cudaMalloc(&t2, 16 * 20 * 4 * sizeof(long long)));
      PB2_CUDA_H(cudaMemset(t2, 0, 16 * 20 * 4 * sizeof(long long)));
      PB2_CUDA_H(cudaMemcpyToSymbol(pb2::g_trace2, &t2, sizeof(t2)));
      h->dTrace2 = t2;
    }
    PB2_CUDA_H(cudaMalloc(&h->dTrace3, (size_t)64 * 148 * 16 * 40 * sizeof(long long)));
    PB2_CUDA_H(cudaMemset(h->dTrace3, 0, (size_t)64 * 148 * 16 * 40 * sizeof(long long)));
#endif
  }
#undef PB2_CUDA_H
  *out = h;
  return PB2_OK;
}

void pb2_destroy(pb2_handle* h) {
  if (!h) return;
  DeviceGuard guard_d(h->d.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (double* p : {h->dG0, h->dGj, h->dZ, h->dDelta, h->dJac, h->dMu, h->dHess})
    if (p) cudaFree(p);
  if (h->dGfrag) cudaFree(h->dGfrag);
  for (void* q : {(void*)h->dHGfrag, (void*)h->dHGfragT, (void*)h->dHNorms, (void*)h->dHEll, (void*)h->dHEllT})
    if (q) cudaFree(q);
  if (h->dEll) cudaFree(h->dEll);
  if (h->dNorms) cudaFree(h->dNorms);
  if (h->dTab) cudaFree(h->dTab);
  if (h->dTables) cudaFree(h->dTables);
  if (h->dTablesP) cudaFree(h->dTablesP);
  if (h->dTablesQ) cudaFree(h->dTablesQ);
  if (h->dSyncWords) cudaFree(h->dSyncWords);
  if (h->dCanon) cudaFree(h->dCanon);
  for (double* q : {h->dCoef, h->dCoefDot, h->dZs})
    if (q) cudaFree(q);
  for (double* q : {h->dRoJac, h->dRoStates, h->dRoX0, h->dRoOut})
    if (q) cudaFree(q);
  if (h->dComp) cudaFree(h->dComp);
  if (h->hComp) cudaFreeHost(h->hComp);
  for (cudaEvent_t e : h->chunk_ev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->in_ev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->k_ev)
    if (e) cudaEventDestroy(e);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->dTrace) cudaFree(h->dTrace);
  if (h->dTrace2) cudaFree(h->dTrace2);
  if (h->dTrace3) cudaFree(h->dTrace3);
  for (double* p : {h->hZ, h->hDelta, h->hJac, h->hMu, h->hHess})
    if (p) cudaFreeHost(p);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

int64_t pb2_dim(const pb2_handle* h) { return h ? (int64_t)h->n_x() * h->nk() : -1; }
int64_t pb2_nnz_jac(const pb2_handle* h) { return h ? (int64_t)h->nnz_jac_user_knot() * h->nk() : -1; }
int64_t pb2_nnz_hess(const pb2_handle* h) { return h ? (int64_t)h->nnz_hess_knot() * h->nk() : -1; }
// which kernel evaluates the Lagrangian Hessian: 1 = jet kernel, 2 = general tensor-core kernel, 3 = 3-qubit kernel
int32_t pb2_hessian_algorithm(const pb2_handle* h) {
  if (!h) return -1;
  if (h->d.time_dependent) return 0;
  if (h->u8h_ok && !(h->hplan.ok && h->hess_prefer_dmmah) && (h->d.D % 2 == 0) && (h->d.x_off % 2 == 0)) return 3;
  return h->hplan.ok ? 2 : 1;
}

int32_t pb2_algorithm(const pb2_handle* h) { return h ? h->alg : -1; }
int64_t pb2_launch_count(const pb2_handle* h) { return h ? h->launches : -1; }
#ifdef PB2_TRACE
extern "C" int pb2_debug_trace(pb2_handle* h, long long* out) {   // 8 knots x 16 warps x 8 stamps
  if (!h || !h->dTrace) return PB2_EINVAL;
  cudaDeviceSynchronize();
  return cudaMemcpy(out, h->dTrace, 8 * 16 * 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : PB2_ECUDA;
}
extern "C" int pb2_debug_trace3(pb2_handle* h, long long* out) {   // 64 launches x 148 blocks x 16 warps x 8 stamps
  if (!h || !h->dTrace3) return PB2_EINVAL;
  cudaDeviceSynchronize();
  const size_t n = (size_t)64 * 148 * 16 * 40 * sizeof(long long);
  if (cudaMemcpy(out, h->dTrace3, n, cudaMemcpyDeviceToHost) != cudaSuccess) return PB2_ECUDA;
  cudaMemset(h->dTrace3, 0, n);
  h->trace_launch = 0;
  return 0;
}
extern "C" int pb2_debug_trace2(pb2_handle* h, long long* out) {   // 16 warps x 20 steps x 4 stamps (last knot)
  if (!h || !h->dTrace2) return PB2_EINVAL;
  cudaDeviceSynchronize();
  return cudaMemcpy(out, h->dTrace2, 16 * 20 * 4 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : PB2_ECUDA;
}
#endif

int pb2_structure_jac(const pb2_handle* h, int64_t* rows, int64_t* cols) {
  if (check(h)) return PB2_EINVAL;
  if (!rows || !cols) return fail(PB2_EINVAL, "pb2_structure_jac: null output");
  const pb2_desc& d = h->d;
  const int64_t b = d.b, n_b = d.n_b, m = d.m, n_x = h->n_x(), D = d.D;
  int64_t o = 0;
  for (int64_t kl = 0; kl < h->nk(); ++kl) {
    const int64_t k = d.knot0 + kl;
    const int64_t r0 = k * n_x + 1, c0 = k * D + 1;
    if (d.dense_blocks) {
      for (int64_t j = 0; j < n_x; ++j)
        for (int64_t i = 0; i < n_x; ++i) {
          rows[o] = r0 + i;
          cols[o++] = c0 + d.x_off + j;
        }
    } else
    for (int64_t c = 0; c < n_b; ++c)
      for (int64_t j = 0; j < b; ++j)
        for (int64_t i = 0; i < b; ++i) {
          rows[o] = r0 + c * b + i;
          cols[o++] = c0 + d.x_off + c * b + j;
        }
    for (int64_t j = 0; j < m; ++j)
      for (int64_t i = 0; i < n_x; ++i) {
        rows[o] = r0 + i;
        cols[o++] = c0 + d.u_off + j;
      }
    for (int64_t i = 0; i < n_x; ++i) {
      rows[o] = r0 + i;
      cols[o++] = c0 + d.dt_off;
    }
    for (int64_t i = 0; i < n_x; ++i) {
      rows[o] = r0 + i;
      cols[o++] = c0 + D + d.x_off + i;
    }
    if (d.time_dependent)
      for (int64_t i = 0; i < n_x; ++i) {
        rows[o] = r0 + i;
        cols[o++] = c0 + d.t_off;
      }
  }
  return PB2_OK;
}

int pb2_structure_hess(const pb2_handle* h, int64_t* rows, int64_t* cols) {
  if (check(h)) return PB2_EINVAL;
  if (!rows || !cols) return fail(PB2_EINVAL, "pb2_structure_hess: null output");
  const pb2_desc& d = h->d;
  const int64_t m = d.m, n_x = h->n_x(), D = d.D;
  int64_t o = 0;
  auto emit = [&](int64_t a, int64_t c) {
    rows[o] = std::min(a, c);
    cols[o++] = std::max(a, c);
  };
  for (int64_t kl = 0; kl < h->nk(); ++kl) {
    const int64_t c0 = (d.knot0 + kl) * D + 1;
    for (int64_t j = 0; j < m; ++j)
      for (int64_t i = 0; i < n_x; ++i) emit(c0 + d.x_off + i, c0 + d.u_off + j);
    for (int64_t i = 0; i < n_x; ++i) emit(c0 + d.x_off + i, c0 + d.dt_off);
    for (int64_t j = 0; j < m; ++j)
      for (int64_t i = 0; i <= j; ++i) emit(c0 + d.u_off + i, c0 + d.u_off + j);
    for (int64_t j = 0; j < m; ++j) emit(c0 + d.u_off + j, c0 + d.dt_off);
    emit(c0 + d.dt_off, c0 + d.dt_off);
  }
  return PB2_OK;
}

int pb2_residual_jacobian_async(pb2_handle* h, const double* dZ, double* ddelta, double* dvals,
                                void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (!dZ) return fail(PB2_EINVAL, "pb2_residual_jacobian_async: null Z");
  DeviceGuard guard_2(h->d.device);
  return launch_resjac(h, dZ, ddelta, dvals, (cudaStream_t)stream);
}

int pb2_hess_lagrangian_async(pb2_handle* h, const double* dZ, const double* dmu, double* dvals,
                              void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (!dZ || !dmu || !dvals) return fail(PB2_EINVAL, "pb2_hess_lagrangian_async: null argument");
  DeviceGuard guard_3(h->d.device);
  return launch_hess(h, dZ, dmu, dvals, (cudaStream_t)stream);
}

int64_t pb2_compact_stride(const pb2_handle* h) {
  if (!h || !(h->alg == PB2_ALG_DMMA && h->u8_ok) || (h->d.D % 2) || (h->d.x_off % 2) || h->d.time_dependent || h->d.dense_blocks) return 0;
  if (h->u8p_ok && h->u8q) {
    // a knot column too long for the slab-staging kernels (even with four knots per CTA) goes through the general
    // kernel, which writes canonical arrays only
    pb2::U8qParams probe{};
    probe.m = h->d.m; probe.zlen = h->d.D + h->d.x_off + 128; probe.n_peers = 8; probe.cstride = (h->d.m + 3) * 128;
    if (pb2::u8q_layout(probe, 4) > kSmemLimit / 2) return 0;
  }
  return (int64_t)(h->d.m + 3) * 128;
}

int pb2_residual_jacobian_compact_async(pb2_handle* h, const double* dZ, double* dcompact, void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (!dZ || !dcompact) return fail(PB2_EINVAL, "pb2_residual_jacobian_compact_async: null argument");
  if (pb2_compact_stride(h) == 0) return fail(PB2_EINVAL, "pb2_residual_jacobian_compact_async: unsupported for this handle");
  DeviceGuard guard_4(h->d.device);
  return launch_resjac(h, dZ, nullptr, dcompact, (cudaStream_t)stream, 1);
}

int pb2_residual_jacobian_exchange_async(pb2_handle* h, const double* dZ, int32_t n_ranks, int32_t rank,
                                         double* const* gather_bufs, int64_t slot_offset, void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (!dZ || !gather_bufs || n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks || slot_offset < 0)
    return fail(PB2_EINVAL, "pb2_residual_jacobian_exchange_async: bad argument");
  if (pb2_compact_stride(h) == 0)
    return fail(PB2_EINVAL, "pb2_residual_jacobian_exchange_async: unsupported for this handle");
  for (int r = 0; r < n_ranks; ++r)
    if (!gather_bufs[r] || ((uintptr_t)gather_bufs[r] % 16) != 0)
      return fail(PB2_EINVAL, "pb2_residual_jacobian_exchange_async: gather buffers must be 16-byte aligned device pointers");
  DeviceGuard guard_5(h->d.device);
  return launch_resjac(h, dZ, nullptr, gather_bufs[rank] + slot_offset, (cudaStream_t)stream, 1, n_ranks, gather_bufs, rank);
}

int pb2_residual_jacobian_exchange_sync_async(pb2_handle* h, const double* dZ, int32_t n_ranks, int32_t rank,
                                              double* const* gather_bufs, int64_t slot_offset, int64_t flag_offset,
                                              void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (flag_offset < 0 || (flag_offset % 2) != 0)
    return fail(PB2_EINVAL, "pb2_residual_jacobian_exchange_sync_async: flag_offset must be a non-negative even number of doubles");
  if (!(h->u8p_ok && h->u8q) || std::getenv("PB2_U8Q_NO_PEERS"))
    return fail(PB2_EINVAL, "pb2_residual_jacobian_exchange_sync_async: the in-kernel step barrier needs the small-CTA kernel");
  h->xchg_flag_off = flag_offset;
  const int rc = pb2_residual_jacobian_exchange_async(h, dZ, n_ranks, rank, gather_bufs, slot_offset, stream);
  h->xchg_flag_off = -1;
  return rc;
}

int pb2_enable_peer_access(int32_t device, int32_t peer) {
  DeviceGuard guard_6(device);
  int can = 0;
  PB2_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
  if (!can) return fail(PB2_EINVAL, "pb2_enable_peer_access: no peer access between these devices");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
    return fail(PB2_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
  cudaGetLastError();
  return PB2_OK;
}

int pb2_expand_compact_async(pb2_handle* h, const double* dcompact, int64_t n_knots, double* ddelta, double* dvals,
                             void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (!dcompact || !dvals || n_knots < 0) return fail(PB2_EINVAL, "pb2_expand_compact_async: bad argument");
  const int64_t cs = pb2_compact_stride(h);
  if (cs == 0) return fail(PB2_EINVAL, "pb2_expand_compact_async: unsupported for this handle");
  if (n_knots == 0) return PB2_OK;
  DeviceGuard guard_7(h->d.device);
  const int n_x = h->n_x();
  const int blocks = (int)std::min<int64_t>(n_knots, (int64_t)h->n_sm * 8);
  pb2::expand_compact_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dcompact, n_knots, (int)cs, (h->d.m + 1) * n_x,
                                                                     h->nnz_jac_knot(), ddelta, dvals);
  PB2_CUDA(cudaGetLastError());
  h->launches++;
  return PB2_OK;
}

void* pb2_stream(const pb2_handle* h) { return h ? (void*)h->stream : nullptr; }

int pb2_set_option(pb2_handle* h, int32_t option, int64_t value) {
  if (check(h)) return PB2_EINVAL;
  switch (option) {
    case PB2_OPT_EARLY_Z: h->early_z = value != 0; return PB2_OK;
    case PB2_OPT_PIPELINED: h->pipelined = value != 0; return PB2_OK;
    case PB2_OPT_HESSIAN_CTAS:
      if (value < 0) return fail(PB2_EINVAL, "pb2_set_option: PB2_OPT_HESSIAN_CTAS needs a count >= 0");
      h->hess_ctas = (int)std::min<int64_t>(value, 1 << 20);
      return PB2_OK;
    default: return fail(PB2_EINVAL, "pb2_set_option: unknown option");
  }
}

int pb2_sync(pb2_handle* h) {
  if (check(h)) return PB2_EINVAL;
  DeviceGuard guard_8(h->d.device);
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  return PB2_OK;
}

static int ensure(double** dev, double** host, size_t n) {
  if (n == 0) n = 1;
  if (dev && !*dev) PB2_CUDA(cudaMalloc(dev, n * sizeof(double)));
  if (host && !*host) PB2_CUDA(cudaMallocHost(host, n * sizeof(double)));
  return PB2_OK;
}

int pb2_residual_jacobian(pb2_handle* h, const double* Z, double* delta, double* vals, int space) {
  if (check(h)) return PB2_EINVAL;
  if (!Z) return fail(PB2_EINVAL, "pb2_residual_jacobian: null Z");
  DeviceGuard guard_9(h->d.device);
  if (space == PB2_DEVICE) {
    int rc = launch_resjac(h, Z, delta, vals, h->stream);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_residual_jacobian: bad space");
  const size_t nZ = (size_t)h->d.D * h->d.K, nD = (size_t)pb2_dim(h), nJ = (size_t)pb2_nnz_jac(h);
  int rc;
  if (vals && h->nk() > 0 && pb2_compact_stride(h) != 0 && !std::getenv("PB2_NO_COMPACT_D2H")) {
    // 3-qubit unitary shape: the d/dx_k block is n_b copies of one b x b block, so only the compact
    // record (9.2 KB instead of 23.5 KB per knot at C3) crosses PCIe; it lands in pinned memory in
    // chunks, and host threads replicate each chunk into the caller's arrays (pb2_structure_jac
    // order) while the next chunk is still in flight.  Data movement only.
    const int64_t nk = h->nk(), cs = pb2_compact_stride(h);   // record: [E cols 0..7 | jets, d/d dt | delta]
    const int n_x = h->n_x(), bb = h->d.b * h->d.b, n_b = h->d.n_b, nJd = (h->d.m + 1) * n_x, nnz = h->nnz_jac_knot();
    if ((rc = ensure(&h->dZ, &h->hZ, nZ))) return rc;
    if ((rc = ensure(&h->dComp, &h->hComp, (size_t)(nk * cs)))) return rc;
    // Chunk boundaries kb[0..nch]: equal chunks, 8 by default (PB2_D2H_CHUNKS=n overrides).  Measured on C3
    // (tools/e2e_timeline.py, profiles/r02_e2e_timeline.txt): the records cross PCIe back to back at 45-51 GB/s
    // (158 us for 7.2 MB) from ~26 us after the call; the host-side replication into the caller's 23.5 MB of COO
    // values runs at 130-170 GB/s on 16 threads and finishes ~28 us after the last chunk lands.  A small first chunk,
    // small last chunks, 6, 12 or 16 chunks were all measured and are no better (per-copy overhead, replication lag).
    const int nch_env = std::getenv("PB2_D2H_CHUNKS") ? std::atoi(std::getenv("PB2_D2H_CHUNKS")) : 8;
    const int nch = (int)std::min<int64_t>(std::min(16, std::max(1, nch_env)), std::max<int64_t>(1, nk / 32));
    int64_t kb[17];
    {
      const int64_t per = (nk + nch - 1) / nch;
      for (int c = 0; c <= nch; ++c) kb[c] = std::min(nk, c * per);
    }
    static const bool piped = !(std::getenv("PB2_E2E_PIPE") && std::atoi(std::getenv("PB2_E2E_PIPE")) == 0);
    const bool tl = std::getenv("PB2_E2E_TIMELINE") != nullptr;     // host-clock stamps on stderr (measurement aid)
    const auto t0 = std::chrono::steady_clock::now();
    auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
    double t_enq = 0.0, t_chunk[16] = {}, t_pool = 0.0;
    if (piped && nch > 1) {
      // chunk c: upload its knot columns (s_in) -> evaluate its knots (h->stream) -> download its records (s_out).
      // The download of the records is the long pole (PCIe, 7 KB per knot); everything else hides under it, and
      // the first records leave ~one chunk's upload + kernel after the call instead of after the whole kernel.
      if (!h->s_in) PB2_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
      if (!h->s_out) PB2_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
      const bool direct_in = is_pinned_or_device(Z);
      const int D = h->d.D;
      static const int h2d_mode = std::getenv("PB2_E2E_H2D") ? std::atoi(std::getenv("PB2_E2E_H2D")) : 0;
      for (int c = 0; c < nch; ++c) {
        const int64_t k0 = kb[c], k1 = kb[c + 1];
        for (cudaEvent_t* e : {&h->chunk_ev[c], &h->in_ev[c], &h->k_ev[c]})
          if (!*e) PB2_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        if (k1 > k0) {
          // knots k0 .. k1-1 read columns k0 .. k1; column k0 came with the previous chunk.
          // h2d_mode 0: every chunk's columns on s_in;  1: on the compute stream (no events);  2: chunk 0 alone,
          // then all the rest in one copy (two uploads per call whatever the number of chunks)
          int64_t c0 = c == 0 ? 0 : k0 + 1, c1 = k1 + 1;
          if (h2d_mode == 2) {
            if (c == 1) c1 = nk + 1;
            if (c >= 2) c1 = c0;
          }
          if (c1 > c0) {
            const double* from = Z + c0 * D;
            if (!direct_in) {
              std::memcpy(h->hZ + c0 * D, from, (size_t)(c1 - c0) * D * sizeof(double));
              from = h->hZ + c0 * D;
            }
            cudaStream_t sh = h2d_mode == 1 ? h->stream : h->s_in;
            PB2_CUDA(cudaMemcpyAsync(h->dZ + c0 * D, from, (size_t)(c1 - c0) * D * sizeof(double), cudaMemcpyHostToDevice, sh));
            if (h2d_mode != 1) {
              PB2_CUDA(cudaEventRecord(h->in_ev[c], h->s_in));
              PB2_CUDA(cudaStreamWaitEvent(h->stream, h->in_ev[c], 0));
            }
          }
          h->nk_sub = k1 - k0;
          rc = launch_resjac(h, h->dZ + k0 * D, nullptr, h->dComp + k0 * cs, h->stream, 1, 0, nullptr, 0, 1);
          h->nk_sub = -1;
          if (rc) {
            // earlier chunks are in flight and read / write the caller's buffers: drain them before reporting
            const std::string msg = g_err;
            cudaStreamSynchronize(h->s_in);
            cudaStreamSynchronize(h->stream);
            cudaStreamSynchronize(h->s_out);
            cudaGetLastError();
            return fail(rc, msg);
          }
          PB2_CUDA(cudaEventRecord(h->k_ev[c], h->stream));
          PB2_CUDA(cudaStreamWaitEvent(h->s_out, h->k_ev[c], 0));
          PB2_CUDA(cudaMemcpyAsync(h->hComp + k0 * cs, h->dComp + k0 * cs, (size_t)(k1 - k0) * cs * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->s_out));
        }
        PB2_CUDA(cudaEventRecord(h->chunk_ev[c], h->s_out));
      }
    } else {
      if ((rc = stage_in(h, Z, h->hZ, h->dZ, nZ))) return rc;
      if ((rc = launch_resjac(h, h->dZ, nullptr, h->dComp, h->stream, 1, 0, nullptr, 0, 1))) return rc;
      for (int c = 0; c < nch; ++c) {
        const int64_t k0 = kb[c], k1 = kb[c + 1];
        if (!h->chunk_ev[c]) PB2_CUDA(cudaEventCreateWithFlags(&h->chunk_ev[c], cudaEventDisableTiming));
        if (k1 > k0)
          PB2_CUDA(cudaMemcpyAsync(h->hComp + k0 * cs, h->dComp + k0 * cs, (size_t)(k1 - k0) * cs * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->stream));
        PB2_CUDA(cudaEventRecord(h->chunk_ev[c], h->stream));
      }
    }
    if (tl) t_enq = us();
    std::atomic<int> ready[16];
    for (auto& r : ready) r.store(0, std::memory_order_relaxed);
    std::atomic<int> failed{0};
    const double* comp = h->hComp;
    // non-temporal stores pay once the caller's arrays no longer fit the host's caches (C5: 188 MB): no read-for-
    // ownership traffic; below that they only compete with the DMA writes (C3: slower).  PB2_HOST_NT=0/1 overrides.
    const bool nt = std::getenv("PB2_HOST_NT") ? std::atoi(std::getenv("PB2_HOST_NT")) != 0
                                               : (size_t)nk * (size_t)nnz * sizeof(double) > ((size_t)96 << 20);
    auto expand = [&](int64_t a, int64_t b) {
      alignas(16) double E[256];
      alignas(16) double ones[128];
      for (double& o : ones) o = 1.0;
      for (int64_t k = a; k < b; ++k) {
        int c = 0;
        while (k >= kb[c + 1]) ++c;
        while (!ready[c].load(std::memory_order_acquire))
          if (failed.load(std::memory_order_relaxed)) return;
        const double* src = comp + k * cs;
        // -E = -[[P, -Q], [Q, P]]: column c + 8 is [-(rows 8..15 of column c); rows 0..7 of column c]
        std::memcpy(E, src, 128 * sizeof(double));
        for (int col = 0; col < 8; ++col)
          for (int r = 0; r < 8; ++r) {
            E[(col + 8) * 16 + r] = -src[col * 16 + 8 + r];
            E[(col + 8) * 16 + 8 + r] = src[col * 16 + r];
          }
        double* out = vals + k * (int64_t)nnz;
        for (int cp = 0; cp < n_b; ++cp) host_copy(out + (size_t)cp * bb, E, bb, nt);
        out += (size_t)n_b * bb;
        host_copy(out, src + 128, nJd, nt);
        host_copy(out + nJd, ones, n_x, nt);               // the constant d/dx_{k+1} identity entries
        if (delta) host_copy(delta + k * n_x, src + 128 + nJd, n_x, nt);
      }
      if (nt) _mm_sfence();
    };
    // the pool's threads start un-packing as chunks land; this thread releases the chunks, then joins in
    std::function<void(int64_t, int64_t)> fn = expand;
    pb2::HostPool& pool = pb2::HostPool::instance();
    pool.begin(nk, 4, fn);
    for (int c = 0; c < nch; ++c) {
      if (cudaEventSynchronize(h->chunk_ev[c]) != cudaSuccess) { failed.store(1); break; }
      ready[c].store(1, std::memory_order_release);
      if (tl) t_chunk[c] = us();
    }
    pool.finish();
    if (tl) {
      t_pool = us();
      std::fprintf(stderr, "pb2 e2e timeline [us]: enqueued %.1f | chunks landed", t_enq);
      for (int c = 0; c < nch; ++c) std::fprintf(stderr, " %.1f", t_chunk[c]);
      std::fprintf(stderr, " | replicated %.1f\n", t_pool);
    }
    if (failed.load()) {
      cudaStreamSynchronize(h->stream);
      if (h->s_out) cudaStreamSynchronize(h->s_out);
      return fail(PB2_ECUDA, "pb2_residual_jacobian: device-to-host copy failed");
    }
    // (the last chunk's event completed: every upload, kernel and download of this call has)
    return PB2_OK;
  }
  if ((rc = ensure(&h->dZ, &h->hZ, nZ))) return rc;
  if (delta && (rc = ensure(&h->dDelta, &h->hDelta, nD))) return rc;
  if (vals && (rc = ensure(&h->dJac, &h->hJac, nJ))) return rc;
  if ((rc = stage_in(h, Z, h->hZ, h->dZ, nZ))) return rc;
  if ((rc = launch_resjac(h, h->dZ, delta ? h->dDelta : nullptr, vals ? h->dJac : nullptr, h->stream, 0, 0, nullptr, 0, 1)))
    return rc;
  PendingOut po1{}, po2{};
  if (delta && (rc = stage_out_begin(h, delta, h->hDelta, h->dDelta, nD, po1))) return rc;
  if (vals && (rc = stage_out_begin(h, vals, h->hJac, h->dJac, nJ, po2))) return rc;
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  stage_out_finish(po1);
  stage_out_finish(po2);
  return PB2_OK;
}

int pb2_residual(pb2_handle* h, const double* Z, double* delta, int space) {
  if (!delta) return fail(PB2_EINVAL, "pb2_residual: null output");
  return pb2_residual_jacobian(h, Z, delta, nullptr, space);
}

int pb2_jacobian(pb2_handle* h, const double* Z, double* vals, int space) {
  if (!vals) return fail(PB2_EINVAL, "pb2_jacobian: null output");
  return pb2_residual_jacobian(h, Z, nullptr, vals, space);
}

int pb2_hess_lagrangian(pb2_handle* h, const double* Z, const double* mu, double* vals, int space) {
  if (check(h)) return PB2_EINVAL;
  if (!Z || !mu || !vals) return fail(PB2_EINVAL, "pb2_hess_lagrangian: null argument");
  DeviceGuard guard_10(h->d.device);
  if (space == PB2_DEVICE) {
    int rc = launch_hess(h, Z, mu, vals, h->stream);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_hess_lagrangian: bad space");
  const size_t nZ = (size_t)h->d.D * h->d.K, nD = (size_t)pb2_dim(h), nH = (size_t)pb2_nnz_hess(h);
  int rc;
  if ((rc = ensure(&h->dZ, &h->hZ, nZ))) return rc;
  if ((rc = ensure(&h->dMu, &h->hMu, nD))) return rc;
  if ((rc = ensure(&h->dHess, &h->hHess, nH))) return rc;
  if ((rc = stage_in(h, Z, h->hZ, h->dZ, nZ))) return rc;
  if ((rc = stage_in(h, mu, h->hMu, h->dMu, nD))) return rc;
  if ((rc = launch_hess(h, h->dZ, h->dMu, h->dHess, h->stream))) return rc;
  PendingOut po{};
  if ((rc = stage_out_begin(h, vals, h->hHess, h->dHess, nH, po))) return rc;
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  stage_out_finish(po);
  return PB2_OK;
}

int pb2_set_time_coefficients(pb2_handle* h, const double* c, const double* cdot, int space) {
  if (check(h)) return PB2_EINVAL;
  if (!h->d.time_dependent) return fail(PB2_EINVAL, "pb2_set_time_coefficients: not a time-dependent handle");
  if (!c || !cdot || (space != PB2_HOST && space != PB2_DEVICE)) return fail(PB2_EINVAL, "pb2_set_time_coefficients: bad argument");
  DeviceGuard guard(h->d.device);
  const size_t n = (size_t)h->d.m * h->d.K * sizeof(double);
  const cudaMemcpyKind kind = space == PB2_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  if (n) {
    PB2_CUDA(cudaMemcpyAsync(h->dCoef, c, n, kind, h->stream));
    PB2_CUDA(cudaMemcpyAsync(h->dCoefDot, cdot, n, kind, h->stream));
    PB2_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->coef_set = true;
  return PB2_OK;
}

static int aux_ensure(double** dev, size_t n) {
  if (!*dev) PB2_CUDA(cudaMalloc(dev, std::max<size_t>(n, 1) * sizeof(double)));
  return PB2_OK;
}

// ---- ensembles: all members in one launch ------------------------------------------------------------------
struct pb2_batch {
  std::vector<pb2_handle*> mem;
  bool fused = false;
  BatchLaunch bl;
  int* dXoffs = nullptr;
  double *dG0 = nullptr, *dGj = nullptr, *dGfrag = nullptr, *dNorms = nullptr;
  pb2::EllEntry* dEll = nullptr;
  double *dHGfrag = nullptr, *dHGfragT = nullptr, *dHNorms = nullptr;   // tensor-core Hessian tables, member-major
  pb2::EllEntry *dHEll = nullptr, *dHEllT = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> done;
  cudaEvent_t fork = nullptr;
  double *dZ = nullptr, *dDelta = nullptr, *dJac = nullptr, *dMu = nullptr, *dHess = nullptr;
};

void pb2_batch_destroy(pb2_batch* b) {
  if (!b) return;
  if (!b->mem.empty()) {
    DeviceGuard guard(b->mem[0]->d.device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (void* q : {(void*)b->dHGfrag, (void*)b->dHGfragT, (void*)b->dHNorms, (void*)b->dHEll, (void*)b->dHEllT})
      if (q) cudaFree(q);
    for (void* q : {(void*)b->dXoffs, (void*)b->dG0, (void*)b->dGj, (void*)b->dGfrag, (void*)b->dNorms, (void*)b->dEll,
                    (void*)b->dZ, (void*)b->dDelta, (void*)b->dJac, (void*)b->dMu, (void*)b->dHess})
      if (q) cudaFree(q);
    for (cudaEvent_t e : b->done)
      if (e) cudaEventDestroy(e);
    if (b->fork) cudaEventDestroy(b->fork);
    if (b->stream) cudaStreamDestroy(b->stream);
  }
  for (pb2_handle* h : b->mem) pb2_destroy(h);
  delete b;
}

int pb2_batch_create(const pb2_desc* descs, int32_t n, pb2_batch** out) {
  if (!descs || !out || n < 1 || n > 65535) return fail(PB2_EINVAL, "pb2_batch_create: bad argument (1 .. 65535 members)");
  *out = nullptr;
  const pb2_desc& d0 = descs[0];
  for (int i = 1; i < n; ++i) {
    const pb2_desc& d = descs[i];
    if (d.kind != d0.kind || d.b != d0.b || d.n_b != d0.n_b || d.m != d0.m || d.K != d0.K || d.D != d0.D ||
        d.dt_off != d0.dt_off || d.u_off != d0.u_off || d.global_dim != d0.global_dim || d.device != d0.device ||
        d.algorithm != d0.algorithm || d.knot0 != d0.knot0)
      return fail(PB2_EINVAL, "pb2_batch_create: members must agree in everything but x_off and the generators");
  }
  for (int i = 0; i < n; ++i)
    if (descs[i].time_dependent || descs[i].dense_blocks)
      return fail(PB2_EINVAL, "pb2_batch_create: time-dependent / dense_blocks members are not supported");
  pb2_batch* b = new (std::nothrow) pb2_batch();
  if (!b) return fail(PB2_ENOMEM, "pb2_batch_create: out of memory");
  for (int i = 0; i < n; ++i) {
    pb2_handle* h = nullptr;
    const int rc = pb2_create(&descs[i], &h);
    if (rc) {
      const std::string msg = g_err;
      pb2_batch_destroy(b);
      return fail(rc, msg);
    }
    b->mem.push_back(h);
  }
  DeviceGuard guard(d0.device);
  auto bail = [&](const char* what) {
    pb2_batch_destroy(b);
    return fail(PB2_ECUDA, std::string("pb2_batch_create: ") + what);
  };
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
  if (cudaEventCreateWithFlags(&b->fork, cudaEventDisableTiming) != cudaSuccess) return bail("event");
  b->done.assign(n, nullptr);
  for (int i = 0; i < n; ++i)
    if (cudaEventCreateWithFlags(&b->done[i], cudaEventDisableTiming) != cudaSuccess) return bail("event");
  // can the members share one launch?  same algorithm, same tensor-core plan shape, not the 3-qubit kernels
  const pb2_handle* h0 = b->mem[0];
  bool same = true;
  for (const pb2_handle* h : b->mem) {
    same = same && h->alg == h0->alg && !h->u8_ok;
    if (h->alg == PB2_ALG_DMMA)
      same = same && h->plan.NT == h0->plan.NT && h->plan.W == h0->plan.W && h->plan.iso == h0->plan.iso &&
             h->plan.ncT == h0->plan.ncT && h->plan.tiles_full == h0->plan.tiles_full;
  }
  if (std::getenv("PB2_BATCH_NO_FUSE")) same = false;
  b->fused = same && n > 1;
  if (b->fused) {
    const size_t bb = (size_t)d0.b * d0.b;
    std::vector<int> xo(n);
    std::vector<double> G0((size_t)n * bb), Gj((size_t)n * std::max<size_t>(1, (size_t)d0.m * bb));
    int mx = 0;
    for (int i = 0; i < n; ++i) {
      xo[i] = descs[i].x_off;
      mx = std::max(mx, xo[i]);
      std::copy(b->mem[i]->G0.begin(), b->mem[i]->G0.end(), G0.begin() + (size_t)i * bb);
      std::copy(b->mem[i]->Gj.begin(), b->mem[i]->Gj.end(), Gj.begin() + (size_t)i * d0.m * bb);
    }
    auto up = [&](void** dst, const void* src, size_t bytes) {
      return cudaMalloc(dst, std::max<size_t>(bytes, 8)) == cudaSuccess &&
             (bytes == 0 || cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess);
    };
    if (!up((void**)&b->dXoffs, xo.data(), n * sizeof(int)) || !up((void**)&b->dG0, G0.data(), G0.size() * sizeof(double)) ||
        !up((void**)&b->dGj, Gj.data(), (size_t)n * d0.m * bb * sizeof(double)))
      return bail("upload");
    BatchLaunch& bl = b->bl;
    bl.n = n; bl.max_xoff = mx; bl.x_offs = b->dXoffs; bl.G0 = b->dG0; bl.Gj = b->dGj;
    bl.mem_G0 = (long long)bb; bl.mem_Gj = (long long)d0.m * bb;
    bl.mem_delta = pb2_dim(h0); bl.mem_jac = pb2_nnz_jac(h0); bl.mem_hess = pb2_nnz_hess(h0);
    if (h0->alg == PB2_ALG_DMMA) {
      const size_t ng = h0->plan.gfrag.size(), ne = h0->plan.ell.size(), nn = h0->plan.norms.size();
      std::vector<double> gf(n * ng), nr(n * nn);
      std::vector<pb2::EllEntry> el(n * ne);
      for (int i = 0; i < n; ++i) {
        std::copy(b->mem[i]->plan.gfrag.begin(), b->mem[i]->plan.gfrag.end(), gf.begin() + i * ng);
        std::copy(b->mem[i]->plan.norms.begin(), b->mem[i]->plan.norms.end(), nr.begin() + i * nn);
        std::copy(b->mem[i]->plan.ell.begin(), b->mem[i]->plan.ell.end(), el.begin() + i * ne);
      }
      if (!up((void**)&b->dGfrag, gf.data(), gf.size() * sizeof(double)) || !up((void**)&b->dNorms, nr.data(), nr.size() * sizeof(double)) ||
          !up((void**)&b->dEll, el.data(), el.size() * sizeof(pb2::EllEntry)))
        return bail("upload");
      bl.gfrag = b->dGfrag; bl.norms = b->dNorms; bl.ell = b->dEll;
      bl.mem_gfrag = (long long)ng; bl.mem_norms = (long long)nn; bl.mem_ell = (long long)ne;
    }
    // the Hessian shares a launch on the tensor cores when every member has the same tile plan
    bool hsame = h0->hplan.ok;
    for (const pb2_handle* h : b->mem)
      hsame = hsame && h->hplan.ok && h->hplan.NT == h0->hplan.NT && h->hplan.W == h0->hplan.W &&
              h->hplan.tiles_f == h0->hplan.tiles_f && h->hplan.tiles_a == h0->hplan.tiles_a;
    if (hsame) {
      const size_t ng = h0->hplan.gfrag.size(), ne = h0->hplan.ell.size(), nn = h0->hplan.norms.size();
      std::vector<double> gf(n * ng), gt(n * ng), nr(n * nn);
      std::vector<pb2::EllEntry> el(n * ne), et(n * ne);
      for (int i = 0; i < n; ++i) {
        const pb2::DmmahPlan& hp = b->mem[i]->hplan;
        std::copy(hp.gfrag.begin(), hp.gfrag.end(), gf.begin() + i * ng);
        std::copy(hp.gfragT.begin(), hp.gfragT.end(), gt.begin() + i * ng);
        std::copy(hp.norms.begin(), hp.norms.end(), nr.begin() + i * nn);
        std::copy(hp.ell.begin(), hp.ell.end(), el.begin() + i * ne);
        std::copy(hp.ellT.begin(), hp.ellT.end(), et.begin() + i * ne);
      }
      if (!up((void**)&b->dHGfrag, gf.data(), gf.size() * sizeof(double)) || !up((void**)&b->dHGfragT, gt.data(), gt.size() * sizeof(double)) ||
          !up((void**)&b->dHNorms, nr.data(), nr.size() * sizeof(double)) || !up((void**)&b->dHEll, el.data(), el.size() * sizeof(pb2::EllEntry)) ||
          !up((void**)&b->dHEllT, et.data(), et.size() * sizeof(pb2::EllEntry)))
        return bail("upload");
      bl.hgfrag = b->dHGfrag; bl.hgfragT = b->dHGfragT; bl.hnorms = b->dHNorms; bl.hell = b->dHEll; bl.hellT = b->dHEllT;
      bl.mem_hgfrag = (long long)ng; bl.mem_hell = (long long)ne; bl.mem_hnorms = (long long)nn;
    }
  }
  *out = b;
  return PB2_OK;
}

int32_t pb2_batch_size(const pb2_batch* b) { return b ? (int32_t)b->mem.size() : -1; }
int32_t pb2_batch_fused(const pb2_batch* b) { return b ? (b->fused ? 1 : 0) : -1; }
int64_t pb2_batch_dim(const pb2_batch* b) { return b ? pb2_dim(b->mem[0]) : -1; }
int64_t pb2_batch_nnz_jac(const pb2_batch* b) { return b ? pb2_nnz_jac(b->mem[0]) : -1; }
int64_t pb2_batch_nnz_hess(const pb2_batch* b) { return b ? pb2_nnz_hess(b->mem[0]) : -1; }
int pb2_batch_structure_jac(const pb2_batch* b, int32_t i, int64_t* rows, int64_t* cols) {
  if (!b || i < 0 || i >= (int)b->mem.size()) return fail(PB2_EINVAL, "pb2_batch_structure_jac: bad member");
  return pb2_structure_jac(b->mem[i], rows, cols);
}
int pb2_batch_structure_hess(const pb2_batch* b, int32_t i, int64_t* rows, int64_t* cols) {
  if (!b || i < 0 || i >= (int)b->mem.size()) return fail(PB2_EINVAL, "pb2_batch_structure_hess: bad member");
  return pb2_structure_hess(b->mem[i], rows, cols);
}

// what == 0: residual + Jacobian (a = delta, c = vals);  1: Hessian (a2 = mu, c = vals)
static int batch_launch(pb2_batch* b, int what, const double* dZ, double* ddelta, const double* dmu, double* dvals,
                        cudaStream_t st) {
  pb2_handle* h0 = b->mem[0];
  const int n = (int)b->mem.size();
  if (b->fused) {
    h0->batch = &b->bl;
    const int rc = what == 0 ? launch_resjac_core(h0, dZ, ddelta, dvals, st, 0, 0, nullptr, 0, 0) : launch_hess(h0, dZ, dmu, dvals, st);
    h0->batch = nullptr;
    return rc;
  }
  // members whose kernels cannot share a grid: fork onto the members' own streams, join on `st`
  const int64_t dim = pb2_dim(h0), nj = pb2_nnz_jac(h0), nh = pb2_nnz_hess(h0);
  PB2_CUDA(cudaEventRecord(b->fork, st));
  for (int i = 0; i < n; ++i) {
    pb2_handle* h = b->mem[i];
    PB2_CUDA(cudaStreamWaitEvent(h->stream, b->fork, 0));
    const int rc = what == 0
        ? launch_resjac(h, dZ, ddelta ? ddelta + (size_t)i * dim : nullptr, dvals ? dvals + (size_t)i * nj : nullptr, h->stream)
        : launch_hess(h, dZ, dmu + (size_t)i * dim, dvals + (size_t)i * nh, h->stream);
    if (rc) return rc;
    PB2_CUDA(cudaEventRecord(b->done[i], h->stream));
    PB2_CUDA(cudaStreamWaitEvent(st, b->done[i], 0));
  }
  return PB2_OK;
}

int pb2_batch_residual_jacobian_async(pb2_batch* b, const double* dZ, double* ddelta, double* dvals, void* stream) {
  if (!b || !dZ) return fail(PB2_EINVAL, "pb2_batch_residual_jacobian_async: null argument");
  DeviceGuard guard(b->mem[0]->d.device);
  return batch_launch(b, 0, dZ, ddelta, nullptr, dvals, (cudaStream_t)stream);
}

int pb2_batch_hess_lagrangian_async(pb2_batch* b, const double* dZ, const double* dmu, double* dvals, void* stream) {
  if (!b || !dZ || !dmu || !dvals) return fail(PB2_EINVAL, "pb2_batch_hess_lagrangian_async: null argument");
  DeviceGuard guard(b->mem[0]->d.device);
  return batch_launch(b, 1, dZ, nullptr, dmu, dvals, (cudaStream_t)stream);
}

static int batch_host(pb2_batch* b, int what, const double* Z, double* delta, const double* mu, double* vals, int space) {
  pb2_handle* h0 = b->mem[0];
  DeviceGuard guard(h0->d.device);
  if (space == PB2_DEVICE) {
    int rc = batch_launch(b, what, Z, delta, mu, vals, b->stream);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(b->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_batch: bad space");
  const size_t n = b->mem.size(), nZ = (size_t)h0->d.D * h0->d.K;
  const size_t nD = n * (size_t)pb2_dim(h0), nJ = n * (size_t)pb2_nnz_jac(h0), nH = n * (size_t)pb2_nnz_hess(h0);
  int rc;
  if ((rc = aux_ensure(&b->dZ, nZ))) return rc;
  PB2_CUDA(cudaMemcpyAsync(b->dZ, Z, nZ * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (what == 0) {
    if ((delta && (rc = aux_ensure(&b->dDelta, nD))) || (vals && (rc = aux_ensure(&b->dJac, nJ)))) return rc;
    if ((rc = batch_launch(b, 0, b->dZ, delta ? b->dDelta : nullptr, nullptr, vals ? b->dJac : nullptr, b->stream))) return rc;
    if (delta) PB2_CUDA(cudaMemcpyAsync(delta, b->dDelta, nD * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (vals) PB2_CUDA(cudaMemcpyAsync(vals, b->dJac, nJ * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  } else {
    if ((rc = aux_ensure(&b->dMu, nD)) || (rc = aux_ensure(&b->dHess, nH))) return rc;
    PB2_CUDA(cudaMemcpyAsync(b->dMu, mu, nD * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    if ((rc = batch_launch(b, 1, b->dZ, nullptr, b->dMu, b->dHess, b->stream))) return rc;
    PB2_CUDA(cudaMemcpyAsync(vals, b->dHess, nH * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  }
  PB2_CUDA(cudaStreamSynchronize(b->stream));
  return PB2_OK;
}

int pb2_batch_residual_jacobian(pb2_batch* b, const double* Z, double* delta, double* vals, int space) {
  if (!b || !Z) return fail(PB2_EINVAL, "pb2_batch_residual_jacobian: null argument");
  return batch_host(b, 0, Z, delta, nullptr, vals, space);
}

int pb2_batch_hess_lagrangian(pb2_batch* b, const double* Z, const double* mu, double* vals, int space) {
  if (!b || !Z || !mu || !vals) return fail(PB2_EINVAL, "pb2_batch_hess_lagrangian: null argument");
  return batch_host(b, 1, Z, nullptr, mu, vals, space);
}

// ---- rollout (knot_rollout.cuh) --------------------------------------------------------------------------
static int launch_rollout(pb2_handle* h, const double* dZ, const double* dx0, double* dstates, double* dout3,
                          cudaStream_t st, int z_stable) {
  const size_t nJ = (size_t)std::max<int64_t>((int64_t)h->nnz_jac_knot() * h->nk(), 1);
  if (!h->dRoJac) PB2_CUDA(cudaMalloc(&h->dRoJac, nJ * sizeof(double)));
  // the propagators: one residual + Jacobian launch (canonical Jacobian values only) into the handle's scratch
  int rc = launch_resjac_td(h, dZ, nullptr, h->dRoJac, st, 0, 0, nullptr, 0, z_stable);
  if (rc) return rc;
  pb2::RolloutParams q{};
  q.b = h->d.b; q.n_b = h->d.n_b; q.K = h->d.K; q.D = h->d.D; q.x_off = h->d.x_off;
  q.nnz_jac = h->nnz_jac_knot();
  q.jac = h->dRoJac; q.Z = dZ; q.x0 = dx0; q.states = dstates; q.out = dout3;
  const size_t smem = pb2::rollout_smem_bytes(q.b, q.n_b);
  if (smem > 48 * 1024) return fail(PB2_EINVAL, "pb2_rollout: state too large for the shared-memory chain");
  pb2::knot_rollout_kernel<<<1, pb2::kRolloutThreads, smem, st>>>(q);
  PB2_CUDA(cudaGetLastError());
  h->launches++;
  return PB2_OK;
}

int pb2_rollout_async(pb2_handle* h, const double* dZ, const double* dx0, double* dstates, double* dout3, void* stream) {
  if (check(h)) return PB2_EINVAL;
  if (!dZ) return fail(PB2_EINVAL, "pb2_rollout_async: null Z");
  DeviceGuard guard(h->d.device);
  return launch_rollout(h, dZ, dx0, dstates, dout3, (cudaStream_t)stream, 0);
}

int pb2_rollout(pb2_handle* h, const double* Z, const double* x0, double* states, double* out3, int space) {
  if (check(h)) return PB2_EINVAL;
  if (!Z) return fail(PB2_EINVAL, "pb2_rollout: null Z");
  DeviceGuard guard(h->d.device);
  if (space == PB2_DEVICE) {
    int rc = launch_rollout(h, Z, x0, states, out3, h->stream, 0);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_rollout: bad space");
  const size_t nZ = (size_t)h->d.D * h->d.K, nx = (size_t)h->n_x(), nS = nx * (size_t)h->d.K;
  int rc;
  if ((rc = ensure(&h->dZ, &h->hZ, nZ))) return rc;
  if (!h->dRoStates) PB2_CUDA(cudaMalloc(&h->dRoStates, std::max<size_t>(nS, 1) * sizeof(double)));
  if (!h->dRoX0) PB2_CUDA(cudaMalloc(&h->dRoX0, std::max<size_t>(nx, 1) * sizeof(double)));
  if (!h->dRoOut) PB2_CUDA(cudaMalloc(&h->dRoOut, 3 * sizeof(double)));
  if ((rc = stage_in(h, Z, h->hZ, h->dZ, nZ))) return rc;
  if (x0) PB2_CUDA(cudaMemcpyAsync(h->dRoX0, x0, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if ((rc = launch_rollout(h, h->dZ, x0 ? h->dRoX0 : nullptr, h->dRoStates, h->dRoOut, h->stream, 1))) return rc;
  if (states) PB2_CUDA(cudaMemcpyAsync(states, h->dRoStates, nS * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out3) PB2_CUDA(cudaMemcpyAsync(out3, h->dRoOut, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  return PB2_OK;
}

int pb2_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes < 0) return fail(PB2_EINVAL, "pb2_host_alloc: bad argument");
  PB2_CUDA(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 8)));
  return PB2_OK;
}

int pb2_host_free(void* ptr) {
  if (!ptr) return PB2_OK;
  PB2_CUDA(cudaFreeHost(ptr));
  return PB2_OK;
}

// ---- gather buffers for a sharded trajectory, set up without any other CUDA binding (one process per GPU) ----
int pb2_device_alloc(void** ptr, int64_t bytes, int32_t device) {
  if (!ptr || bytes < 0) return fail(PB2_EINVAL, "pb2_device_alloc: bad argument");
  DeviceGuard guard(device);
  PB2_CUDA(cudaMalloc(ptr, (size_t)std::max<int64_t>(bytes, 8)));
  PB2_CUDA(cudaMemset(*ptr, 0, (size_t)std::max<int64_t>(bytes, 8)));
  return PB2_OK;
}

int pb2_device_free(void* ptr) {
  if (!ptr) return PB2_OK;
  PB2_CUDA(cudaFree(ptr));
  return PB2_OK;
}

int pb2_device_copy(void* dst, const void* src, int64_t bytes) {
  if (!dst || !src || bytes < 0) return fail(PB2_EINVAL, "pb2_device_copy: bad argument");
  PB2_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDefault));
  return PB2_OK;
}

int pb2_ipc_export(const void* dptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == PB2_IPC_HANDLE_BYTES, "handle size");
  if (!dptr || !handle64) return fail(PB2_EINVAL, "pb2_ipc_export: null argument");
  cudaIpcMemHandle_t hd;
  PB2_CUDA(cudaIpcGetMemHandle(&hd, const_cast<void*>(dptr)));
  std::memcpy(handle64, &hd, sizeof hd);
  return PB2_OK;
}

int pb2_ipc_open(const void* handle64, int32_t device, void** dptr) {
  if (!dptr || !handle64) return fail(PB2_EINVAL, "pb2_ipc_open: null argument");
  DeviceGuard guard(device);
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle64, sizeof hd);
  PB2_CUDA(cudaIpcOpenMemHandle(dptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return PB2_OK;
}

int pb2_ipc_close(void* dptr) {
  if (!dptr) return PB2_OK;
  PB2_CUDA(cudaIpcCloseMemHandle(dptr));
  return PB2_OK;
}


// ---- linear knot constraints (DerivativeIntegrator pairs + time consistency) -----------------------
struct pb2_aux {
  pb2_aux_desc d{};
  pb2::AuxParams p{};
  cudaStream_t stream = nullptr;
  double *dZ = nullptr, *dDelta = nullptr, *dJac = nullptr, *dMu = nullptr, *dHess = nullptr;
  int64_t n_der = 0, n_time = 0, n_eq = 0;
};

int pb2_aux_create(const pb2_aux_desc* desc, pb2_aux** out) {
  if (!desc || !out) return fail(PB2_EINVAL, "pb2_aux_create: null argument");
  *out = nullptr;
  const pb2_aux_desc& d = *desc;
  if (d.K < 1 || d.D < 1 || d.n_pairs < 0 || d.n_pairs > PB2_AUX_MAX_PAIRS)
    return fail(PB2_EINVAL, "pb2_aux_create: bad sizes");
  auto inside = [&](int off, int len) { return off >= 0 && len >= 1 && off + len <= d.D; };
  if (!inside(d.dt_off, 1) || (d.t_off >= 0 && !inside(d.t_off, 1)))
    return fail(PB2_EINVAL, "pb2_aux_create: timestep / time offsets outside the knot column");
  for (int i = 0; i < d.n_pairs; ++i)
    if (!inside(d.x_off[i], d.dim[i]) || !inside(d.xdot_off[i], d.dim[i]))
      return fail(PB2_EINVAL, "pb2_aux_create: component offsets outside the knot column");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(PB2_ENODEVICE, "pb2_aux_create: no CUDA device (this library has no CPU path)");
  }
  if (d.device < 0 || d.device >= ndev) return fail(PB2_EINVAL, "pb2_aux_create: bad device ordinal");
  DeviceGuard guard_11(d.device);
  pb2_aux* h = new (std::nothrow) pb2_aux();
  if (!h) return fail(PB2_ENOMEM, "pb2_aux_create: out of memory");
  h->d = d;
  pb2::AuxParams& p = h->p;
  p.K = d.K; p.D = d.D; p.dt_off = d.dt_off; p.t_off = d.t_off; p.n_pairs = d.n_pairs;
  const long long nk = d.K - 1;
  long long row = 0;
  for (int i = 0; i < d.n_pairs; ++i) {
    p.x_off[i] = d.x_off[i]; p.xdot_off[i] = d.xdot_off[i]; p.dim[i] = d.dim[i];
    p.row0[i] = row;
    row += nk * d.dim[i];
  }
  p.row0[d.n_pairs] = row;
  h->n_der = row;
  h->n_time = d.t_off >= 0 ? nk : 0;
  h->n_eq = d.timesteps_all_equal ? nk : 0;
  p.row_eq = row + h->n_time;
  p.n_rows = row + h->n_time + h->n_eq;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return fail(PB2_ECUDA, "pb2_aux_create: cudaStreamCreate failed");
  }
  *out = h;
  return PB2_OK;
}

void pb2_aux_destroy(pb2_aux* h) {
  if (!h) return;
  DeviceGuard guard_d(h->d.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (double* q : {h->dZ, h->dDelta, h->dJac, h->dMu, h->dHess})
    if (q) cudaFree(q);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

int64_t pb2_aux_dim(const pb2_aux* h) { return h ? h->p.n_rows : -1; }
int64_t pb2_aux_nnz_jac(const pb2_aux* h) { return h ? 4 * h->n_der + 3 * h->n_time + 2 * h->n_eq : -1; }
int64_t pb2_aux_nnz_hess(const pb2_aux* h) { return h ? h->n_der : -1; }

int pb2_aux_structure_jac(const pb2_aux* h, int64_t* rows, int64_t* cols) {
  if (!h || !rows || !cols) return fail(PB2_EINVAL, "pb2_aux_structure_jac: null argument");
  const pb2_aux_desc& d = h->d;
  const int64_t nk = d.K - 1, D = d.D;
  int64_t o = 0, r = 1;
  for (int pr = 0; pr < d.n_pairs; ++pr)
    for (int64_t k = 0; k < nk; ++k)
      for (int64_t i = 0; i < d.dim[pr]; ++i, ++r) {
        const int64_t c0 = k * D + 1;
        rows[o] = r; cols[o++] = c0 + d.x_off[pr] + i;
        rows[o] = r; cols[o++] = c0 + d.xdot_off[pr] + i;
        rows[o] = r; cols[o++] = c0 + d.dt_off;
        rows[o] = r; cols[o++] = c0 + D + d.x_off[pr] + i;
      }
  if (d.t_off >= 0)
    for (int64_t k = 0; k < nk; ++k, ++r) {
      const int64_t c0 = k * D + 1;
      rows[o] = r; cols[o++] = c0 + d.t_off;
      rows[o] = r; cols[o++] = c0 + d.dt_off;
      rows[o] = r; cols[o++] = c0 + D + d.t_off;
    }
  if (d.timesteps_all_equal)
    for (int64_t k = 0; k < nk; ++k, ++r) {
      const int64_t c0 = k * D + 1;
      rows[o] = r; cols[o++] = c0 + d.dt_off;
      rows[o] = r; cols[o++] = c0 + D + d.dt_off;
    }
  return PB2_OK;
}

int pb2_aux_structure_hess(const pb2_aux* h, int64_t* rows, int64_t* cols) {
  if (!h || !rows || !cols) return fail(PB2_EINVAL, "pb2_aux_structure_hess: null argument");
  const pb2_aux_desc& d = h->d;
  const int64_t nk = d.K - 1, D = d.D;
  int64_t o = 0;
  for (int pr = 0; pr < d.n_pairs; ++pr)
    for (int64_t k = 0; k < nk; ++k)
      for (int64_t i = 0; i < d.dim[pr]; ++i) {
        const int64_t a = k * D + 1 + d.xdot_off[pr] + i, c = k * D + 1 + d.dt_off;
        rows[o] = std::min(a, c);
        cols[o++] = std::max(a, c);
      }
  return PB2_OK;
}

static int aux_launch(pb2_aux* h, const double* dZ, const double* dmu, double* ddelta, double* djac, double* dhess,
                      cudaStream_t st) {
  if (h->p.n_rows <= 0) return PB2_OK;
  pb2::AuxParams p = h->p;
  p.Z = dZ; p.mu = dmu; p.delta = ddelta; p.jac = djac; p.hess = dhess;
  const int blocks = (int)std::min<long long>((p.n_rows + 255) / 256, 148 * 8);
  pb2::knot_aux_kernel<<<blocks, 256, 0, st>>>(p);
  PB2_CUDA(cudaGetLastError());
  return PB2_OK;
}

int pb2_aux_residual_jacobian_async(pb2_aux* h, const double* dZ, double* ddelta, double* dvals, void* stream) {
  if (!h || !dZ) return fail(PB2_EINVAL, "pb2_aux_residual_jacobian_async: null argument");
  DeviceGuard guard_12(h->d.device);
  return aux_launch(h, dZ, nullptr, ddelta, dvals, nullptr, (cudaStream_t)stream);
}

int pb2_aux_residual_jacobian(pb2_aux* h, const double* Z, double* delta, double* vals, int space) {
  if (!h || !Z) return fail(PB2_EINVAL, "pb2_aux_residual_jacobian: null argument");
  DeviceGuard guard_13(h->d.device);
  if (space == PB2_DEVICE) {
    int rc = aux_launch(h, Z, nullptr, delta, vals, nullptr, h->stream);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_aux_residual_jacobian: bad space");
  const size_t nZ = (size_t)h->d.D * h->d.K, nD = (size_t)h->p.n_rows, nJ = (size_t)pb2_aux_nnz_jac(h);
  int rc;
  if ((rc = aux_ensure(&h->dZ, nZ)) || (delta && (rc = aux_ensure(&h->dDelta, nD))) ||
      (vals && (rc = aux_ensure(&h->dJac, nJ))))
    return rc;
  PB2_CUDA(cudaMemcpyAsync(h->dZ, Z, nZ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if ((rc = aux_launch(h, h->dZ, nullptr, delta ? h->dDelta : nullptr, vals ? h->dJac : nullptr, nullptr, h->stream)))
    return rc;
  if (delta) PB2_CUDA(cudaMemcpyAsync(delta, h->dDelta, nD * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (vals) PB2_CUDA(cudaMemcpyAsync(vals, h->dJac, nJ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  return PB2_OK;
}

int pb2_aux_hess_lagrangian(pb2_aux* h, const double* mu, double* vals, int space) {
  if (!h || !mu || !vals) return fail(PB2_EINVAL, "pb2_aux_hess_lagrangian: null argument");
  DeviceGuard guard_14(h->d.device);
  if (h->n_der <= 0) return PB2_OK;
  pb2::AuxParams p = h->p;
  p.n_rows = h->n_der;   // time rows have no second derivatives
  p.Z = nullptr; p.delta = nullptr; p.jac = nullptr;
  const int blocks = (int)std::min<long long>((p.n_rows + 255) / 256, 148 * 8);
  if (space == PB2_DEVICE) {
    p.mu = mu; p.hess = vals;
    pb2::knot_aux_kernel<<<blocks, 256, 0, h->stream>>>(p);
    PB2_CUDA(cudaGetLastError());
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_aux_hess_lagrangian: bad space");
  const size_t n = (size_t)h->n_der;
  int rc;
  if ((rc = aux_ensure(&h->dMu, (size_t)h->p.n_rows)) || (rc = aux_ensure(&h->dHess, n))) return rc;
  PB2_CUDA(cudaMemcpyAsync(h->dMu, mu, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  p.mu = h->dMu; p.hess = h->dHess;
  pb2::knot_aux_kernel<<<blocks, 256, 0, h->stream>>>(p);
  PB2_CUDA(cudaGetLastError());
  PB2_CUDA(cudaMemcpyAsync(vals, h->dHess, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  return PB2_OK;
}

// ---- objective value + gradient ---------------------------------------------------------------------
struct pb2_obj {
  int K = 0, D = 0, dt_off = 0, n_terms = 0, n_regs = 0, device = 0;
  pb2::ObjParams p{};
  cudaStream_t stream = nullptr;
  std::vector<void*> owned;            // device allocations behind p
  double *dZ = nullptr, *dGrad = nullptr, *dJ = nullptr, *dHess = nullptr;
  double* hJ = nullptr;                // pinned
  size_t smem = 0;
  std::vector<int64_t> hrows, hcols;   // Hessian structure (1-based, upper triangle)
};

extern "C++" {
template <class T>
static int obj_upload(pb2_obj* h, const std::vector<T>& v, const T** out) {
  void* d = nullptr;
  PB2_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  h->owned.push_back(d);
  if (!v.empty()) PB2_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)d;
  return PB2_OK;
}
}  // extern "C++"

static int obj_build(pb2_obj* h, const pb2_obj_desc& d) {
  const int K = d.K;
  std::vector<int> t_off(1, 0), t_flags, rows, r_off(1, 0), r_pow, r_rows;
  std::vector<double> t_scale, a_re, a_im, a_sq, a_lin, r_R, r_w((size_t)d.n_regs * K, 0.0);
  std::vector<std::vector<std::pair<int, double>>> per_knot(K);
  for (int t = 0; t < d.n_terms; ++t) {
    const pb2_obj_term& T = d.terms[t];
    for (int i = 0; i < T.n_rows; ++i) {
      rows.push_back(T.rows[i]);
      a_re.push_back(T.a_re ? T.a_re[i] : 0.0);
      a_im.push_back(T.a_im ? T.a_im[i] : 0.0);
      a_sq.push_back(T.a_sq ? T.a_sq[i] : 0.0);
      a_lin.push_back(T.a_lin ? T.a_lin[i] : 0.0);
    }
    t_off.push_back((int)rows.size());
    t_flags.push_back(T.flags);
    t_scale.push_back(T.scale);
    if (T.n_times == 0) per_knot[K - 1].push_back({t, T.Q[0]});
    else
      for (int i = 0; i < T.n_times; ++i) per_knot[T.times[i]].push_back({t, T.Q[i]});
  }
  std::vector<int> knot_ptr(1, 0), item_term;
  std::vector<double> item_q;
  for (int k = 0; k < K; ++k) {
    for (auto& it : per_knot[k]) {
      item_term.push_back(it.first);
      item_q.push_back(it.second);
    }
    knot_ptr.push_back((int)item_term.size());
  }
  std::vector<const double*> r_base(std::max(d.n_regs, 1), nullptr);
  for (int r = 0; r < d.n_regs; ++r) {
    const pb2_obj_reg& R = d.regs[r];
    for (int i = 0; i < R.n_rows; ++i) {
      r_rows.push_back(R.rows[i]);
      r_R.push_back(R.R[i]);
    }
    r_off.push_back((int)r_rows.size());
    r_pow.push_back(R.dt_power);
    if (R.n_times == 0) std::fill(r_w.begin() + (size_t)r * K, r_w.begin() + (size_t)(r + 1) * K, 1.0);
    else
      for (int i = 0; i < R.n_times; ++i) r_w[(size_t)r * K + R.times[i]] = 1.0;
    if (R.baseline) {
      std::vector<double> b(R.baseline, R.baseline + (size_t)R.n_rows * K);
      int rc = obj_upload(h, b, &r_base[r]);
      if (rc) return rc;
    }
  }
  pb2::ObjParams& p = h->p;
  p.K = K; p.D = d.D; p.dt_off = d.dt_off; p.n_regs = d.n_regs;
  int rc;
  if ((rc = obj_upload(h, knot_ptr, &p.knot_ptr)) || (rc = obj_upload(h, item_term, &p.item_term)) ||
      (rc = obj_upload(h, item_q, &p.item_q)) || (rc = obj_upload(h, t_off, &p.t_off)) ||
      (rc = obj_upload(h, t_flags, &p.t_flags)) || (rc = obj_upload(h, t_scale, &p.t_scale)) ||
      (rc = obj_upload(h, rows, &p.rows)) || (rc = obj_upload(h, a_re, &p.a_re)) ||
      (rc = obj_upload(h, a_im, &p.a_im)) || (rc = obj_upload(h, a_sq, &p.a_sq)) ||
      (rc = obj_upload(h, a_lin, &p.a_lin)) || (rc = obj_upload(h, r_off, &p.r_off)) ||
      (rc = obj_upload(h, r_pow, &p.r_pow)) || (rc = obj_upload(h, r_rows, &p.r_rows)) ||
      (rc = obj_upload(h, r_R, &p.r_R)) || (rc = obj_upload(h, r_w, &p.r_w)))
    return rc;
  const double* const* dbase = nullptr;
  if ((rc = obj_upload(h, r_base, (const double* const**)&dbase))) return rc;
  p.r_base = dbase;
  std::vector<double> zeros((size_t)K + 1, 0.0);
  const double* jp = nullptr;
  if ((rc = obj_upload(h, zeros, &jp))) return rc;
  p.Jpart = const_cast<double*>(jp);
  std::vector<unsigned int> cz(1, 0u);
  const unsigned int* cp = nullptr;
  if ((rc = obj_upload(h, cz, &cp))) return rc;
  p.counter = const_cast<unsigned int*>(cp);
  // Hessian structure, in the order knot_objective_hess_kernel writes the values
  std::vector<int> t_kind(std::max(d.n_terms, 1), 0);
  for (int t = 0; t < d.n_terms; ++t) {
    bool dense = false, diag = false;
    for (int i = t_off[t]; i < t_off[t + 1]; ++i) {
      dense = dense || a_re[i] != 0.0 || a_im[i] != 0.0;
      diag = diag || a_sq[i] != 0.0;
    }
    t_kind[t] = dense ? 2 : (diag ? 1 : 0);
  }
  std::vector<long long> hess_ptr(1, 0);
  for (int k = 0; k < K; ++k) {
    const int64_t base = (int64_t)k * d.D + 1;
    for (auto& it : per_knot[k]) {
      const int t = it.first, o0 = t_off[t], n = t_off[t + 1] - o0;
      if (t_kind[t] == 1)
        for (int i = 0; i < n; ++i) {
          h->hrows.push_back(base + rows[o0 + i]);
          h->hcols.push_back(base + rows[o0 + i]);
        }
      if (t_kind[t] == 2)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i <= j; ++i) {
            h->hrows.push_back(base + std::min(rows[o0 + i], rows[o0 + j]));
            h->hcols.push_back(base + std::max(rows[o0 + i], rows[o0 + j]));
          }
    }
    for (int r = 0; r < d.n_regs; ++r) {
      if (r_w[(size_t)r * K + k] == 0.0) continue;
      const int o0 = r_off[r], n = r_off[r + 1] - o0;
      for (int i = 0; i < n; ++i) {
        h->hrows.push_back(base + r_rows[o0 + i]);
        h->hcols.push_back(base + r_rows[o0 + i]);
      }
      if (r_pow[r]) {
        for (int i = 0; i < n; ++i) {
          h->hrows.push_back(base + std::min(r_rows[o0 + i], d.dt_off));
          h->hcols.push_back(base + std::max(r_rows[o0 + i], d.dt_off));
        }
        h->hrows.push_back(base + d.dt_off);
        h->hcols.push_back(base + d.dt_off);
      }
    }
    hess_ptr.push_back((long long)h->hrows.size());
  }
  if ((rc = obj_upload(h, t_kind, &p.t_kind)) || (rc = obj_upload(h, hess_ptr, &p.hess_ptr))) return rc;
  return PB2_OK;
}

int pb2_obj_create(const pb2_obj_desc* desc, pb2_obj** out) {
  if (!desc || !out) return fail(PB2_EINVAL, "pb2_obj_create: null argument");
  *out = nullptr;
  const pb2_obj_desc& d = *desc;
  if (d.K < 1 || d.D < 1 || d.n_terms < 0 || d.n_regs < 0 || (d.n_terms && !d.terms) || (d.n_regs && !d.regs))
    return fail(PB2_EINVAL, "pb2_obj_create: bad sizes");
  if (d.dt_off < 0 || d.dt_off >= d.D) return fail(PB2_EINVAL, "pb2_obj_create: timestep offset outside the knot column");
  auto rows_ok = [&](const int32_t* rows, int n) {
    if (n < 0 || (n && !rows)) return false;
    std::vector<char> seen(d.D, 0);
    for (int i = 0; i < n; ++i) {
      if (rows[i] < 0 || rows[i] >= d.D || seen[rows[i]]) return false;
      seen[rows[i]] = 1;
    }
    return true;
  };
  auto times_ok = [&](const int32_t* times, int n) {
    if (n < 0 || (n && !times)) return false;
    for (int i = 0; i < n; ++i)
      if (times[i] < 0 || times[i] >= d.K) return false;
    return true;
  };
  for (int t = 0; t < d.n_terms; ++t) {
    const pb2_obj_term& T = d.terms[t];
    if (!rows_ok(T.rows, T.n_rows) || !times_ok(T.times, T.n_times) || !T.Q)
      return fail(PB2_EINVAL, "pb2_obj_create: term rows must be distinct and inside the knot column, times inside 0..K-1");
  }
  for (int r = 0; r < d.n_regs; ++r) {
    const pb2_obj_reg& R = d.regs[r];
    if (!rows_ok(R.rows, R.n_rows) || !times_ok(R.times, R.n_times) || (R.n_rows && !R.R) || R.dt_power < 0 ||
        R.dt_power > 2)
      return fail(PB2_EINVAL, "pb2_obj_create: regularizer rows / times / dt_power out of range");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(PB2_ENODEVICE, "pb2_obj_create: no CUDA device (this library has no CPU path)");
  }
  if (d.device < 0 || d.device >= ndev) return fail(PB2_EINVAL, "pb2_obj_create: bad device ordinal");
  DeviceGuard guard_15(d.device);
  pb2_obj* h = new (std::nothrow) pb2_obj();
  if (!h) return fail(PB2_ENOMEM, "pb2_obj_create: out of memory");
  h->K = d.K; h->D = d.D; h->dt_off = d.dt_off; h->n_terms = d.n_terms; h->n_regs = d.n_regs; h->device = d.device;
  h->smem = 2 * (size_t)d.D * sizeof(double);
  int rc = obj_build(h, d);
  if (!rc && h->smem > 200 * 1024) rc = fail(PB2_EINVAL, "pb2_obj_create: knot column too large for shared memory");
  // (function attribute, process-wide: always the device limit -- see pb2_create)
  if (!rc && (raise_dynamic_smem(pb2::knot_objective_kernel) != cudaSuccess ||
              raise_dynamic_smem(pb2::knot_objective_hess_kernel) != cudaSuccess))
    rc = fail(PB2_ECUDA, "pb2_obj_create: cudaFuncSetAttribute failed");
  if (!rc && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    rc = fail(PB2_ECUDA, "pb2_obj_create: cudaStreamCreate failed");
  if (!rc && cudaMallocHost((void**)&h->hJ, sizeof(double)) != cudaSuccess)
    rc = fail(PB2_ECUDA, "pb2_obj_create: cudaMallocHost failed");
  if (rc) {
    pb2_obj_destroy(h);
    return rc;
  }
  *out = h;
  return PB2_OK;
}

void pb2_obj_destroy(pb2_obj* h) {
  if (!h) return;
  DeviceGuard guard_d(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* q : h->owned) cudaFree(q);
  for (double* q : {h->dZ, h->dGrad, h->dJ, h->dHess})
    if (q) cudaFree(q);
  if (h->hJ) cudaFreeHost(h->hJ);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

static int obj_launch(pb2_obj* h, const double* dZ, double* dJ, double* dgrad, cudaStream_t st) {
  pb2::ObjParams p = h->p;
  p.Z = dZ; p.J = dJ; p.grad = dgrad;
  pb2::knot_objective_kernel<<<h->K, pb2::kObjThreads, h->smem, st>>>(p);
  PB2_CUDA(cudaGetLastError());
  return PB2_OK;
}

int64_t pb2_obj_nnz_hess(const pb2_obj* h) { return h ? (int64_t)h->hrows.size() : -1; }

int pb2_obj_structure_hess(const pb2_obj* h, int64_t* rows, int64_t* cols) {
  if (!h || !rows || !cols) return fail(PB2_EINVAL, "pb2_obj_structure_hess: null argument");
  std::copy(h->hrows.begin(), h->hrows.end(), rows);
  std::copy(h->hcols.begin(), h->hcols.end(), cols);
  return PB2_OK;
}

static int obj_launch_hess(pb2_obj* h, const double* dZ, double sigma, double* dvals, cudaStream_t st) {
  if (h->hrows.empty()) return PB2_OK;
  pb2::ObjParams p = h->p;
  p.Z = dZ; p.hess = dvals; p.sigma = sigma;
  pb2::knot_objective_hess_kernel<<<h->K, pb2::kObjThreads, h->smem, st>>>(p);
  PB2_CUDA(cudaGetLastError());
  return PB2_OK;
}

int pb2_obj_hessian_async(pb2_obj* h, const double* dZ, double sigma, double* dvals, void* stream) {
  if (!h || !dZ || !dvals) return fail(PB2_EINVAL, "pb2_obj_hessian_async: null argument");
  DeviceGuard guard(h->device);
  return obj_launch_hess(h, dZ, sigma, dvals, (cudaStream_t)stream);
}

int pb2_obj_hessian(pb2_obj* h, const double* Z, double sigma, double* vals, int space) {
  if (!h || !Z || !vals) return fail(PB2_EINVAL, "pb2_obj_hessian: null argument");
  DeviceGuard guard(h->device);
  if (space == PB2_DEVICE) {
    int rc = obj_launch_hess(h, Z, sigma, vals, h->stream);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_obj_hessian: bad space");
  const size_t nZ = (size_t)h->D * h->K, nH = h->hrows.size();
  int rc;
  if ((rc = aux_ensure(&h->dZ, nZ)) || (rc = aux_ensure(&h->dHess, nH))) return rc;
  PB2_CUDA(cudaMemcpyAsync(h->dZ, Z, nZ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if ((rc = obj_launch_hess(h, h->dZ, sigma, h->dHess, h->stream))) return rc;
  PB2_CUDA(cudaMemcpyAsync(vals, h->dHess, nH * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  return PB2_OK;
}

int pb2_obj_value_gradient_async(pb2_obj* h, const double* dZ, double* dJ, double* dgrad, void* stream) {
  if (!h || !dZ || !dJ) return fail(PB2_EINVAL, "pb2_obj_value_gradient_async: null argument");
  DeviceGuard guard_16(h->device);
  return obj_launch(h, dZ, dJ, dgrad, (cudaStream_t)stream);
}

int pb2_obj_value_gradient(pb2_obj* h, const double* Z, double* J, double* grad, int space) {
  if (!h || !Z || !J) return fail(PB2_EINVAL, "pb2_obj_value_gradient: null argument");
  DeviceGuard guard_17(h->device);
  if (space == PB2_DEVICE) {
    int rc = obj_launch(h, Z, J, grad, h->stream);
    if (rc) return rc;
    PB2_CUDA(cudaStreamSynchronize(h->stream));
    return PB2_OK;
  }
  if (space != PB2_HOST) return fail(PB2_EINVAL, "pb2_obj_value_gradient: bad space");
  const size_t nZ = (size_t)h->D * h->K;
  int rc;
  if ((rc = aux_ensure(&h->dZ, nZ)) || (rc = aux_ensure(&h->dJ, 1)) || (grad && (rc = aux_ensure(&h->dGrad, nZ))))
    return rc;
  PB2_CUDA(cudaMemcpyAsync(h->dZ, Z, nZ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if ((rc = obj_launch(h, h->dZ, h->dJ, grad ? h->dGrad : nullptr, h->stream))) return rc;
  PB2_CUDA(cudaMemcpyAsync(h->hJ, h->dJ, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (grad) PB2_CUDA(cudaMemcpyAsync(grad, h->dGrad, nZ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PB2_CUDA(cudaStreamSynchronize(h->stream));
  *J = *h->hJ;
  return PB2_OK;
}

}  // extern "C"
