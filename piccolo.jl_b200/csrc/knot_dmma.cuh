// Tensor-core knot kernel (residual + Jacobian): persistent CTAs, one warp per 8-column tile,
// one warp GROUP per knot, TMA-staged input and output.
//
// Per knot k the reference's BilinearIntegrator evaluates (constructed at
// /root/reference/src/control/integrators.jl:35-95; constraint docs/src/concepts/index.md:21,62)
//     delta_k = x_{k+1} - expv(dt_k, Ghat(u_k), x_k)
// and differentiates it with forward duals over u.  ExponentialAction.expv is the
// Al-Mohy--Higham truncated-Taylor *action*: n_sub sub-steps of a degree-M Horner recurrence
// applied to the vectors.  This kernel runs that same recurrence -- on the state columns X, on
// their first-order jets in every drive direction, and on the columns of the identity (which
// yields the propagator E itself for the d/dx_k block) -- with FP64 tensor-core MMAs
// (DMMA.8x8x4, the only native FP64 MMA shape on sm_100a):
//
//     S   <-  a_k B   + G(u) S                       k = M-1 .. 0,  a_k = dt'^k / k!
//     S_j <-  a_k B_j + G(u) S_j + G_j S             (jet of drive j; G_j sparse, ELL in registers)
//
// Work decomposition (B200: 148 SMs, 4 DMMA pipes per SM, one DMMA per 16 cycles per pipe):
//   * the stacked columns [I (E) | X | jet_1 .. jet_m] are cut into 8-column tiles; ONE WARP owns
//     one tile for the whole recurrence, so the iterate never leaves registers: the warp holds
//     the TRANSPOSED tile T^T (8 columns x b) as mma.m8n8k4 accumulators and computes
//     T^T <- T^T A^T.  The accumulator fragment of one step (lane (g,q) holds rows 2q,2q+1 of
//     column g) is exactly the A-operand fragment of the next step when the contraction index is
//     enumerated as pi(4 kt + q) = 8 (kt/2) + 2q + (kt%2); G's B-fragments are stored in that
//     permuted order once at setup.  A Horner step is 4 NT^2 DMMAs per warp and nothing else.
//   * the warps of one knot (a "group") meet once per Horner step on a named barrier to exchange
//     the current X tile (1 KB through shared memory) for the sparse G_j S coupling.
//   * CTAs are persistent (one per SM, `gpc` groups each); a group walks its knots in a loop.
//     The (z_k, x_{k+1}) slab of the NEXT knot is prefetched by cp.async.bulk (TMA, mbarrier
//     completion) while the current knot computes; results are staged in shared memory in
//     their final COO order and leave with cp.async.bulk shared->global (the n_b replicated
//     propagator blocks are n_b bulk copies of one 2 KB staging block), so the store of knot i
//     overlaps the math of knot i+1.
//
// Generators with the real-isomorphism structure G = [[S, R], [-R, S]] (every ket / unitary
// generator, isomorphisms.jl:350,359) need only the first b/2 columns of E: E = [[P,-Q],[Q,P]].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <vector>

#include "knot_generic.cuh"

namespace pb2 {

struct __align__(16) EllEntry {
  double val;
  int idx;
  int pad;
};

struct DmmaParams {
  int b, n_b, m, D, x_off, dt_off, u_off, nnz_jac;
  int ncT;       // identity columns carried (b/2 for iso generators, b otherwise, 0 = residual only)
  int m_jets;    // drive jets carried (m, or 0 for residual only)
  int iso;       // 1: E is rebuilt from its first b/2 columns
  int max_sub;   // bound on the number of Taylor sub-steps
  int tiles;     // warps per knot group (8-column tiles of the stacked columns)
  int gpc;       // knot groups per CTA
  int nk;        // knots to evaluate (K - 1)
  int zlen;      // doubles staged per knot: z_k and the state of z_{k+1}
  int bulk_in, bulk_out;   // alignment allows cp.async.bulk for the slab load / the result stores
  // shared-memory layout, in doubles (dmma_layout): CTA-wide tables, then gpc group regions
  int o_norm, o_grp, grp_stride;
  int zpad, o_sA, o_sY, o_sC, o_meta, o_mbar;       // inside a group region
  int o_stg, stg_stride, o_sJ, o_sD;                // two staging buffers [E | J | D]
  const double* Gfrag;   // (m+1) * KT*NT*32 doubles, B-fragment order
  const EllEntry* ell;   // (m+1) * Bp * W   (drive m = all-zero dummy)
  const double* norms;   // m+1: ||G_0||_1, ||G_j||_1
  const double* Z;
  double* delta;
  double* jac;
  long long* trace;      // debug: per-phase clock64 stamps of block 0 (null in production)
  // batched launch (pb2_batch_*, blockIdx.y = member of a SamplingTrajectory ensemble): own tables and state block
  // per member, one shared trajectory
  int mem_n;                                   // 0 / 1: not batched
  const int* x_offs;                           // [mem_n]
  long long mem_gfrag, mem_ell, mem_norms;     // strides (elements) between the members' tables
  long long mem_delta, mem_jac;                // strides (doubles) between the members' outputs
};

constexpr int kDmmaMaxTiles = 8;
constexpr int kDmmaMaxW = 4;
constexpr int kDmmaMaxGroups = 7;     // 2 named barriers per group, ids 1..14
constexpr int kDmmaMaxThreads = 512;
// Threads per CTA by generator size: the 16 x 16 variants (e.g. the compact
// Lindbladians, BASELINE config C4) keep 2 x 4 x W coupling values and addresses per lane next to 8 generator fragments; at 512 threads
// (128 registers) they spilled 0.5 - 1 KB per thread into the Horner loop, at 256 threads they fit.
__host__ __device__ constexpr int dmma_max_threads(int NT, int W) { return NT == 2 ? 256 : kDmmaMaxThreads; }

// c_invfact[k] = 1 / k!
__constant__ double c_invfact[kMaxDeg + 1];

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA (bulk async copy) global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// TMA (bulk async copy) shared -> global, bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


__device__ __forceinline__ void dmma884z(double (&c)[2], double a, double b) {   // c = a b (no accumulate)
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%4};"
      : "=d"(c[0]), "=d"(c[1])
      : "d"(a), "d"(b), "d"(0.0));
}
template <int OFF>
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}

// One Horner step of one warp's tile:  t <- a_k B + G(u) t (+ G_j S for jet lanes).
// MODE 0: B = unit columns (propagator tile, first sub-step)   1: general B in `base`
//      2: B = 0 (jet-only tile, first sub-step).   PAR: which half of the exchange buffer.
// The DMMAs start from a zero accumulator and depend only on registers, so the exchange
// barrier, the Y loads and the coefficient load overlap them; the additive terms come last.
#ifdef PB2_TRACE
__device__ long long* g_trace2;   // [warp 16][step 20][4]
#define PB2_STEP_STAMP(i) do { if (g_trace2 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_trace2[(((threadIdx.x >> 5) * 20 + ((ck_addr >> 3) & 15)) * 4) + (i)] = clock64(); } while (0)
#else
#define PB2_STEP_STAMP(i) do { } while (0)
#endif

template <int NT, int W, int MODE, int PAR>
__device__ __forceinline__ void horner_step(double (&t)[2 * NT], const double (&base)[2 * NT],
                                            const double (&A)[2 * NT][NT], const double (&ev)[2 * NT][W],
                                            const uint32_t (&yrd)[2 * NT][W], uint32_t ypub, uint32_t ck_addr,
                                            bool in_x, bool pub, bool cpl, int bar_x, int nx, unsigned diag) {
  constexpr int KT = 2 * NT, YB = KT * 32 * 8;
  PB2_STEP_STAMP(0);
  if (in_x) {
    if (pub) {
#pragma unroll
      for (int i = 0; i < KT; ++i) sts_f64<0>(ypub + PAR * YB + i * 256, t[i]);
    }
    PB2_STEP_STAMP(1);
    bar_sync(bar_x, nx);
  }
  double y[KT][W];
  if (cpl) {
#pragma unroll
    for (int i = 0; i < KT; ++i)
#pragma unroll
      for (int ww = 0; ww < W; ++ww) y[i][ww] = lds_f64<PAR * YB>(yrd[i][ww]);
  }
  double ck = 0.0;
  if (MODE != 2) ck = lds_f64<0>(ck_addr);
  // every additive term goes into the accumulators BEFORE the products (see knot_u8.cuh, u8_mma_acc:
  // an FP64 CUDA-core instruction after the DMMAs would queue behind the other warps' tensor work)
  double d[NT][2];
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    double v = 0.0;
    if (MODE == 1) v = ck * base[i];
    if (MODE == 0 && ((diag >> i) & 1u)) v = ck;
    if (cpl) {
#pragma unroll
      for (int ww = 0; ww < W; ++ww) v = fma(ev[i][ww], y[i][ww], v);
    }
    d[i >> 1][i & 1] = v;
  }
#pragma unroll
  for (int kt = 0; kt < KT; ++kt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) dmma884(d[nt], t[kt], A[kt][nt]);
#pragma unroll
  for (int i = 0; i < KT; ++i) t[i] = d[i >> 1][i & 1];
  PB2_STEP_STAMP(2);
}

// the M steps of one sub-step, buffer halves alternating 0,1,0,...
template <int NT, int W, int MODE>
__device__ __forceinline__ void horner_run(int M, double (&t)[2 * NT], const double (&base)[2 * NT],
                                           const double (&A)[2 * NT][NT], const double (&ev)[2 * NT][W],
                                           const uint32_t (&yrd)[2 * NT][W], uint32_t ypub, uint32_t sC_addr,
                                           bool in_x, bool pub, bool cpl, int bar_x, int nx, unsigned diag) {
  int kq = M - 1;
  for (; kq >= 1; kq -= 2) {
    horner_step<NT, W, MODE, 0>(t, base, A, ev, yrd, ypub, sC_addr + 8 * kq, in_x, pub, cpl, bar_x, nx, diag);
    horner_step<NT, W, MODE, 1>(t, base, A, ev, yrd, ypub, sC_addr + 8 * kq - 8, in_x, pub, cpl, bar_x, nx, diag);
  }
  if (kq == 0) horner_step<NT, W, MODE, 0>(t, base, A, ev, yrd, ypub, sC_addr, in_x, pub, cpl, bar_x, nx, diag);
}

#ifdef PB2_TRACE
#define PB2_STAMP(i) do { if (p.trace && blockIdx.x == 0 && lane == 0 && it < 8) p.trace[((it * 16 + wcta) * 8) + (i)] = clock64(); } while (0)
#else
#define PB2_STAMP(i) do { } while (0)
#endif

// NT = padded generator size / 8 (1 or 2); W = ELL width of the drive generators (1, 2 or 4).
template <int NT, int W>
__global__ void __launch_bounds__(dmma_max_threads(NT, W), 1) knot_dmma_kernel(DmmaParams p) {
  constexpr int KT = 2 * NT, Bp = 8 * NT, FR = KT * NT * 32;
  extern __shared__ __align__(16) double dmma_smem[];
  if (p.mem_n > 1) {
    const long long mi = blockIdx.y;
    p.Gfrag += mi * p.mem_gfrag;
    p.ell += mi * p.mem_ell;
    p.norms += mi * p.mem_norms;
  }
  apply_member(p);
  const int lane = threadIdx.x & 31, wcta = threadIdx.x >> 5;
  const int tiles = p.tiles;
  const int group = wcta / tiles, w = wcta - group * tiles;
  const int g = lane >> 2, q = lane & 3;
  const int b = p.b, n_b = p.n_b, m = p.m, ncT = p.ncT, mj = p.m_jets;
  const int n_x = b * n_b, half = b >> 1, bb = b * b;
  const int gthreads = 32 * tiles, gtid = w * 32 + lane;
  const int bar_all = 1 + group, bar_x = 1 + p.gpc + group;

  // byte addresses (shared window) of the CTA-wide tables and of this group's regions
  const uint32_t a_cG = smem_u32(dmma_smem);
  const uint32_t a_grp = a_cG + 8u * (uint32_t)(p.o_grp + group * p.grp_stride);
  const uint32_t a_sA = a_grp + 8u * p.o_sA, a_sY = a_grp + 8u * p.o_sY, a_sC = a_grp + 8u * p.o_sC;
  const uint32_t a_meta = a_grp + 8u * p.o_meta, a_mbar = a_grp + 8u * p.o_mbar;
  const uint32_t a_stg = a_grp + 8u * p.o_stg;

  // ---- once per CTA: fragment tables, norms, the constant identity entries, mbarriers -------
  {
    const int ncg = (m + 1) * FR;
    for (int e = threadIdx.x; e < ncg; e += blockDim.x) dmma_smem[e] = p.Gfrag[e];
    for (int e = threadIdx.x; e <= m; e += blockDim.x) dmma_smem[p.o_norm + e] = p.norms[e];
    if (p.jac) {
      double* grp = dmma_smem + p.o_grp + group * p.grp_stride + p.o_stg + p.o_sJ + (mj + 1) * n_x;
      for (int e = gtid; e < n_x; e += gthreads) { grp[e] = 1.0; grp[p.stg_stride + e] = 1.0; }
    }
    if (p.bulk_in && gtid == 0) {
      mbar_init(a_mbar, 1);
      mbar_init(a_mbar + 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();

  // ---- column bookkeeping: this warp's tile holds stacked columns 8w .. 8w+7 ---------------
  // kind 0: identity column (propagator E)   1: state column (Y = E x)   2: jet of drive jd
  // 3: padding
  int kind = 3, ccs = 0, jd = m;
  {
    const int c = 8 * w + g;
    if (c < ncT) {
      kind = 0;
      ccs = c;
    } else {
      const int rel = c - ncT, slab = rel / n_b;
      const int cc = rel - slab * n_b;
      if (slab == 0) { kind = 1; ccs = cc; }
      else if (slab <= mj) { kind = 2; jd = slab - 1; ccs = cc; }
    }
  }
  const bool tile_e = 8 * w < ncT;
  const bool tile_y = (8 * w + 7 >= ncT) && (8 * w < ncT + n_b);
  const bool tile_cpl = (mj > 0) && (8 * w + 7 >= ncT + n_b) && (8 * w < ncT + (1 + mj) * n_b);
  const bool in_x = (mj > 0) && (tile_y || tile_cpl);
  int n_xw = 0;   // warps of the group that take part in the X exchange
  for (int w2 = 0; w2 < tiles; ++w2)
    n_xw += ((8 * w2 + 7 >= ncT) && (8 * w2 < ncT + (1 + mj) * n_b)) ? 1 : 0;
  const int nx = 32 * n_xw;
  const bool pub = kind == 1;

  // rows held by this lane: element i = 2 nt + s  <->  row 8 nt + 2 q + s  (byte offset kOff(i))
  // sparse drive coupling of a jet lane, in registers: value and the Y-exchange slot it reads
  double ev[KT][W];
  uint32_t yrd[KT][W];
  unsigned diag = 0, rowok = 0;
  const uint32_t ypub = a_sY + 8u * (uint32_t)(ccs * 4 + q);
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    const int r = 8 * (i >> 1) + 2 * q + (i & 1);
    if (kind == 0 && r == ccs) diag |= 1u << i;
    if (r < b && kind != 3) rowok |= 1u << i;
#pragma unroll
    for (int ww = 0; ww < W; ++ww) {
      EllEntry en{0.0, 0, 0};
      if (kind == 2 && r < b) en = p.ell[((size_t)jd * Bp + r) * W + ww];
      ev[i][ww] = en.val;
      // lanes that are not jets read slot 0 (always published) and multiply it by zero
      const int eo = kind == 2 ? (2 * (en.idx >> 3) + (en.idx & 1)) * 32 + ccs * 4 + ((en.idx & 7) >> 1) : 0;
      yrd[i][ww] = a_sY + 8u * (uint32_t)eo;
    }
  }
  // where this lane's four values go in a staging buffer (buffer 0; element i at + kOff(i)):
  //   kind 0 -> E block, kind 1 -> delta (o_out) and the d/d dt column (o_out2), kind 2 -> its jet column
  const uint32_t lane_off = 8u * (uint32_t)(ccs * b + 2 * q);
  uint32_t o_out = a_stg + lane_off, o_out2 = 0;
  if (kind == 1) { o_out += 8u * p.o_sD; o_out2 = a_stg + 8u * (uint32_t)(p.o_sJ + mj * n_x) + lane_off; }
  if (kind == 2) o_out += 8u * (uint32_t)(p.o_sJ + jd * n_x);
  const uint32_t x_lane = 8u * (uint32_t)p.x_off + lane_off;   // state rows inside the staged slab

  // per-lane constants of the scalar pass (warp 0 of the group): theta_l, 1/l!
  const double th_l = c_theta[lane <= kMaxDeg ? lane : kMaxDeg];
  const double if_l = c_invfact[lane <= kMaxDeg ? lane : kMaxDeg];
  const double th_max = c_theta[kMaxDeg];

  const int TG = gridDim.x * p.gpc, gg = group * gridDim.x + blockIdx.x;
  const uint32_t zbytes = (uint32_t)p.zlen * 8u;
  if (p.bulk_in && gtid == 0 && gg < p.nk) {
    mbar_expect_tx(a_mbar, zbytes);
    bulk_g2s(a_grp, p.Z + (size_t)gg * p.D, zbytes, a_mbar);
  }

  int it = 0;
  for (int k = gg; k < p.nk; k += TG, ++it) {
    const int slot = it & 1;
    const uint32_t a_z = a_grp + (uint32_t)slot * 8u * (uint32_t)p.zpad;
    const uint32_t stg = (uint32_t)slot * 8u * (uint32_t)p.stg_stride;
    // ---- the (z_k, x_{k+1}) slab --------------------------------------------------------------
    PB2_STAMP(0);
    if (p.bulk_in) {
      mbar_wait(a_mbar + 8u * slot, (it >> 1) & 1);
      PB2_STAMP(1);
      if (gtid == 0) {
        // every warp left the previous knot at its last barrier: the other slab slot is free
        if (k + TG < p.nk) {
          mbar_expect_tx(a_mbar + 8u * (slot ^ 1), zbytes);
          bulk_g2s(a_grp + (uint32_t)(slot ^ 1) * 8u * (uint32_t)p.zpad, p.Z + (size_t)(k + TG) * p.D, zbytes,
                   a_mbar + 8u * (slot ^ 1));
        }
      }
    } else {
      const double* src = p.Z + (size_t)k * p.D;
      for (int e = gtid; e < p.zlen; e += gthreads) sts_f64<0>(a_z + 8u * e, src[e]);
      bar_sync(bar_all, gthreads);
    }
    // staging buffer `slot` was last used two knots ago: its bulk stores must have read it
    if (p.bulk_out && gtid == 0) bulk_wait_read1();

    // ---- G(u) = G0 + sum_j u_j G_j in permuted B-fragment order (shared by the group) --------
    for (int e = gtid; e < FR; e += gthreads) {
      double acc = lds_f64<0>(a_cG + 8u * e);
      for (int j = 0; j < m; ++j)
        acc = fma(lds_f64<0>(a_z + 8u * (p.u_off + j)), lds_f64<0>(a_cG + 8u * ((1 + j) * FR + e)), acc);
      sts_f64<0>(a_sA + 8u * e, acc);
    }
    // ---- warp 0: ||dt G(u)||_1 <= |dt| (||G_0|| + sum |u_j| ||G_j||) -> Taylor degree M, sub-steps,
    //      coefficients a_k = dt'^k / k!  (lane k) ----------------------------------------------------
    if (w == 0) {
      double dt = lds_f64<0>(a_z + 8u * p.dt_off);
      double nrm = lds_f64<0>(a_cG + 8u * p.o_norm);
      for (int j = 0; j < m; ++j)
        nrm = fma(fabs(lds_f64<0>(a_z + 8u * (p.u_off + j))), lds_f64<0>(a_cG + 8u * (p.o_norm + 1 + j)), nrm);
      nrm *= fabs(dt);
      int n_sub = 1;
      double per = nrm;
      if (nrm > th_max) {
        const double ns = ceil(nrm / th_max);
        if (ns <= (double)p.max_sub) {
          n_sub = (int)ns;
          dt = dt / ns;
          per = nrm / ns;
        } else {
          dt = __longlong_as_double(0x7ff8000000000000LL);  // norm beyond the supported range: NaN out
        }
      }
      // M = 1 + #{ l in 1..kMaxDeg-1 : theta_l < per }   (theta increasing; NaN -> M = 1)
      const unsigned below = __ballot_sync(0xffffffffu, lane >= 1 && lane < kMaxDeg && th_l < per);
      const int M = 1 + __popc(below);
      // dt^lane by binary powering
      double pw = 1.0, sq = dt;
#pragma unroll
      for (int bit = 0; bit < 5; ++bit) {
        if ((lane >> bit) & 1) pw *= sq;
        sq *= sq;
      }
      if (lane <= kMaxDeg) sts_f64<0>(a_sC + 8u * lane, lane <= M ? if_l * pw : 0.0);
      if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_meta), "r"(M), "r"(n_sub) : "memory");
    }
    PB2_STAMP(2);
    bar_sync(bar_all, gthreads);
    PB2_STAMP(3);

    int M, n_sub;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(M), "=r"(n_sub) : "r"(a_meta) : "memory");
    double A[KT][NT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) A[kt][nt] = lds_f64<0>(a_sA + 8u * ((kt * NT + nt) * 32 + lane));

    // ---- B = [I | X | 0 ...] ---------------------------------------------------------------------
    double t[KT], base[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      double v = ((diag >> i) & 1u) ? 1.0 : 0.0;
      if (kind == 1 && ((rowok >> i) & 1u)) v = lds_f64<0>(a_z + x_lane + 8u * (8 * (i >> 1) + (i & 1)));
      base[i] = v;
    }

    // ---- Taylor action: n_sub sub-steps of a degree-M Horner recurrence ----------------------
    PB2_STAMP(4);
    for (int sub = 0; sub < n_sub; ++sub) {
      // tiles without state columns start from B = unit columns / zero: no multiply needed
      const int mode = ((sub > 0) || tile_y) ? 1 : (tile_e ? 0 : 2);
      {
        const double cM = lds_f64<0>(a_sC + 8u * M);
#pragma unroll
        for (int i = 0; i < KT; ++i)
          t[i] = mode == 1 ? cM * base[i] : ((mode == 0 && ((diag >> i) & 1u)) ? cM : 0.0);
      }
      if (mode == 1) horner_run<NT, W, 1>(M, t, base, A, ev, yrd, ypub, a_sC, in_x, pub, tile_cpl, bar_x, nx, diag);
      else if (mode == 0) horner_run<NT, W, 0>(M, t, base, A, ev, yrd, ypub, a_sC, in_x, pub, tile_cpl, bar_x, nx, diag);
      else horner_run<NT, W, 2>(M, t, base, A, ev, yrd, ypub, a_sC, in_x, pub, tile_cpl, bar_x, nx, diag);
      if (sub + 1 < n_sub) {
#pragma unroll
        for (int i = 0; i < KT; ++i) base[i] = t[i];
        if (in_x) bar_sync(bar_x, nx);   // every reader of the exchange buffer is done before it restarts
      }
    }

    // ---- results into staging buffer `slot`, in final COO order ---------------------------------
    PB2_STAMP(5);
    if (kind == 0) {
      if (p.jac) {
#pragma unroll
        for (int i = 0; i < KT; ++i) {
          if ((rowok >> i) & 1u) {
            // -E (the d/dx_k block is I (x) E: one staging copy, n_b stores)
            const double v = -t[i];
            sts_f64<0>(o_out + stg + 8u * (8 * (i >> 1) + (i & 1)), v);
            if (p.iso) {
              // E = [[P, -Q], [Q, P]] : column half + c from column c
              const int r = 8 * (i >> 1) + 2 * q + (i & 1);
              const bool top = r < half;
              sts_f64<0>(a_stg + stg + 8u * (uint32_t)((ccs + half) * b + (top ? r + half : r - half)), top ? v : -v);
            }
          }
        }
      }
    } else if (kind == 1) {
      if (p.delta) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if ((rowok >> i) & 1u) {
            const uint32_t off = 8u * (8 * (i >> 1) + (i & 1));
            sts_f64<0>(o_out + stg + off, lds_f64<0>(a_z + 8u * p.D + x_lane + off) - t[i]);
          }
      }
    } else if (kind == 2) {
      if (p.jac) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if ((rowok >> i) & 1u) sts_f64<0>(o_out + stg + 8u * (8 * (i >> 1) + (i & 1)), -t[i]);
      }
    }
    if (tile_y && p.jac) {
      // d/d dt = -G(u) E x  : one more generator product on the state tile
      double d[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884z(d[nt], t[0], A[0][nt]);
#pragma unroll
      for (int kt = 1; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(d[nt], t[kt], A[kt][nt]);
      if (kind == 1) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if ((rowok >> i) & 1u) sts_f64<0>(o_out2 + stg + 8u * (8 * (i >> 1) + (i & 1)), -d[i >> 1][i & 1]);
      }
    }
    if (p.bulk_out) fence_proxy_async();
    PB2_STAMP(6);
    bar_sync(bar_all, gthreads);
    PB2_STAMP(7);

    // ---- staging -> HBM ---------------------------------------------------------------------------
    double* jac = p.jac ? p.jac + (size_t)k * p.nnz_jac : nullptr;
    double* dl = p.delta ? p.delta + (size_t)k * n_x : nullptr;
    if (p.bulk_out) {
      if (gtid == 0) {
        if (jac) {
          for (int copy = 0; copy < n_b; ++copy) bulk_s2g(jac + (size_t)copy * bb, a_stg + stg, (uint32_t)bb * 8u);
          bulk_s2g(jac + (size_t)n_b * bb, a_stg + stg + 8u * p.o_sJ, (uint32_t)((mj + 2) * n_x) * 8u);
        }
        if (dl) bulk_s2g(dl, a_stg + stg + 8u * p.o_sD, (uint32_t)n_x * 8u);
        bulk_commit();
      }
    } else {
      const double* sE = dmma_smem + p.o_grp + group * p.grp_stride + p.o_stg + slot * p.stg_stride;
      if (jac) {
        for (int copy = 0; copy < n_b; ++copy)
          for (int e = gtid; e < bb; e += gthreads) jac[(size_t)copy * bb + e] = sE[e];
        double* out = jac + (size_t)n_b * bb;
        for (int e = gtid; e < (mj + 2) * n_x; e += gthreads) out[e] = sE[p.o_sJ + e];
      }
      if (dl)
        for (int e = gtid; e < n_x; e += gthreads) dl[e] = sE[p.o_sD + e];
    }
  }
  if (p.bulk_out && gtid == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------
// host side: fragment / ELL tables, support test, launch
// ------------------------------------------------------------------------------------------
struct DmmaPlan {
  bool ok = false;
  int NT = 0, Bp = 0, W = 1, iso = 0, ncT = 0;
  bool persistent_ok = false;   // knot_dmma_kernel (NT <= 2, n_b <= 8) can take the shape; knot_dmmaq takes all of them
  int tiles_full = 0, tiles_res = 0;
  std::vector<double> gfrag;
  std::vector<EllEntry> ell;
  std::vector<double> norms;
};

inline int dmma_perm(int kt, int q) { return 8 * (kt / 2) + 2 * q + (kt % 2); }

// G0, Gj: host, column-major b x b.  allow_iso: the state uses the half (real-isomorphism) layout.
inline DmmaPlan dmma_plan(int b, int n_b, int m, bool allow_iso, const double* G0, const double* Gj) {
  DmmaPlan pl;
  if (b < 1 || b > 24 || n_b > 12) return pl;
  pl.NT = b <= 8 ? 1 : (b <= 16 ? 2 : 3);
  pl.persistent_ok = pl.NT <= 2 && n_b <= 8;
  pl.Bp = 8 * pl.NT;
  const int KT = 2 * pl.NT, NT = pl.NT, Bp = pl.Bp;
  auto at = [&](int mat, int r, int c) -> double {
    if (r >= b || c >= b) return 0.0;
    const double* A = mat == 0 ? G0 : Gj + (size_t)(mat - 1) * b * b;
    return A[r + (size_t)c * b];
  };
  // real-isomorphism structure [[S, R], [-R, S]] of every matrix?
  bool iso = allow_iso && (b % 2 == 0);
  if (iso) {
    const int h = b / 2;
    for (int mat = 0; mat <= m && iso; ++mat)
      for (int i = 0; i < h && iso; ++i)
        for (int j = 0; j < h; ++j)
          if (at(mat, i, j) != at(mat, i + h, j + h) || at(mat, i, j + h) != -at(mat, i + h, j)) {
            iso = false;
            break;
          }
  }
  pl.iso = iso ? 1 : 0;
  pl.ncT = iso ? b / 2 : b;
  const int cols_full = pl.ncT + (1 + m) * n_b, cols_res = n_b;
  pl.tiles_full = (cols_full + 7) / 8;
  pl.tiles_res = (cols_res + 7) / 8;
  if (pl.tiles_full > kDmmaMaxTiles) return pl;
  // ELL width of the drive generators, padded to 1, 2 or 4
  int W = 1;
  for (int j = 0; j < m; ++j)
    for (int r = 0; r < b; ++r) {
      int cnt = 0;
      for (int c = 0; c < b; ++c) cnt += at(1 + j, r, c) != 0.0;
      W = cnt > W ? cnt : W;
    }
  if (W > kDmmaMaxW) return pl;
  W = W <= 1 ? 1 : (W <= 2 ? 2 : 4);
  pl.W = W;
  pl.gfrag.assign((size_t)(m + 1) * KT * NT * 32, 0.0);
  for (int mat = 0; mat <= m; ++mat)
    for (int kt = 0; kt < KT; ++kt)
      for (int nt = 0; nt < NT; ++nt)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, q = lane & 3;
          pl.gfrag[((size_t)mat * KT * NT + kt * NT + nt) * 32 + lane] = at(mat, 8 * nt + g, dmma_perm(kt, q));
        }
  pl.ell.assign((size_t)(m + 1) * Bp * W, EllEntry{0.0, 0, 0});
  for (int j = 0; j < m; ++j)
    for (int r = 0; r < b; ++r) {
      int w = 0;
      for (int c = 0; c < b; ++c)
        if (at(1 + j, r, c) != 0.0) pl.ell[((size_t)j * Bp + r) * W + w++] = EllEntry{at(1 + j, r, c), c, 0};
    }
  // 1-norms (max absolute column sum) for the per-knot bound on ||dt G(u)||_1
  pl.norms.assign(m + 1, 0.0);
  for (int mat = 0; mat <= m; ++mat)
    for (int c = 0; c < b; ++c) {
      double cs = 0.0;
      for (int r = 0; r < b; ++r) cs += std::fabs(at(mat, r, c));
      pl.norms[mat] = cs > pl.norms[mat] ? cs : pl.norms[mat];
    }
  pl.ok = true;
  return pl;
}

// Shared-memory layout (doubles; every region 16-byte aligned).  CTA-wide: fragment tables
// [(m+1) FR], norms.  Per group: slab x2 | G(u) fragments | X exchange x2 | coefficients | meta |
// mbarriers x2 | staging x2 of [E (b*b) | J ((m_jets+2) n_x) | D (n_x)].
inline size_t dmma_layout(DmmaParams& q, int NT, int gpc) {
  auto even = [](int v) { return (v + 1) & ~1; };
  const int KT = 2 * NT, FR = KT * NT * 32, n_x = q.b * q.n_b;
  q.o_norm = (q.m + 1) * FR;
  q.o_grp = q.o_norm + even(q.m + 1);
  q.zpad = even(q.zlen);
  q.o_sA = 2 * q.zpad;
  q.o_sY = q.o_sA + FR;
  q.o_sC = q.o_sY + 2 * KT * 32;
  q.o_meta = q.o_sC + even(kMaxDeg + 1);
  q.o_mbar = q.o_meta + 2;
  q.o_stg = q.o_mbar + 2;
  q.o_sJ = even(q.b * q.b);
  q.o_sD = q.o_sJ + even((q.m_jets + 2) * n_x);
  q.stg_stride = q.o_sD + even(n_x);
  q.grp_stride = q.o_stg + 2 * q.stg_stride;
  return sizeof(double) * ((size_t)q.o_grp + (size_t)gpc * q.grp_stride);
}


using DmmaKernel = void (*)(DmmaParams);

inline DmmaKernel dmma_kernel(int NT, int W) {
  if (NT == 1) return W == 1 ? knot_dmma_kernel<1, 1> : (W == 2 ? knot_dmma_kernel<1, 2> : knot_dmma_kernel<1, 4>);
  return W == 1 ? knot_dmma_kernel<2, 1> : (W == 2 ? knot_dmma_kernel<2, 2> : knot_dmma_kernel<2, 4>);
}

}  // namespace pb2
