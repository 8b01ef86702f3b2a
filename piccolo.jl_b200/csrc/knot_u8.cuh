// Warp-specialised tensor-core knot kernel for the 3-qubit unitary shape (generator 16 x 16 with
// real-isomorphism structure, 8 state columns: BASELINE configs C3 / C5), residual + Jacobian.
//
// Same mathematics as knot_dmma.cuh (truncated-Taylor action of exp(dt G(u)) on the stacked
// columns [I | X | jet_1 .. jet_m] by DMMA.8x8x4 with register-resident transposed tiles; the
// reference path it replaces is DirectTrajOpt's BilinearIntegrator as built at
// /root/reference/src/control/integrators.jl:35-51), specialised so that every shared-memory
// offset is an immediate, and organised around what the first profiles showed: one knot is far
// too little work to hide its own latencies, and the FP64 tensor pipe (one DMMA per 16 cycles per
// SM sub-partition) idles whenever all warps of a knot wait on the same barrier.
//
//   * a knot GROUP = 1 producer warp + 1 warp for the tiles (E, X) + one warp per PAIR of drive
//     jets; up to 4 groups per persistent CTA work on 4 different knots, so while one group sits
//     in its exchange barrier or epilogue the others keep the pipe busy.  Roles are rotated by
//     group so that every sub-partition gets the same number of DMMA-issuing warps.
//   * the producer warp runs ahead of its group: it receives the (z_k, x_{k+1}) slab by TMA
//     (cp.async.bulk + mbarrier, 3-deep ring), builds G(u_k) in B-fragment order, picks the Taylor
//     degree and writes the coefficients a_k = dt^k / k!  -- all for knot i+1 while the compute
//     warps are inside knot i -- and hands it over through an mbarrier.  After the compute warps
//     have staged their results it replicates the propagator block (the Jacobian's d/dx_k block
//     is I (x) E: n_b = 8 copies) inside shared memory and issues ONE bulk store of the knot's
//     whole COO row segment (22.5 KB) plus one for delta.
//   * compute warps: per Horner step 16 DMMAs from registers, one named barrier for the X
//     exchange, then the additive terms (a_k B, sparse G_j S) -- nothing else.
#pragma once
#include "knot_dmma.cuh"

namespace pb2 {

struct U8Params {
  int m, D, x_off, dt_off, u_off, nnz_jac, max_sub, gpc, nk, zlen;
  int gw;                // warps per group: producer + (E,X) + ceil(m/2) jet warps
  int stagger;           // cycles by which a group with fewer knots than group 0 delays its start
  int stagger_g;         // additional delay of group g: g * stagger_g cycles
  int compact;           // 1: records [E columns 0..7 | jets, d/d dt | delta] of cstride doubles go to `jac`
  int cstride;
  int n_peers;           // > 0 (sharded run): every record is written into each of peers[0..n_peers-1] (this
                         // rank's gather buffer and, over NVLink, every other rank's) at p.jac's slot offset
  double* peers[8];
  int self;              // this rank's index in peers[]
  int peer_tma;          // 1: peers are written with bulk (TMA) stores too (needs cuMem-mapped peer memory)
  int direct_last;       // 1: a group's LAST knot leaves straight from the accumulator registers (st.global.v2):
                         // shortens the kernel's tail; every other knot goes through the stage + bulk store
  int dry;               // debug: 1 = exit after the prologue, 2 = exit immediately (launch-floor measurement)
  // shared-memory layout in doubles (u8_layout)
  int o_norm, o_tab, o_grp, grp_stride, zpad, o_prep, o_y, o_stage, o_mbar;
  const double* tables;  // [G fragments (m+1) 256 | norms (padded even) | theta_0..19 | 1/0! .. 1/19!], smem order
  const EllEntry* ell;   // (m+1) * 16 * W   (drive m = all-zero dummy)
  const double* Z;
  double* delta;         // may be null
  double* jac;
  long long* trace;      // debug build only: clock stamps of block 0
};

constexpr int kU8Prep = 280;        // doubles per prepared knot: G(u) frags 256 | a_k 20 | M, n_sub | pad
constexpr int kU8MaxGroups = 4;
constexpr int kU8MaxThreads = 512;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void lds_v2u32(uint32_t addr, int& a, int& b) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
}
template <int OFF>
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void stg_f64x2(double* ptr, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(ptr), "d"(a), "d"(b) : "memory");
}

// the per-step exchange barrier; the debug build can drop it (PB2_DRY=3: wrong numbers, upper bound
// on what decoupling the warps of a group could gain)
#ifdef PB2_TRACE
#define U8_BAR(id, n) do { if ((id) >= 0) bar_sync((id), (n)); } while (0)
#else
#define U8_BAR(id, n) bar_sync((id), (n))
#endif

// byte offset of element i (row 8 (i>>1) + 2q + (i&1)) relative to the lane's (column, 2q) address
#define U8_OFF(i) (((i) >> 1) * 64 + ((i) & 1) * 8)

// t <- G t for one 8-column tile (8 DMMAs, two independent accumulator chains)
__device__ __forceinline__ void u8_mma(double (&d)[2][2], const double (&t)[4], const double (&A)[4][2]) {
  dmma884z(d[0], t[0], A[0][0]);
  dmma884z(d[1], t[0], A[0][1]);
#pragma unroll
  for (int kt = 1; kt < 4; ++kt) {
    dmma884(d[0], t[kt], A[kt][0]);
    dmma884(d[1], t[kt], A[kt][1]);
  }
}

// d <- d + G t: the accumulators arrive pre-loaded with everything else the Horner step adds, so that
// no FP64 CUDA-core instruction sits between two steps' tensor instructions.  (A dependent DFMA / DADD
// issued while other warps of the sub-partition have DMMAs queued waits for the whole backlog: 265
// cycles per busy warp measured with tools/dfma_lat.cu, against 8.5 on an idle pipe.)
__device__ __forceinline__ void u8_mma_acc(double (&d)[2][2], const double (&t)[4], const double (&A)[4][2]) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    dmma884(d[0], t[kt], A[kt][0]);
    dmma884(d[1], t[kt], A[kt][1]);
  }
}

// G(u) = G0 + sum_j u_j G_j for this lane's 8 B-fragment slots (all loads first, then the arithmetic)
template <bool LOWREG>
__device__ __forceinline__ void u8_build_G(uint32_t a_z, uint32_t a_cG, int lane, int m, int u_off, double (&acc)[8]) {
  if (LOWREG) {   // one drive at a time: a few registers, used once per group at start-up
#pragma unroll
    for (int s = 0; s < 8; ++s) acc[s] = lds_f64<0>(a_cG + 8u * (uint32_t)(s * 32 + lane));
    for (int j = 0; j < m; ++j) {
      const double u = lds_f64<0>(a_z + 8u * (uint32_t)(u_off + j));
      const uint32_t a_gj = a_cG + 8u * (uint32_t)((1 + j) * 256 + lane);
      double gvj[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) gvj[s] = lds_f64<0>(a_gj + 256u * s);
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[s] = fma(u, gvj[s], acc[s]);
    }
    return;
  }
  double uj[6], gv[6][8];
#pragma unroll
  for (int j = 0; j < 6; ++j) uj[j] = j < m ? lds_f64<0>(a_z + 8u * (uint32_t)(u_off + j)) : 0.0;
#pragma unroll
  for (int s = 0; s < 8; ++s) acc[s] = lds_f64<0>(a_cG + 8u * (uint32_t)(s * 32 + lane));
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < m) {
      const uint32_t a_gj = a_cG + 8u * (uint32_t)((1 + j) * 256 + lane);
#pragma unroll
      for (int s = 0; s < 8; ++s) gv[j][s] = lds_f64<0>(a_gj + 256u * s);
    }
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j < m) {
      if (j >= 4) {   // drives 5, 6: loaded late to bound register use
        const uint32_t a_gj = a_cG + 8u * (uint32_t)((1 + j) * 256 + lane);
#pragma unroll
        for (int s = 0; s < 8; ++s) gv[j][s] = lds_f64<0>(a_gj + 256u * s);
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[s] = fma(uj[j], gv[j][s], acc[s]);
    }
}

// ---- one Horner step of the (E, X) warp:  t <- c_k b + G t.  The caller keeps `acc` = c_k b for THIS
// step (computed while the previous step's tensor instructions were in flight) and gets it back holding
// c_{k-1} b for the next one; nothing but the publish and the barrier separates two steps' products.
// FIRST: first sub-step (b_E = unit columns: the coefficient lands on one element, no arithmetic).
template <int PAR, bool FIRST>
__device__ __forceinline__ void u8_step_ex(double (&tE)[4], double (&tX)[4], const double (&bE)[4],
                                           const double (&bX)[4], const double (&A)[4][2], int iE, uint32_t ypub,
                                           uint32_t ck_next_addr, int xbar, int nx, double (&accE)[4],
                                           double (&accX)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) sts_f64<PAR * 1024>(ypub + i * 256, tX[i]);
  double dE[2][2], dX[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dX[i >> 1][i & 1] = accX[i];
    dE[i >> 1][i & 1] = accE[i];
  }
  U8_BAR(xbar, nx);
  u8_mma_acc(dX, tX, A);
  u8_mma_acc(dE, tE, A);
  // the next step's additive terms, queued behind the products just issued
  const double ckn = lds_f64<0>(ck_next_addr);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accX[i] = ckn * bX[i];
    if (FIRST) accE[i] = (i == iE) ? ckn : 0.0;
    else accE[i] = ckn * bE[i];
    tX[i] = dX[i >> 1][i & 1];
    tE[i] = dE[i >> 1][i & 1];
  }
}

// ---- one Horner step of a jet warp (two jet tiles):  t <- c_k b + G t + G_j y,  y the state iterate
// the (E, X) warp published for this step.  The coupling term (and c_k b) is formed first and handed
// to the tensor instructions as their accumulator.  FIRST: B_j = 0; NOMMA: the iterate is zero.
template <int W, int PAR, bool FIRST, bool NOMMA>
__device__ __forceinline__ void u8_step_jets(double (&t)[2][4], const double (&bJ)[2][4], const double (&A)[4][2],
                                             const double (&ev)[2][4][W], const uint32_t (&yad)[2][4][W],
                                             uint32_t ck_addr, bool two, int xbar, int nx) {
  double ck = 0.0;
  if (!FIRST) ck = lds_f64<0>(ck_addr);
  U8_BAR(xbar, nx);
  double y[2][4][W];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int ww = 0; ww < W; ++ww) y[a][i][ww] = lds_f64<PAR * 1024>(yad[a][i][ww]);
  double d[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double v = FIRST ? ev[a][i][0] * y[a][i][0] : fma(ev[a][i][0], y[a][i][0], ck * bJ[a][i]);
#pragma unroll
      for (int ww = 1; ww < W; ++ww) v = fma(ev[a][i][ww], y[a][i][ww], v);
      d[a][i >> 1][i & 1] = v;
    }
  if (!NOMMA) {
    u8_mma_acc(d[0], t[0], A);
    if (two) u8_mma_acc(d[1], t[1], A);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i) t[a][i] = d[a][i >> 1][i & 1];
}

#ifdef PB2_TRACE
#define U8_STAMP(i) do { if (blockIdx.x == 0 && lane == 0 && i_knot < 4) p.trace[((wcta * 4 + i_knot) * 8) + (i)] = clock64(); } while (0)
#define U8_STAMPX(i) do { if (blockIdx.x == 0 && lane == 0 && i_knot == 0) p.trace[((wcta * 4 + 3) * 8) + (i)] = clock64(); } while (0)
#else
#define U8_STAMPX(i) do { } while (0)
#define U8_STAMP(i) do { } while (0)
#endif

template <int W>
__global__ void __launch_bounds__(kU8MaxThreads, 1) knot_u8_kernel(const __grid_constant__ U8Params p) {
  extern __shared__ __align__(16) double u8_smem[];
  const int lane = threadIdx.x & 31, wcta = threadIdx.x >> 5;
  const int gw = p.gw, group = wcta / gw, wg = wcta - group * gw;
  const int role = (wg + group) % gw;   // 0 producer, 1 (E, X) tiles, 2 + jw: jets 2 jw, 2 jw + 1
  const int g = lane >> 2, q = lane & 3;
#ifdef PB2_TRACE
  const int m = p.m, ncw = gw - 1, nx = 32 * ncw, xbar = p.dry == 3 ? -1 : 1 + group;
#else
  const int m = p.m, ncw = gw - 1, nx = 32 * ncw, xbar = 1 + group;
#endif

  const uint32_t a_cG = smem_u32(u8_smem);
  const uint32_t a_grp = a_cG + 8u * (uint32_t)(p.o_grp + group * p.grp_stride);
  const uint32_t a_prep = a_grp + 8u * p.o_prep, a_y = a_grp + 8u * p.o_y, a_stage = a_grp + 8u * p.o_stage;
  const uint32_t a_mbar = a_grp + 8u * p.o_mbar;
  const uint32_t mb_zfull = a_mbar, mb_ready = a_mbar + 24, mb_staged = a_mbar + 40, mb_free = a_mbar + 48;
  const uint32_t o_J = 8u * 2048u, o_D = 8u * (2048u + (uint32_t)(m + 2) * 128u);   // inside the stage

  const int TG = gridDim.x * p.gpc, gg = group * gridDim.x + blockIdx.x;
  const int n_my = gg < p.nk ? (p.nk - gg + TG - 1) / TG : 0;
  const int n_max = ((int)blockIdx.x < p.nk) ? (p.nk - (int)blockIdx.x + TG - 1) / TG : 0;   // group 0's count
  const uint32_t zbytes = (uint32_t)p.zlen * 8u;

  // ---- once per CTA ------------------------------------------------------------------------------
  // the producer lane first arms its group's mbarriers and starts the first two slab loads, so that
  // the HBM latency of the slabs overlaps the table fill below
  if (p.dry == 2) return;
  // programmatic dependent launch: let the next grid on the stream start its own prologue as
  // SMs drain, and do not touch trajectory / output memory before the previous grid is complete
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t mb_tab = a_cG + 8u * (uint32_t)(p.o_tab + 40);
  if (threadIdx.x == 0) {
    // the handle's constant tables (G fragments, norms, theta / factorial tables: contiguous in HBM
    // in shared-memory order) arrive by ONE bulk copy; only the producer warps ever wait for it
    mbar_init(mb_tab, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(mb_tab, 8u * (uint32_t)(p.o_tab + 40));
    bulk_g2s(a_cG, p.tables, 8u * (uint32_t)(p.o_tab + 40), mb_tab);
  }
  if (role == 0 && lane == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(mb_zfull + 8 * i, 1);
    mbar_init(mb_ready, 1);
    mbar_init(mb_ready + 8, 1);
    mbar_init(mb_staged, ncw);
    mbar_init(mb_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    double* ones = u8_smem + p.o_grp + group * p.grp_stride + p.o_stage + 2048 + (m + 1) * 128;
    for (int e = wg * 32 + lane; e < 128; e += 32 * gw) ones[e] = 1.0;
  }
  __syncthreads();
  if (p.dry == 1) return;

  if (role == 0) {
    // =============================== producer warp ===============================================
    if (lane == 0) {
      // pull the first slabs towards L2 while the previous grid drains (a hint: no data is consumed
      // before the dependency wait), then wait for the previous grid and start the slab ring
      for (int i = 0; i < 2 && i < n_my; ++i)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.Z + (size_t)(gg + i * TG) * p.D), "r"(zbytes)
                     : "memory");
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int i = 0; i < 2 && i < n_my; ++i) {
        mbar_expect_tx(mb_zfull + 8 * i, zbytes);
        bulk_g2s(a_grp + 8u * (uint32_t)(i * p.zpad), p.Z + (size_t)(gg + i * TG) * p.D, zbytes, mb_zfull + 8 * i);
      }
      mbar_arrive(mb_free);   // the stage starts free
    }
    mbar_wait(mb_tab, 0);
    const double th_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double if_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + 20 + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double th_max = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + kMaxDeg));
    const long long t_begin = clock64();
    int s3 = 0;
    for (int i = 0; i <= n_my; ++i) {
      if (i < n_my) {
        // ---- prepare knot i: G(u), Taylor degree, coefficients ------------------------------------
        const uint32_t a_z = a_grp + 8u * (uint32_t)(s3 * p.zpad);
        const uint32_t a_p = a_prep + 8u * (uint32_t)((i & 1) * kU8Prep);
        const int i_knot = i;
        U8_STAMP(0);
        mbar_wait(mb_zfull + 8 * s3, (uint32_t)((i / 3) & 1));
        U8_STAMP(1);
        // the first knot's G(u) is built by the compute warps themselves (they are idle anyway, and it
        // takes the hand-over off the start-up path); from then on this warp runs ahead of them
        U8_STAMPX(0);
        double dt = lds_f64<0>(a_z + 8u * p.dt_off);
        double nrm = lds_f64<0>(a_cG + 8u * p.o_norm);
        for (int j = 0; j < m; ++j)
          nrm = fma(fabs(lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + j))),
                    lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_norm + 1 + j)), nrm);
        U8_STAMPX(1);
        if (i > 0) {
          double acc[8];
          u8_build_G<false>(a_z, a_cG, lane, m, p.u_off, acc);
#pragma unroll
          for (int s = 0; s < 8; ++s) sts_f64<0>(a_p + 8u * (uint32_t)(s * 32 + lane), acc[s]);
        }
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
        U8_STAMPX(2);
        const unsigned below = __ballot_sync(0xffffffffu, lane >= 1 && lane < kMaxDeg && th_l < per);
        const int M = 1 + __popc(below);
        U8_STAMPX(3);
        double pw = 1.0, sq = dt;   // dt^lane by binary powering
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
          if ((lane >> bit) & 1) pw *= sq;
          sq *= sq;
        }
        U8_STAMPX(4);
        if (lane <= kMaxDeg) sts_f64<0>(a_p + 8u * 256u + 8u * lane, lane <= M ? if_l * pw : 0.0);
        if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_p + 8u * 276u), "r"(M), "r"(n_sub) : "memory");
        __syncwarp();
        U8_STAMPX(5);
        if (i == 0) {
          // groups that run in lockstep leave the tensor pipe idle, and flood L2 with their stores,
          // all at the same time: de-phase them.  A group with one knot less than its neighbours
          // starts later still, so that its knot fills the pipe while the others sit between two knots.
          const long long t_end = t_begin + (long long)group * p.stagger_g + (n_my < n_max ? (long long)p.stagger : 0);
          while (clock64() < t_end) { }
        }
        if (lane == 0) mbar_arrive(mb_ready + 8 * (i & 1));
        U8_STAMP(2);
      }
      const bool direct_prev = p.direct_last && p.n_peers == 0 && (i - 1) == n_my - 1;   // that knot's warps stored it themselves
      if (i >= 1 && !direct_prev) {
        // ---- finish knot i-1: replicate the propagator block, one bulk store per output --------
        const int kprev = gg + (i - 1) * TG;
        const int i_knot = i - 1;
        U8_STAMP(3);
        mbar_wait(mb_staged, (uint32_t)((i - 1) & 1));
        U8_STAMP(4);
        if (!p.compact) {
          double2 v[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) v[jj] = lds_f64x2<0>(a_stage + 16u * (uint32_t)(lane + 32 * jj));
#pragma unroll
          for (int c = 1; c < 8; ++c)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) sts_f64x2<0>(a_stage + 2048u * c + 16u * (uint32_t)(lane + 32 * jj), v[jj]);
        }
        if (p.n_peers > 1 && !p.peer_tma) {
          // fused exchange: the record also goes straight into every other rank's gather buffer over
          // NVLink.  Bulk (TMA) stores fault on CUDA-IPC peer mappings, so these are 16-byte st.global
          // by this warp, read from the stage; they overlap the compute warps' next knot.
          const uint32_t a_rec = a_stage + o_J - 1024u;
          const size_t off = (size_t)(p.jac - p.peers[p.self]) + (size_t)kprev * p.cstride;
          for (int e = lane; e < p.cstride / 2; e += 32) {
            const double2 v2 = lds_f64x2<0>(a_rec + 16u * (uint32_t)e);
            for (int r = 0; r < p.n_peers; ++r)
              if (r != p.self) stg_f64x2(p.peers[r] + off + 2 * e, v2.x, v2.y);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.compact) {
            // record = [E columns 0..7 (128) | jets, d/d dt | delta]: the least that has to cross
            // NVLink / PCIe (the other half of E is its mirror image, the identity entries are constant).
            // In this mode the compute warps stage it contiguously (E half right below the jets, delta in
            // the slot of the identity entries), so it leaves with one bulk store per destination.
            const uint32_t a_rec = a_stage + o_J - 1024u, nbytes = (uint32_t)p.cstride * 8u;
            bulk_s2g(p.jac + (size_t)kprev * p.cstride, a_rec, nbytes);
            if (p.n_peers > 1 && p.peer_tma) {
              const size_t off = (size_t)(p.jac - p.peers[p.self]) + (size_t)kprev * p.cstride;
              for (int r = 0; r < p.n_peers; ++r)
                if (r != p.self) bulk_s2g(p.peers[r] + off, a_rec, nbytes);
            }
          } else {
            bulk_s2g(p.jac + (size_t)kprev * p.nnz_jac, a_stage, (uint32_t)p.nnz_jac * 8u);
            if (p.delta) bulk_s2g(p.delta + (size_t)kprev * 128, a_stage + o_D, 1024u);
          }
          bulk_commit();
        }
        U8_STAMP(5);
      }
      if (lane == 0) {
        // slab (i+2)%3 held knot i-1, which is finished
        if (i + 2 < n_my) {
          const int s = s3 == 0 ? 2 : s3 - 1;   // (i + 2) % 3
          mbar_expect_tx(mb_zfull + 8 * s, zbytes);
          bulk_g2s(a_grp + 8u * (uint32_t)(s * p.zpad), p.Z + (size_t)(gg + (i + 2) * TG) * p.D, zbytes,
                   mb_zfull + 8 * s);
        }
        if (i >= 1 && !direct_prev) {
          bulk_wait_read0();        // the stage has been read: the compute warps may refill it
          mbar_arrive(mb_free);
          { const int i_knot = i - 1; U8_STAMP(6); }
        }
      }
      __syncwarp();
      s3 = s3 == 2 ? 0 : s3 + 1;
    }
    // the stage must outlive the bulk stores' reads of it; their writes complete with the grid
    if (p.n_peers > 1) __threadfence_system();   // remote writes are performed before the grid completes
    if (lane == 0) {
      if (p.n_peers > 1 && p.peer_tma) bulk_wait0();
      else bulk_wait_read0();
    }
    return;
  }

  // ================================= compute warps ==================================================
  const uint32_t lane_col = 8u * (uint32_t)(g * 16 + 2 * q);   // (column g, row 2q) inside a 16 x 8 block
  const uint32_t a_z0 = a_grp;                                  // slab slot 0 holds the group's first knot
  if (role == 1) {
    // ---- tiles E (columns 0..7 of the propagator) and X (the 8 state columns) -------------------
    const int iE = (g == 2 * q) ? 0 : ((g == 2 * q + 1) ? 1 : -1);   // which element is the unit entry
    const uint32_t ypub0 = a_y + 8u * (uint32_t)(g * 4 + q);
    const uint32_t oE1 = a_stage + lane_col, oE2 = a_stage + 8u * 128u + lane_col;
    const uint32_t oD = a_stage + o_D + lane_col, oT = a_stage + o_J + 8u * (uint32_t)(m * 128) + lane_col;
    const uint32_t xl = 8u * (uint32_t)p.x_off + lane_col;
    int s3 = 0;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t a_z = a_grp + 8u * (uint32_t)(s3 * p.zpad);
      const uint32_t a_p = a_prep + 8u * (uint32_t)((i & 1) * kU8Prep);
      const uint32_t a_c = a_p + 8u * 256u;
      const int i_knot = i;
      U8_STAMP(0);
      double A[4][2];
      if (i == 0) {
        // first knot: build G(u) here while the producer derives the Taylor degree and coefficients
        double acc[8];
        mbar_wait(mb_tab, 0);
        mbar_wait(mb_zfull, 0);
        u8_build_G<true>(a_z0, a_cG, lane, m, p.u_off, acc);
#pragma unroll
        for (int s = 0; s < 8; ++s) A[s >> 1][s & 1] = acc[s];
      }
      mbar_wait(mb_ready + 8 * (i & 1), (uint32_t)((i >> 1) & 1));
      U8_STAMP(1);
      if (i > 0) {
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) A[kt][nt] = lds_f64<0>(a_p + 8u * (uint32_t)((kt * 2 + nt) * 32 + lane));
      }
      int M, n_sub;
      lds_v2u32(a_p + 8u * 276u, M, n_sub);
      double bX[4], tE[4], tX[4];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) bX[i4] = lds_f64<0>(a_z + xl + U8_OFF(i4));
      {
        const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          tE[i4] = (i4 == iE) ? cM : 0.0;
          tX[i4] = cM * bX[i4];
        }
      }
      // exchange buffers alternate with the knot's parity: readers of the previous knot are never
      // overtaken (this warp passes a knot's step barriers only together with them)
      const uint32_t ypub = ypub0 + (uint32_t)(i & 1) * 2048u;
      U8_STAMP(2);
      double accE[4], accX[4];   // c_k b of the coming step
      {
        int kq = M - 1;
        const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)kq);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          accE[i4] = (i4 == iE) ? c0 : 0.0;
          accX[i4] = c0 * bX[i4];
        }
        // (a step asks for the coefficient after its own; the last one re-reads c_0, unused)
        for (; kq >= 1; kq -= 2) {
          u8_step_ex<0, true>(tE, tX, bX, bX, A, iE, ypub, a_c + 8u * kq - 8u, xbar, nx, accE, accX);
          u8_step_ex<1, true>(tE, tX, bX, bX, A, iE, ypub, a_c + 8u * (uint32_t)(kq >= 2 ? kq - 2 : 0), xbar, nx, accE, accX);
        }
        if (kq == 0) u8_step_ex<0, true>(tE, tX, bX, bX, A, iE, ypub, a_c, xbar, nx, accE, accX);
      }
      // further sub-steps (||dt G|| beyond the largest tabulated radius: rare), general B
      for (int sub = 1; sub < n_sub; ++sub) {
        double bE2[4], bX2[4];
        const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          bE2[i4] = tE[i4];
          bX2[i4] = tX[i4];
          tE[i4] *= cM;
          tX[i4] *= cM;
        }
        U8_BAR(xbar, nx);
        int kq = M - 1;
        const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)kq);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          accE[i4] = c0 * bE2[i4];
          accX[i4] = c0 * bX2[i4];
        }
        for (; kq >= 1; kq -= 2) {
          u8_step_ex<0, false>(tE, tX, bE2, bX2, A, iE, ypub, a_c + 8u * kq - 8u, xbar, nx, accE, accX);
          u8_step_ex<1, false>(tE, tX, bE2, bX2, A, iE, ypub, a_c + 8u * (uint32_t)(kq >= 2 ? kq - 2 : 0), xbar, nx, accE, accX);
        }
        if (kq == 0) u8_step_ex<0, false>(tE, tX, bE2, bX2, A, iE, ypub, a_c, xbar, nx, accE, accX);
      }
      // ---- d/d dt = -G(u) E x : one more generator product on the state tile --------------------
      U8_STAMP(3);
      double dT[2][2];
      u8_mma(dT, tX, A);
      double xn[4];
      const bool want_delta = p.delta != nullptr || p.compact;
      if (want_delta) {
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) xn[i4] = lds_f64<0>(a_z + 8u * p.D + xl + U8_OFF(i4));
      }
      U8_STAMP(4);
      if (p.direct_last && p.n_peers == 0 && i == n_my - 1) {
        // the group's last knot: results straight from the accumulator registers, 16-byte stores
        const size_t kk = (size_t)(gg + i * TG);
        double* jk = p.jac + kk * (size_t)(p.compact ? p.cstride : p.nnz_jac);
        const int lc = g * 16 + 2 * q;
        if (p.compact) {
          stg_f64x2(jk + lc, -tE[0], -tE[1]);
          stg_f64x2(jk + lc + 8, -tE[2], -tE[3]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            // -E = -[[P, -Q], [Q, P]]: own column g, mirrored column g + 8
            stg_f64x2(jk + c * 256 + lc, -tE[0], -tE[1]);
            stg_f64x2(jk + c * 256 + lc + 8, -tE[2], -tE[3]);
            stg_f64x2(jk + c * 256 + 128 + lc + 8, -tE[0], -tE[1]);
            stg_f64x2(jk + c * 256 + 128 + lc, tE[2], tE[3]);
          }
        }
        double* jj = jk + (p.compact ? 128 : 2048);
        stg_f64x2(jj + m * 128 + lc, -dT[0][0], -dT[0][1]);
        stg_f64x2(jj + m * 128 + lc + 8, -dT[1][0], -dT[1][1]);
        if (want_delta) {
          double* dd = (p.compact ? jj + (m + 1) * 128 : p.delta + kk * 128) + lc;
          stg_f64x2(dd, xn[0] - tX[0], xn[1] - tX[1]);
          stg_f64x2(dd + 8, xn[2] - tX[2], xn[3] - tX[3]);
        }
        break;
      }
      mbar_wait(mb_free, (uint32_t)(i & 1));
      U8_STAMP(5);
      if (p.compact) {
        // contiguous record [E columns 0..7 | jets | d/d dt | delta]: E half right below the jets, delta
        // where the identity entries live in the canonical layout
        const uint32_t oEc = a_stage + o_J - 1024u + lane_col, oDc = a_stage + o_J + 8u * (uint32_t)((m + 1) * 128) + lane_col;
        sts_f64<0>(oEc, -tE[0]);   sts_f64<8>(oEc, -tE[1]);   sts_f64<64>(oEc, -tE[2]);  sts_f64<72>(oEc, -tE[3]);
        sts_f64<0>(oDc, xn[0] - tX[0]);   sts_f64<8>(oDc, xn[1] - tX[1]);
        sts_f64<64>(oDc, xn[2] - tX[2]);  sts_f64<72>(oDc, xn[3] - tX[3]);
      } else {
        // -E = -[[P, -Q], [Q, P]]: own column g, mirrored column g + 8
        sts_f64<0>(oE1, -tE[0]);   sts_f64<8>(oE1, -tE[1]);   sts_f64<64>(oE1, -tE[2]);  sts_f64<72>(oE1, -tE[3]);
        sts_f64<64>(oE2, -tE[0]);  sts_f64<72>(oE2, -tE[1]);  sts_f64<0>(oE2, tE[2]);    sts_f64<8>(oE2, tE[3]);
        if (want_delta) {
          sts_f64<0>(oD, xn[0] - tX[0]);   sts_f64<8>(oD, xn[1] - tX[1]);
          sts_f64<64>(oD, xn[2] - tX[2]);  sts_f64<72>(oD, xn[3] - tX[3]);
        }
      }
      sts_f64<0>(oT, -dT[0][0]);   sts_f64<8>(oT, -dT[0][1]);
      sts_f64<64>(oT, -dT[1][0]);  sts_f64<72>(oT, -dT[1][1]);
      __syncwarp();   // the producer orders these writes before its bulk store (fence.proxy.async after the acquire)
      if (lane == 0) mbar_arrive(mb_staged);
      U8_STAMP(6);
      s3 = s3 == 2 ? 0 : s3 + 1;
    }
    return;
  }

  {
    // ---- jets of drives jd0 = 2 jw and jd1 = 2 jw + 1 ---------------------------------------------
    const int jw = role - 2;
    const int jd0 = 2 * jw, jd1 = 2 * jw + 1;
    const bool two = jd1 < m;
    double ev[2][4][W];
    uint32_t yad[2][4][W];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int jd = a == 0 ? jd0 : (two ? jd1 : m);   // drive m is the all-zero dummy
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const int r = 8 * (i4 >> 1) + 2 * q + (i4 & 1);
#pragma unroll
        for (int ww = 0; ww < W; ++ww) {
          const EllEntry en = p.ell[((size_t)jd * 16 + r) * W + ww];
          ev[a][i4][ww] = en.val;
          yad[a][i4][ww] = a_y + 8u * (uint32_t)((2 * (en.idx >> 3) + (en.idx & 1)) * 32 + g * 4 + ((en.idx & 7) >> 1));
        }
      }
    }
    const uint32_t oJ0 = a_stage + o_J + 8u * (uint32_t)(jd0 * 128) + lane_col;
    const uint32_t oJ1 = a_stage + o_J + 8u * (uint32_t)(jd1 * 128) + lane_col;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t a_p = a_prep + 8u * (uint32_t)((i & 1) * kU8Prep);
      const uint32_t a_c = a_p + 8u * 256u;
      const int i_knot = i;
      U8_STAMP(0);
      double A[4][2];
      if (i == 0) {
        // first knot: build G(u) here while the producer derives the Taylor degree and coefficients
        double acc[8];
        mbar_wait(mb_tab, 0);
        mbar_wait(mb_zfull, 0);
        u8_build_G<true>(a_z0, a_cG, lane, m, p.u_off, acc);
#pragma unroll
        for (int s = 0; s < 8; ++s) A[s >> 1][s & 1] = acc[s];
      }
      mbar_wait(mb_ready + 8 * (i & 1), (uint32_t)((i >> 1) & 1));
      U8_STAMP(1);
      if (i > 0) {
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) A[kt][nt] = lds_f64<0>(a_p + 8u * (uint32_t)((kt * 2 + nt) * 32 + lane));
      }
      int M, n_sub;
      lds_v2u32(a_p + 8u * 276u, M, n_sub);
      double t[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) t[a][i4] = 0.0;
      if (i > 0) {
        // switch to the other pair of exchange buffers
        const uint32_t dy = (i & 1) ? 2048u : (uint32_t)-2048;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
            for (int ww = 0; ww < W; ++ww) yad[a][i4][ww] += dy;
      }
      U8_STAMP(2);
      {
        // the jets start from zero: the first step is the coupling term alone
        int kq = M - 2;
        u8_step_jets<W, 0, true, true>(t, t, A, ev, yad, a_c, two, xbar, nx);
        for (; kq >= 1; kq -= 2) {
          u8_step_jets<W, 1, true, false>(t, t, A, ev, yad, a_c, two, xbar, nx);
          u8_step_jets<W, 0, true, false>(t, t, A, ev, yad, a_c, two, xbar, nx);
        }
        if (kq == 0) u8_step_jets<W, 1, true, false>(t, t, A, ev, yad, a_c, two, xbar, nx);
      }
      for (int sub = 1; sub < n_sub; ++sub) {
        double bJ[2][4];
        const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            bJ[a][i4] = t[a][i4];
            t[a][i4] *= cM;
          }
        U8_BAR(xbar, nx);
        int kq = M - 1;
        for (; kq >= 1; kq -= 2) {
          u8_step_jets<W, 0, false, false>(t, bJ, A, ev, yad, a_c + 8u * kq, two, xbar, nx);
          u8_step_jets<W, 1, false, false>(t, bJ, A, ev, yad, a_c + 8u * kq - 8u, two, xbar, nx);
        }
        if (kq == 0) u8_step_jets<W, 0, false, false>(t, bJ, A, ev, yad, a_c, two, xbar, nx);
      }
      U8_STAMP(4);
      if (p.direct_last && p.n_peers == 0 && i == n_my - 1) {
        const size_t kk = (size_t)(gg + i * TG);
        double* jj = p.jac + kk * (size_t)(p.compact ? p.cstride : p.nnz_jac) + (p.compact ? 128 : 2048);
        const int lc = g * 16 + 2 * q;
        stg_f64x2(jj + jd0 * 128 + lc, -t[0][0], -t[0][1]);
        stg_f64x2(jj + jd0 * 128 + lc + 8, -t[0][2], -t[0][3]);
        if (two) {
          stg_f64x2(jj + jd1 * 128 + lc, -t[1][0], -t[1][1]);
          stg_f64x2(jj + jd1 * 128 + lc + 8, -t[1][2], -t[1][3]);
        }
        if (jw == 0 && !p.compact) {   // the constant d/dx_{k+1} identity entries
          stg_f64x2(jj + (m + 1) * 128 + 2 * lane, 1.0, 1.0);
          stg_f64x2(jj + (m + 1) * 128 + 64 + 2 * lane, 1.0, 1.0);
        }
        break;
      }
      mbar_wait(mb_free, (uint32_t)(i & 1));
      U8_STAMP(5);
      sts_f64<0>(oJ0, -t[0][0]);   sts_f64<8>(oJ0, -t[0][1]);
      sts_f64<64>(oJ0, -t[0][2]);  sts_f64<72>(oJ0, -t[0][3]);
      if (two) {
        sts_f64<0>(oJ1, -t[1][0]);   sts_f64<8>(oJ1, -t[1][1]);
        sts_f64<64>(oJ1, -t[1][2]);  sts_f64<72>(oJ1, -t[1][3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(mb_staged);
      U8_STAMP(6);
    }
  }
}

// Shared-memory layout (doubles; every region 16-byte aligned).  CTA-wide: fragment tables
// [(m+1) 256], norms.  Per group: slab x3 | prepared knot x2 | X exchange x4 | stage
// [E x8 (2048) | jets, d/d dt, ones ((m+2) 128) | delta (128)] | mbarriers.
inline size_t u8_layout(U8Params& q, int gpc) {
  auto even = [](int v) { return (v + 1) & ~1; };
  q.o_norm = (q.m + 1) * 256;
  q.o_tab = q.o_norm + even(q.m + 1);
  q.o_grp = q.o_tab + 40 + 2;   // + the tables' mbarrier
  q.zpad = even(q.zlen);
  q.o_prep = 3 * q.zpad;
  q.o_y = q.o_prep + 2 * kU8Prep;
  q.o_stage = q.o_y + 4 * 128;
  q.o_mbar = q.o_stage + 2048 + (q.m + 2) * 128 + 128;
  q.grp_stride = q.o_mbar + 8;
  return sizeof(double) * ((size_t)q.o_grp + (size_t)gpc * q.grp_stride);
}

// Compact records -> canonical arrays.  Record = [E columns 0..7 (128) | jets, d/d dt ((m+1) 128) |
// delta (128)]: the propagator block is rebuilt from its first half (-E = -[[P,-Q],[Q,P]]), written
// n_b = 8 times, and the constant identity entries are filled in.  Pure data movement (HBM-bound);
// one CTA per knot, grid-stride.
__global__ void __launch_bounds__(256) expand_compact_kernel(const double* __restrict__ comp, long long n_knots,
                                                             int cstride, int nJd, int nnz_jac,
                                                             double* __restrict__ delta, double* __restrict__ jac) {
  for (long long r = blockIdx.x; r < n_knots; r += gridDim.x) {
    const double* src = comp + r * cstride;
    double* out = jac + r * nnz_jac;
    {
      const int e = threadIdx.x, col = e >> 4, row = e & 15;   // 256 threads <-> the 16 x 16 block
      double v;
      if (col < 8) v = src[e];
      else v = row < 8 ? -src[(col - 8) * 16 + 8 + row] : src[(col - 8) * 16 + row - 8];
#pragma unroll
      for (int c = 0; c < 8; ++c) out[c * 256 + e] = v;
    }
    out += 2048;
    for (int e = threadIdx.x; e < nJd; e += blockDim.x) out[e] = src[128 + e];
    if (threadIdx.x < 128) out[nJd + threadIdx.x] = 1.0;
    if (delta && threadIdx.x < 128) delta[r * 128 + threadIdx.x] = src[128 + nJd + threadIdx.x];
  }
}

using U8Kernel = void (*)(U8Params);
inline U8Kernel u8_kernel(int W) {
  return W == 1 ? knot_u8_kernel<1> : (W == 2 ? knot_u8_kernel<2> : knot_u8_kernel<4>);
}

}  // namespace pb2
