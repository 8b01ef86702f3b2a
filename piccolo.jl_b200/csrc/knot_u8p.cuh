// Single-round residual + Jacobian kernel for the 3-qubit unitary shape (generator 16 x 16 with the
// real-isomorphism structure, 8 state columns, m = 3 or 4 drives) when an SM owns at most SEVEN knots
// (BASELINE config C3: 999 knot evaluations on 148 SMs).  Same mathematics as knot_u8.cuh -- the
// truncated-Taylor action of exp(dt G(u)) on the stacked columns [I | X | jet_1 .. jet_m] by DMMA.8x8x4 on
// register-resident transposed tiles; it replaces DirectTrajOpt's BilinearIntegrator evaluation as built at
// /root/reference/src/control/integrators.jl:35-51 -- organised around what the traces of knot_u8 and of the
// first single-round draft showed (tools/trace_u8s.py):
//
//   * with 6.75 knots per SM a persistent kernel that keeps four knots in flight runs two rounds, the second
//     one latency-bound; here EVERY knot of the SM is in flight at once, one slot per knot;
//   * all seven knots then finish together and their 7 x 22.5 KB of results hit the LSU / L2 in one burst
//     (measured: up to 4 800 cycles for one warp's 38 16-byte stores).  The propagator tile E does not depend
//     on the state or the jets, and the d/dx_k block -- n_b = 8 replicated, mirrored copies of it -- is 70 % of
//     a knot's bytes: so E runs FIRST (phase 1: seven tiles spread (2, 2, 2, 1) over the sub-partitions), is
//     staged once in shared memory and leaves by eight bulk (TMA) stores per knot that drain underneath
//     phase 2 (state + jet tiles); the tail only carries jets, d/d dt and delta (6 KB per knot);
//   * phase 2 is bound by the busiest sub-partition's tile count: a slot is warp A (tiles X, J_1, J_2) and
//     warp B (J_3 [, J_4]); the warp -> (slot, role) table below puts (9, 9, 9, 8) tiles on the four
//     sub-partitions (sub-partition = warp id mod 4) instead of the (12, 12, 9, 9) of whole-knot slots;
//   * nothing before the first global WRITE depends on the previous grid of the stream, so with
//     `early_z` (the caller's promise that the trajectory buffer is not produced by the kernel enqueued just
//     before -- true whenever Z arrives by a copy, pb2_set_option) the slab loads, the generator build and
//     the whole of phase 1 run before `griddepcontrol.wait`: back-to-back callbacks overlap one launch's
//     drain with the next one's prologue on every SM that is already free.
#pragma once
#include "knot_u8.cuh"

namespace pb2 {

struct U8pParams {
  int m, D, x_off, dt_off, u_off, nnz_jac, max_sub, nk, zlen;
  int early_z;            // 1: Z may be read before the programmatic dependency wait (see above)
  int compact;            // 1: records [E columns 0..7 | jets, d/d dt | delta] of cstride doubles go to `jac`
  int cstride;
  // shared-memory layout in doubles (u8p_layout)
  int o_norm, o_tab, o_slot, slot_stride, zpad, o_prep, o_y, o_est, o_mbar;
  const double* tables;   // as U8Params::tables
  const EllEntry* ell;    // (m+1) * 16 * W, drive m = all-zero dummy
  const double* Z;
  double* delta;          // may be null
  double* jac;
  long long* trace;       // debug build: [launch][block][warp 16][8] clock64 stamps (entry 6 of helper warp 12: %smid)
  int trace_id;
};

constexpr int kU8pSlots = 7;
constexpr int kU8pThreads = 512;

#ifdef PB2_TRACE
#define U8P_STAMP(i) do { if (p.trace && lane == 0) p.trace[((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + w) * 8) + (i)] = clock64(); } while (0)
#else
#define U8P_STAMP(i) do { } while (0)
#endif

// warp -> slot / role.  Sub-partition = warp id mod 4:
//   sub-partition 0: A0 A1 A2 (9 tiles)   1: A3 A4 A5 (9)   2: A6 B0 B1 B2 (3 + 6)   3: B3 B4 B5 B6 (8)
// role 0 = warp A (X, J_1, J_2), 1 = warp B (J_3, J_4), 2 = helper (constant identity entries).
// Phase 1 (E tiles): A0 A1 | A3 A4 | A6 B2 | B5  ->  (2, 2, 2, 1) per sub-partition.
__device__ __forceinline__ void u8p_role(int w, int& slot, int& role, bool& doE) {
  // one byte per warp: slot | role << 4 | doE << 6, packed into two 64-bit immediates (no local array)
  //   w0 A0* w1 A3* w2 A6* w3 B3 | w4 A1* w5 A4* w6 B0 w7 B4 | w8 A2 w9 A5 w10 B1 w11 B5* | w12 H w13 H w14 B2* w15 B6
  const unsigned long long lo = 0x1410444113464340ull, hi = 0x1652212055110502ull;
  const unsigned v = (unsigned)(((w & 8) ? hi : lo) >> (8 * (w & 7))) & 0xffu;
  slot = (int)(v & 15u);
  role = (int)((v >> 4) & 3u);
  doE = (v & 0x40u) != 0;
}

// one jet tile of one step: additive terms into the accumulator, then the product
template <int W, int PAR, bool FIRST, bool NOMMA>
__device__ __forceinline__ void u8p_jet_tile(double (&t)[4], const double (&bJ)[4], const double (&A)[4][2],
                                             const double (&ev)[4][W], const uint32_t (&yad)[4][W], double ck) {
  double d[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double y0 = lds_f64<PAR * 1024>(yad[i][0]);
    double v = FIRST ? ev[i][0] * y0 : fma(ev[i][0], y0, ck * bJ[i]);
#pragma unroll
    for (int ww = 1; ww < W; ++ww) v = fma(ev[i][ww], lds_f64<PAR * 1024>(yad[i][ww]), v);
    d[i >> 1][i & 1] = v;
  }
  if (!NOMMA) u8_mma_acc(d, t, A);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = d[i >> 1][i & 1];
}

// warp A, one step: publish X, X product (accumulator pre-loaded with c_k b), barrier, next c_k b, J_1, J_2
template <int W, int PAR, bool FIRST, bool NOMMA>
__device__ __forceinline__ void u8p_step_A(double (&tX)[4], double (&tJ)[2][4], const double (&bX)[4],
                                           const double (&bJ)[2][4], const double (&A)[4][2], uint32_t ypub,
                                           uint32_t ck_addr, uint32_t ck_next_addr, int bar, double (&accX)[4],
                                           const double (&ev)[2][4][W], const uint32_t (&yad)[2][4][W]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) sts_f64<PAR * 1024>(ypub + i * 256, tX[i]);
  double dX[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) dX[i >> 1][i & 1] = accX[i];
  double ck = 0.0;
  if (!FIRST) ck = lds_f64<0>(ck_addr);
  const double ckn = lds_f64<0>(ck_next_addr);
  u8_mma_acc(dX, tX, A);
  bar_sync(bar, 64);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accX[i] = ckn * bX[i];
    tX[i] = dX[i >> 1][i & 1];
  }
  u8p_jet_tile<W, PAR, FIRST, NOMMA>(tJ[0], bJ[0], A, ev[0], yad[0], ck);
  u8p_jet_tile<W, PAR, FIRST, NOMMA>(tJ[1], bJ[1], A, ev[1], yad[1], ck);
}

// warp B, one step: barrier, then its jet tiles
template <int W, int PAR, bool FIRST, bool NOMMA>
__device__ __forceinline__ void u8p_step_B(double (&t)[2][4], const double (&bJ)[2][4], const double (&A)[4][2],
                                           uint32_t ck_addr, int bar, bool two, const double (&ev)[2][4][W],
                                           const uint32_t (&yad)[2][4][W]) {
  double ck = 0.0;
  if (!FIRST) ck = lds_f64<0>(ck_addr);
  bar_sync(bar, 64);
  u8p_jet_tile<W, PAR, FIRST, NOMMA>(t[0], bJ[0], A, ev[0], yad[0], ck);
  if (two) u8p_jet_tile<W, PAR, FIRST, NOMMA>(t[1], bJ[1], A, ev[1], yad[1], ck);
}

// the sparse drive-generator rows of one jet tile: value and the exchange slot it reads
template <int W>
__device__ __forceinline__ void u8p_ell(const EllEntry* ell, int jd, int g, int q, uint32_t a_y, double (&ev)[4][W],
                                        uint32_t (&yad)[4][W]) {
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const int r = 8 * (i4 >> 1) + 2 * q + (i4 & 1);
#pragma unroll
    for (int ww = 0; ww < W; ++ww) {
      const EllEntry en = ell[((size_t)jd * 16 + r) * W + ww];
      ev[i4][ww] = en.val;
      yad[i4][ww] = a_y + 8u * (uint32_t)((2 * (en.idx >> 3) + (en.idx & 1)) * 32 + g * 4 + ((en.idx & 7) >> 1));
    }
  }
}

template <int W>
__global__ void __launch_bounds__(kU8pThreads, 1) knot_u8p_kernel(const __grid_constant__ U8pParams p) {
  extern __shared__ __align__(16) double u8p_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int slot, role;
  bool doE;
  u8p_role(w, slot, role, doE);
  const int g = lane >> 2, q = lane & 3, m = p.m, bar = 1 + slot;

  const uint32_t a_cG = smem_u32(u8p_smem);
  const uint32_t a_slot = a_cG + 8u * (uint32_t)(p.o_slot + slot * p.slot_stride);
  const uint32_t a_z = a_slot, a_p = a_slot + 8u * p.o_prep, a_c = a_p + 8u * 256u, a_y = a_slot + 8u * p.o_y;
  const uint32_t a_est = a_slot + 8u * p.o_est;
  const uint32_t mb_z = a_slot + 8u * p.o_mbar, mb_tab = a_cG + 8u * (uint32_t)(p.o_tab + 40);
  const int k = slot * gridDim.x + blockIdx.x;            // this slot's knot
  const bool have = role != 2 && k < p.nk;
  const uint32_t zbytes = (uint32_t)p.zlen * 8u;

  U8P_STAMP(0);
#ifdef PB2_TRACE
  if (p.trace && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + 12) * 8) + 6] = (long long)smid;
  }
#endif
  // programmatic dependent launch: the next grid of the stream may start its own prologue as SMs drain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    // the handle's constant tables (never written after pb2_create) arrive by one bulk copy
    mbar_init(mb_tab, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(mb_tab, 8u * (uint32_t)(p.o_tab + 40));
    bulk_g2s(a_cG, p.tables, 8u * (uint32_t)(p.o_tab + 40), mb_tab);
  }
  if (role == 0 && lane == 0) {
    mbar_init(mb_z, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (have) {
      if (!p.early_z) asm volatile("griddepcontrol.wait;" ::: "memory");
      U8P_STAMP(6);
      mbar_expect_tx(mb_z, zbytes);
      bulk_g2s(a_z, p.Z + (size_t)k * p.D, zbytes, mb_z);
    }
  }
  __syncthreads();
  U8P_STAMP(1);

  const int rec = p.compact ? p.cstride : p.nnz_jac;      // doubles per knot in `jac`
  const int o_jets = p.compact ? 128 : 2048;              // where the jet columns start inside a knot's segment
  if (role == 2) {
    // ---- helper warps: the constant d/dx_{k+1} identity entries of every knot of this CTA ---------------
    if (!p.compact) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      const int hl = (w & 1) * 32 + lane;                 // 64 lanes, 16 bytes each = one knot's 128 entries
      for (int s = 0; s < kU8pSlots; ++s) {
        const int ks = s * gridDim.x + blockIdx.x;
        if (ks < p.nk) stg_f64x2(p.jac + (size_t)ks * rec + o_jets + (m + 1) * 128 + 2 * hl, 1.0, 1.0);
      }
    }
    return;
  }
  if (!have) return;                                       // both warps of an empty slot leave together

  mbar_wait(mb_tab, 0);
  mbar_wait(mb_z, 0);
  U8P_STAMP(2);
  double A[4][2];
  if (role == 1) {
    // ---- warp B: G(u) for the slot, kept in registers and published for warp A -------------------------
    double acc[8];
    u8_build_G<false>(a_z, a_cG, lane, m, p.u_off, acc);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      sts_f64<0>(a_p + 8u * (uint32_t)(s * 32 + lane), acc[s]);
      A[s >> 1][s & 1] = acc[s];
    }
  } else {
    // ---- warp A: Taylor degree and coefficients (lane l holds degree l) --------------------------------
    const double th_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double if_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + 20 + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double th_max = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + kMaxDeg));
    double dt = lds_f64<0>(a_z + 8u * p.dt_off);
    // |u_j| ||G_j||_1 summed over the lanes (one batch of independent loads, then shuffles: no chain of
    // dependent shared-memory loads behind the other warps' generator build)
    double nj = 0.0, uj = 0.0;
    if (lane <= m) {
      nj = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_norm + lane));
      if (lane >= 1) uj = fabs(lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + lane - 1)));
    }
    // same operation order as the other kernels (fma chain over the drives): the same degree M everywhere
    double nrm = __shfl_sync(0xffffffffu, nj, 0);
    for (int j = 1; j <= m; ++j) nrm = fma(__shfl_sync(0xffffffffu, uj, j), __shfl_sync(0xffffffffu, nj, j), nrm);
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
        dt = __longlong_as_double(0x7ff8000000000000LL);   // norm beyond the supported range: NaN out
      }
    }
    const unsigned below = __ballot_sync(0xffffffffu, lane >= 1 && lane < kMaxDeg && th_l < per);
    const int Mq = 1 + __popc(below);
    double pw = 1.0, sq = dt;
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
      if ((lane >> bit) & 1) pw *= sq;
      sq *= sq;
    }
    if (lane <= kMaxDeg) sts_f64<0>(a_c + 8u * lane, lane <= Mq ? if_l * pw : 0.0);
    if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_p + 8u * 276u), "r"(Mq), "r"(n_sub) : "memory");
  }
  bar_sync(bar, 64);                                       // hand-over in both directions
  if (role == 0) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) A[kt][nt] = lds_f64<0>(a_p + 8u * (uint32_t)((kt * 2 + nt) * 32 + lane));
  }
  int M, n_sub;
  lds_v2u32(a_p + 8u * 276u, M, n_sub);
  U8P_STAMP(3);

  const uint32_t lane_col = 8u * (uint32_t)(g * 16 + 2 * q);
  const int lc = g * 16 + 2 * q;
  double* jk = p.jac + (size_t)k * (size_t)rec;
  double* jj = jk + o_jets;

  // ======================== phase 1: the propagator tile E (columns 0..7) =================================
  if (doE) {
    const int iE = (g == 2 * q) ? 0 : ((g == 2 * q + 1) ? 1 : -1);   // which element is the unit entry
    double tE[4], acc[4];
    {
      const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
      const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        tE[i4] = (i4 == iE) ? cM : 0.0;
        acc[i4] = (i4 == iE) ? c0 : 0.0;
      }
    }
    for (int kq = M - 1; kq >= 0; --kq) {
      double d[2][2];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) d[i4 >> 1][i4 & 1] = acc[i4];
      u8_mma_acc(d, tE, A);
      const double ckn = lds_f64<0>(a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0));
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        acc[i4] = (i4 == iE) ? ckn : 0.0;
        tE[i4] = d[i4 >> 1][i4 & 1];
      }
    }
    for (int sub = 1; sub < n_sub; ++sub) {                 // further sub-steps: general B
      double bE[4];
      const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
      const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        bE[i4] = tE[i4];
        tE[i4] *= cM;
        acc[i4] = c0 * bE[i4];
      }
      for (int kq = M - 1; kq >= 0; --kq) {
        double d[2][2];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) d[i4 >> 1][i4 & 1] = acc[i4];
        u8_mma_acc(d, tE, A);
        const double ckn = lds_f64<0>(a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0));
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          acc[i4] = ckn * bE[i4];
          tE[i4] = d[i4 >> 1][i4 & 1];
        }
      }
    }
    U8P_STAMP(4);
    // the first global write of this warp: from here on the previous grid of the stream must be complete
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.compact) {
      stg_f64x2(jk + lc, -tE[0], -tE[1]);
      stg_f64x2(jk + lc + 8, -tE[2], -tE[3]);
    } else {
      // -E = -[[P, -Q], [Q, P]] staged once (own column g, mirrored column g + 8); the d/dx_k block is
      // I (x) E: eight bulk stores of the same 2 KB, draining underneath phase 2
      sts_f64x2<0>(a_est + lane_col, make_double2(-tE[0], -tE[1]));
      sts_f64x2<64>(a_est + lane_col, make_double2(-tE[2], -tE[3]));
      sts_f64x2<1024 + 64>(a_est + lane_col, make_double2(-tE[0], -tE[1]));
      sts_f64x2<1024>(a_est + lane_col, make_double2(tE[2], tE[3]));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) bulk_s2g(jk + c * 256, a_est, 2048u);
        bulk_commit();
      }
    }
  }

  if (role == 0) {
    // ================================ warp A: X, J_1, J_2 ==============================================
    const uint32_t ypub = a_y + 8u * (uint32_t)(g * 4 + q);
    const uint32_t xl = 8u * (uint32_t)p.x_off + lane_col;
    double ev[2][4][W];
    uint32_t yad[2][4][W];
    u8p_ell<W>(p.ell, 0, g, q, a_y, ev[0], yad[0]);
    u8p_ell<W>(p.ell, 1, g, q, a_y, ev[1], yad[1]);
    double bX[4], tX[4], tJ[2][4], accX[4];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      bX[i4] = lds_f64<0>(a_z + xl + U8_OFF(i4));
      tJ[0][i4] = 0.0;
      tJ[1][i4] = 0.0;
    }
    {
      const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
      const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        tX[i4] = cM * bX[i4];
        accX[i4] = c0 * bX[i4];
      }
    }
    {
      // first sub-step: state columns as B, jets start from zero (first step: coupling term alone)
      int kq = M - 1;
      u8p_step_A<W, 0, true, true>(tX, tJ, bX, tJ, A, ypub, a_c, a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar, accX, ev, yad);
      --kq;
      for (; kq >= 1; kq -= 2) {
        u8p_step_A<W, 1, true, false>(tX, tJ, bX, tJ, A, ypub, a_c, a_c + 8u * (uint32_t)(kq - 1), bar, accX, ev, yad);
        u8p_step_A<W, 0, true, false>(tX, tJ, bX, tJ, A, ypub, a_c, a_c + 8u * (uint32_t)(kq >= 2 ? kq - 2 : 0), bar, accX, ev, yad);
      }
      if (kq == 0) u8p_step_A<W, 1, true, false>(tX, tJ, bX, tJ, A, ypub, a_c, a_c, bar, accX, ev, yad);
    }
    for (int sub = 1; sub < n_sub; ++sub) {
      double bX2[4], bJ2[2][4];
      const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
      const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        bX2[i4] = tX[i4];
        bJ2[0][i4] = tJ[0][i4];
        bJ2[1][i4] = tJ[1][i4];
        tX[i4] *= cM;
        tJ[0][i4] *= cM;
        tJ[1][i4] *= cM;
        accX[i4] = c0 * bX2[i4];
      }
      bar_sync(bar, 64);    // every reader of the exchange buffers is done before the parities restart
      int kq = M - 1;
      for (; kq >= 1; kq -= 2) {
        u8p_step_A<W, 0, false, false>(tX, tJ, bX2, bJ2, A, ypub, a_c + 8u * (uint32_t)kq, a_c + 8u * (uint32_t)(kq - 1), bar, accX, ev, yad);
        u8p_step_A<W, 1, false, false>(tX, tJ, bX2, bJ2, A, ypub, a_c + 8u * (uint32_t)(kq - 1),
                                       a_c + 8u * (uint32_t)(kq >= 2 ? kq - 2 : 0), bar, accX, ev, yad);
      }
      if (kq == 0) u8p_step_A<W, 0, false, false>(tX, tJ, bX2, bJ2, A, ypub, a_c, a_c, bar, accX, ev, yad);
    }
    // ---- d/d dt = -G(u) E x, delta, jets: straight from the registers -----------------------------------
    U8P_STAMP(5);
    double dT[2][2];
    u8_mma(dT, tX, A);
    const bool want_delta = p.delta != nullptr || p.compact;
    double xn[4] = {0.0, 0.0, 0.0, 0.0};
    if (want_delta) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) xn[i4] = lds_f64<0>(a_z + 8u * p.D + xl + U8_OFF(i4));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    stg_f64x2(jj + lc, -tJ[0][0], -tJ[0][1]);
    stg_f64x2(jj + lc + 8, -tJ[0][2], -tJ[0][3]);
    stg_f64x2(jj + 128 + lc, -tJ[1][0], -tJ[1][1]);
    stg_f64x2(jj + 128 + lc + 8, -tJ[1][2], -tJ[1][3]);
    stg_f64x2(jj + m * 128 + lc, -dT[0][0], -dT[0][1]);
    stg_f64x2(jj + m * 128 + lc + 8, -dT[1][0], -dT[1][1]);
    if (want_delta) {
      double* dd = (p.compact ? jj + (m + 1) * 128 : p.delta + (size_t)k * 128) + lc;
      stg_f64x2(dd, xn[0] - tX[0], xn[1] - tX[1]);
      stg_f64x2(dd + 8, xn[2] - tX[2], xn[3] - tX[3]);
    }
  } else {
    // ================================== warp B: J_3 [, J_4] =============================================
    const bool two = m >= 4;
    double ev[2][4][W];
    uint32_t yad[2][4][W];
    u8p_ell<W>(p.ell, 2, g, q, a_y, ev[0], yad[0]);
    u8p_ell<W>(p.ell, two ? 3 : m, g, q, a_y, ev[1], yad[1]);   // drive m is the all-zero dummy
    double t[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) t[a][i4] = 0.0;
    {
      int kq = M - 1;
      u8p_step_B<W, 0, true, true>(t, t, A, a_c, bar, two, ev, yad);
      --kq;
      for (; kq >= 1; kq -= 2) {
        u8p_step_B<W, 1, true, false>(t, t, A, a_c, bar, two, ev, yad);
        u8p_step_B<W, 0, true, false>(t, t, A, a_c, bar, two, ev, yad);
      }
      if (kq == 0) u8p_step_B<W, 1, true, false>(t, t, A, a_c, bar, two, ev, yad);
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
      bar_sync(bar, 64);
      int kq = M - 1;
      for (; kq >= 1; kq -= 2) {
        u8p_step_B<W, 0, false, false>(t, bJ, A, a_c + 8u * (uint32_t)kq, bar, two, ev, yad);
        u8p_step_B<W, 1, false, false>(t, bJ, A, a_c + 8u * (uint32_t)(kq - 1), bar, two, ev, yad);
      }
      if (kq == 0) u8p_step_B<W, 0, false, false>(t, bJ, A, a_c, bar, two, ev, yad);
    }
    U8P_STAMP(5);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    stg_f64x2(jj + 2 * 128 + lc, -t[0][0], -t[0][1]);
    stg_f64x2(jj + 2 * 128 + lc + 8, -t[0][2], -t[0][3]);
    if (two) {
      stg_f64x2(jj + 3 * 128 + lc, -t[1][0], -t[1][1]);
      stg_f64x2(jj + 3 * 128 + lc + 8, -t[1][2], -t[1][3]);
    }
  }
  // the staged propagator block must outlive the bulk stores' reads of it; their writes complete with the grid
  if (doE && !p.compact && lane == 0) bulk_wait_read0();
  U8P_STAMP(7);
}

// Shared-memory layout (doubles): tables as in u8_layout, then per slot
//   slab | prepared knot | X exchange x2 | staged propagator block (16 x 16) | mbarrier.
inline size_t u8p_layout(U8pParams& q) {
  auto even = [](int v) { return (v + 1) & ~1; };
  q.o_norm = (q.m + 1) * 256;
  q.o_tab = q.o_norm + even(q.m + 1);
  q.o_slot = q.o_tab + 40 + 2;
  q.zpad = even(q.zlen);
  q.o_prep = q.zpad;
  q.o_y = q.o_prep + kU8Prep;
  q.o_est = q.o_y + 2 * 128;
  q.o_mbar = q.o_est + 256;
  q.slot_stride = q.o_mbar + 2;
  return sizeof(double) * ((size_t)q.o_slot + (size_t)kU8pSlots * q.slot_stride);
}

inline auto u8p_kernel(int W) -> void (*)(const U8pParams) {
  return W == 1 ? knot_u8p_kernel<1> : (W == 2 ? knot_u8p_kernel<2> : knot_u8p_kernel<4>);
}

}  // namespace pb2
