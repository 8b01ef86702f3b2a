// Single-round residual + Jacobian kernel for the 3-qubit unitary shape (generator 16 x 16 with the
// real-isomorphism structure, 8 state columns, m = 3 or 4 drives) when an SM owns at most SEVEN knots
// (BASELINE config C3: 999 knot evaluations on 148 SMs).  Same mathematics as knot_u8.cuh -- the
// truncated-Taylor action of exp(dt G(u)) on the stacked columns [I | X | jet_1 .. jet_m] by DMMA.8x8x4 on
// register-resident transposed tiles; it replaces DirectTrajOpt's BilinearIntegrator evaluation as built at
// /root/reference/src/control/integrators.jl:35-51 -- organised around what the traces of knot_u8 and of the
// first single-round draft showed (tools/trace_u8p.py):
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
//   * the drive generators of this path have at most one nonzero per row (ELL width 1), so G(u) is built from
//     the dense drift fragments plus ONE value per (drive, output half) and lane instead of dense per-drive
//     fragment tables (the generator build is shared-memory-bandwidth bound: 24 loads per lane instead of 46);
//     when all nonzeros of a drive generator have one magnitude c_j (Pauli-type drives: every 2-level transmon
//     drive) the jets are carried as J_j / c_j, their coupling term G_j S is a sign flip of the exchanged
//     state iterate (an integer instruction: nothing FP64 queues between the exchange and the products) and
//     c_j is applied once at the end (template UNIT);
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
  int o_dval, o_dsel, o_norm, o_tab, o_slot, slot_stride, zpad, o_prep, o_y, o_est, o_mbar;
  double cj[4];           // UNIT: the common magnitude of drive generator j's nonzeros
  const double* tables;   // [G0 fragments 256 | drive values m*64 | drive slots m*32 (int32) | norms | theta | 1/k!]
  const EllEntry* ell;    // (m+1) * 16, drive m = all-zero dummy
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
#define U8P_STEP_STAMP(tr, idx) do { if ((tr) && (threadIdx.x & 31) == 0 && (idx) < 32) (tr)[idx] = clock64(); } while (0)
#else
#define U8P_STAMP(i) do { } while (0)
#define U8P_STEP_STAMP(tr, idx) do { } while (0)
#endif

// warp -> slot / role.  Sub-partition = warp id mod 4:
//   sub-partition 0: A0 A1 A2 (9 tiles)   1: A3 A4 A5 (9)   2: A6 B0 B1 B2 (3 + 6)   3: B3 B4 B5 B6 (8)
// role 0 = warp A (X, J_1, J_2), 1 = warp B (J_3, J_4), 2 = helper (tables, constant identity entries).
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

struct U8pStepTrace {
  long long* base;   // this warp's 32 step stamps (null outside the debug build)
  int n;
};

// One jet tile of one step: additive terms into the accumulator, then the product.  The exchange addresses
// already point at the current parity's buffer.  UNIT: the coupling term is +-y (sign flip on the high word) and
// the tile carries J / c_j.  GEN: general sub-step (B_j != 0: c_k b_J joins the additive term).
// The code is deliberately NOT specialised per step parity / first step: one copy of the step body keeps the
// kernel small enough for the instruction cache to hold the once-per-launch prologue code across launches
// (the fully specialised version was 64 KB and its cold prologue ran ~3x slower than its instruction count).
template <bool UNIT, bool GEN>
__device__ __forceinline__ void u8p_jet_tile(double (&t)[4], const double (&bJ)[4], const double (&A)[4][2],
                                             const double (&ev)[4], const int (&sg)[4], const uint32_t (&yad)[4],
                                             double ck, bool mma) {
  double d[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double v;
    if (UNIT) {
      int lo, hi;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(yad[i]) : "memory");
      v = __hiloint2double(hi ^ sg[i], lo);
      if (GEN) v = fma(ck, bJ[i], v);
    } else {
      const double y0 = lds_f64<0>(yad[i]);
      v = GEN ? fma(ev[i], y0, ck * bJ[i]) : ev[i] * y0;
    }
    d[i >> 1][i & 1] = v;
  }
  if (mma) u8_mma_acc(d, t, A);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = d[i >> 1][i & 1];
}

// warp A, one step: publish X, X product (accumulator pre-loaded with c_k b), barrier, next c_k b, J_1, J_2.
// `mma`: false on the very first step (the jets are still zero: the coupling term alone).
// `dT_out` non-null: the final step of the knot -- d/d dt = -G(u) E x is one more generator product on the NEW
// state tile, issued between the two jet tiles so that it queues with the step's own products, and stored.
template <bool UNIT, bool GEN>
__device__ __forceinline__ void u8p_step_A(double (&tX)[4], double (&tJ)[2][4], const double (&bX)[4], uint32_t a_bx,
                                           const double (&bJ)[2][4], const double (&A)[4][2], uint32_t ypub,
                                           uint32_t ck_addr, uint32_t ck_next_addr, int bar, double (&accX)[4],
                                           const double (&ev)[2][4], const int (&sg)[2][4], const uint32_t (&yad)[2][4],
                                           bool mma, double* dT_out, U8pStepTrace& st, bool depwait = true,
                                           uint32_t a_graw = 0) {
  U8P_STEP_STAMP(st.base, st.n); ++st.n;
#pragma unroll
  for (int i = 0; i < 4; ++i) sts_f64<0>(ypub + i * 256, tX[i]);
  double dX[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) dX[i >> 1][i & 1] = accX[i];
  double ck = 0.0;
  if (GEN) ck = lds_f64<0>(ck_addr);
  const double ckn = lds_f64<0>(ck_next_addr);
  u8_mma_acc(dX, tX, A);
  bar_sync(bar, 64);
  U8P_STEP_STAMP(st.base, st.n); ++st.n;
  // c_k b for the next step.  Common case (one sub-step): b = x_k, re-read from the slab instead of being held
  // in eight registers through the loop (the step loop of this warp is at the 128-register limit)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accX[i] = ckn * (GEN ? bX[i] : lds_f64<0>(a_bx + U8_OFF(i)));
    tX[i] = dX[i >> 1][i & 1];
  }
  u8p_jet_tile<UNIT, GEN>(tJ[0], bJ[0], A, ev[0], sg[0], yad[0], ck, mma);
  if (dT_out) {
    // the knot's final step: every earlier grid of the stream has long completed (the wait returns at once)
    double dT[2][2];
    if (a_graw) {
      // the caller iterates on a scaled generator (knot_u8q: dt G): this one product takes the unscaled
      // fragments, kept in shared memory for it
      double Gr[4][2];
#pragma unroll
      for (int s8 = 0; s8 < 8; ++s8) Gr[s8 >> 1][s8 & 1] = lds_f64<0>(a_graw + 256u * (uint32_t)s8);
      u8_mma(dT, tX, Gr);
    } else {
      u8_mma(dT, tX, A);
    }
    if (depwait) asm volatile("griddepcontrol.wait;" ::: "memory");
    stg_f64x2(dT_out, -dT[0][0], -dT[0][1]);
    stg_f64x2(dT_out + 8, -dT[1][0], -dT[1][1]);
  }
  u8p_jet_tile<UNIT, GEN>(tJ[1], bJ[1], A, ev[1], sg[1], yad[1], ck, mma);
}

// warp B, one step: barrier, then its jet tiles
template <bool UNIT, bool GEN>
__device__ __forceinline__ void u8p_step_B(double (&t)[2][4], const double (&bJ)[2][4], const double (&A)[4][2],
                                           uint32_t ck_addr, int bar, bool two, const double (&ev)[2][4],
                                           const int (&sg)[2][4], const uint32_t (&yad)[2][4], bool mma,
                                           U8pStepTrace& st) {
  U8P_STEP_STAMP(st.base, st.n); ++st.n;
  double ck = 0.0;
  if (GEN) ck = lds_f64<0>(ck_addr);
  bar_sync(bar, 64);
  U8P_STEP_STAMP(st.base, st.n); ++st.n;
  u8p_jet_tile<UNIT, GEN>(t[0], bJ[0], A, ev[0], sg[0], yad[0], ck, mma);
  if (two) u8p_jet_tile<UNIT, GEN>(t[1], bJ[1], A, ev[1], sg[1], yad[1], ck, mma);
}

// the exchange buffers alternate with the step: all of a warp's exchange addresses move by +-1 KB
__device__ __forceinline__ void u8p_flip(uint32_t (&yad)[2][4], int& dy) {
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i) yad[a][i] += (uint32_t)dy;
  dy = -dy;
}

// the sparse drive-generator rows of one jet tile: value (or its sign), and the exchange slot it reads
__device__ __forceinline__ void u8p_ell(const EllEntry* ell, int jd, int g, int q, uint32_t a_y, double (&ev)[4],
                                        int (&sg)[4], uint32_t (&yad)[4]) {
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const int r = 8 * (i4 >> 1) + 2 * q + (i4 & 1);
    const int4 en = __ldg(reinterpret_cast<const int4*>(ell + (size_t)jd * 16 + r));   // {val lo, val hi, idx, pad}
    ev[i4] = __hiloint2double(en.y, en.x);
    sg[i4] = en.y & (int)0x80000000;
    yad[i4] = a_y + 8u * (uint32_t)((2 * (en.z >> 3) + (en.z & 1)) * 32 + g * 4 + ((en.z & 7) >> 1));
  }
}

// ||dt G|| beyond the largest tabulated radius: number of sub-steps (rare; kept out of the instruction stream)
__device__ __noinline__ void u8p_substeps(double nrm, double th_max, int max_sub, double& dt, double& per, int& n_sub) {
  const double ns = ceil(nrm / th_max);
  if (ns <= (double)max_sub) {
    n_sub = (int)ns;
    dt = dt / ns;
    per = nrm / ns;
  } else {
    dt = __longlong_as_double(0x7ff8000000000000LL);   // norm beyond the supported range: NaN out
  }
}

template <bool UNIT>
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
  const uint32_t mb_est = mb_z + 8u;                      // the slot's propagator block is staged (one arrival)
  const int k = slot * gridDim.x + blockIdx.x;            // this slot's knot
  const bool have = role != 2 && k < p.nk;
  const uint32_t zbytes = (uint32_t)p.zlen * 8u;

  U8P_STAMP(0);
  if (role == 0 && lane == 0) {
    // first thing: the slab of this slot's knot (HBM latency is the longest item of the prologue)
    mbar_init(mb_z, 1);
    mbar_init(mb_est, 1);
    fence_proxy_async();   // the init is visible to the async proxy (no cluster-scope fence: that one invalidates L1)
    if (have) {
      if (!p.early_z) asm volatile("griddepcontrol.wait;" ::: "memory");
      U8P_STAMP(6);
      mbar_expect_tx(mb_z, zbytes);
      bulk_g2s(a_z, p.Z + (size_t)k * p.D, zbytes, mb_z);
    }
  }
  if (w == 12 && lane == 0) {
    // the handle's constant tables (never written after pb2_create) arrive by one bulk copy
    mbar_init(mb_tab, 1);
    fence_proxy_async();   // the init is visible to the async proxy (no cluster-scope fence: that one invalidates L1)
    mbar_expect_tx(mb_tab, 8u * (uint32_t)(p.o_tab + 40));
    bulk_g2s(a_cG, p.tables, 8u * (uint32_t)(p.o_tab + 40), mb_tab);
#ifdef PB2_TRACE
    if (p.trace) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + 12) * 8) + 6] = (long long)smid;
    }
#endif
  }
  // programmatic dependent launch: the next grid of the stream may start its own prologue as SMs drain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // the sparse rows of this warp's jet tiles (global loads: issued here, under the slab's latency)
  double ev[2][4];
  int sg[2][4];
  uint32_t yad[2][4];
  if (role == 0) {
    u8p_ell(p.ell, 0, g, q, a_y, ev[0], sg[0], yad[0]);
    u8p_ell(p.ell, 1, g, q, a_y, ev[1], sg[1], yad[1]);
  } else {
    u8p_ell(p.ell, 2, g, q, a_y, ev[0], sg[0], yad[0]);
    u8p_ell(p.ell, m >= 4 ? 3 : m, g, q, a_y, ev[1], sg[1], yad[1]);   // drive m is the all-zero dummy
  }
  __syncthreads();
  U8P_STAMP(1);

  const int rec = p.compact ? p.cstride : p.nnz_jac;      // doubles per knot in `jac`
  const int o_jets = p.compact ? 128 : 2048;              // where the jet columns start inside a knot's segment
  if (role == 2) {
    // ---- helper warps.  They are the only ones that wait for the previous grid of the stream before the
    //      tail: in back-to-back callbacks that grid is still draining on other SMs when this CTA's
    //      propagator tiles are done, and a compute warp blocked in the dependency wait would hold up
    //      its whole slot (measured: 2 600 cycles).
    if (!p.compact) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      if (w == 12) {
        // the constant d/dx_{k+1} identity entries of every knot of this CTA
#pragma unroll 1
        for (int s = 0; s < kU8pSlots; ++s) {
          const int ks = s * gridDim.x + blockIdx.x;
          if (ks < p.nk) {
            double* o = p.jac + (size_t)ks * rec + o_jets + (m + 1) * 128 + 4 * lane;
            stg_f64x2(o, 1.0, 1.0);
            stg_f64x2(o + 2, 1.0, 1.0);
          }
        }
      } else if (lane == 0) {
        // the d/dx_k block of every knot: I (x) E, eight bulk (TMA) stores of the 2 KB block its slot staged,
        // in the order the slots finish phase 1 (sub-partition 3 hosts one E tile, the others two)
        const unsigned order = 0x2416305u;   // slots 5, 0, 3, 6, 1, 4, 2 (one hex digit each, low digit first)
#pragma unroll 1
        for (int i = 0; i < kU8pSlots; ++i) {
          const int s = (int)((order >> (4 * i)) & 15u);
          const int ks = s * gridDim.x + blockIdx.x;
          if (ks >= p.nk) continue;
          const uint32_t a_s = a_cG + 8u * (uint32_t)(p.o_slot + s * p.slot_stride);
          mbar_wait(a_s + 8u * p.o_mbar + 8u, 0);
          fence_proxy_async();
          double* o = p.jac + (size_t)ks * rec;
#pragma unroll
          for (int c = 0; c < 8; ++c) bulk_s2g(o + c * 256, a_s + 8u * p.o_est, 2048u);
          bulk_commit();
        }
        bulk_wait_read0();   // the staged blocks must outlive the bulk stores' reads of them
      }
    }
    return;
  }
  if (!have) return;                                       // both warps of an empty slot leave together

  mbar_wait(mb_tab, 0);
  mbar_wait(mb_z, 0);
  U8P_STAMP(2);
#ifdef PB2_TRACE
  long long* ptr_ = p.trace ? p.trace + (size_t)64 * 148 * 16 * 8 + ((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + w) * 32) : nullptr;
#define U8P_PSTAMP(i) do { if (ptr_ && lane == 0) ptr_[i] = clock64(); } while (0)
#else
#define U8P_PSTAMP(i) do { } while (0)
#endif
  double A[4][2];
  if (role == 1) {
    // ---- warp B: G(u) = G0 + sum_j u_j G_j for the slot, kept in registers and published for warp A.
    //      Lane (g, q) holds G[8 nt + g][pi(4 kt + q)]; a drive generator has at most one nonzero per row, so
    //      per (drive, nt) the lane holds at most one nonzero slot kt: one value and its slot index.
    double acc[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) acc[s] = lds_f64<0>(a_cG + 8u * (uint32_t)(s * 32 + lane));
    double uj[4], dv[4][2];
    int sel[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool on = j < m;
      uj[j] = on ? lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + j)) : 0.0;
      dv[j][0] = on ? lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_dval + (j * 2 + 0) * 32 + lane)) : 0.0;
      dv[j][1] = on ? lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_dval + (j * 2 + 1) * 32 + lane)) : 0.0;
      sel[j] = 0xff;
      if (on) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sel[j]) : "r"(a_cG + 8u * (uint32_t)p.o_dsel + 4u * (uint32_t)(j * 32 + lane)) : "memory");
    }
    U8P_PSTAMP(24);
    {
      double sink = acc[0] + acc[7] + uj[0] + uj[3] + dv[3][1] + (double)sel[3];   // trace build only: loads have landed
      if (sink == 1.2345e300) U8P_PSTAMP(29);
    }
    U8P_PSTAMP(25);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        // one predicated FMA per slot, written in PTX: left to the compiler, `if (code == kt) acc = fma(..)`
        // becomes a divergent jump table (measured: + 2 000 cycles per launch)
        const int code = (sel[j] >> (4 * nt)) & 15;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
          asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %3, %4;\n\t@p fma.rn.f64 %0, %1, %2, %0;\n\t}"
              : "+d"(acc[kt * 2 + nt])
              : "d"(uj[j]), "d"(dv[j][nt]), "r"(code), "r"(kt));
      }
    U8P_PSTAMP(26);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      sts_f64<0>(a_p + 8u * (uint32_t)(s * 32 + lane), acc[s]);
      A[s >> 1][s & 1] = acc[s];
    }
    U8P_PSTAMP(27);
  } else {
    // ---- warp A: Taylor degree and coefficients (lane l holds degree l) --------------------------------
    const double th_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double if_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + 20 + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double th_max = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + kMaxDeg));
    double dt = lds_f64<0>(a_z + 8u * p.dt_off);
    // |u_j| ||G_j||_1 over the lanes: one batch of independent loads, then shuffles (no chain of dependent
    // shared-memory loads queued behind the other warps' generator build)
    double nj = 0.0, uj = 0.0;
    if (lane <= m) {
      nj = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_norm + lane));
      if (lane >= 1) uj = fabs(lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + lane - 1)));
    }
    // same operation order as the other kernels (fma chain over the drives): the same degree M everywhere
    double nrm = __shfl_sync(0xffffffffu, nj, 0);
    for (int j = 1; j <= m; ++j) nrm = fma(__shfl_sync(0xffffffffu, uj, j), __shfl_sync(0xffffffffu, nj, j), nrm);
    nrm *= fabs(dt);
    U8P_PSTAMP(24);
    int n_sub = 1;
    double per = nrm;
    if (nrm > th_max) u8p_substeps(nrm, th_max, p.max_sub, dt, per, n_sub);
    const unsigned below = __ballot_sync(0xffffffffu, lane >= 1 && lane < kMaxDeg && th_l < per);
    const int Mq = 1 + __popc(below);
    U8P_PSTAMP(25);
    double pw = 1.0, sq = dt;
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
      if ((lane >> bit) & 1) pw *= sq;
      sq *= sq;
    }
    U8P_PSTAMP(26);
    if (lane <= kMaxDeg) sts_f64<0>(a_c + 8u * lane, lane <= Mq ? if_l * pw : 0.0);
    if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_p + 8u * 276u), "r"(Mq), "r"(n_sub) : "memory");
  }
#ifdef PB2_TRACE
  if (p.trace && lane == 0) p.trace[(size_t)64 * 148 * 16 * 8 + ((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + w) * 32) + 30] = clock64();
#endif
  bar_sync(bar, 64);                                       // hand-over in both directions
#ifdef PB2_TRACE
  if (p.trace && lane == 0) p.trace[(size_t)64 * 148 * 16 * 8 + ((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + w) * 32) + 31] = clock64();
#endif
  if (role == 0) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) A[kt][nt] = lds_f64<0>(a_p + 8u * (uint32_t)((kt * 2 + nt) * 32 + lane));
  }
  int M, n_sub;
  lds_v2u32(a_p + 8u * 276u, M, n_sub);
  U8P_STAMP(3);

  U8pStepTrace stt{nullptr, 0};
#ifdef PB2_TRACE
  if (p.trace) stt.base = p.trace + (size_t)64 * 148 * 16 * 8 + ((((size_t)p.trace_id * gridDim.x + blockIdx.x) * 16 + w) * 32);
#endif
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
#pragma unroll 1
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
#pragma unroll 1
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
#pragma unroll 1
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
    if (p.compact) {
      // compact records (host-pointer path, sharded runs: no back-to-back grids to overlap with): the half
      // block leaves right away from the registers
      asm volatile("griddepcontrol.wait;" ::: "memory");
      stg_f64x2(jk + lc, -tE[0], -tE[1]);
      stg_f64x2(jk + lc + 8, -tE[2], -tE[3]);
    } else {
      // -E = -[[P, -Q], [Q, P]] staged once (own column g, mirrored column g + 8); helper warp 13 turns it
      // into the eight copies of the d/dx_k block by bulk stores that drain underneath phase 2
      sts_f64x2<0>(a_est + lane_col, make_double2(-tE[0], -tE[1]));
      sts_f64x2<64>(a_est + lane_col, make_double2(-tE[2], -tE[3]));
      sts_f64x2<1024 + 64>(a_est + lane_col, make_double2(-tE[0], -tE[1]));
      sts_f64x2<1024>(a_est + lane_col, make_double2(tE[2], tE[3]));
      __syncwarp();
      if (lane == 0) mbar_arrive(mb_est);   // release; helper warp 13 acquires, fences towards the async proxy, stores
    }
  }

  if (role == 0) {
    // ================================ warp A: X, J_1, J_2 ==============================================
    const uint32_t ypub = a_y + 8u * (uint32_t)(g * 4 + q);
    const uint32_t xl = 8u * (uint32_t)p.x_off + lane_col;
    double bX[4], tX[4], tJ[2][4], accX[4];
    double* const dT_out = jj + m * 128 + lc;
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
    uint32_t yp = ypub;
    int dy = 1024;
    if (n_sub == 1) {
      // the common case: one sub-step, state columns as B, jets start from zero
#pragma unroll 1
      for (int kq = M - 1; kq >= 0; --kq) {
        u8p_step_A<UNIT, false>(tX, tJ, bX, a_z + xl, tJ, A, yp, a_c, a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar, accX, ev, sg, yad,
                                kq != M - 1, kq == 0 ? dT_out : nullptr, stt);
        yp += (uint32_t)dy;
        u8p_flip(yad, dy);
      }
    } else {
      // ||dt G|| beyond the largest tabulated radius (rare): n_sub sub-steps, general B from the second one on
      double bX2[4], bJ2[2][4];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        bX2[i4] = bX[i4];
        bJ2[0][i4] = 0.0;
        bJ2[1][i4] = 0.0;
      }
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
        if (sub > 0) {
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
        }
#pragma unroll 1
        for (int kq = M - 1; kq >= 0; --kq) {
          u8p_step_A<UNIT, true>(tX, tJ, bX2, a_z + xl, bJ2, A, yp, a_c + 8u * (uint32_t)kq, a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar,
                                 accX, ev, sg, yad, sub > 0 || kq != M - 1, (sub == n_sub - 1 && kq == 0) ? dT_out : nullptr, stt);
          yp += (uint32_t)dy;
          u8p_flip(yad, dy);
        }
      }
    }
    // ---- jets, d/d dt = -G(u) E x, delta: straight from the registers -----------------------------------
    U8P_STAMP(5);
    const bool want_delta = p.delta != nullptr || p.compact;
    double xn[4] = {0.0, 0.0, 0.0, 0.0};
    if (want_delta) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) xn[i4] = lds_f64<0>(a_z + 8u * p.D + xl + U8_OFF(i4));
    }
    const double s0 = UNIT ? -p.cj[0] : -1.0, s1 = UNIT ? -p.cj[1] : -1.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    stg_f64x2(jj + lc, s0 * tJ[0][0], s0 * tJ[0][1]);
    stg_f64x2(jj + lc + 8, s0 * tJ[0][2], s0 * tJ[0][3]);
    stg_f64x2(jj + 128 + lc, s1 * tJ[1][0], s1 * tJ[1][1]);
    stg_f64x2(jj + 128 + lc + 8, s1 * tJ[1][2], s1 * tJ[1][3]);
    if (want_delta) {
      double* dd = (p.compact ? jj + (m + 1) * 128 : p.delta + (size_t)k * 128) + lc;
      stg_f64x2(dd, xn[0] - tX[0], xn[1] - tX[1]);
      stg_f64x2(dd + 8, xn[2] - tX[2], xn[3] - tX[3]);
    }
  } else {
    // ================================== warp B: J_3 [, J_4] =============================================
    const bool two = m >= 4;
    double t[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) t[a][i4] = 0.0;
    int dy = 1024;
    if (n_sub == 1) {
#pragma unroll 1
      for (int kq = M - 1; kq >= 0; --kq) {
        u8p_step_B<UNIT, false>(t, t, A, a_c, bar, two, ev, sg, yad, kq != M - 1, stt);
        u8p_flip(yad, dy);
      }
    } else {
      double bJ[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) bJ[a][i4] = 0.0;
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
        if (sub > 0) {
          const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              bJ[a][i4] = t[a][i4];
              t[a][i4] *= cM;
            }
        }
#pragma unroll 1
        for (int kq = M - 1; kq >= 0; --kq) {
          u8p_step_B<UNIT, true>(t, bJ, A, a_c + 8u * (uint32_t)kq, bar, two, ev, sg, yad, sub > 0 || kq != M - 1, stt);
          u8p_flip(yad, dy);
        }
      }
    }
    U8P_STAMP(5);
    const double s2 = UNIT ? -p.cj[2] : -1.0, s3 = UNIT ? -p.cj[3] : -1.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    stg_f64x2(jj + 2 * 128 + lc, s2 * t[0][0], s2 * t[0][1]);
    stg_f64x2(jj + 2 * 128 + lc + 8, s2 * t[0][2], s2 * t[0][3]);
    if (two) {
      stg_f64x2(jj + 3 * 128 + lc, s3 * t[1][0], s3 * t[1][1]);
      stg_f64x2(jj + 3 * 128 + lc + 8, s3 * t[1][2], s3 * t[1][3]);
    }
  }
  U8P_STAMP(7);
}

// Shared-memory layout (doubles): [G0 fragments 256 | drive values m*64 | drive slots (m*32 int32) | norms |
// theta, 1/k! | tables' mbarrier], then per slot
//   slab | prepared knot | X exchange x2 | staged propagator block (16 x 16) | mbarrier.
inline size_t u8p_layout(U8pParams& q) {
  auto even = [](int v) { return (v + 1) & ~1; };
  q.o_dval = 256;
  q.o_dsel = q.o_dval + q.m * 64;
  q.o_norm = q.o_dsel + q.m * 16;
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

// Host side: the table blob of this kernel, from the dense fragment tables of the DmmaPlan.  Returns false when
// a drive generator has a row with two nonzeros in one lane's slots (cannot happen for ELL width 1).
// unit / cj: all nonzeros of drive generator j have the magnitude cj[j] and no row is empty.
inline bool u8p_tables(const DmmaPlan& pl, int m, std::vector<double>& blob, bool& unit, double (&cj)[4]) {
  if (pl.W != 1 || pl.NT != 2 || m < 1 || m > 4) return false;
  const int o_dval = 256, o_dsel = o_dval + m * 64, o_norm = o_dsel + m * 16;
  const int o_tab = o_norm + ((m + 2) & ~1);
  blob.assign((size_t)o_tab + 40, 0.0);
  for (int e = 0; e < 256; ++e) blob[e] = pl.gfrag[e];
  int32_t* sel = reinterpret_cast<int32_t*>(blob.data() + o_dsel);
  for (int j = 0; j < m; ++j)
    for (int lane = 0; lane < 32; ++lane) {
      int code[2] = {15, 15};
      for (int nt = 0; nt < 2; ++nt)
        for (int kt = 0; kt < 4; ++kt) {
          const double v = pl.gfrag[((size_t)(1 + j) * 8 + kt * 2 + nt) * 32 + lane];
          if (v != 0.0) {
            if (code[nt] != 15) return false;
            code[nt] = kt;
            blob[o_dval + (j * 2 + nt) * 32 + lane] = v;
          }
        }
      sel[j * 32 + lane] = code[0] | (code[1] << 4);
    }
  for (int j = 0; j <= m; ++j) blob[o_norm + j] = pl.norms[j];
  unit = true;
  for (int j = 0; j < 4; ++j) cj[j] = 1.0;
  for (int j = 0; j < m && unit; ++j) {
    double c = 0.0;
    for (int r = 0; r < 16; ++r) {
      const double v = std::fabs(pl.ell[(size_t)j * 16 + r].val);
      if (v == 0.0 || (c != 0.0 && v != c)) { unit = false; break; }
      c = v;
    }
    cj[j] = c;
  }
  return true;
}

inline auto u8p_kernel(bool unit) -> void (*)(const U8pParams) {
  return unit ? knot_u8p_kernel<true> : knot_u8p_kernel<false>;
}

}  // namespace pb2
