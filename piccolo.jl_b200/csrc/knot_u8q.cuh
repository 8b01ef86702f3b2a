// Residual + Jacobian kernel for the 3-qubit unitary shape (generator 16 x 16 with the real-isomorphism
// structure, 8 state columns, m = 3 or 4 drives with one nonzero per generator row): TWO small CTAs per SM.
//
// Same mathematics as knot_u8.cuh / knot_u8p.cuh (truncated-Taylor action of exp(dt G(u)) on the stacked columns
// [I | X | jet_1 .. jet_m] by DMMA.8x8x4 on register-resident transposed tiles; it replaces DirectTrajOpt's
// BilinearIntegrator evaluation as built at /root/reference/src/control/integrators.jl:35-51).  What the traces of
// the one-CTA-per-SM kernels showed (tools/trace_u8p.py): with every knot of an SM in flight
// the Horner phase runs at ~86 % of the FP64 tensor pipe, but a quarter of a launch is spent where the pipe idles --
// the launch gap between two grids on an SM (~750 cycles), the HBM latency of the knot slabs (~1 800), the
// generator build (~1 300) and the tail -- and a 512-thread CTA that owns the whole register file cannot overlap
// any of that with another CTA's products.
//
// Here a CTA is 8 warps (256 threads x 128 registers = half the register file, ~40 KB of shared memory) and owns
// at most FOUR knots; two CTAs share an SM.  They drift out of phase (a grid mixes CTAs of four and of three
// knots, and the hardware hands a freed half-SM to the next grid of the stream at once -- programmatic dependent
// launch), so one CTA's prologue and tail run underneath the other one's products:
//   * slot s of a CTA = warp s (tiles X, J_1, J_2) + warp s + 4 (tiles E, J_3 [, J_4]): three tiles per warp,
//     both warps of a slot on sub-partition s, every sub-partition carries 6 tiles per CTA -- perfectly balanced;
//   * the propagator tile E rides along in warp B (it needs no exchange), so there is no separate phase;
//   * a CTA's prologue now runs while other CTAs keep the FP64 pipe full, and that pipe is in-order: every
//     DEPENDENT FP64 instruction of the prologue would queue behind hundreds of cycles of other warps' DMMAs
//     (profiles/r01_fp64_queue_latency.txt).  So the prologue has no FP64 chain: the Taylor degree comes from an
//     FP32 bound on ||dt G(u)||_1 (rounded up: the degree is never too small), and the series is run on
//     G' = dt G(u) -- exp(dt G) = sum_k G'^k / k! -- so the Horner coefficients are the constants 1 / k! and no
//     powers of dt are formed; d/d dt = -G E x uses the unscaled generator, kept in shared memory for that one
//     product; the jets are carried without their factor dt (and c_j), applied once at the end;
//   * nothing is written to global memory before a knot's last Horner step, and every warp passes
//     `griddepcontrol.wait` only then: with the early-Z promise (pb2_set_option) a CTA of the NEXT callback starts
//     computing while the previous grid still runs on the other half of the SM.  The d/dx_k block (I (x) E, 70 % of
//     a knot's bytes) is staged once in shared memory and leaves by eight bulk (TMA) stores.
#pragma once
#include "knot_u8p.cuh"

namespace pb2 {

struct U8qParams {
  int m, D, x_off, dt_off, u_off, nnz_jac, max_sub, nk, zlen;
  int early_z;            // 1: Z may be read before the programmatic dependency wait
  int compact;            // 1: records [E columns 0..7 | jets, d/d dt | delta] of cstride doubles go to `jac`
  int cstride;
  int split;              // > 0: CTAs >= split own one slot less than the others (knot = slot * gridDim + block)
  int n_peers;            // > 1 (sharded run, compact records): every record also goes into the gather buffer of each
                          // other rank, peers[r] + slot_off (+ knot * cstride), over NVLink; peers[self] is the local one
  int self;
  long long slot_off;     // doubles: this rank's slot inside every gather buffer (jac == peers[self] + slot_off)
  double* peers[8];
  long long flag_off;     // >= 0: doubles offset of 16 64-bit words inside every gather buffer (0..7 arrival words, one per
                          // source rank; 8, 9 local counters): the grid ends with an all-to-all exchange of arrival
                          // words, so when it completes every other rank's records of this step have landed here as
                          // well (no separate barrier kernel)
  int nowait;             // 1: no programmatic dependency wait at all (the caller's promise: PB2_OPT_PIPELINED)
  // shared-memory layout in doubles (u8q_layout)
  int o_norm, o_tab, o_f32, o_slot, slot_stride, zpad, o_prep, o_y, o_est, o_rec, o_mbar;
  double cj[4];           // UNIT: the common magnitude of drive generator j's nonzeros
  const double* tables;   // u8q_tables: [G fragments (m+1) 256 | norms | theta | 1/k! | float norms (8) | float theta (20)]
  const EllEntry* ell;
  const double* Z;
  double* delta;          // may be null
  double* jac;
  long long* trace;       // debug build: [launch 64][cta 304][warp 8][8] clock64 stamps
  int trace_id;
};

constexpr int kU8qTraceCtas = 608;
constexpr int kU8qPrep = 520;   // doubles per prepared knot: G' = dt G(u) fragments 256 | G(u) fragments 256 | M, n_sub | dt' | pad

// |x| as a float that is never smaller than |x|, by bit manipulation on the integer pipe (a cvt.f32.f64 would be
// one more instruction queued on the FP64 pipe).  Below the float range: 0 (irrelevant for a norm bound); NaN stays
// NaN, overflow becomes +inf.
__device__ __forceinline__ float u8q_absf_up(double x) {
  const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu, lo = (unsigned)__double2loint(x);
  const unsigned e = hi >> 20;
  const unsigned base = ((hi << 3) | (lo >> 29)) - 0xC0000000u + 1u;   // re-biased exponent | top 23 mantissa bits, + 1 ulp
  const unsigned big = ((hi & 0xfffffu) | lo) && e == 0x7ffu ? 0x7fc00000u : 0x7f800000u;
  return __uint_as_float(e < 897u ? 0u : (e >= 1151u ? big : base));
}

// ||dt G|| beyond the largest tabulated radius (rare; kept out of the instruction stream): number of sub-steps
__device__ __noinline__ void u8q_substeps(float nrm, float th_max, int max_sub, double& dt, float& per, int& n_sub) {
  const float ns = ceilf(nrm / th_max * 1.000001f);
  if (ns <= (float)max_sub) {
    n_sub = (int)ns;
    dt = dt / (double)n_sub;
    per = nrm / ns * 1.000001f;
  } else {
    dt = __longlong_as_double(0x7ff8000000000000LL);   // norm beyond the supported range: NaN out
    n_sub = 2;                                          // (takes the re-scaling path, which spreads the NaN into G')
  }
}

#ifdef PB2_TRACE
#define U8Q_STAMP(i) do { if (p.trace && lane == 0 && blockIdx.x < kU8qTraceCtas) p.trace[((((size_t)(p.trace_id & 15) * kU8qTraceCtas + blockIdx.x) * 8 + w) * 8) + (i)] = clock64(); } while (0)
#else
#define U8Q_STAMP(i) do { } while (0)
#endif

// Sharded runs: the knot's compact record [E columns 0..7 | jets, d/d dt | delta] is staged in shared memory by the
// slot's two warps and leaves by ONE bulk (TMA) store per destination -- the local gather buffer and, over NVLink,
// every other rank's (cuMem-mapped peer memory): 7 KB contiguous per transfer instead of 64-byte fragments from
// the register layout (measured at N = 2: 290 GB/s with per-lane st.global, link rate with bulk stores).
// warp B, one step: E product (accumulator pre-loaded with the unit columns' coefficient / c_k b_E), barrier,
// next additive term, J_3 [, J_4]
template <bool UNIT, bool GEN>
__device__ __forceinline__ void u8q_step_B(double (&tE)[4], double (&t)[2][4], const double (&bE)[4], const double (&bJ)[2][4],
                                           const double (&A)[4][2], int iE, uint32_t ck_addr, uint32_t ck_next_addr, int bar,
                                           bool two, double (&accE)[4], const double (&ev)[2][4], const int (&sg)[2][4],
                                           const uint32_t (&yad)[2][4], bool mma) {
  double dE[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) dE[i >> 1][i & 1] = accE[i];
  double ck = 0.0;
  if (GEN) ck = lds_f64<0>(ck_addr);
  const double ckn = lds_f64<0>(ck_next_addr);
  u8_mma_acc(dE, tE, A);
  bar_sync(bar, 64);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accE[i] = GEN ? ckn * bE[i] : ((i == iE) ? ckn : 0.0);
    tE[i] = dE[i >> 1][i & 1];
  }
  u8p_jet_tile<UNIT, GEN>(t[0], bJ[0], A, ev[0], sg[0], yad[0], ck, mma);
  if (two) u8p_jet_tile<UNIT, GEN>(t[1], bJ[1], A, ev[1], sg[1], yad[1], ck, mma);
}

// NS = knots (slots) per CTA: 4 -> 256 threads, two CTAs per SM, slot s on sub-partition s (6 tiles each);
//                             2 -> 128 threads, four CTAs per SM, warps A0 B0 A1 B1 on sub-partitions 0..3 (3 tiles each);
//                             1 -> 64 threads, eight CTAs per SM: every knot its own CTA, so all of an SM's 6.75 knots
//                                  (C3) are resident at once and the scheduler refills a slot the moment ONE knot
//                                  retires (measured: C3 8.53 -> 7.4 us per callback, C5 56.7 -> 54.3 us; the default).
template <bool UNIT, int NS>
__global__ void __launch_bounds__(64 * NS, 8 / NS) knot_u8q_kernel(const __grid_constant__ U8qParams p) {
  extern __shared__ __align__(16) double u8q_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // role 0: warp A (X, J_1, J_2), 1: warp B (E, J_3, J_4)
  const int slot = NS == 4 ? (w & 3) : (w >> 1), role = NS == 4 ? (w >> 2) : (w & 1);
  const int g = lane >> 2, q = lane & 3, m = p.m, bar = 1 + slot;

  const uint32_t a_cG = smem_u32(u8q_smem);
  const uint32_t a_slot = a_cG + 8u * (uint32_t)(p.o_slot + slot * p.slot_stride);
  const uint32_t a_z = a_slot, a_p = a_slot + 8u * p.o_prep, a_y = a_slot + 8u * p.o_y;
  const uint32_t a_est = a_slot + 8u * p.o_est;
  const uint32_t mb_z = a_slot + 8u * p.o_mbar, mb_tab = a_cG + 8u * (uint32_t)(p.o_f32 + 14);
  // CTAs >= split own NS - 1 knots; WHICH slot stays empty rotates with the block index, so that over the grid
  // every sub-partition carries the same number of tiles
  const bool shortcta = p.split > 0 && (int)blockIdx.x >= p.split;
  const int skip = shortcta ? ((int)blockIdx.x % NS) : NS;
  const int kslot = slot - (slot > skip ? 1 : 0);         // the slot's rank among the CTA's occupied slots
  const int k = kslot * gridDim.x + blockIdx.x;           // this slot's knot
  const bool have = slot != skip && k < p.nk;
  const uint32_t zbytes = (uint32_t)p.zlen * 8u;

  U8Q_STAMP(0);
  if (role == 0 && lane == 0) {
    // first thing: the slab of this slot's knot (its HBM latency is the longest item of the prologue)
    mbar_init(mb_z, 1);
    fence_proxy_async();
    if (have) {
      if (!p.early_z && !p.nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
      mbar_expect_tx(mb_z, zbytes);
      bulk_g2s(a_z, p.Z + (size_t)k * p.D, zbytes, mb_z);
    }
  }
  if (w == 1 && lane == 0) {
    // the handle's constant tables (never written after pb2_create) arrive by one bulk copy
    mbar_init(mb_tab, 1);
    fence_proxy_async();
    mbar_expect_tx(mb_tab, 8u * (uint32_t)(p.o_f32 + 14));
    bulk_g2s(a_cG, p.tables, 8u * (uint32_t)(p.o_f32 + 14), mb_tab);
#ifdef PB2_TRACE
    if (p.trace && blockIdx.x < kU8qTraceCtas) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[((((size_t)(p.trace_id & 15) * kU8qTraceCtas + blockIdx.x) * 8 + 1) * 8) + 6] = (long long)smid;
    }
#endif
  }
  // programmatic dependent launch: the next grid of the stream may take a half-SM as soon as one is free
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
  U8Q_STAMP(1);
  if (!have) return;                                       // both warps of an empty slot leave together

  // ---- the knot's preparation.  Everything that does not depend on the trajectory is in registers BEFORE the slab
  //      lands (the tables come from L2, the slab from HBM): what is left afterwards is a handful of loads, one
  //      level of FMAs and the hand-over.  Warp B: G(u) = G0 + sum_j u_j G_j and G' = dt G(u) in B-fragment order;
  //      warp A: the Taylor degree from an FP32 bound on ||dt G(u)||_1 (every operand rounded up).
  const uint32_t a_meta = a_p + 8u * 512u;                 // {M, n_sub} (two int32), then dt' (double)
  const uint32_t a_c = a_cG + 8u * (uint32_t)(p.o_tab + 20);   // Horner coefficients: the constants 1 / k!
  double A[4][2];
  mbar_wait(mb_tab, 0);
  if (role == 1) {
    double g0[8], gv[4][8];
#pragma unroll
    for (int s = 0; s < 8; ++s) g0[s] = lds_f64<0>(a_cG + 8u * (uint32_t)(s * 32 + lane));
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 0; s < 8; ++s) gv[j][s] = j < m ? lds_f64<0>(a_cG + 8u * (uint32_t)((1 + j) * 256 + s * 32 + lane)) : 0.0;
    mbar_wait(mb_z, 0);
    U8Q_STAMP(2);
    double uj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) uj[j] = j < m ? lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + j)) : 0.0;
    const double dt = lds_f64<0>(a_z + 8u * p.dt_off);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      double acc = g0[s];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc = fma(uj[j], gv[j][s], acc);
      const double gp = dt * acc;
      sts_f64<0>(a_p + 8u * (uint32_t)(s * 32 + lane), gp);
      sts_f64<0>(a_p + 8u * (uint32_t)(256 + s * 32 + lane), acc);
      A[s >> 1][s & 1] = gp;
    }
  } else {
    const uint32_t a_f = a_cG + 8u * (uint32_t)p.o_f32;
    float nf[5], th_l;
#pragma unroll
    for (int j = 0; j < 5; ++j) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(nf[j]) : "r"(a_f + 4u * (uint32_t)j) : "memory");
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(th_l) : "r"(a_f + 32u + 4u * (uint32_t)(lane <= kMaxDeg ? lane : kMaxDeg)) : "memory");
    const float th_max = __shfl_sync(0xffffffffu, th_l, kMaxDeg);
    mbar_wait(mb_z, 0);
    U8Q_STAMP(2);
    double uj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) uj[j] = j < m ? lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + j)) : 0.0;
    double dt = lds_f64<0>(a_z + 8u * p.dt_off);
    float nrm = nf[0];
#pragma unroll
    for (int j = 0; j < 4; ++j) nrm = fmaf(u8q_absf_up(uj[j]), nf[1 + j], nrm);
    nrm = nrm * u8q_absf_up(dt) * 1.000001f;
    int n_sub = 1;
    float per = nrm;
    if (nrm > th_max) u8q_substeps(nrm, th_max, p.max_sub, dt, per, n_sub);
    // M = 1 + #{ l in 1..kMaxDeg-1 : theta_l < per }   (theta increasing; NaN -> M = 1, and G' = NaN poisons the knot)
    const unsigned below = __ballot_sync(0xffffffffu, lane >= 1 && lane < kMaxDeg && th_l < per);
    const int Mq = 1 + __popc(below);
    if (lane == 0) {
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_meta), "r"(Mq), "r"(n_sub) : "memory");
      sts_f64<0>(a_meta + 8u, dt);
    }
  }
  bar_sync(bar, 64);                                       // hand-over
  int M, n_sub;
  lds_v2u32(a_meta, M, n_sub);
  const double dts = lds_f64<0>(a_meta + 8u);              // dt / n_sub: the factor the jets are carried without
  if (n_sub > 1) {
    // rare: the generator was scaled by dt, the sub-steps need dt / n_sub (NaN when the norm is out of range)
    if (role == 1) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const double gp = dts * lds_f64<0>(a_p + 8u * (uint32_t)(256 + s * 32 + lane));
        sts_f64<0>(a_p + 8u * (uint32_t)(s * 32 + lane), gp);
        A[s >> 1][s & 1] = gp;
      }
    }
    bar_sync(bar, 64);
  }
  if (role == 0) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) A[kt][nt] = lds_f64<0>(a_p + 8u * (uint32_t)((kt * 2 + nt) * 32 + lane));
  }
  if (!UNIT) {
    // general drive magnitudes: the coupling term is dt' G_j S
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) ev[a][i4] *= dts;
  }
  U8Q_STAMP(3);

  U8pStepTrace stt{nullptr, 0};
  const uint32_t lane_col = 8u * (uint32_t)(g * 16 + 2 * q);
  const int lc = g * 16 + 2 * q;
  const int rec = p.compact ? p.cstride : p.nnz_jac;      // doubles per knot in `jac`
  const int o_jets = p.compact ? 128 : 2048;              // where the jet columns start inside a knot's segment
  double* jk = p.jac + (size_t)k * (size_t)rec;
  double* jj = jk + o_jets;

  if (role == 0) {
    // ================================ warp A: X, J_1, J_2 ==============================================
    const uint32_t ypub = a_y + 8u * (uint32_t)(g * 4 + q);
    const uint32_t xl = 8u * (uint32_t)p.x_off + lane_col;
    double bX[4], tX[4], tJ[2][4], accX[4];
    // (sharded runs compute d/d dt after the loop: its value goes to every rank, like the rest of the record)
    double* const dT_out = p.n_peers > 1 ? nullptr : jj + m * 128 + lc;
    const uint32_t a_graw = a_p + 8u * (uint32_t)(256 + lane);   // the unscaled G(u) fragments (d/d dt product)
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
        u8p_step_A<UNIT, false>(tX, tJ, bX, a_z + xl, tJ, A, yp, a_c, a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar, accX, ev,
                                sg, yad, kq != M - 1, kq == 0 ? dT_out : nullptr, stt, !p.nowait, a_graw);
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
          u8p_step_A<UNIT, true>(tX, tJ, bX2, a_z + xl, bJ2, A, yp, a_c + 8u * (uint32_t)kq,
                                 a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar, accX, ev, sg, yad, sub > 0 || kq != M - 1,
                                 (sub == n_sub - 1 && kq == 0) ? dT_out : nullptr, stt, !p.nowait, a_graw);
          yp += (uint32_t)dy;
          u8p_flip(yad, dy);
        }
      }
    }
    // ---- jets and delta straight from the registers (d/d dt left in the last step, after the dependency wait)
    U8Q_STAMP(5);
    const bool want_delta = p.delta != nullptr || p.compact;
    double xn[4] = {0.0, 0.0, 0.0, 0.0};
    if (want_delta) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) xn[i4] = lds_f64<0>(a_z + 8u * p.D + xl + U8_OFF(i4));
    }
    const double s0 = UNIT ? -p.cj[0] * dts : -1.0, s1 = UNIT ? -p.cj[1] * dts : -1.0;
    if (p.n_peers > 1) {
      double Gr[4][2], dT[2][2];
#pragma unroll
      for (int s8 = 0; s8 < 8; ++s8) Gr[s8 >> 1][s8 & 1] = lds_f64<0>(a_graw + 256u * (uint32_t)s8);
      u8_mma(dT, tX, Gr);
      const uint32_t o = a_slot + 8u * (uint32_t)p.o_rec + 1024u + lane_col;   // jets start after the E half
      sts_f64x2<0>(o, make_double2(s0 * tJ[0][0], s0 * tJ[0][1]));
      sts_f64x2<64>(o, make_double2(s0 * tJ[0][2], s0 * tJ[0][3]));
      sts_f64x2<1024>(o, make_double2(s1 * tJ[1][0], s1 * tJ[1][1]));
      sts_f64x2<1024 + 64>(o, make_double2(s1 * tJ[1][2], s1 * tJ[1][3]));
      const uint32_t oT = o + 1024u * (uint32_t)m;
      sts_f64x2<0>(oT, make_double2(-dT[0][0], -dT[0][1]));
      sts_f64x2<64>(oT, make_double2(-dT[1][0], -dT[1][1]));
      sts_f64x2<1024>(oT, make_double2(xn[0] - tX[0], xn[1] - tX[1]));
      sts_f64x2<1024 + 64>(oT, make_double2(xn[2] - tX[2], xn[3] - tX[3]));
      fence_proxy_async();
      bar_sync(bar, 64);        // warp B sends the staged record
    } else {
      stg_f64x2(jj + lc, s0 * tJ[0][0], s0 * tJ[0][1]);
      stg_f64x2(jj + lc + 8, s0 * tJ[0][2], s0 * tJ[0][3]);
      stg_f64x2(jj + 128 + lc, s1 * tJ[1][0], s1 * tJ[1][1]);
      stg_f64x2(jj + 128 + lc + 8, s1 * tJ[1][2], s1 * tJ[1][3]);
      if (want_delta) {
        double* dd = (p.compact ? jj + (m + 1) * 128 : p.delta + (size_t)k * 128) + lc;
        stg_f64x2(dd, xn[0] - tX[0], xn[1] - tX[1]);
        stg_f64x2(dd + 8, xn[2] - tX[2], xn[3] - tX[3]);
      }
    }
  } else {
    // ================================ warp B: E, J_3 [, J_4] ============================================
    const bool two = m >= 4;
    const int iE = (g == 2 * q) ? 0 : ((g == 2 * q + 1) ? 1 : -1);   // which element is the unit entry
    double tE[4], accE[4], t[2][4];
    {
      const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
      const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        tE[i4] = (i4 == iE) ? cM : 0.0;
        accE[i4] = (i4 == iE) ? c0 : 0.0;
        t[0][i4] = 0.0;
        t[1][i4] = 0.0;
      }
    }
    int dy = 1024;
    if (n_sub == 1) {
#pragma unroll 1
      for (int kq = M - 1; kq >= 0; --kq) {
        u8q_step_B<UNIT, false>(tE, t, tE, t, A, iE, a_c, a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar, two, accE, ev, sg, yad,
                                kq != M - 1);
        u8p_flip(yad, dy);
      }
    } else {
      double bE[4], bJ[2][4];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        bE[i4] = (i4 == iE) ? 1.0 : 0.0;
        bJ[0][i4] = 0.0;
        bJ[1][i4] = 0.0;
      }
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
        if (sub > 0) {
          const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
          const double c0 = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            bE[i4] = tE[i4];
            tE[i4] *= cM;
            accE[i4] = c0 * bE[i4];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              bJ[a][i4] = t[a][i4];
              t[a][i4] *= cM;
            }
          }
        }
#pragma unroll 1
        for (int kq = M - 1; kq >= 0; --kq) {
          u8q_step_B<UNIT, true>(tE, t, bE, bJ, A, iE, a_c + 8u * (uint32_t)kq, a_c + 8u * (uint32_t)(kq >= 1 ? kq - 1 : 0), bar,
                                 two, accE, ev, sg, yad, sub > 0 || kq != M - 1);
          u8p_flip(yad, dy);
        }
      }
    }
    U8Q_STAMP(5);
    // the first global writes of this warp: from here on the previous grid of the stream must be complete
    if (!p.nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
    const double s2 = UNIT ? -p.cj[2] * dts : -1.0, s3 = UNIT ? -p.cj[3] * dts : -1.0;
    if (p.n_peers > 1) {
      const uint32_t a_rec = a_slot + 8u * (uint32_t)p.o_rec, o = a_rec + lane_col;
      sts_f64x2<0>(o, make_double2(-tE[0], -tE[1]));
      sts_f64x2<64>(o, make_double2(-tE[2], -tE[3]));
      sts_f64x2<1024 + 2 * 1024>(o, make_double2(s2 * t[0][0], s2 * t[0][1]));
      sts_f64x2<1024 + 2 * 1024 + 64>(o, make_double2(s2 * t[0][2], s2 * t[0][3]));
      if (two) {
        sts_f64x2<1024 + 3 * 1024>(o, make_double2(s3 * t[1][0], s3 * t[1][1]));
        sts_f64x2<1024 + 3 * 1024 + 64>(o, make_double2(s3 * t[1][2], s3 * t[1][3]));
      }
      fence_proxy_async();
      bar_sync(bar, 64);        // warp A's part of the record is staged as well
      if (lane == 0) {
        const size_t go = (size_t)k * (size_t)rec;
        const uint32_t nbytes = 8u * (uint32_t)rec;
        bulk_s2g(p.jac + go, a_rec, nbytes);
        // destinations in a rank-dependent rotation (self + 1, self + 2, ...), further rotated by the knot: at any
        // moment the ranks' transfers fan out over all peers instead of converging on one ingress port
        for (int i = 0; i < p.n_peers - 1; ++i) {
          int r = p.self + 1 + ((i + k) % (p.n_peers - 1));
          if (r >= p.n_peers) r -= p.n_peers;
          bulk_s2g(p.peers[r] + p.slot_off + go, a_rec, nbytes);
        }
        bulk_commit();
        bulk_wait0();            // the remote writes are performed before the grid completes
        if (p.flag_off >= 0) {
          // the knot whose record leaves last closes the step: tell every rank that all of this rank's records
          // are on their way (release: they are performed, see above), then wait for the same word from everyone
          __threadfence_system();
          // words 0..7: arrivals (written by the peers), 8: ticket of finished knots, 9: how often THIS buffer has
          // been used -- the step's epoch.  Both counters live in the buffer, so launches that overlap (different
          // buffers) never share them, and every rank derives the same epoch for the same step.
          unsigned long long* mine = reinterpret_cast<unsigned long long*>(p.peers[p.self] + p.flag_off);
          if (atomicAdd(mine + 8, 1ull) == (unsigned long long)(p.nk - 1)) {
            mine[8] = 0ull;
            const unsigned long long e = mine[9] + 1ull;
            mine[9] = e;
            for (int r = 0; r < p.n_peers; ++r)
              if (r != p.self) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(p.peers[r] + p.flag_off) + p.self;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(e) : "memory");
              }
            for (int r = 0; r < p.n_peers; ++r)
              if (r != p.self) {
                unsigned long long v;
                do {
                  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + r) : "memory");
                } while (v < e);
              }
          }
        }
      }
      U8Q_STAMP(7);
      return;
    }
    if (p.compact) {
      stg_f64x2(jk + lc, -tE[0], -tE[1]);
      stg_f64x2(jk + lc + 8, -tE[2], -tE[3]);
    } else {
      // -E = -[[P, -Q], [Q, P]] staged once (own column g, mirrored column g + 8); the d/dx_k block is
      // I (x) E: eight bulk (TMA) stores of the same 2 KB
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
      // the constant d/dx_{k+1} identity entries
      stg_f64x2(jj + (m + 1) * 128 + 4 * lane, 1.0, 1.0);
      stg_f64x2(jj + (m + 1) * 128 + 4 * lane + 2, 1.0, 1.0);
    }
    stg_f64x2(jj + 2 * 128 + lc, s2 * t[0][0], s2 * t[0][1]);
    stg_f64x2(jj + 2 * 128 + lc + 8, s2 * t[0][2], s2 * t[0][3]);
    if (two) {
      stg_f64x2(jj + 3 * 128 + lc, s3 * t[1][0], s3 * t[1][1]);
      stg_f64x2(jj + 3 * 128 + lc + 8, s3 * t[1][2], s3 * t[1][3]);
    }
    // the staged propagator block must outlive the bulk stores' reads of it; their writes complete with the grid
    if (!p.compact && lane == 0) bulk_wait_read0();
  }
  U8Q_STAMP(7);
}

// Shared-memory layout (doubles): the table blob [G fragments (m+1) 256 | norms | theta, 1 / k! | float norms,
// float theta], the tables' mbarrier, the CTA's product-phase start time, then per slot
//   slab | prepared knot (G' fragments, G fragments, M, n_sub, dt') | X exchange x2 | staged propagator block | mbarriers.
inline size_t u8q_layout(U8qParams& q, int ns) {
  auto even = [](int v) { return (v + 1) & ~1; };
  q.o_norm = (q.m + 1) * 256;
  q.o_tab = q.o_norm + even(q.m + 1);
  q.o_f32 = q.o_tab + 40;
  q.o_slot = q.o_f32 + 16;
  q.zpad = even(q.zlen);
  q.o_prep = q.zpad;
  q.o_y = q.o_prep + kU8qPrep;
  q.o_est = q.o_y + 2 * 128;
  q.o_rec = q.o_est + 256;                                    // sharded runs: the knot's compact record, staged
  q.o_mbar = q.o_rec + (q.n_peers > 1 ? q.cstride : 0);
  q.slot_stride = q.o_mbar + 2;
  return sizeof(double) * ((size_t)q.o_slot + (size_t)ns * q.slot_stride);
}

// Host side: the table blob.  `tab40` = theta_0..19 | 1/0! .. 1/19!.  The float copies bound from the safe side:
// norms rounded up, theta rounded down.
inline std::vector<double> u8q_tables(const DmmaPlan& pl, int m, const double* tab40) {
  std::vector<double> blob(pl.gfrag);
  blob.insert(blob.end(), pl.norms.begin(), pl.norms.end());
  if (blob.size() % 2) blob.push_back(0.0);
  blob.insert(blob.end(), tab40, tab40 + 40);
  float f32[28] = {0};
  for (int j = 0; j <= m && j < 8; ++j) {
    float v = (float)pl.norms[j];
    if ((double)v < pl.norms[j]) v = std::nextafter(v, INFINITY);
    f32[j] = v * 1.000001f;
  }
  for (int q = 0; q < 20; ++q) {
    float v = (float)tab40[q];
    if ((double)v > tab40[q]) v = std::nextafter(v, 0.0f);
    f32[8 + q] = v * 0.999999f;
  }
  const size_t o = blob.size();
  blob.resize(o + 14);
  std::memcpy(blob.data() + o, f32, sizeof(f32));
  return blob;
}

inline auto u8q_kernel(bool unit, int ns) -> void (*)(const U8qParams) {
  if (ns == 2) return unit ? knot_u8q_kernel<true, 2> : knot_u8q_kernel<false, 2>;
  if (ns == 1) return unit ? knot_u8q_kernel<true, 1> : knot_u8q_kernel<false, 1>;
  return unit ? knot_u8q_kernel<true, 4> : knot_u8q_kernel<false, 4>;
}

}  // namespace pb2
