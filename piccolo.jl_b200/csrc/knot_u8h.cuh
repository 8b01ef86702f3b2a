// Tensor-core Lagrangian-Hessian kernel for the 3-qubit unitary shape (generator 16 x 16, real
// isomorphism of an anti-Hermitian matrix, 8 state columns, 1..4 drives: BASELINE configs C3 / C5).
//
// What it replaces: DirectTrajOpt's Hessian of the Lagrangian for a BilinearIntegrator (enabled through
// `eval_hessian`, call site /root/reference/src/control/templates/spline_pulse_problem.jl:96; the
// integrator is built at src/control/integrators.jl:35-51) -- nested ForwardDiff through
// Ghat(u) and expv in the reference.  With M = reshape(mu_k) and f = -<M, exp(dt G(u)) X> per knot:
//     (x, u_j)   = -F_j^T M          (x, dt)   = -(G E)^T M
//     (u_i, u_j) = -<M, L2(dt G; dt G_i, dt G_j) X>
//     (u_j, dt)  = -<M, (G_j E + G F_j) X>          (dt, dt) = -<M, G^2 E X>
// All of it comes out of ONE truncated-Taylor Horner recurrence (the same a_k = dt^k / k!, degree M
// and sub-steps as the residual / Jacobian kernels) run on 8-column tiles with DMMA.8x8x4:
//   forward tiles   X,  J_j (first-order jets),  P_ij (second-order jets, i <= j):
//        S    <- a_k B    + G S
//        S_j  <- a_k B_j  + G S_j  + G_j S
//        S_ij <- a_k B_ij + G S_ij + G_i S_j + G_j S_i
//   adjoint tiles   Mt, JA_j  -- for an anti-symmetric generator E^T = exp(-dt G), so the adjoint
//   quantities E^T M and F_j^T M are the same recurrence with G -> -G, G_j -> -G_j started from mu.
// 2 + 2m + m(m+1)/2 tiles (20 at m = 4), two per warp, one knot per persistent CTA at a time: a knot
// of the Hessian has enough independent tiles to fill the four tensor pipes by itself.  A producer warp
// receives (z_k, x_{k+1}) and mu_k by TMA, builds G(u_k) in B-fragment order, the Taylor degree and
// coefficients for the next knot while the compute warps work, and writes the staged 655 values
// of the finished knot to HBM.
#pragma once
#include "knot_u8.cuh"

namespace pb2 {

struct U8hParams {
  int m, D, x_off, dt_off, u_off, nnz_hess, max_sub, nk, zlen;
  int ntiles, ncw;       // tiles per knot, compute warps (forward tiles two per warp, then adjoint tiles two per warp)
  // shared-memory layout in doubles (u8h_layout)
  int o_norm, o_tab, o_slab, zpad, o_prep, o_xch, o_stage, o_mbar;
  const double* tables;  // [G fragments (m+1) 256 | norms (padded even) | theta | 1/k!], smem order
  const EllEntry* ell;   // (m+1) * 16 * W
  const double* Z;
  const double* mu;
  double* hess;
  long long* trace;      // debug build only
};

constexpr int kU8hXchBytes = 6 * 1024;   // exchange buffers per step parity: X, Mt, J_0..J_3 (1 KB each)

// tile kinds
enum { U8H_X = 0, U8H_J = 1, U8H_P = 2, U8H_MT = 3, U8H_JA = 4, U8H_NONE = 5 };

struct U8hSlot {
  int kind, i, j;
};

// tile order: X, J_0.., P_ij (j major, i <= j: the COO order of the (u,u) block), Mt, JA_0..
__host__ __device__ inline U8hSlot u8h_tile(int t, int m) {
  U8hSlot s{U8H_NONE, 0, 0};
  const int npair = m * (m + 1) / 2;
  if (t == 0) { s.kind = U8H_X; return s; }
  t -= 1;
  if (t < m) { s.kind = U8H_J; s.j = t; return s; }
  t -= m;
  if (t < npair) {
    int j = 0;
    while ((j + 1) * (j + 2) / 2 <= t) ++j;
    s.kind = U8H_P; s.j = j; s.i = t - j * (j + 1) / 2;
    return s;
  }
  t -= npair;
  if (t == 0) { s.kind = U8H_MT; return s; }
  t -= 1;
  if (t < m) { s.kind = U8H_JA; s.j = t; return s; }
  return s;
}

// One Horner step of a warp's two tiles,  t <- c_k base +- G t + (coupling terms read from the
// exchange buffers).  PAR: exchange-buffer half.  Everything but the product is formed first and handed
// to the tensor instructions as their accumulator, so no FP64 CUDA-core instruction sits between two
// steps' products (such an instruction waits for every DMMA other warps have queued on the
// sub-partition: tools/dfma_lat.cu).  The sign of an adjoint tile (G^T = -G) goes onto the operand by
// flipping its sign bit, an integer instruction.
template <int W, int PAR>
__device__ __forceinline__ void u8h_step(double (&t)[2][4], const double (&base)[2][4], const double (&A)[4][2],
                                         const double (&ev)[2][2][4][W], const uint32_t (&yad)[2][2][4][W],
                                         const uint32_t (&pub)[2], const int (&skip)[2], const double (&sgn)[2],
                                         const bool (&act)[2], int s, uint32_t ck_addr, int bar, int nthreads) {
#pragma unroll
  for (int a = 0; a < 2; ++a)
    if (pub[a]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) sts_f64<PAR * kU8hXchBytes>(pub[a] + i * 256, t[a][i]);
    }
  const double ck = lds_f64<0>(ck_addr);
  bar_sync(bar, nthreads);
  double d[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double v = ck * base[a][i];
      if (act[a]) {
#pragma unroll
        for (int term = 0; term < 2; ++term)
#pragma unroll
          for (int ww = 0; ww < W; ++ww)
            v = fma(ev[a][term][i][ww], lds_f64<PAR * kU8hXchBytes>(yad[a][term][i][ww]), v);
      }
      d[a][i >> 1][i & 1] = v;
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    if (act[a] && s >= skip[a]) {
      const int flip = sgn[a] < 0.0 ? (int)0x80000000 : 0;
      double tn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tn[i] = __hiloint2double(__double2hiint(t[a][i]) ^ flip, __double2loint(t[a][i]));
      u8_mma_acc(d[a], tn, A);
    }
    if (act[a]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) t[a][i] = d[a][i >> 1][i & 1];
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <int W>
__global__ void __launch_bounds__(384, 1) knot_u8h_kernel(const __grid_constant__ U8hParams p) {
  extern __shared__ __align__(16) double u8_smem[];
  const int lane = threadIdx.x & 31, wcta = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m = p.m, ncw = p.ncw;
  const int npair = m * (m + 1) / 2;

  const uint32_t a_cG = smem_u32(u8_smem);
  const uint32_t a_slab = a_cG + 8u * p.o_slab, a_prep = a_cG + 8u * p.o_prep, a_xch = a_cG + 8u * p.o_xch;
  const uint32_t a_stage = a_cG + 8u * p.o_stage, a_mbar = a_cG + 8u * p.o_mbar;
  const uint32_t mb_zfull = a_mbar, mb_ready = a_mbar + 24, mb_staged = a_mbar + 40, mb_free = a_mbar + 48;
  const uint32_t zbytes = (uint32_t)p.zlen * 8u, slab_stride = 8u * (uint32_t)(p.zpad + 128);

  const int stride = gridDim.x;
  const int n_my = (int)blockIdx.x < p.nk ? (p.nk - (int)blockIdx.x + stride - 1) / stride : 0;

  // ---- once per CTA ------------------------------------------------------------------------------
  const uint32_t mb_tab = a_cG + 8u * (uint32_t)(p.o_tab + 40);
  if (threadIdx.x == 0) {
    mbar_init(mb_tab, 1);
    for (int i = 0; i < 3; ++i) mbar_init(mb_zfull + 8 * i, 1);
    mbar_init(mb_ready, 1);
    mbar_init(mb_ready + 8, 1);
    mbar_init(mb_staged, ncw);
    mbar_init(mb_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // constant tables (G fragments, norms, theta / factorial tables) by one bulk copy
    mbar_expect_tx(mb_tab, 8u * (uint32_t)(p.o_tab + 40));
    bulk_g2s(a_cG, p.tables, 8u * (uint32_t)(p.o_tab + 40), mb_tab);
    for (int i = 0; i < 2 && i < n_my; ++i) {
      const size_t k = (size_t)blockIdx.x + (size_t)i * stride;
      mbar_expect_tx(mb_zfull + 8 * i, zbytes + 1024u);
      bulk_g2s(a_slab + i * slab_stride, p.Z + k * p.D, zbytes, mb_zfull + 8 * i);
      bulk_g2s(a_slab + i * slab_stride + 8u * p.zpad, p.mu + k * 128, 1024u, mb_zfull + 8 * i);
    }
    mbar_arrive(mb_free);   // the stage starts free
  }
  __syncthreads();

  if (wcta == 0) {
    // =============================== producer warp ===============================================
    mbar_wait(mb_tab, 0);
    const double th_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double if_l = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + 20 + (lane <= kMaxDeg ? lane : kMaxDeg)));
    const double th_max = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_tab + kMaxDeg));
    int s3 = 0;
    for (int i = 0; i <= n_my; ++i) {
      if (i < n_my) {
        const uint32_t a_z = a_slab + (uint32_t)s3 * slab_stride;
        const uint32_t a_p = a_prep + 8u * (uint32_t)((i & 1) * kU8Prep);
        mbar_wait(mb_zfull + 8 * s3, (uint32_t)((i / 3) & 1));
        double uj[4], nj[4], gv[4][8], acc[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uj[j] = 0.0;
          nj[j] = 0.0;
          if (j < m) {
            uj[j] = lds_f64<0>(a_z + 8u * (uint32_t)(p.u_off + j));
            nj[j] = lds_f64<0>(a_cG + 8u * (uint32_t)(p.o_norm + 1 + j));
          }
        }
        double dt = lds_f64<0>(a_z + 8u * p.dt_off);
        double nrm = lds_f64<0>(a_cG + 8u * p.o_norm);
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
        for (int j = 0; j < 4; ++j)
          if (j < m) {
#pragma unroll
            for (int s = 0; s < 8; ++s) acc[s] = fma(uj[j], gv[j][s], acc[s]);
            nrm = fma(fabs(uj[j]), nj[j], nrm);
          }
#pragma unroll
        for (int s = 0; s < 8; ++s) sts_f64<0>(a_p + 8u * (uint32_t)(s * 32 + lane), acc[s]);
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
        const unsigned below = __ballot_sync(0xffffffffu, lane >= 1 && lane < kMaxDeg && th_l < per);
        const int M = 1 + __popc(below);
        double pw = 1.0, sq = dt;
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
          if ((lane >> bit) & 1) pw *= sq;
          sq *= sq;
        }
        if (lane <= kMaxDeg) sts_f64<0>(a_p + 8u * 256u + 8u * lane, lane <= M ? if_l * pw : 0.0);
        if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_p + 8u * 276u), "r"(M), "r"(n_sub) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(mb_ready + 8 * (i & 1));
      }
      if (i >= 1) {
        // ---- finish knot i-1: staged values -> HBM (655 doubles, 8-byte aligned: plain stores) -----
        const size_t kprev = (size_t)blockIdx.x + (size_t)(i - 1) * stride;
        mbar_wait(mb_staged, (uint32_t)((i - 1) & 1));
        double* out = p.hess + kprev * p.nnz_hess;
        for (int e = lane; e < p.nnz_hess; e += 32) out[e] = lds_f64<0>(a_stage + 8u * e);
        __syncwarp();
        if (lane == 0) mbar_arrive(mb_free);
      }
      if (lane == 0 && i + 2 < n_my) {
        const int s = s3 == 0 ? 2 : s3 - 1;   // (i + 2) % 3: held knot i-1, which is finished
        const size_t k = (size_t)blockIdx.x + (size_t)(i + 2) * stride;
        mbar_expect_tx(mb_zfull + 8 * s, zbytes + 1024u);
        bulk_g2s(a_slab + s * slab_stride, p.Z + k * p.D, zbytes, mb_zfull + 8 * s);
        bulk_g2s(a_slab + s * slab_stride + 8u * p.zpad, p.mu + k * 128, 1024u, mb_zfull + 8 * s);
      }
      __syncwarp();
      s3 = s3 == 2 ? 0 : s3 + 1;
    }
    return;
  }

  // ================================= compute warps ==================================================
  const int cw = wcta - 1;
  const uint32_t lane_col = 8u * (uint32_t)(g * 16 + 2 * q);
  const uint32_t lane_x = 8u * (uint32_t)(g * 4 + q);   // this lane's slot inside a 1 KB exchange buffer
  U8hSlot sl[2];
  bool act[2];
  int skip[2];
  double sgn[2];
  uint32_t pub[2];
  double ev[2][2][4][W];
  uint32_t yad[2][2][4][W];
#pragma unroll
  // forward tiles (X, J_j, P_ij) and adjoint tiles (Mt, JA_j) never exchange anything: they live in separate warps
  // and meet on separate named barriers, so the two recurrences drift apart and fill each other's barrier gaps
  const int nf = 1 + m + m * (m + 1) / 2, nfw = (nf + 1) / 2, naw = ncw - nfw;
  const bool adj = cw >= nfw;
  const int bar = adj ? 2 : 1, nthr = 32 * (adj ? naw : nfw);
  for (int a = 0; a < 2; ++a) {
    const int tix = adj ? nf + 2 * (cw - nfw) + a : (2 * cw + a < nf ? 2 * cw + a : p.ntiles);
    sl[a] = tix < p.ntiles ? u8h_tile(tix, m) : U8hSlot{U8H_NONE, 0, 0};
    const int kind = sl[a].kind;
    act[a] = kind != U8H_NONE;
    skip[a] = (kind == U8H_J || kind == U8H_JA) ? 1 : (kind == U8H_P ? 2 : 0);
    sgn[a] = (kind == U8H_MT || kind == U8H_JA) ? -1.0 : 1.0;
    // exchange buffers: 0 = X, 1 = Mt, 2 + j = J_j
    pub[a] = kind == U8H_X ? a_xch + lane_x : (kind == U8H_MT ? a_xch + 1024u + lane_x
             : (kind == U8H_J ? a_xch + 1024u * (uint32_t)(2 + sl[a].j) + lane_x : 0u));
    // coupling terms: (drive, source buffer, scale)
    // (unused terms carry a zero coefficient and read the tile's OWN group's state buffer: the other group may be in
    // a different knot, and 0 * NaN of a poisoned knot must not leak into this one)
    int drv[2] = {m, m}, src[2] = {adj ? 1 : 0, adj ? 1 : 0};
    double sc[2] = {0.0, 0.0};
    if (kind == U8H_J) { drv[0] = sl[a].j; src[0] = 0; sc[0] = 1.0; }
    if (kind == U8H_JA) { drv[0] = sl[a].j; src[0] = 1; sc[0] = -1.0; }
    if (kind == U8H_P) {
      if (sl[a].i == sl[a].j) { drv[0] = sl[a].i; src[0] = 2 + sl[a].i; sc[0] = 2.0; }
      else {
        drv[0] = sl[a].i; src[0] = 2 + sl[a].j; sc[0] = 1.0;
        drv[1] = sl[a].j; src[1] = 2 + sl[a].i; sc[1] = 1.0;
      }
    }
#pragma unroll
    for (int term = 0; term < 2; ++term)
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const int r = 8 * (i4 >> 1) + 2 * q + (i4 & 1);
#pragma unroll
        for (int ww = 0; ww < W; ++ww) {
          const EllEntry en = p.ell[((size_t)drv[term] * 16 + r) * W + ww];   // drive m: all-zero dummy
          ev[a][term][i4][ww] = en.val * sc[term];
          yad[a][term][i4][ww] = a_xch + 1024u * (uint32_t)src[term] +
                                 8u * (uint32_t)((2 * (en.idx >> 3) + (en.idx & 1)) * 32 + g * 4 + ((en.idx & 7) >> 1));
        }
      }
  }
  const uint32_t xl = 8u * (uint32_t)p.x_off + lane_col;

  int s3 = 0;
  for (int i = 0; i < n_my; ++i) {
    const uint32_t a_z = a_slab + (uint32_t)s3 * slab_stride;
    const uint32_t a_mu = a_z + 8u * p.zpad + lane_col;
    const uint32_t a_p = a_prep + 8u * (uint32_t)((i & 1) * kU8Prep);
    const uint32_t a_c = a_p + 8u * 256u;
    const int i_knot = i;
    U8_STAMP(0);
    mbar_wait(mb_ready + 8 * (i & 1), (uint32_t)((i >> 1) & 1));
    U8_STAMP(1);
    double A[4][2];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) A[kt][nt] = lds_f64<0>(a_p + 8u * (uint32_t)((kt * 2 + nt) * 32 + lane));
    int M, n_sub;
    lds_v2u32(a_p + 8u * 276u, M, n_sub);
    double mu4[4];   // mu at this lane's (row, column) positions
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) mu4[i4] = lds_f64<0>(a_mu + U8_OFF(i4));
    double t[2][4], base[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        double v = 0.0;
        if (sl[a].kind == U8H_X) v = lds_f64<0>(a_z + xl + U8_OFF(i4));
        if (sl[a].kind == U8H_MT) v = mu4[i4];
        base[a][i4] = v;
      }
    bar_sync(bar, nthr);   // exchange buffers free (readers of the previous knot are done)
    U8_STAMP(2);
    for (int sub = 0; sub < n_sub; ++sub) {
      const double cM = lds_f64<0>(a_c + 8u * (uint32_t)M);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          if (sub > 0) base[a][i4] = t[a][i4];
          t[a][i4] = cM * base[a][i4];
        }
      if (sub > 0) bar_sync(bar, nthr);
      // the structural zeros of the first sub-step (jets start from zero) are skipped only there
      int skp[2] = {sub == 0 ? skip[0] : 0, sub == 0 ? skip[1] : 0};
      int s = 0, kq = M - 1;
      for (; kq >= 1; kq -= 2, s += 2) {
        u8h_step<W, 0>(t, base, A, ev, yad, pub, skp, sgn, act, s, a_c + 8u * kq, bar, nthr);
        u8h_step<W, 1>(t, base, A, ev, yad, pub, skp, sgn, act, s + 1, a_c + 8u * kq - 8u, bar, nthr);
      }
      if (kq == 0) u8h_step<W, 0>(t, base, A, ev, yad, pub, skp, sgn, act, s, a_c, bar, nthr);
    }

    // ---- final products and contractions ------------------------------------------------------------
    U8_STAMP(3);
    // the state tile publishes Y = E X once more: (u_j, dt) needs G_j Y.  It goes into the buffer
    // half the last Horner step did NOT use (that one may still be read by slower warps).
    const uint32_t fo = (uint32_t)(M & 1) * (uint32_t)kU8hXchBytes;
#pragma unroll
    for (int a = 0; a < 2; ++a)
      if (sl[a].kind == U8H_X) {
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) sts_f64<0>(pub[a] + fo + i4 * 256, t[a][i4]);
      }
    bar_sync(bar, nthr);
    double outv[2][4];     // tile-shaped results (adjoint tiles)
    double outs[2];        // scalar results (forward tiles)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      outs[a] = 0.0;
      const int kind = sl[a].kind;
      if (kind == U8H_NONE) continue;
      if (kind == U8H_P) {
        double acc = 0.0;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) acc = fma(mu4[i4], t[a][i4], acc);
        outs[a] = -warp_sum(acc);
      } else if (kind == U8H_X) {
        double d1[2][2], d2[2][2], t1[4];
        u8_mma(d1, t[a], A);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) t1[i4] = d1[i4 >> 1][i4 & 1];
        u8_mma(d2, t1, A);
        double acc = 0.0;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) acc = fma(mu4[i4], d2[i4 >> 1][i4 & 1], acc);
        outs[a] = -warp_sum(acc);   // (dt, dt) = -<M, G^2 E X>
      } else if (kind == U8H_J) {
        double d1[2][2];
        u8_mma(d1, t[a], A);         // G F_j X
        double acc = 0.0;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          double v = d1[i4 >> 1][i4 & 1];
#pragma unroll
          for (int ww = 0; ww < W; ++ww) v = fma(ev[a][0][i4][ww], lds_f64<0>(yad[a][0][i4][ww] + fo), v);   // + G_j Y
          acc = fma(mu4[i4], v, acc);
        }
        outs[a] = -warp_sum(acc);   // (u_j, dt)
      } else if (kind == U8H_MT) {
        double d1[2][2];
        u8_mma(d1, t[a], A);         // G P;  (x, dt) = -G^T P = +G P
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) outv[a][i4] = d1[i4 >> 1][i4 & 1];
      } else {                       // JA: (x, u_j) = -F_j^T M
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) outv[a][i4] = -t[a][i4];
      }
    }
    U8_STAMP(4);
    mbar_wait(mb_free, (uint32_t)(i & 1));
    U8_STAMP(5);
    // stage: [(x,u_j) m x 128 | (x,dt) 128 | (u_i,u_j) npair | (u_j,dt) m | (dt,dt)]
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kind = sl[a].kind;
      if (kind == U8H_JA || kind == U8H_MT) {
        const uint32_t o = a_stage + 8u * (uint32_t)((kind == U8H_JA ? sl[a].j : m) * 128) + lane_col;
        sts_f64<0>(o, outv[a][0]);   sts_f64<8>(o, outv[a][1]);
        sts_f64<64>(o, outv[a][2]);  sts_f64<72>(o, outv[a][3]);
      } else if (kind != U8H_NONE && lane == 0) {
        const int sidx = kind == U8H_P ? sl[a].j * (sl[a].j + 1) / 2 + sl[a].i
                                       : (kind == U8H_J ? npair + sl[a].j : npair + m);
        sts_f64<0>(a_stage + 8u * (uint32_t)((m + 1) * 128 + sidx), outs[a]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(mb_staged);
    U8_STAMP(6);
    s3 = s3 == 2 ? 0 : s3 + 1;
  }
}

// Shared-memory layout (doubles): fragment tables | norms | theta / factorial tables | 3 slabs of
// (z_k, x_{k+1} | mu_k) | 2 prepared knots | exchange buffers 2 x 6 KB | stage | mbarriers.
inline size_t u8h_layout(U8hParams& q) {
  auto even = [](int v) { return (v + 1) & ~1; };
  q.o_norm = (q.m + 1) * 256;
  q.o_tab = q.o_norm + even(q.m + 1);
  q.o_slab = q.o_tab + 40 + 2;   // + the tables' mbarrier
  q.zpad = even(q.zlen);
  q.o_prep = q.o_slab + 3 * (q.zpad + 128);
  q.o_xch = q.o_prep + 2 * kU8Prep;
  q.o_stage = q.o_xch + 2 * (kU8hXchBytes / 8);
  q.o_mbar = q.o_stage + even(q.nnz_hess);
  return sizeof(double) * ((size_t)q.o_mbar + 8);
}

using U8hKernel = void (*)(U8hParams);
inline U8hKernel u8h_kernel(int W) {
  return W == 1 ? knot_u8h_kernel<1> : (W == 2 ? knot_u8h_kernel<2> : knot_u8h_kernel<4>);
}

}  // namespace pb2
