// Residual + Jacobian on the tensor cores for general generators with b <= 24, many small CTAs
// (SURVEY.md section 8, rows a1-a3, a7).
//
// Same mathematics and the same tables as knot_dmma.cuh (the reference's BilinearIntegrator,
// /root/reference/src/control/integrators.jl:35-95: delta_k = x_{k+1} - expv(dt_k, Ghat(u_k), x_k) and its
// derivatives, as ExponentialAction's truncated-Taylor action run on the stacked columns [I (E) | X | jet_1 .. jet_m]
// held as transposed 8-column DMMA.8x8x4 tiles), but the work decomposition of knot_dmmah.cuh: ONE CTA PER KNOT made of
// one warp per tile, no producer, no staging -- several CTAs resident per SM, so one knot's prologue (the dependent
// global loads of dt, u, x) and its stores run under the other knots' Horner steps.  knot_dmma.cuh's persistent
// one-CTA-per-SM pipeline (TMA slab prefetch, staged bulk stores) pays off when there are many knots per SM; BASELINE's
// small systems have 0.3 - 3.4 knots per SM, and ensembles of them (pb2_batch_*, blockIdx.y = member) want many
// independent knots in flight rather than a deep pipeline per SM.
//
//     S   <- a_k B   + A S                     (identity columns -> E, state columns -> E x)
//     S_j <- a_k B_j + A S_j + G_j S           (jet of drive j: -d delta / d u_j)
// then one more product on the state columns for d/d dt = -A E x.  Results go from the accumulator registers straight
// to their COO positions (16 B per lane and column; the n_b copies of -E and, for iso generators [[P,-Q],[Q,P]], the
// mirrored half are written by the lane that holds the value).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "knot_dmma.cuh"

namespace pb2 {

struct DmmaqParams {
  int b, n_b, m, D, x_off, dt_off, u_off, nnz_jac, max_sub, nk;
  int ncT, iso, tiles;        // identity columns carried (0: residual only), iso layout, warps per knot
  int m_jets;                 // m, or 0 for the residual-only call
  int o_norm, o_sC, o_sY, ybuf;   // shared-memory layout (doubles)
  const double* Gfrag;
  const EllEntry* ell;        // m * Bp * W
  const double* norms;
  const double* Z;
  double* delta;
  double* jac;
  int mem_n;
  const int* x_offs;
  long long mem_gfrag, mem_ell, mem_norms, mem_delta, mem_jac;
};

template <int NT, int W>
__global__ void __launch_bounds__(32 * kDmmaMaxTiles, NT == 1 ? 2 : 1) knot_dmmaq_kernel(DmmaqParams p) {
  constexpr int KT = 2 * NT, Bp = 8 * NT, FR = KT * NT * 32;
  extern __shared__ __align__(16) double qs[];
  if (p.mem_n > 1) {
    const long long mi = blockIdx.y;
    p.Gfrag += mi * p.mem_gfrag;
    p.ell += mi * p.mem_ell;
    p.norms += mi * p.mem_norms;
    p.x_off = p.x_offs[mi];
    if (p.delta) p.delta += mi * p.mem_delta;
    if (p.jac) p.jac += mi * p.mem_jac;
  }
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int b = p.b, n_b = p.n_b, m = p.m, mj = p.m_jets, ncT = p.ncT;
  const int n_x = b * n_b, half = b >> 1, bb = b * b;

  // ---- once per CTA: fragment tables, norms ------------------------------------------------------------------
  {
    const int ntab = (m + 1) * FR;
    for (int e = tid; e < ntab; e += blockDim.x) qs[e] = p.Gfrag[e];
    for (int e = tid; e <= m; e += blockDim.x) qs[p.o_norm + e] = p.norms[e];
  }

  // ---- column bookkeeping: kind 0 padding, 1 state, 2 jet of drive jd, 3 identity column --------------------
  int kind = 0, cc = 0, jd = 0;
  {
    const int c = 8 * w + g;
    if (c < ncT) {
      kind = 3;
      cc = c;
    } else {
      const int rel = c - ncT, slab = rel / n_b;
      cc = rel - slab * n_b;
      if (slab == 0) kind = 1;
      else if (slab <= mj) { kind = 2; jd = slab - 1; }
    }
  }
  const uint32_t a_sY = smem_u32(qs + p.o_sY), ybytes = 8u * (uint32_t)p.ybuf;
  const uint32_t ypub = a_sY + 8u * (uint32_t)(cc * Bp + 2 * q);
  const bool pub = kind == 1 && mj > 0;
  const bool tile_cpl = __any_sync(0xffffffffu, kind == 2);
  const bool tile_x = __any_sync(0xffffffffu, kind == 1);

  double ev[KT][W];
  uint32_t yrd[KT][W];
  unsigned rowok = 0, diag = 0;
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    const int r = 8 * (i >> 1) + 2 * q + (i & 1);
    if (r < b && kind != 0) rowok |= 1u << i;
    if (kind == 3 && r == cc) diag |= 1u << i;
#pragma unroll
    for (int ww = 0; ww < W; ++ww) {
      EllEntry en{0.0, 0, 0};
      if (kind == 2 && r < b) en = p.ell[((size_t)jd * Bp + r) * W + ww];
      ev[i][ww] = en.val;
      yrd[i][ww] = a_sY + 8u * (uint32_t)((kind == 2 ? cc : 0) * Bp + en.idx);
    }
  }
  __syncthreads();

  const double th_max = c_theta[kMaxDeg];
  unsigned step = 0;

  for (int k = blockIdx.x; k < p.nk; k += gridDim.x) {
    const double* z = p.Z + (size_t)k * p.D;
    // ---- A(u) in B-fragment order ----------------------------------------------------------------------------
    double A[KT][NT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int e = (kt * NT + nt) * 32 + lane;
        double acc = qs[e];
        for (int j = 0; j < m; ++j) acc = fma(__ldg(z + p.u_off + j), qs[(1 + j) * FR + e], acc);
        A[kt][nt] = acc;
      }
    // ---- Taylor degree, sub-steps, coefficients (same arithmetic as knot_dmma) -------------------------------
    double dt = __ldg(z + p.dt_off);
    double nrm = qs[p.o_norm];
    for (int j = 0; j < m; ++j) nrm = fma(fabs(__ldg(z + p.u_off + j)), qs[p.o_norm + 1 + j], nrm);
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
        dt = __longlong_as_double(0x7ff8000000000000LL);
      }
    }
    int M = 1;
    for (int l = 1; l < kMaxDeg; ++l) M += c_theta[l] < per ? 1 : 0;
    if (tid <= kMaxDeg) {
      double pw = 1.0, sq = dt;
#pragma unroll
      for (int bit = 0; bit < 5; ++bit) {
        if ((tid >> bit) & 1) pw *= sq;
        sq *= sq;
      }
      qs[p.o_sC + tid] = tid <= M ? c_invfact[tid] * pw : 0.0;
    }
    // ---- B = [I | X | 0 ...]; the next knot's state for the residual -------------------------------------------
    double t[KT], base[KT], xn[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int r = 8 * (i >> 1) + 2 * q + (i & 1);
      const bool ok = (rowok >> i) & 1u;
      double v = ((diag >> i) & 1u) ? 1.0 : 0.0;
      xn[i] = 0.0;
      if (kind == 1 && ok) {
        v = __ldg(z + p.x_off + cc * b + r);
        xn[i] = __ldg(z + p.D + p.x_off + cc * b + r);
      }
      base[i] = v;
    }
    __syncthreads();   // coefficients visible

    for (int sub = 0; sub < n_sub; ++sub) {
      const double cM = qs[p.o_sC + M];
#pragma unroll
      for (int i = 0; i < KT; ++i) t[i] = cM * base[i];
      for (int kq = M - 1; kq >= 0; --kq) {
        const double ck = qs[p.o_sC + kq];
        const uint32_t par = (step & 1u) * ybytes;
        double cb[KT];
#pragma unroll
        for (int i = 0; i < KT; ++i) cb[i] = ck * base[i];
        if (mj > 0) {      // the jets read the state iterate of this step
          if (pub) {
#pragma unroll
            for (int i = 0; i < KT; ++i) sts_f64<0>(ypub + par + 8u * (8 * (i >> 1) + (i & 1)), t[i]);
          }
          __syncthreads();
          ++step;
        }
        double d[NT][2];
#pragma unroll
        for (int i = 0; i < KT; ++i) {
          double v = cb[i];
          if (tile_cpl) {
#pragma unroll
            for (int ww = 0; ww < W; ++ww) v = fma(ev[i][ww], lds_f64<0>(yrd[i][ww] + par), v);
          }
          d[i >> 1][i & 1] = v;
        }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(d[nt], t[kt], A[kt][nt]);
#pragma unroll
        for (int i = 0; i < KT; ++i) t[i] = d[i >> 1][i & 1];
      }
#pragma unroll
      for (int i = 0; i < KT; ++i) base[i] = t[i];
    }

    // ---- results, straight from the accumulators ------------------------------------------------------------------
    double* jac = p.jac ? p.jac + (size_t)k * p.nnz_jac : nullptr;
    if (kind == 3 && jac) {
#pragma unroll
      for (int i = 0; i < KT; ++i)
        if ((rowok >> i) & 1u) {
          const int r = 8 * (i >> 1) + 2 * q + (i & 1);
          const double v = -t[i];
          for (int copy = 0; copy < n_b; ++copy) {
            double* blk = jac + (size_t)copy * bb;
            blk[cc * b + r] = v;
            if (p.iso) {      // E = [[P, -Q], [Q, P]]: column half + c from column c
              const bool top = r < half;
              blk[(cc + half) * b + (top ? r + half : r - half)] = top ? v : -v;
            }
          }
        }
    }
    if (kind == 2 && jac) {
      double* blk = jac + (size_t)n_b * bb + (size_t)jd * n_x + cc * b;
#pragma unroll
      for (int i = 0; i < KT; ++i)
        if ((rowok >> i) & 1u) blk[8 * (i >> 1) + 2 * q + (i & 1)] = -t[i];
    }
    if (kind == 1 && p.delta) {
      double* dl = p.delta + (size_t)k * n_x + cc * b;
#pragma unroll
      for (int i = 0; i < KT; ++i)
        if ((rowok >> i) & 1u) dl[8 * (i >> 1) + 2 * q + (i & 1)] = xn[i] - t[i];
    }
    if (tile_x && jac) {
      // d/d dt = -A E x: one more product on the state tile; the constant d/dx_{k+1} entries
      double d[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) d[nt][0] = d[nt][1] = 0.0;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(d[nt], t[kt], A[kt][nt]);
      if (kind == 1) {
        double* blk = jac + (size_t)n_b * bb + (size_t)mj * n_x + cc * b;
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if ((rowok >> i) & 1u) {
            const int r = 8 * (i >> 1) + 2 * q + (i & 1);
            blk[r] = -d[i >> 1][i & 1];
            blk[n_x + r] = 1.0;
          }
      }
    }
    __syncthreads();   // the coefficient table is rewritten by the next knot
  }
}

// shared memory (doubles): [tables | norms | coefficients | state exchange x2]
inline size_t dmmaq_layout(DmmaqParams& q, int NT) {
  auto even = [](int v) { return (v + 1) & ~1; };
  const int KT = 2 * NT, FR = KT * NT * 32, Bp = 8 * NT;
  q.o_norm = (q.m + 1) * FR;
  q.o_sC = q.o_norm + even(q.m + 1);
  q.o_sY = q.o_sC + even(kMaxDeg + 1);
  q.ybuf = q.n_b * Bp;
  return sizeof(double) * (size_t)(q.o_sY + 2 * q.ybuf);
}

using DmmaqKernel = void (*)(DmmaqParams);

inline DmmaqKernel dmmaq_kernel(int NT, int W) {
  if (NT == 3) return W == 1 ? knot_dmmaq_kernel<3, 1> : (W == 2 ? knot_dmmaq_kernel<3, 2> : knot_dmmaq_kernel<3, 4>);
  if (NT == 1) return W == 1 ? knot_dmmaq_kernel<1, 1> : (W == 2 ? knot_dmmaq_kernel<1, 2> : knot_dmmaq_kernel<1, 4>);
  return W == 1 ? knot_dmmaq_kernel<2, 1> : (W == 2 ? knot_dmmaq_kernel<2, 2> : knot_dmmaq_kernel<2, 4>);
}

}  // namespace pb2
