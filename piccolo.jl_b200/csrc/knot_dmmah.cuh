// Tensor-core Lagrangian Hessian for general generators with b <= 24.
//
// What it replaces: the Hessian of sum_k mu_k . delta_k that DirectTrajOpt's BilinearIntegrator hands to Ipopt's
// eval_h (constructed at /root/reference/src/control/integrators.jl:35-95; `hessian_structure`, test/aqua.jl:6-9),
// for every shape the 3-qubit kernel (knot_u8h.cuh) does not take: C1, C2, C4, kets, compact Lindbladians.
//
// With phi(x, u, dt) = <M, exp(dt A(u)) X>,  M = reshape(mu_k),  A(u) = G0 + sum_j u_j G_j,  the blocks are
//     (x, u_j)   = -F_j^T M            F_j = d/du_j exp(dt A)
//     (x, dt)    = -(A E)^T M
//     (u_i, u_j) = -<M, E_ij X>        second derivative of the propagator applied to the state
//     (u_j, dt)  = -<M, (G_j E + A F_j) X>
//     (dt, dt)   = -<M, A^2 E X>
// ExponentialAction's truncated-Taylor action gives all of them from ONE Horner recurrence run on stacked columns
// (a_k = dt'^k / k!, k = M-1 .. 0, n_sub sub-steps):
//   forward, generator A:    S    <- a_k B    + A S                              (state columns X)
//                            S_j  <- a_k B_j  + A S_j  + G_j S                    (first-order jets)
//                            S_ij <- a_k B_ij + A S_ij + G_i S_j + G_j S_i        (second-order jets, i <= j)
//   adjoint, generator A^T:  V    <- a_k C    + A^T V                             (started from M:  E^T M)
//                            V_j  <- a_k C_j  + A^T V_j + G_j^T V                 (F_j^T M)
// No symmetry of the generators is assumed (the compact Lindbladian of C4 has none): the adjoint columns use the
// transposed fragment tables.  After the recurrence one more product gives A^T V and Q_j = A S_j + G_j S, a second
// one A^2 E X; the scalar blocks are reductions of mu (.) tile, summed in a fixed order (bitwise reproducible).
//
// Work decomposition: one CTA per knot (several resident per SM; a grid-stride loop when there are more knots than
// resident CTAs), ONE WARP per 8-column tile holding its transposed tile as DMMA.8x8x4 accumulators exactly as in
// knot_dmma.cuh; the published iterates (X, S_j, V) go through a double-buffered exchange area in shared memory,
// one CTA barrier per Horner step; the sparse couplings G_j S are ELL entries held in registers.  Members of an
// ensemble (pb2_batch_*) are blockIdx.y with their own tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "knot_dmma.cuh"

namespace pb2 {

struct DmmahParams {
  int b, n_b, m, D, x_off, dt_off, u_off, nnz_hess, max_sub, nk;
  int tiles_f, tiles_a;                 // forward / adjoint warps of a knot
  int o_tabT, o_norm, o_sC, o_sY, ybuf, o_part;   // shared-memory layout (doubles): see dmmah_layout
  const double* Gfrag;    // (m+1) * FR, B-fragment order of G0, G_j
  const double* GfragT;   // the same for the transposed matrices
  const EllEntry* ell;    // m * Bp * W   rows of G_j
  const EllEntry* ellT;   // m * Bp * W   rows of G_j^T
  const double* norms;    // m+1
  const double* Z;
  const double* mu;
  double* hess;
  // ensemble launch (blockIdx.y = member)
  int mem_n;
  const int* x_offs;
  long long mem_gfrag, mem_ell, mem_norms, mem_mu, mem_hess;
};

constexpr int kDmmahMaxWarps = 12;
// the 3-qubit unitary shape (15 forward + 5 adjoint tiles) fits as 20 warps when the couplings are one entry wide
// ... and the widest variant (24 x 24 generators, four coupling entries per row) keeps 48 + 48 coupling values and
// addresses per lane: compiled for 12 warps (170 registers) it spills into the step loop (ncu: long-scoreboard stalls,
// 324 us for a qutrit-pair ket problem); compiled for 8 warps (255 registers) it does not (74 us).  Knots that need
// 9 - 12 warps of that variant still take the spilling build, which beats the jet kernel.
__host__ __device__ constexpr int dmmah_max_warps(int NT, int W) { return (NT == 2 && W == 1) ? 20 : kDmmahMaxWarps; }

template <int NT, int W, int MAXW = dmmah_max_warps(NT, W)>
__global__ void __launch_bounds__(32 * MAXW, NT == 1 ? 2 : 1) knot_dmmah_kernel(DmmahParams p) {
  constexpr int KT = 2 * NT, Bp = 8 * NT, FR = KT * NT * 32, W2 = 2 * W;
  extern __shared__ __align__(16) double hs[];
  if (p.mem_n > 1) {
    const long long mi = blockIdx.y;
    p.Gfrag += mi * p.mem_gfrag;
    p.GfragT += mi * p.mem_gfrag;
    p.ell += mi * p.mem_ell;
    p.ellT += mi * p.mem_ell;
    p.norms += mi * p.mem_norms;
    p.x_off = p.x_offs[mi];
    p.mu += mi * p.mem_mu;
    p.hess += mi * p.mem_hess;
  }
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int b = p.b, n_b = p.n_b, m = p.m, n_x = b * n_b;
  const int npair = m * (m + 1) / 2, nscal = npair + m + 1;
  const bool fwd = w < p.tiles_f;

  // ---- once per CTA: both fragment tables and the norms ----------------------------------------------------
  {
    const int ntab = (m + 1) * FR;
    for (int e = tid; e < ntab; e += blockDim.x) {
      hs[e] = p.Gfrag[e];
      hs[p.o_tabT + e] = p.GfragT[e];
    }
    for (int e = tid; e <= m; e += blockDim.x) hs[p.o_norm + e] = p.norms[e];
  }

  // ---- column bookkeeping -----------------------------------------------------------------------------------
  // kind 0: padding   1: state X   2: first-order jet S_ja   3: second-order jet S_(pi,pj)
  //      4: adjoint V (starts from mu)   5: adjoint jet V_ja
  int kind = 0, cc = 0, ja = 0, pi = 0, pj = 0, fcol = 0;
  {
    const int c = 8 * (fwd ? w : w - p.tiles_f) + g;
    const int slab = c / n_b;
    cc = c - slab * n_b;
    fcol = c;
    if (fwd) {
      if (slab == 0) kind = 1;
      else if (slab <= m) { kind = 2; ja = slab - 1; }
      else if (slab < 1 + m + npair) {
        kind = 3;
        const int pidx = slab - 1 - m;
        while ((pj + 1) * (pj + 2) / 2 <= pidx) ++pj;
        pi = pidx - pj * (pj + 1) / 2;
      }
    } else {
      if (slab == 0) kind = 4;
      else if (slab <= m) { kind = 5; ja = slab - 1; }
    }
  }
  // published columns of the exchange area: X at cc, S_j at n_b (1 + j) + cc, V at n_b (1 + m) + cc
  const int pcol = kind == 1 ? cc : (kind == 2 ? n_b * (1 + ja) + cc : (kind == 4 ? n_b * (1 + m) + cc : -1));
  const uint32_t a_sY = smem_u32(hs + p.o_sY), ybytes = 8u * (uint32_t)p.ybuf;
  const uint32_t ypub = a_sY + 8u * (uint32_t)((pcol < 0 ? 0 : pcol) * Bp + 2 * q);
  const bool tile_cpl = __any_sync(0xffffffffu, kind == 2 || kind == 3 || kind == 5);

  double ev[KT][W2];
  uint32_t yrd[KT][W2];
  unsigned rowok = 0;
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    const int r = 8 * (i >> 1) + 2 * q + (i & 1);
    if (r < b && kind != 0) rowok |= 1u << i;
#pragma unroll
    for (int ww = 0; ww < W2; ++ww) {
      EllEntry en{0.0, 0, 0};
      int parent = 0;
      if (r < b) {
        if (kind == 2 && ww < W) { en = p.ell[((size_t)ja * Bp + r) * W + ww]; parent = cc; }
        if (kind == 5 && ww < W) { en = p.ellT[((size_t)ja * Bp + r) * W + ww]; parent = n_b * (1 + m) + cc; }
        if (kind == 3) {
          if (ww < W) {                      // G_pi S_pj  (twice when pi == pj)
            en = p.ell[((size_t)pi * Bp + r) * W + ww];
            if (pi == pj) en.val *= 2.0;
            parent = n_b * (1 + pj) + cc;
          } else if (pi != pj) {             // G_pj S_pi
            en = p.ell[((size_t)pj * Bp + r) * W + (ww - W)];
            parent = n_b * (1 + pi) + cc;
          }
        }
      }
      ev[i][ww] = en.val;
      yrd[i][ww] = a_sY + 8u * (uint32_t)(parent * Bp + en.idx);
    }
  }
  __syncthreads();

  const double* tab = fwd ? hs : hs + p.o_tabT;
  const double th_max = c_theta[kMaxDeg];
  unsigned step = 0;   // parity of the exchange buffer: one barrier per published step, alternating halves

  for (int k = blockIdx.x; k < p.nk; k += gridDim.x) {
    const double* z = p.Z + (size_t)k * p.D;
    const double* muk = p.mu + (size_t)k * n_x;
    // ---- A(u) (or its transpose) in B-fragment order, straight into registers ------------------------------
    double A[KT][NT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int e = (kt * NT + nt) * 32 + lane;
        double acc = tab[e];
        for (int j = 0; j < m; ++j) acc = fma(__ldg(z + p.u_off + j), tab[(1 + j) * FR + e], acc);
        A[kt][nt] = acc;
      }
    // ---- Taylor degree, sub-steps (every thread, same arithmetic as knot_dmma) ---------------------------------
    double dt = __ldg(z + p.dt_off);
    double nrm = hs[p.o_norm];
    for (int j = 0; j < m; ++j) nrm = fma(fabs(__ldg(z + p.u_off + j)), hs[p.o_norm + 1 + j], nrm);
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
      hs[p.o_sC + tid] = tid <= M ? c_invfact[tid] * pw : 0.0;
    }
    // ---- start values: X, mu --------------------------------------------------------------------------------
    double t[KT], base[KT], mur[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int r = 8 * (i >> 1) + 2 * q + (i & 1);
      const bool ok = (rowok >> i) & 1u;
      mur[i] = ok ? __ldg(muk + cc * b + r) : 0.0;
      double v = 0.0;
      if (kind == 1 && ok) v = __ldg(z + p.x_off + cc * b + r);
      if (kind == 4) v = mur[i];
      base[i] = v;
    }
    __syncthreads();   // coefficients visible

    for (int sub = 0; sub < n_sub; ++sub) {
      const double cM = hs[p.o_sC + M];
#pragma unroll
      for (int i = 0; i < KT; ++i) t[i] = cM * base[i];
      for (int kq = M - 1; kq >= 0; --kq) {
        const double ck = hs[p.o_sC + kq];
        const uint32_t par = (step & 1u) * ybytes;
        // a_k B before the barrier.  (Splitting the coupling sum into two shorter chains, or skipping the second
        // parent's terms on tiles without mixed second-order lanes, was measured: no gain / 8 % slower on C4.)
        double cb[KT];
#pragma unroll
        for (int i = 0; i < KT; ++i) cb[i] = ck * base[i];
        if (pcol >= 0) {
#pragma unroll
          for (int i = 0; i < KT; ++i) sts_f64<0>(ypub + par + 8u * (8 * (i >> 1) + (i & 1)), t[i]);
        }
        __syncthreads();
        ++step;
        double d[NT][2];
#pragma unroll
        for (int i = 0; i < KT; ++i) {
          double v = cb[i];
          if (tile_cpl) {
#pragma unroll
            for (int ww = 0; ww < W2; ++ww) v = fma(ev[i][ww], lds_f64<0>(yrd[i][ww] + par), v);
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

    // ---- one more product: X -> A Y, S_j -> Q_j = A S_j + G_j Y, V -> A^T V; the others keep their values ----
    {
      const uint32_t par = (step & 1u) * ybytes;
      if (pcol >= 0) {
#pragma unroll
        for (int i = 0; i < KT; ++i) sts_f64<0>(ypub + par + 8u * (8 * (i >> 1) + (i & 1)), t[i]);
      }
      __syncthreads();
      ++step;
      double d[NT][2];
#pragma unroll
      for (int i = 0; i < KT; ++i) {
        double v = 0.0;
        if (tile_cpl) {
#pragma unroll
          for (int ww = 0; ww < W2; ++ww) v = fma(ev[i][ww], lds_f64<0>(yrd[i][ww] + par), v);
        }
        d[i >> 1][i & 1] = v;
      }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(d[nt], t[kt], A[kt][nt]);
      const bool take = kind == 1 || kind == 2 || kind == 4;
#pragma unroll
      for (int i = 0; i < KT; ++i) t[i] = take ? d[i >> 1][i & 1] : t[i];
    }
    // ---- and a second one for the state columns: A^2 Y -----------------------------------------------------
    if (__any_sync(0xffffffffu, kind == 1)) {
      double d[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) d[nt][0] = d[nt][1] = 0.0;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(d[nt], t[kt], A[kt][nt]);
#pragma unroll
      for (int i = 0; i < KT; ++i) t[i] = kind == 1 ? d[i >> 1][i & 1] : t[i];
    }

    // ---- results ---------------------------------------------------------------------------------------------
    double* out = p.hess + (size_t)k * p.nnz_hess;
    if (fwd) {
      if (kind != 0) {
        double part = 0.0;
#pragma unroll
        for (int i = 0; i < KT; ++i) part = fma(mur[i], t[i], part);
        hs[p.o_part + fcol * 4 + q] = part;
      }
    } else if (kind != 0) {
      // (x, u_j) = -F_j^T M from the adjoint jets, (x, dt) = -A^T E^T M from the adjoint state columns
      double* blk = out + (size_t)(kind == 5 ? ja : m) * n_x + cc * b;
#pragma unroll
      for (int i = 0; i < KT; ++i)
        if ((rowok >> i) & 1u) blk[8 * (i >> 1) + 2 * q + (i & 1)] = -t[i];
    }
    __syncthreads();
    if (tid < nscal) {
      // (u_i, u_j) <- second-order slabs, (u_j, dt) <- first-order slabs (Q_j), (dt, dt) <- the state slab
      const int slab = tid < npair ? 1 + m + tid : (tid < npair + m ? 1 + (tid - npair) : 0);
      double acc = 0.0;
      for (int e = 0; e < 4 * n_b; ++e) acc += hs[p.o_part + slab * n_b * 4 + e];
      out[(size_t)(m + 1) * n_x + tid] = -acc;
    }
    __syncthreads();   // the coefficient table and the partial sums are rewritten by the next knot
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct DmmahPlan {
  bool ok = false;
  int NT = 0, Bp = 0, W = 1, tiles_f = 0, tiles_a = 0;
  std::vector<double> gfrag, gfragT, norms;
  std::vector<EllEntry> ell, ellT;
};

// G0, Gj: host, column-major b x b
inline DmmahPlan dmmah_plan(int b, int n_b, int m, const double* G0, const double* Gj) {
  DmmahPlan pl;
  if (b < 1 || b > 24 || n_b < 1) return pl;
  pl.NT = b <= 8 ? 1 : (b <= 16 ? 2 : 3);
  pl.Bp = 8 * pl.NT;
  const int KT = 2 * pl.NT, NT = pl.NT, Bp = pl.Bp, npair = m * (m + 1) / 2;
  pl.tiles_f = (n_b * (1 + m + npair) + 7) / 8;
  pl.tiles_a = (n_b * (1 + m) + 7) / 8;
  auto at = [&](int mat, int r, int c) -> double {
    if (r >= b || c >= b) return 0.0;
    const double* A = mat == 0 ? G0 : Gj + (size_t)(mat - 1) * b * b;
    return A[r + (size_t)c * b];
  };
  int W = 1;
  for (int j = 0; j < m; ++j)
    for (int r = 0; r < b; ++r) {
      int cr = 0, ct = 0;
      for (int c = 0; c < b; ++c) {
        cr += at(1 + j, r, c) != 0.0;
        ct += at(1 + j, c, r) != 0.0;
      }
      W = std::max(W, std::max(cr, ct));
    }
  if (W > 4) return pl;
  W = W <= 1 ? 1 : (W <= 2 ? 2 : 4);
  pl.W = W;
  if (pl.tiles_f + pl.tiles_a > dmmah_max_warps(pl.NT, W) || npair + m + 1 > 32 * (pl.tiles_f + pl.tiles_a)) return pl;
  const size_t FR = (size_t)KT * NT * 32;
  pl.gfrag.assign((m + 1) * FR, 0.0);
  pl.gfragT.assign((m + 1) * FR, 0.0);
  for (int mat = 0; mat <= m; ++mat)
    for (int kt = 0; kt < KT; ++kt)
      for (int nt = 0; nt < NT; ++nt)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, q = lane & 3;
          const size_t e = ((size_t)mat * KT * NT + kt * NT + nt) * 32 + lane;
          pl.gfrag[e] = at(mat, 8 * nt + g, dmma_perm(kt, q));
          pl.gfragT[e] = at(mat, dmma_perm(kt, q), 8 * nt + g);
        }
  pl.ell.assign((size_t)std::max(m, 1) * Bp * W, EllEntry{0.0, 0, 0});
  pl.ellT.assign((size_t)std::max(m, 1) * Bp * W, EllEntry{0.0, 0, 0});
  for (int j = 0; j < m; ++j)
    for (int r = 0; r < b; ++r) {
      int w = 0, wt = 0;
      for (int c = 0; c < b; ++c) {
        if (at(1 + j, r, c) != 0.0) pl.ell[((size_t)j * Bp + r) * W + w++] = EllEntry{at(1 + j, r, c), c, 0};
        if (at(1 + j, c, r) != 0.0) pl.ellT[((size_t)j * Bp + r) * W + wt++] = EllEntry{at(1 + j, c, r), c, 0};
      }
    }
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

// shared memory (doubles): [tables | transposed tables | norms | coefficients | exchange x2 | partial sums]
inline size_t dmmah_layout(DmmahParams& q, int NT) {
  auto even = [](int v) { return (v + 1) & ~1; };
  const int KT = 2 * NT, FR = KT * NT * 32, Bp = 8 * NT, npair = q.m * (q.m + 1) / 2;
  q.o_tabT = (q.m + 1) * FR;
  q.o_norm = 2 * q.o_tabT;
  q.o_sC = q.o_norm + even(q.m + 1);
  q.o_sY = q.o_sC + even(kMaxDeg + 1);
  q.ybuf = q.n_b * (2 + q.m) * Bp;
  q.o_part = q.o_sY + 2 * q.ybuf;
  return sizeof(double) * (size_t)(q.o_part + 4 * 8 * q.tiles_f + even(npair + q.m + 1));
}

using DmmahKernel = void (*)(DmmahParams);

inline DmmahKernel dmmah_kernel(int NT, int W, int warps) {
  if (NT == 3 && W == 4 && warps <= 8) return knot_dmmah_kernel<3, 4, 8>;
  if (NT == 3) return W == 1 ? knot_dmmah_kernel<3, 1> : (W == 2 ? knot_dmmah_kernel<3, 2> : knot_dmmah_kernel<3, 4>);
  if (NT == 1) return W == 1 ? knot_dmmah_kernel<1, 1> : (W == 2 ? knot_dmmah_kernel<1, 2> : knot_dmmah_kernel<1, 4>);
  return W == 1 ? knot_dmmah_kernel<2, 1> : (W == 2 ? knot_dmmah_kernel<2, 2> : knot_dmmah_kernel<2, 4>);
}

}  // namespace pb2
