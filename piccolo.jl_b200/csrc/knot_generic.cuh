// Generic knot kernels: any real b x b generator (ket, unitary and the non-normal compact
// Lindbladian).  One thread group per knot; exp(dt G(u)) and its first / second Frechet
// derivatives in the drive directions are obtained by pushing truncated-Taylor "jets"
// (value, d/du_j, d2/du_i du_j) through a scaling-and-squaring recurrence, all slabs
// resident in shared memory.
//
// Replaces, per knot, what DirectTrajOpt's BilinearIntegrator does in the reference's Ipopt
// callbacks (constructed at /root/reference/src/control/integrators.jl:35-95): assemble
// Ghat(u) (quantum_systems.jl:226 / open_quantum_systems.jl:607-636), expv, ForwardDiff duals.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb2 {

struct KnotParams {
  int b, n_b, m, K, D, x_off, dt_off, u_off;
  int nnz_jac, nnz_hess;
  const double* G0;   // device, b*b col-major
  const double* Gj;   // device, m*b*b
  const double* Z;    // device, D*K
  double* delta;      // device or null
  double* jac;        // device or null
  const double* mu;   // device (hessian only)
  double* hess;       // device (hessian only)
  // batched launch (pb2_batch_*: the members of a SamplingTrajectory ensemble, blockIdx.y = member): every member
  // has its own generator factors and state block, all read the same trajectory
  int mem_n;                    // 0 / 1: not batched
  const int* x_offs;            // [mem_n] state-block row of each member
  long long mem_G0, mem_Gj;     // strides (doubles) between the members' generator factors
  long long mem_delta, mem_jac, mem_hess;   // strides between the members' outputs (mem_delta also strides mu)
};

// batched launches: shift the member-dependent fields of a by-value parameter block
template <class P>
__device__ __forceinline__ void apply_member(P& p) {
  if (p.mem_n > 1) {
    const long long mi = blockIdx.y;
    p.x_off = p.x_offs[mi];
    if (p.delta) p.delta += mi * p.mem_delta;
    if (p.jac) p.jac += mi * p.mem_jac;
  }
}

constexpr int kMaxDeg = 18;
// c_theta[q]: largest ||A||_1 for which the degree-q Taylor polynomial of exp is accurate
// to 2^-53 (forward bound theta^(q+1)/(q+1)! / (1 - theta/(q+2))).  Filled at library init.
__constant__ double c_theta[kMaxDeg + 1];

// choose Taylor degree M and squarings s minimising M + s  subject to nrm / 2^s <= theta_M
__device__ inline void pick_degree(double nrm, int& M, int& s) {
  M = kMaxDeg;
  s = 0;
  if (!(nrm > 0.0) || !(nrm < 1e300)) {  // zero, NaN or inf: nothing sensible to choose
    M = 2;
    return;
  }
  int best = 1 << 30;
  for (int q = 2; q <= kMaxDeg; ++q) {
    int sq = 0;
    if (nrm > c_theta[q]) {
      sq = (int)ceil(log2(nrm / c_theta[q]));
      if (sq < 0) sq = 0;
    }
    if (q + sq < best) {
      best = q + sq;
      M = q;
      s = sq;
    }
  }
}

__device__ inline int pair_index(int i, int j) { return j * (j + 1) / 2 + i; }  // i <= j

// ORDER = 1: residual + Jacobian values.   ORDER = 2: Lagrangian-Hessian values.
// Block = NT threads = KPC groups of GS threads; group g of block blk owns knot blk*KPC+g.
template <int ORDER, int NT>
__global__ void __launch_bounds__(NT) knot_generic_kernel(KnotParams p, int GS, int KPC,
                                                          int gj_in_smem) {
  extern __shared__ double sm[];
  __shared__ int s_M, s_s;

  if (p.mem_n > 1) {
    const long long mi = blockIdx.y;
    p.G0 += mi * p.mem_G0;
    p.Gj += mi * p.mem_Gj;
    if (p.mu) p.mu += mi * p.mem_delta;
    if (p.hess) p.hess += mi * p.mem_hess;
  }
  apply_member(p);
  const int b = p.b, bb = b * b, n_b = p.n_b, n_x = b * n_b, m = p.m;
  const int npair = (ORDER == 2) ? m * (m + 1) / 2 : 0;
  const int S = 1 + m + npair;                 // jet slabs
  const int nvec = (ORDER == 2) ? (5 + m) : 2; // small b x n_b vectors
  const int nscal = npair + m + 1;
  const int nwarp_g = (GS + 31) / 32;

  const int tid = threadIdx.x;
  const int g = tid / GS, t = tid % GS;
  const bool in_group = g < KPC;
  const int k = blockIdx.x * KPC + g;
  const bool active = in_group && (k < p.K - 1);

  // ---- shared-memory carve-up --------------------------------------------------------
  double* smGj = sm;
  const size_t per_group = (size_t)bb * (1 + 2 * S) + (size_t)nvec * n_x +
                           (ORDER == 2 ? (size_t)(nwarp_g + 1) * nscal : 0) + b;
  double* base = sm + (gj_in_smem ? (size_t)m * bb : 0) + (size_t)(in_group ? g : 0) * per_group;
  double* G = base;                // bb
  double* Ta = G + bb;             // S * bb
  double* Tb = Ta + (size_t)S * bb;
  double* vec = Tb + (size_t)S * bb;  // nvec * n_x
  double* colsum = vec + (size_t)nvec * n_x;  // b
  double* red = colsum + b;        // (nwarp_g + 1) * nscal   (ORDER 2)
  const double* Gd = gj_in_smem ? smGj : p.Gj;

  if (gj_in_smem)
    for (int e = tid; e < m * bb; e += NT) smGj[e] = p.Gj[e];

  const double* z = p.Z + (size_t)(active ? k : 0) * p.D;
  const double dt = active ? z[p.dt_off] : 0.0;

  // ---- G(u) = G0 + sum_j u_j G_j ; ||dt G||_1 ----------------------------------------
  if (in_group) {
    for (int e = t; e < bb; e += GS) {
      double acc = 0.0;
      if (active) {
        acc = p.G0[e];
        for (int j = 0; j < m; ++j) acc = fma(z[p.u_off + j], p.Gj[(size_t)j * bb + e], acc);
      }
      G[e] = acc;
    }
  }
  __syncthreads();
  if (in_group) {
    for (int c = t; c < b; c += GS) {
      double acc = 0.0;
      for (int i = 0; i < b; ++i) acc += fabs(G[i + c * b]);
      colsum[c] = acc * fabs(dt);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double nrm = 0.0;
    bool bad = false;
    for (int gg = 0; gg < KPC; ++gg) {
      const double* cs = sm + (gj_in_smem ? (size_t)m * bb : 0) + (size_t)gg * per_group +
                         (size_t)bb * (1 + 2 * S) + (size_t)nvec * n_x;
      for (int c = 0; c < b; ++c) {
        if (!(cs[c] == cs[c])) bad = true;
        nrm = fmax(nrm, cs[c]);
      }
    }
    int M, s;
    pick_degree(bad ? 0.0 : nrm, M, s);
    if (s > 64) s = 64;
    s_M = M;
    s_s = s;
  }
  __syncthreads();
  const int M = s_M, s = s_s;
  const double scale = ldexp(dt, -s);  // A' = scale * G

  // ---- Horner in jet arithmetic: T <- I + (A' T)/q,  q = M .. 1 ------------------------
  double* Tc = Ta;  // current
  double* Tn = Tb;  // next
  if (in_group)
    for (int e = t; e < S * bb; e += GS) Tc[e] = (e < bb && (e % b) == (e / b)) ? 1.0 : 0.0;
  __syncthreads();
  for (int q = M; q >= 1; --q) {
    const double c = scale / (double)q;
    if (in_group) {
      for (int w = t; w < S * bb; w += GS) {
        const int slab = w / bb, e = w - slab * bb;
        const int i = e % b, j = e / b;
        double acc = 0.0;
        if (slab == 0) {
          for (int r = 0; r < b; ++r) acc = fma(G[i + r * b], Tc[r + j * b], acc);
          acc = fma(c, acc, (i == j) ? 1.0 : 0.0);
        } else if (slab <= m) {
          const double* Gs = Gd + (size_t)(slab - 1) * bb;
          const double* Ts = Tc + (size_t)slab * bb;
          for (int r = 0; r < b; ++r) {
            acc = fma(G[i + r * b], Ts[r + j * b], acc);
            acc = fma(Gs[i + r * b], Tc[r + j * b], acc);
          }
          acc *= c;
        } else {
          // pair (pi <= pj)
          int pr = slab - 1 - m, pj = 0;
          while ((pj + 1) * (pj + 2) / 2 <= pr) ++pj;
          const int pi = pr - pj * (pj + 1) / 2;
          const double* Gi = Gd + (size_t)pi * bb;
          const double* Gjj = Gd + (size_t)pj * bb;
          const double* Ti = Tc + (size_t)(1 + pi) * bb;
          const double* Tj = Tc + (size_t)(1 + pj) * bb;
          const double* Ts = Tc + (size_t)slab * bb;
          for (int r = 0; r < b; ++r) {
            acc = fma(G[i + r * b], Ts[r + j * b], acc);
            acc = fma(Gi[i + r * b], Tj[r + j * b], acc);
            acc = fma(Gjj[i + r * b], Ti[r + j * b], acc);
          }
          acc *= c;
        }
        Tn[w] = acc;
      }
    }
    __syncthreads();
    double* tmp = Tc; Tc = Tn; Tn = tmp;
  }
  // ---- squarings -----------------------------------------------------------------------
  for (int sq = 0; sq < s; ++sq) {
    if (in_group) {
      for (int w = t; w < S * bb; w += GS) {
        const int slab = w / bb, e = w - slab * bb;
        const int i = e % b, j = e / b;
        double acc = 0.0;
        if (slab == 0) {
          for (int r = 0; r < b; ++r) acc = fma(Tc[i + r * b], Tc[r + j * b], acc);
        } else if (slab <= m) {
          const double* Ts = Tc + (size_t)slab * bb;
          for (int r = 0; r < b; ++r) {
            acc = fma(Tc[i + r * b], Ts[r + j * b], acc);
            acc = fma(Ts[i + r * b], Tc[r + j * b], acc);
          }
        } else {
          int pr = slab - 1 - m, pj = 0;
          while ((pj + 1) * (pj + 2) / 2 <= pr) ++pj;
          const int pi = pr - pj * (pj + 1) / 2;
          const double* Ti = Tc + (size_t)(1 + pi) * bb;
          const double* Tj = Tc + (size_t)(1 + pj) * bb;
          const double* Ts = Tc + (size_t)slab * bb;
          for (int r = 0; r < b; ++r) {
            acc = fma(Tc[i + r * b], Ts[r + j * b], acc);
            acc = fma(Ts[i + r * b], Tc[r + j * b], acc);
            acc = fma(Ti[i + r * b], Tj[r + j * b], acc);
            acc = fma(Tj[i + r * b], Ti[r + j * b], acc);
          }
        }
        Tn[w] = acc;
      }
    }
    __syncthreads();
    double* tmp = Tc; Tc = Tn; Tn = tmp;
  }
  // Tc: slab 0 = E, slab 1+j = F_j = L(A; dt G_j), slab 1+m+pair = L2(A; dt G_i, dt G_j)
  const double* E = Tc;

  double* X = vec;            // x_k as b x n_b
  double* Y0 = vec + n_x;     // E X
  if (active)
    for (int e = t; e < n_x; e += GS) X[e] = z[p.x_off + e];
  __syncthreads();
  if (active) {
    for (int e = t; e < n_x; e += GS) {
      const int i = e % b, c = e / b;
      double acc = 0.0;
      for (int r = 0; r < b; ++r) acc = fma(E[i + r * b], X[r + c * b], acc);
      Y0[e] = acc;
    }
  }
  __syncthreads();

  if (ORDER == 1) {
    if (active) {
      if (p.delta) {
        const double* zn = z + p.D;
        double* out = p.delta + (size_t)k * n_x;
        for (int e = t; e < n_x; e += GS) out[e] = zn[p.x_off + e] - Y0[e];
      }
      if (p.jac) {
        double* out = p.jac + (size_t)k * p.nnz_jac;
        for (int w = t; w < n_b * bb; w += GS) out[w] = -E[w % bb];
        out += (size_t)n_b * bb;
        for (int w = t; w < m * n_x; w += GS) {
          const int j = w / n_x, e = w - j * n_x;
          const int i = e % b, c = e / b;
          const double* F = Tc + (size_t)(1 + j) * bb;
          double acc = 0.0;
          for (int r = 0; r < b; ++r) acc = fma(F[i + r * b], X[r + c * b], acc);
          out[w] = -acc;
        }
        out += (size_t)m * n_x;
        for (int e = t; e < n_x; e += GS) {
          const int i = e % b, c = e / b;
          double acc = 0.0;
          for (int r = 0; r < b; ++r) acc = fma(G[i + r * b], Y0[r + c * b], acc);
          out[e] = -acc;
        }
        out += n_x;
        for (int e = t; e < n_x; e += GS) out[e] = 1.0;
      }
    }
    return;
  }

  // ---- ORDER 2: Hessian of mu . delta ------------------------------------------------------
  double* Mu = vec + 2 * (size_t)n_x;
  double* W = vec + 3 * (size_t)n_x;    // G^T Mu
  double* GY = vec + 4 * (size_t)n_x;   // G Y0
  double* Yj = vec + 5 * (size_t)n_x;   // F_j X, j = 0..m-1
  if (active)
    for (int e = t; e < n_x; e += GS) Mu[e] = p.mu[(size_t)k * n_x + e];
  __syncthreads();
  if (active) {
    for (int w = t; w < (2 + m) * n_x; w += GS) {
      const int which = w / n_x, e = w - which * n_x;
      const int i = e % b, c = e / b;
      double acc = 0.0;
      if (which == 0) {
        for (int r = 0; r < b; ++r) acc = fma(G[r + i * b], Mu[r + c * b], acc);
        W[e] = acc;
      } else if (which == 1) {
        for (int r = 0; r < b; ++r) acc = fma(G[i + r * b], Y0[r + c * b], acc);
        GY[e] = acc;
      } else {
        const double* F = Tc + (size_t)(which - 1) * bb;
        for (int r = 0; r < b; ++r) acc = fma(F[i + r * b], X[r + c * b], acc);
        Yj[(size_t)(which - 2) * n_x + e] = acc;
      }
    }
  }
  __syncthreads();
  double* out = p.hess + (size_t)(active ? k : 0) * p.nnz_hess;
  if (active) {
    // (x, u_j) = -F_j^T Mu ;  (x, dt) = -E^T W
    for (int w = t; w < (m + 1) * n_x; w += GS) {
      const int j = w / n_x, e = w - j * n_x;
      const int i = e % b, c = e / b;
      double acc = 0.0;
      if (j < m) {
        const double* F = Tc + (size_t)(1 + j) * bb;
        for (int r = 0; r < b; ++r) acc = fma(F[r + i * b], Mu[r + c * b], acc);
      } else {
        for (int r = 0; r < b; ++r) acc = fma(E[r + i * b], W[r + c * b], acc);
      }
      out[w] = -acc;
    }
  }
  // scalar blocks: per-thread partials -> warp shuffle -> fixed-order sum over warps
  const int wig = t / 32, lane = t % 32;
  for (int sidx = 0; sidx < nscal; ++sidx) {
    double part = 0.0;
    if (active) {
      if (sidx < npair) {
        const double* L2 = Tc + (size_t)(1 + m + sidx) * bb;
        for (int e = t; e < n_x; e += GS) {
          const int i = e % b, c = e / b;
          double acc = 0.0;
          for (int r = 0; r < b; ++r) acc = fma(L2[i + r * b], X[r + c * b], acc);
          part = fma(Mu[e], acc, part);
        }
      } else if (sidx < npair + m) {
        const int j = sidx - npair;
        const double* Gs = Gd + (size_t)j * bb;
        for (int e = t; e < n_x; e += GS) {
          const int i = e % b, c = e / b;
          double acc = 0.0;
          for (int r = 0; r < b; ++r) acc = fma(Gs[i + r * b], Y0[r + c * b], acc);
          part = fma(Mu[e], acc, part);
          part = fma(W[e], Yj[(size_t)j * n_x + e], part);
        }
      } else {
        for (int e = t; e < n_x; e += GS) part = fma(W[e], GY[e], part);
      }
    }
    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    if (in_group && lane == 0) red[(size_t)wig * nscal + sidx] = part;
  }
  __syncthreads();
  if (active) {
    double* sc = out + (size_t)(m + 1) * n_x;
    for (int sidx = t; sidx < nscal; sidx += GS) {
      double acc = 0.0;
      for (int w = 0; w < nwarp_g; ++w) acc += red[(size_t)w * nscal + sidx];
      sc[sidx] = -acc;
    }
  }
}

// shared memory (bytes) the generic kernel needs for one block
inline size_t generic_smem_bytes(int order, int b, int n_b, int m, int GS, int KPC, bool gj_in_smem) {
  const size_t bb = (size_t)b * b, n_x = (size_t)b * n_b;
  const int npair = (order == 2) ? m * (m + 1) / 2 : 0;
  const size_t S = 1 + m + npair;
  const size_t nvec = (order == 2) ? (5 + m) : 2;
  const size_t nscal = npair + m + 1;
  const size_t nwarp_g = (GS + 31) / 32;
  size_t per_group = bb * (1 + 2 * S) + nvec * n_x + (order == 2 ? (nwarp_g + 1) * nscal : 0) + b;
  return 8 * ((gj_in_smem ? (size_t)m * bb : 0) + (size_t)KPC * per_group);
}

}  // namespace pb2
