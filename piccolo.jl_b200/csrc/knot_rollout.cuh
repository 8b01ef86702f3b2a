// Rollout of a trajectory's piecewise-constant controls:
//     x_1 = x0,   x_{k+1} = exp(dt_k Ghat(u_k)) x_k,   k = 1 .. K-1.
// What it replaces: `rollout!(qtraj, pulse)` of the reference
// (/root/reference/src/quantum/trajectories/rollouts_extensions.jl:46-92) for the zero-order-hold pulse that
// `extract_pulse` builds from the optimizer's trajectory, as called by `sync_trajectory!`
// (src/control/problems.jl:186-208), and the diagnostic `rollout_divergence` (problems.jl:336-356)
//     eps = || x_K^rollout - x_K^collocation ||_2 / max(|| x_K^collocation ||_2, 1).
// The reference integrates the ODE adaptively; for a zero-order-hold pulse the exact flow over a knot interval is
// the propagator E_k = exp(dt_k Ghat(u_k)) -- exactly what the knot kernels already compute: the d/dx_k block of
// the Jacobian values is -(I (x) E_k).  So a rollout is one residual+Jacobian launch followed by THIS kernel, a
// chain of K-1 small matrix products.  The chain is sequential in k; one CTA walks it with the next propagator
// block staged into shared memory (double buffered) while the current product runs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb2 {

struct RolloutParams {
  int b, n_b, K, D, x_off;
  long long nnz_jac;        // values per knot; the first b*b of them are -E_k, column-major
  const double* jac;        // Jacobian values of the K-1 knots
  const double* Z;          // trajectory (collocation states for the divergence; x0 when `x0` is null)
  const double* x0;         // n_x initial state, or null: Z's first state column
  double* states;           // n_x * K, column k = rolled-out state at knot k; may be null
  double* out;              // [0] = rollout divergence, [1] = ||x_K^rollout - x_K^collocation||_2, [2] = ||x_K^collocation||_2
};

constexpr int kRolloutThreads = 256;

// dynamic shared memory: 2 * b*b (propagator blocks) + 2 * n_x (state ping-pong) + 64 (reduction) doubles
__global__ void __launch_bounds__(kRolloutThreads) knot_rollout_kernel(const RolloutParams p) {
  extern __shared__ double ro_smem[];
  const int b = p.b, bb = b * b, n_x = b * p.n_b, tid = threadIdx.x, nt = blockDim.x;
  double* sE = ro_smem;                 // [2][bb]
  double* sx = ro_smem + 2 * bb;        // [2][n_x]
  double* red = sx + 2 * n_x;           // [64]
  for (int e = tid; e < n_x; e += nt) {
    const double v = p.x0 ? p.x0[e] : p.Z[p.x_off + e];
    sx[e] = v;
    if (p.states) p.states[e] = v;
  }
  if (p.K > 1)
    for (int e = tid; e < bb; e += nt) sE[e] = p.jac[e];
  __syncthreads();
  for (int k = 0; k + 1 < p.K; ++k) {
    const double* Ec = sE + (k & 1) * bb;
    const double* xc = sx + (k & 1) * n_x;
    double* xn = sx + ((k + 1) & 1) * n_x;
    // stage the next propagator block while this product runs
    if (k + 2 < p.K) {
      const double* src = p.jac + (long long)(k + 1) * p.nnz_jac;
      double* dst = sE + ((k + 1) & 1) * bb;
      for (int e = tid; e < bb; e += nt) dst[e] = src[e];
    }
    for (int e = tid; e < n_x; e += nt) {
      const int c = e / b, i = e - c * b;
      // x_{k+1}[i, c] = sum_j E[i, j] x_k[j, c], summed in index order (the stored block is -E)
      double acc = 0.0;
      for (int j = 0; j < b; ++j) acc = fma(-Ec[i + j * b], xc[c * b + j], acc);
      xn[e] = acc;
      if (p.states) p.states[(long long)(k + 1) * n_x + e] = acc;
    }
    __syncthreads();
  }
  // rollout divergence against the trajectory's terminal state (problems.jl:336-356)
  const double* xf = sx + ((p.K - 1) & 1) * n_x;
  const double* zc = p.Z + (long long)(p.K - 1) * p.D + p.x_off;
  double sd = 0.0, sc = 0.0;
  for (int e = tid; e < n_x; e += nt) {
    const double d = xf[e] - zc[e];
    sd = fma(d, d, sd);
    sc = fma(zc[e], zc[e], sc);
  }
  for (int o = 16; o > 0; o >>= 1) {
    sd += __shfl_down_sync(0xffffffffu, sd, o);
    sc += __shfl_down_sync(0xffffffffu, sc, o);
  }
  if ((tid & 31) == 0) { red[tid >> 5] = sd; red[32 + (tid >> 5)] = sc; }
  __syncthreads();
  if (tid == 0 && p.out) {
    double a = 0.0, c = 0.0;
    for (int wv = 0; wv < (nt + 31) / 32; ++wv) { a += red[wv]; c += red[32 + wv]; }
    const double nd = sqrt(a), nc = sqrt(c);
    p.out[0] = nd / fmax(nc, 1.0);
    p.out[1] = nd;
    p.out[2] = nc;
  }
}

inline size_t rollout_smem_bytes(int b, int n_b) { return sizeof(double) * (size_t)(2 * b * b + 2 * b * n_b + 64); }

}  // namespace pb2
