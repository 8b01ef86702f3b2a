// Time-dependent (carrier-modulated) drives: the reference builds
// TimeDependentBilinearIntegrator(Ghat, x, u, :t, traj) when sys.time_dependent
// (/root/reference/src/control/integrators.jl:38-46, 63, 110), Ghat(u, t) = sys.G(u, t); for
// ModulatedDrive(LinearDrive(H_j, j), c_j) drives (src/quantum/systems/drives.jl:342-388,
// quantum_systems.jl:575-597) that is  Ghat(u, t) = G_drift + sum_j c_j(t) u_j G_j.  The modulation functions are
// Julia closures and cannot cross a C ABI; the time dependence is separable, so the host evaluates
// c_j(t_k), c_j'(t_k) at the trajectory's current time row (m x K numbers per callback) and the device does the rest:
//   * td_scale_controls_kernel: Z' = Z with the drive rows multiplied by c_j(t_k) -- the knot kernels then
//     evaluate exp(dt Ghat(u_k, t_k)) and the jets with respect to the EFFECTIVE controls c_j u_j unchanged;
//   * td_finish_kernel: chain rule on the finished Jacobian values of every knot,
//         d delta / d t_k  = sum_j c_j'(t_k) u_j * (jet_j)        (the extra column, last n_x values of the knot)
//         d delta / d u_j  = c_j(t_k) * (jet_j).
// Both are streaming passes (HBM-bound, a few microseconds at BASELINE sizes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb2 {

struct TdParams {
  int K, D, m, n_x, u_off;
  long long nnz_jac;       // values per knot (including the trailing d/dt column)
  long long o_jets;        // offset of the first jet column inside a knot's values (n_b * b * b)
  const double* c;         // m x K, column k = c_j(t_k)
  const double* cdot;      // m x K
  const double* Z;
  double* Zs;              // D x K scaled copy
  double* jac;
};

__global__ void __launch_bounds__(256) td_scale_controls_kernel(const TdParams p) {
  const long long n = (long long)p.K * p.D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / p.D;
    const int r = (int)(e - k * p.D), j = r - p.u_off;
    double v = p.Z[e];
    if (j >= 0 && j < p.m) v *= p.c[j + (long long)p.m * k];
    p.Zs[e] = v;
  }
}

// one CTA per knot (grid-stride): thread i handles row i of the n_x-row columns
__global__ void __launch_bounds__(128) td_finish_kernel(const TdParams p) {
  for (long long k = blockIdx.x; k + 1 < p.K; k += gridDim.x) {
    double* jk = p.jac + k * p.nnz_jac;
    const double* zk = p.Z + k * p.D;
    for (int i = threadIdx.x; i < p.n_x; i += blockDim.x) {
      double tcol = 0.0;
      for (int j = 0; j < p.m; ++j) {
        double* jet = jk + p.o_jets + (long long)j * p.n_x + i;
        const double v = *jet;                       // -(F_j X)[i] with respect to the effective control
        tcol = fma(p.cdot[j + (long long)p.m * k] * zk[p.u_off + j], v, tcol);
        *jet = p.c[j + (long long)p.m * k] * v;
      }
      jk[p.nnz_jac - p.n_x + i] = tcol;
    }
  }
}

}  // namespace pb2
