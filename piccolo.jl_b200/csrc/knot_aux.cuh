// Linear knot constraints that ride along with the dynamics in every Ipopt callback
// (SURVEY.md section 8f, rank 1): DirectTrajOpt's DerivativeIntegrator(x, xdot, traj)
//     r_k = x_{k+1} - x_k - dt_k * xdot_k                  (u -> du, du -> ddu:
//     /root/reference/src/control/templates/smooth_pulse_problem.jl:267-275,
//     spline_pulse_problem.jl:363-366, bang_bang_pulse_problem.jl:213)
// and the time-consistency constraint  t_{k+1} - t_k - dt_k = 0  that DirectTrajOpt applies
// whenever :t and :dt are present (smooth_pulse_problem.jl:277), and TimeStepsAllEqualConstraint
//     dt_{k+1} - dt_k = 0                                  (pushed by the templates when
//     piccolo_options.timesteps_all_equal is set: src/control/templates/_problem_templates.jl:175-180; every
//     reference solution under docs/data satisfies it exactly).  All pairs of one trajectory are
// evaluated by ONE launch: residual, Jacobian values and (optionally) Lagrangian-Hessian values.
// Pure streaming work: one thread per constraint row, coalesced over the component index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb2 {

constexpr int kAuxMaxPairs = 8;

struct AuxParams {
  int K, D, dt_off, t_off;        // t_off < 0: no time-consistency rows
  int n_pairs;
  int x_off[kAuxMaxPairs], xdot_off[kAuxMaxPairs], dim[kAuxMaxPairs];
  long long row0[kAuxMaxPairs + 1];   // first row of each pair's block; [n_pairs] = time rows
  long long row_eq;                   // first timesteps-all-equal row (= n_rows when the constraint is off)
  long long n_rows;
  const double* Z;
  const double* mu;    // Hessian launch only
  double* delta;       // n_rows
  double* jac;         // 4 per derivative row (d x_k, d xdot_k, d dt_k, d x_{k+1}); 3 per time row; 2 per equal-dt row
  double* hess;        // 1 per derivative row: (xdot_k[i], dt_k) = -mu
};

__global__ void __launch_bounds__(256) knot_aux_kernel(AuxParams p) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.n_rows;
       r += (long long)gridDim.x * blockDim.x) {
    int pr = 0;
    while (pr < p.n_pairs && r >= p.row0[pr + 1]) ++pr;
    const long long loc = r - p.row0[pr];
    if (pr < p.n_pairs) {
      const int dim = p.dim[pr];
      const long long k = loc / dim;
      const int i = (int)(loc - k * dim);
      if (p.delta || p.jac) {
        const double* z = p.Z + k * p.D;
        const double dt = z[p.dt_off], x = z[p.x_off[pr] + i], xd = z[p.xdot_off[pr] + i];
        const double xn = z[p.D + p.x_off[pr] + i];
        // (x_{k+1} - x_k) - dt * xdot with individually rounded operations: bit-identical to the
        // reference's left-to-right Float64 evaluation
        if (p.delta) p.delta[r] = __dsub_rn(__dsub_rn(xn, x), __dmul_rn(dt, xd));
        if (p.jac) {
          double* v = p.jac + 4 * r;
          v[0] = -1.0; v[1] = -dt; v[2] = -xd; v[3] = 1.0;
        }
      }
      if (p.hess) p.hess[r] = -p.mu[r];   // (xdot_k[i], dt_k): the only second derivative
    } else if (r < p.row_eq) {
      const double* z = p.Z + loc * p.D;   // loc = knot
      if (p.delta) p.delta[r] = __dsub_rn(__dsub_rn(z[p.D + p.t_off], z[p.t_off]), z[p.dt_off]);
      if (p.jac) {
        double* v = p.jac + 4 * p.row0[p.n_pairs] + 3 * loc;
        v[0] = -1.0; v[1] = -1.0; v[2] = 1.0;
      }
    } else {
      const long long k = r - p.row_eq;    // dt_{k+1} - dt_k
      const double* z = p.Z + k * p.D;
      if (p.delta) p.delta[r] = __dsub_rn(z[p.D + p.dt_off], z[p.dt_off]);
      if (p.jac) {
        double* v = p.jac + 4 * p.row0[p.n_pairs] + 3 * (p.row_eq - p.row0[p.n_pairs]) + 2 * k;
        v[0] = -1.0; v[1] = 1.0;
      }
    }
  }
}

}  // namespace pb2
