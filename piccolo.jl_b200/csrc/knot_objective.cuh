// Objective value and dense gradient over a knot-major trajectory.
//
// J(Z) = sum over "terms" (terminal / knot-point losses of the real-form
//          F = scale*((a_re.z)^2 + (a_im.z)^2 + sum_i a_sq_i z_i^2) + a_lin.z,
//          loss = Q*|1 - F| or Q*F:   /root/reference/src/control/objectives.jl:24-38, 96-121, 330-345,
//          387-394, 412-419, 464-474)
//      + sum over quadratic regularizers 1/2 sum_i R_i (z_i - b_i)^2 dt^p
//          (DirectTrajOpt's QuadraticRegularizer as smooth_pulse_problem.jl:248-250 adds it).
//
// One CTA per knot: the knot column and its gradient column live in shared memory, every term that
// is active at the knot is a block-wide dot product followed by an axpy into the gradient column, so
// each byte of Z is read once and each byte of the gradient written once (no memset, no atomics on
// values).  The last CTA to finish adds the per-knot partial sums in knot order, so J is bitwise
// reproducible from run to run.  HBM-bound: 16 bytes per trajectory entry.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pb2 {

struct ObjParams {
  int K, D, dt_off, n_regs;
  // per-knot CSR of active term instances
  const int* knot_ptr;      // K+1
  const int* item_term;     // term id
  const double* item_q;     // weight Q of that instance
  // per-term
  const int* t_off;         // n_terms+1, offsets into the pooled row / coefficient arrays
  const int* t_flags;
  const double* t_scale;
  const int* rows;          // pooled
  const double *a_re, *a_im, *a_sq, *a_lin;   // pooled
  // per-regularizer
  const int* r_off;         // n_regs+1, offsets into r_rows / r_R
  const int* r_pow;
  const int* r_rows;
  const double* r_R;
  const double* r_w;        // n_regs x K activity (0/1)
  const double* const* r_base;  // n_regs pointers (n_rows x K column-major) or null
  // data
  const double* Z;
  double* grad;             // K*D or null
  double* J;                // 1
  double* Jpart;            // K
  unsigned int* counter;
  // Hessian (knot_objective_hess_kernel)
  const int* t_kind;        // per term: 0 = no second derivative (a_lin only), 1 = diagonal (a_sq only), 2 = dense block
  const long long* hess_ptr;   // K+1: first Hessian entry of each knot
  double* hess;
  double sigma;             // Ipopt's obj_factor
};

constexpr int kObjThreads = 128;
constexpr int kObjRegs = 8;       // regularizers whose metadata is staged in shared memory

// block-wide sums of N values; every thread returns with the totals (fixed order => deterministic)
template <int N>
__device__ __forceinline__ void obj_block_sum(double (&v)[N], double* red) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int o = 16; o; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  const int w = threadIdx.x >> 5, nw = kObjThreads / 32;
  __syncthreads();   // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int i = 0; i < N; ++i) red[w * N + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = red[i];
    for (int q = 1; q < nw; ++q) s += red[q * N + i];
    v[i] = s;
  }
}

__global__ void __launch_bounds__(kObjThreads) knot_objective_kernel(ObjParams p) {
  extern __shared__ double smem_obj[];
  double* z = smem_obj;
  double* g = smem_obj + p.D;
  __shared__ double red[4 * (kObjThreads / 32)];
  __shared__ bool last;
  // per-knot metadata of the regularizers, fetched in parallel with the knot column (one round trip
  // instead of a chain of dependent loads per regularizer)
  __shared__ int r_o0[kObjRegs], r_n[kObjRegs], r_pw[kObjRegs];
  __shared__ double r_act[kObjRegs];
  __shared__ const double* r_b[kObjRegs];
  __shared__ int it_range[2];
  const int k = blockIdx.x, tid = threadIdx.x;
  const double* Zk = p.Z + (size_t)k * p.D;
  if (tid < p.n_regs && tid < kObjRegs) {
    r_o0[tid] = p.r_off[tid];
    r_n[tid] = p.r_off[tid + 1] - p.r_off[tid];
    r_pw[tid] = p.r_pow[tid];
    r_act[tid] = p.r_w[(size_t)tid * p.K + k];
    r_b[tid] = p.r_base[tid];
  }
  if (tid >= kObjThreads - 2) it_range[tid - (kObjThreads - 2)] = p.knot_ptr[k + tid - (kObjThreads - 2)];
  for (int i = tid; i < p.D; i += kObjThreads) {
    z[i] = Zk[i];
    g[i] = 0.0;
  }
  __syncthreads();
  double Jk = 0.0;

  for (int it = it_range[0]; it < it_range[1]; ++it) {
    const int t = p.item_term[it];
    const double Q = p.item_q[it];
    const int o0 = p.t_off[t], n = p.t_off[t + 1] - o0;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += kObjThreads) {
      const double zi = z[p.rows[o0 + i]];
      s[0] = fma(p.a_re[o0 + i], zi, s[0]);
      s[1] = fma(p.a_im[o0 + i], zi, s[1]);
      s[2] = fma(p.a_sq[o0 + i] * zi, zi, s[2]);
      s[3] = fma(p.a_lin[o0 + i], zi, s[3]);
    }
    obj_block_sum(s, red);
    const double sc = p.t_scale[t];
    const double F = sc * (s[0] * s[0] + s[1] * s[1] + s[2]) + s[3];
    double c;
    if (p.t_flags[t] & 1) {
      const double d = 1.0 - F;
      Jk += Q * fabs(d);
      c = signbit(d) ? Q : -Q;    // d|x|/dx = +1 at +0, like ForwardDiff
    } else {
      Jk += Q * F;
      c = Q;
    }
    const double cr = 2.0 * c * sc * s[0], ci = 2.0 * c * sc * s[1], cq = 2.0 * c * sc;
    for (int i = tid; i < n; i += kObjThreads) {
      const int r = p.rows[o0 + i];
      g[r] += cr * p.a_re[o0 + i] + ci * p.a_im[o0 + i] + cq * p.a_sq[o0 + i] * z[r] + c * p.a_lin[o0 + i];
    }
    __syncthreads();
  }

  const double dt = z[p.dt_off];
  for (int r = 0; r < p.n_regs; ++r) {
    const bool cached = r < kObjRegs;
    if ((cached ? r_act[r] : p.r_w[(size_t)r * p.K + k]) == 0.0) continue;
    const int o0 = cached ? r_o0[r] : p.r_off[r], n = cached ? r_n[r] : p.r_off[r + 1] - o0;
    const int pw = cached ? r_pw[r] : p.r_pow[r];
    const double dtp = pw == 0 ? 1.0 : (pw == 1 ? dt : dt * dt);
    const double ddtp = pw == 0 ? 0.0 : (pw == 1 ? 1.0 : 2.0 * dt);
    const double* base = cached ? r_b[r] : p.r_base[r];
    double q[1] = {0.0};
    for (int i = tid; i < n; i += kObjThreads) {
      const int row = p.r_rows[o0 + i];
      const double dv = base ? z[row] - base[(size_t)k * n + i] : z[row];
      const double Rd = p.r_R[o0 + i] * dv;
      q[0] = fma(Rd, dv, q[0]);
      g[row] += Rd * dtp;
    }
    obj_block_sum(q, red);
    Jk += 0.5 * q[0] * dtp;
    if (tid == 0 && pw) g[p.dt_off] += 0.5 * q[0] * ddtp;
    __syncthreads();
  }

  if (p.grad) {
    double* Gk = p.grad + (size_t)k * p.D;
    for (int i = tid; i < p.D; i += kObjThreads) Gk[i] = g[i];
  }
  if (tid == 0) {
    p.Jpart[k] = Jk;
    __threadfence();
    last = atomicAdd(p.counter, 1u) == (unsigned)p.K - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // knot-ordered final sum: thread t owns the contiguous slice [t*c, (t+1)*c)
  const int c = (p.K + kObjThreads - 1) / kObjThreads;
  double s1[1] = {0.0};
  for (int i = tid * c; i < min(p.K, (tid + 1) * c); ++i) s1[0] += __ldcg(p.Jpart + i);
  obj_block_sum(s1, red);
  if (tid == 0) {
    *p.J = s1[0];
    *p.counter = 0u;
  }
}

// Hessian of sigma * J over the trajectory entries, upper triangle, COO values in the order pb2_obj_structure_hess
// lists them: knot-major; inside a knot first the term instances (a dense term: the upper triangle of its
// n_rows x n_rows block, column-major; an a_sq-only term: its diagonal; an a_lin-only term: nothing), then the
// active regularizers (n_rows diagonal entries, and for dt_power > 0 n_rows (z_i, dt) entries and one (dt, dt)).
//   loss = Q |1 - F| :  d2 loss = c d2F,  c = -Q sign(1 - F)   (almost everywhere; ForwardDiff's convention at 0)
//   d2F[i,j] = 2 scale (a_re_i a_re_j + a_im_i a_im_j + [i == j] a_sq_i)
//   reg = 1/2 sum_i R_i dv_i^2 dt^p :  (z_i,z_i) R_i dt^p;  (z_i,dt) R_i dv_i p dt^(p-1);  (dt,dt) 1/2 q p (p-1) dt^(p-2)
// Same one-CTA-per-knot shape as the gradient kernel; the only large block is the terminal one.
__global__ void __launch_bounds__(kObjThreads) knot_objective_hess_kernel(ObjParams p) {
  extern __shared__ double smem_obj[];
  double* z = smem_obj;
  __shared__ double red[4 * (kObjThreads / 32)];
  const int k = blockIdx.x, tid = threadIdx.x;
  const double* Zk = p.Z + (size_t)k * p.D;
  for (int i = tid; i < p.D; i += kObjThreads) z[i] = Zk[i];
  __syncthreads();
  long long o = p.hess_ptr[k];
  if (o == p.hess_ptr[k + 1]) return;
  for (int it = p.knot_ptr[k]; it < p.knot_ptr[k + 1]; ++it) {
    const int t = p.item_term[it], kind = p.t_kind[t];
    if (kind == 0) continue;
    const double Q = p.item_q[it];
    const int o0 = p.t_off[t], n = p.t_off[t + 1] - o0;
    const double sc = p.t_scale[t];
    double c = Q;
    if (p.t_flags[t] & 1) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n; i += kObjThreads) {
        const double zi = z[p.rows[o0 + i]];
        s[0] = fma(p.a_re[o0 + i], zi, s[0]);
        s[1] = fma(p.a_im[o0 + i], zi, s[1]);
        s[2] = fma(p.a_sq[o0 + i] * zi, zi, s[2]);
        s[3] = fma(p.a_lin[o0 + i], zi, s[3]);
      }
      obj_block_sum(s, red);
      const double F = sc * (s[0] * s[0] + s[1] * s[1] + s[2]) + s[3];
      c = signbit(1.0 - F) ? Q : -Q;
    }
    const double w = 2.0 * p.sigma * c * sc;
    if (kind == 1) {
      for (int i = tid; i < n; i += kObjThreads) p.hess[o + i] = w * p.a_sq[o0 + i];
      o += n;
    } else {
      const long long ne = (long long)n * (n + 1) / 2;
      for (long long e = tid; e < ne; e += kObjThreads) {
        int j = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
        while ((long long)(j + 1) * (j + 2) / 2 <= e) ++j;
        while ((long long)j * (j + 1) / 2 > e) --j;
        const int i = (int)(e - (long long)j * (j + 1) / 2);
        double v = p.a_re[o0 + i] * p.a_re[o0 + j] + p.a_im[o0 + i] * p.a_im[o0 + j];
        if (i == j) v += p.a_sq[o0 + i];
        p.hess[o + e] = w * v;
      }
      o += ne;
    }
  }
  const double dt = z[p.dt_off];
  for (int r = 0; r < p.n_regs; ++r) {
    if (p.r_w[(size_t)r * p.K + k] == 0.0) continue;
    const int o0 = p.r_off[r], n = p.r_off[r + 1] - o0, pw = p.r_pow[r];
    const double dtp = pw == 0 ? 1.0 : (pw == 1 ? dt : dt * dt);
    const double ddtp = pw == 0 ? 0.0 : (pw == 1 ? 1.0 : 2.0 * dt);
    const double* base = p.r_base[r];
    double q[1] = {0.0};
    for (int i = tid; i < n; i += kObjThreads) {
      const int row = p.r_rows[o0 + i];
      const double dv = base ? z[row] - base[(size_t)k * n + i] : z[row];
      const double Rd = p.r_R[o0 + i] * dv;
      q[0] = fma(Rd, dv, q[0]);
      p.hess[o + i] = p.sigma * p.r_R[o0 + i] * dtp;
      // a regularised timestep row meets itself in the cross term: both orders land on (dt, dt)
      if (pw) p.hess[o + n + i] = p.sigma * Rd * ddtp * (row == p.dt_off ? 2.0 : 1.0);
    }
    o += n;
    if (pw) {
      obj_block_sum(q, red);
      if (tid == 0) p.hess[o + n] = pw == 2 ? p.sigma * q[0] : 0.0;
      o += n + 1;
    }
  }
}

}  // namespace pb2
