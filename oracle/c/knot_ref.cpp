// CPU port of the reference's per-knot algorithm.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
//
// Never linked into, loaded by, or called from the shipped library
// (piccolo.jl_b200/csrc).  Used by tests/ as a second, algorithmically independent
// checker next to oracle/knot.py (SciPy Pade expm / expm_frechet), and by bench.py as
// the timed CPU baseline ("kind": "port").
//
// What it restates.  The reference evaluates, per knot k,
//     delta_k = x_{k+1} - expv(dt_k, Ghat(u_k), x_k)
// (/root/reference/src/control/integrators.jl:48-49,71-72,88-94 build Ghat and hand it to
//  DirectTrajOpt's BilinearIntegrator; constraint docs/src/concepts/index.md:21,62;
//  G(u) = G_drift + sum_j u_j G_j  at src/quantum/systems/quantum_systems.jl:226) and
// differentiates it with forward-mode duals pushed through Ghat and expv.  DirectTrajOpt.jl
// and ExponentialAction.jl (Project.toml:8,10; "0.9.5, 0.10" / "0.2") are not vendored, so
// this file restates the PUBLISHED algorithm of expv -- Al-Mohy & Higham (2011),
// "Computing the action of the matrix exponential": truncated Taylor series applied to the
// vectors, s scaling steps of degree m chosen from the 1-norm -- and carries first- and
// second-order directional derivatives (jets) through the same recurrence, which is what
// dual numbers do.  Differences from the real Julia path, stated so nobody mistakes this
// for it: dense generator instead of sparse, a forward-error theta_m table computed here
// instead of the paper's backward-error table, no trace shift, jets included in the
// early-exit test.
//
// Build: see oracle/c/Makefile (g++ -O3 -pthread -shared).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <thread>

namespace {

struct Prob {
  int b, n_b, m, K, D, x_off, dt_off, u_off;
  const double* G0;  // b*b col-major
  const double* Gj;  // m * b*b
  int n_x() const { return b * n_b; }
};

// largest theta with theta^(m+1)/(m+1)! * 1/(1-theta/(m+2)) <= 2^-53  (forward bound)
double theta_of(int m) {
  static double tab[64];
  static bool init = false;
  if (!init) {
    for (int q = 1; q < 64; ++q) {
      double lo = 0.0, hi = q + 1.0;
      for (int it = 0; it < 200; ++it) {
        double th = 0.5 * (lo + hi);
        double lg = (q + 1) * std::log(th) - std::lgamma(q + 2.0) - std::log1p(-th / (q + 2.0));
        if (lg <= std::log(std::ldexp(1.0, -53))) lo = th; else hi = th;
      }
      tab[q] = lo;
    }
    init = true;
  }
  return tab[m];
}

void pick_ms(double nrm, int& m_star, int& s) {
  m_star = 1; s = 1;
  if (nrm == 0.0) return;
  double best = 1e300;
  for (int m = 2; m <= 55; ++m) {
    double sc = std::ceil(nrm / theta_of(m));
    if (sc < 1) sc = 1;
    if (m * sc < best) { best = m * sc; m_star = m; s = (int)sc; }
  }
}

double norm1(const double* A, int b) {
  double mx = 0;
  for (int j = 0; j < b; ++j) {
    double c = 0;
    for (int i = 0; i < b; ++i) c += std::fabs(A[i + j * b]);
    mx = std::max(mx, c);
  }
  return mx;
}

// C (b x n) (+)= alpha * op(A) * B ; op = transpose if tr
inline void gemm(int b, int n, double alpha, const double* A, bool tr, const double* B, double* C,
                 bool accumulate) {
  for (int c = 0; c < n; ++c) {
    double* Cc = C + c * b;
    const double* Bc = B + c * b;
    if (!accumulate) std::fill(Cc, Cc + b, 0.0);
    if (!tr) {
      for (int k = 0; k < b; ++k) {
        double bk = alpha * Bc[k];
        if (bk == 0.0) continue;
        const double* Ak = A + k * b;
        for (int i = 0; i < b; ++i) Cc[i] += Ak[i] * bk;
      }
    } else {
      for (int i = 0; i < b; ++i) {
        const double* Ai = A + i * b;
        double acc = 0;
        for (int k = 0; k < b; ++k) acc += Ai[k] * Bc[k];
        Cc[i] += alpha * acc;
      }
    }
  }
}

inline double ninf(const double* v, int n) {
  double mx = 0;
  for (int i = 0; i < n; ++i) mx = std::max(mx, std::fabs(v[i]));
  return mx;
}

// Jet action: directions p = 0..P-1; A(p) = A + sum eps_p Ad[p] + sum eps_p eps_q Add[p][q]
// Computes F = e^{A} X, F_p = d/dp, F_pq = d2/dpdq (p<=q) when order2.
struct JetAction {
  int b, n, P;
  bool tr, order2;
  const double* A;
  std::vector<const double*> Ad;            // P first-order generator directions
  std::vector<const double*> Add;           // P*P second-order (may be null)
  std::vector<double> F, B, T;              // (1+P+P(P+1)/2) slabs of b*n each
  int slabs() const { return 1 + P + (order2 ? P * (P + 1) / 2 : 0); }
  int idx2(int p, int q) const { return 1 + P + q * (q + 1) / 2 + p; }  // p<=q

  void run(const double* X) {
    const int sz = b * n, S = slabs();
    F.assign((size_t)S * sz, 0.0);
    B.assign((size_t)S * sz, 0.0);
    T.assign((size_t)S * sz, 0.0);
    std::copy(X, X + sz, F.begin());
    std::copy(X, X + sz, B.begin());
    int m_star, s;
    pick_ms(norm1(A, b), m_star, s);
    const double tol = std::ldexp(1.0, -53);
    for (int i = 0; i < s; ++i) {
      double c1 = ninf(B.data(), S * sz);
      for (int j = 1; j <= m_star; ++j) {
        const double f = 1.0 / ((double)s * j);
        // T = f * (A-jet * B-jet)
        gemm(b, n, f, A, tr, &B[0], &T[0], false);
        for (int p = 0; p < P; ++p) {
          double* Tp = &T[(size_t)(1 + p) * sz];
          gemm(b, n, f, A, tr, &B[(size_t)(1 + p) * sz], Tp, false);
          gemm(b, n, f, Ad[p], tr, &B[0], Tp, true);
        }
        if (order2) {
          for (int q = 0; q < P; ++q)
            for (int p = 0; p <= q; ++p) {
              double* Tpq = &T[(size_t)idx2(p, q) * sz];
              gemm(b, n, f, A, tr, &B[(size_t)idx2(p, q) * sz], Tpq, false);
              gemm(b, n, f, Ad[p], tr, &B[(size_t)(1 + q) * sz], Tpq, true);
              gemm(b, n, f, Ad[q], tr, &B[(size_t)(1 + p) * sz], Tpq, true);
              const double* App = Add.empty() ? nullptr : Add[p * P + q];
              if (App) gemm(b, n, f, App, tr, &B[0], Tpq, true);  // Add = true mixed partial of A
            }
        }
        B.swap(T);
        double c2 = ninf(B.data(), S * sz);
        for (size_t e = 0; e < (size_t)S * sz; ++e) F[e] += B[e];
        if (c1 + c2 <= tol * ninf(F.data(), S * sz)) break;
        c1 = c2;
      }
      B = F;
    }
  }
  const double* slab(int i) const { return &F[(size_t)i * b * n]; }
};

void assemble(const Prob& p, const double* u, double* G) {
  const int bb = p.b * p.b;
  std::copy(p.G0, p.G0 + bb, G);
  for (int j = 0; j < p.m; ++j)
    for (int e = 0; e < bb; ++e) G[e] += u[j] * p.Gj[(size_t)j * bb + e];
}

// static partition of [0, n) over `threads` std::threads (no OpenMP dependency)
template <class Body>
void parallel_knots(int n, int threads, Body body) {
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  if (threads <= 0 || threads > hw) threads = hw;
  if (threads > n) threads = n > 0 ? n : 1;
  if (threads == 1) { body(0, n); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    int lo = (int)((long long)n * t / threads), hi = (int)((long long)n * (t + 1) / threads);
    pool.emplace_back([=] { body(lo, hi); });
  }
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// status 0 = ok.  Z: D*K col-major.  threads<=0 -> all.
int pbo_residual(int b, int n_b, int m, int K, int D, int x_off, int dt_off, int u_off,
                 const double* G0, const double* Gj, const double* Z, double* delta, int threads) {
  Prob p{b, n_b, m, K, D, x_off, dt_off, u_off, G0, Gj};
  const int n_x = p.n_x(), bb = b * b;
  parallel_knots(K - 1, threads, [&](int k_lo, int k_hi) {
    std::vector<double> G(bb), A(bb);
    JetAction ja;
    for (int k = k_lo; k < k_hi; ++k) {
      const double* z = Z + (size_t)k * D;
      const double* zn = z + D;
      const double dt = z[dt_off];
      assemble(p, z + u_off, G.data());
      for (int e = 0; e < bb; ++e) A[e] = dt * G[e];
      ja.b = b; ja.n = n_b; ja.P = 0; ja.tr = false; ja.order2 = false; ja.A = A.data();
      ja.Ad.clear(); ja.Add.clear();
      ja.run(z + x_off);
      for (int i = 0; i < n_x; ++i) delta[(size_t)k * n_x + i] = zn[x_off + i] - ja.slab(0)[i];
    }
  });
  return 0;
}

// vals: (K-1) * (n_b*b*b + n_x*m + n_x + n_x) in the canonical order of oracle/knot.py
int pbo_jacobian(int b, int n_b, int m, int K, int D, int x_off, int dt_off, int u_off,
                 const double* G0, const double* Gj, const double* Z, double* vals, int threads) {
  Prob p{b, n_b, m, K, D, x_off, dt_off, u_off, G0, Gj};
  const int n_x = p.n_x(), bb = b * b;
  const size_t nnz = (size_t)n_b * bb + (size_t)n_x * m + 2 * n_x;
  parallel_knots(K - 1, threads, [&](int k_lo, int k_hi) {
    std::vector<double> G(bb), A(bb), Ad((size_t)(m + 1) * bb), I(bb, 0.0);
    for (int i = 0; i < b; ++i) I[i + i * b] = 1.0;
    JetAction ja, je;
    for (int k = k_lo; k < k_hi; ++k) {
      const double* z = Z + (size_t)k * D;
      const double dt = z[dt_off];
      assemble(p, z + u_off, G.data());
      for (int e = 0; e < bb; ++e) A[e] = dt * G[e];
      for (int j = 0; j < m; ++j)
        for (int e = 0; e < bb; ++e) Ad[(size_t)j * bb + e] = dt * Gj[(size_t)j * bb + e];
      std::copy(G.begin(), G.end(), Ad.begin() + (size_t)m * bb);
      double* v = vals + (size_t)k * nnz;
      // d/dx_k : dense exp = action on the identity (what "dense exp for dx_k" costs)
      je.b = b; je.n = b; je.P = 0; je.tr = false; je.order2 = false; je.A = A.data();
      je.Ad.clear(); je.Add.clear();
      je.run(I.data());
      for (int c = 0; c < n_b; ++c)
        for (int e = 0; e < bb; ++e) v[(size_t)c * bb + e] = -je.slab(0)[e];
      v += (size_t)n_b * bb;
      // d/du_j , d/ddt : first-order jets through the action on x_k
      ja.b = b; ja.n = n_b; ja.P = m + 1; ja.tr = false; ja.order2 = false; ja.A = A.data();
      ja.Ad.resize(m + 1);
      for (int j = 0; j <= m; ++j) ja.Ad[j] = &Ad[(size_t)j * bb];
      ja.Add.clear();
      ja.run(z + x_off);
      for (int j = 0; j <= m; ++j)
        for (int i = 0; i < n_x; ++i) v[(size_t)j * n_x + i] = -ja.slab(1 + j)[i];
      v += (size_t)(m + 1) * n_x;
      for (int i = 0; i < n_x; ++i) v[i] = 1.0;
    }
  });
  return 0;
}

// vals: (K-1) * (n_x*m + n_x + m(m+1)/2 + m + 1)
int pbo_hessian(int b, int n_b, int m, int K, int D, int x_off, int dt_off, int u_off,
                const double* G0, const double* Gj, const double* Z, const double* mu,
                double* vals, int threads) {
  Prob p{b, n_b, m, K, D, x_off, dt_off, u_off, G0, Gj};
  const int n_x = p.n_x(), bb = b * b, P = m + 1;
  const size_t nnz = (size_t)n_x * m + n_x + (size_t)m * (m + 1) / 2 + m + 1;
  parallel_knots(K - 1, threads, [&](int k_lo, int k_hi) {
    std::vector<double> G(bb), A(bb), Ad((size_t)P * bb);
    JetAction ja, jt;
    for (int k = k_lo; k < k_hi; ++k) {
      const double* z = Z + (size_t)k * D;
      const double* M = mu + (size_t)k * n_x;
      const double dt = z[dt_off];
      assemble(p, z + u_off, G.data());
      for (int e = 0; e < bb; ++e) A[e] = dt * G[e];
      for (int j = 0; j < m; ++j)
        for (int e = 0; e < bb; ++e) Ad[(size_t)j * bb + e] = dt * Gj[(size_t)j * bb + e];
      std::copy(G.begin(), G.end(), Ad.begin() + (size_t)m * bb);
      double* v = vals + (size_t)k * nnz;
      // (x, u_j), (x, dt): -(dE/dp)^T M  = first-order jets of the transposed action on M
      jt.b = b; jt.n = n_b; jt.P = P; jt.tr = true; jt.order2 = false; jt.A = A.data();
      jt.Ad.resize(P);
      for (int j = 0; j < P; ++j) jt.Ad[j] = &Ad[(size_t)j * bb];
      jt.Add.clear();
      jt.run(M);
      for (int j = 0; j < P; ++j)
        for (int i = 0; i < n_x; ++i) v[(size_t)j * n_x + i] = -jt.slab(1 + j)[i];
      v += (size_t)P * n_x;
      // (p, q) in (u, dt)^2: second-order jets of the action on x_k, contracted with M
      ja.b = b; ja.n = n_b; ja.P = P; ja.tr = false; ja.order2 = true; ja.A = A.data();
      ja.Ad = jt.Ad;
      ja.Add.assign((size_t)P * P, nullptr);
      for (int j = 0; j < m; ++j) ja.Add[(size_t)j * P + m] = &Gj[(size_t)j * bb];  // d2A/du_j ddt = G_j
      ja.run(z + x_off);
      auto dot = [&](int p_, int q_) {
        const double* s = ja.slab(ja.idx2(p_, q_));
        double acc = 0;
        for (int i = 0; i < n_x; ++i) acc += M[i] * s[i];
        return -acc;
      };
      for (int j = 0; j < m; ++j)
        for (int i = 0; i <= j; ++i) *v++ = dot(i, j);
      for (int j = 0; j < m; ++j) *v++ = dot(j, m);
      *v++ = dot(m, m);
    }
  });
  return 0;
}

int pbo_max_threads() {
  int hw = (int)std::thread::hardware_concurrency();
  return hw < 1 ? 1 : hw;
}

}  // extern "C"
