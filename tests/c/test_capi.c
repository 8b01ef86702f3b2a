/* A plain C consumer of the drop-in boundary: links libpiccolo_b200.so through include/piccolo_b200.h only (no
 * Python, no ctypes), evaluates BASELINE config C1 (single-qubit X gate: H_drift = Z, H_1 = X,
 * /root/reference/src/control/templates/smooth_pulse_problem.jl:792-793; K = 50 knots, SmoothPulseProblem layout
 * [U (8) | dt | t | u | du | ddu]) and compares every entry point with the CPU port of the reference algorithm
 * (oracle/_build/libknot_ref.so, test infrastructure).  Built and run by tests/test_gpu_parity.py
 * (test_c_program_through_the_abi); exit code 0 = every check passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/piccolo_b200.h"

int pbo_residual(int b, int n_b, int m, int K, int D, int x_off, int dt_off, int u_off, const double* G0,
                 const double* Gj, const double* Z, double* delta, int threads);
int pbo_jacobian(int b, int n_b, int m, int K, int D, int x_off, int dt_off, int u_off, const double* G0,
                 const double* Gj, const double* Z, double* vals, int threads);
int pbo_hessian(int b, int n_b, int m, int K, int D, int x_off, int dt_off, int u_off, const double* G0,
                const double* Gj, const double* Z, const double* mu, double* vals, int threads);

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); ++fails; } } while (0)

static double maxdiff(const double* a, const double* b, long n) {
  double m = 0.0;
  for (long i = 0; i < n; ++i) { const double d = fabs(a[i] - b[i]); if (d > m || d != d) m = d != d ? INFINITY : d; }
  return m;
}

static unsigned long long rng_state = 20261017ull;
static double urand(void) {   /* xorshift: reproducible without libc's rand */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}

int main(void) {
  enum { b = 4, n_b = 2, m = 1, K = 50, n_x = 8, D = n_x + 2 + 3 * m };
  /* G(H) = [[Im H, Re H], [-Re H, Im H]] (src/quantum/primitives/isomorphisms.jl:350,359), column-major */
  const double G0[16] = {0, 0, -1, 0,  0, 0, 0, 1,  1, 0, 0, 0,  0, -1, 0, 0};   /* G(Z) */
  const double G1[16] = {0, 0, 0, -1,  0, 0, -1, 0,  0, 1, 0, 0,  1, 0, 0, 0};   /* G(X) */
  double* Z = (double*)calloc((size_t)D * K, sizeof(double));
  for (int k = 0; k < K; ++k) {
    double* z = Z + (size_t)k * D;
    for (int i = 0; i < n_x; ++i) z[i] = 2.0 * urand() - 1.0;   /* any state: parity holds on every input */
    z[n_x] = 10.0 / (K - 1);
    z[n_x + 1] = k * 10.0 / (K - 1);
    z[n_x + 2] = 2.0 * urand() - 1.0;
    z[n_x + 3] = 0.01 * (urand() - 0.5);
    z[n_x + 4] = 0.01 * (urand() - 0.5);
  }

  CHECK(pb2_version() == PB2_VERSION, "version");
  pb2_desc d;
  memset(&d, 0, sizeof(d));
  d.kind = PB2_UNITARY; d.b = b; d.n_b = n_b; d.m = m; d.K = K; d.D = D;
  d.x_off = 0; d.dt_off = n_x; d.u_off = n_x + 2; d.global_dim = 0; d.knot0 = 0; d.device = 0;
  d.algorithm = PB2_ALG_AUTO; d.G0 = G0; d.Gj = G1;
  pb2_handle* h = NULL;
  /* error behaviour: a bad descriptor is refused with a message, nothing is thrown across the boundary */
  pb2_desc bad = d;
  bad.n_b = 3;
  CHECK(pb2_create(&bad, &h) == PB2_EINVAL && h == NULL && strlen(pb2_last_error()) > 0, "bad descriptor accepted");
  int rc = pb2_create(&d, &h);
  if (rc != PB2_OK) { printf("pb2_create: %d %s\n", rc, pb2_last_error()); return 2; }

  const long dim = (long)pb2_dim(h), nj = (long)pb2_nnz_jac(h), nh = (long)pb2_nnz_hess(h);
  CHECK(dim == (long)n_x * (K - 1), "dim %ld", dim);                                   /* integrators.jl:309 */
  CHECK(nj == (long)(n_b * b * b + n_x * m + 2 * n_x) * (K - 1), "nnz_jac %ld", nj);
  CHECK(nh == (long)(n_x * m + n_x + m * (m + 1) / 2 + m + 1) * (K - 1), "nnz_hess %ld", nh);

  int64_t *rows = (int64_t*)malloc(sizeof(int64_t) * nj), *cols = (int64_t*)malloc(sizeof(int64_t) * nj);
  CHECK(pb2_structure_jac(h, rows, cols) == PB2_OK, "structure_jac");
  CHECK(rows[0] == 1 && cols[0] == 1 && rows[1] == 2 && cols[b] == 2, "first entries of the COO structure (1-based)");
  long bad_idx = 0;
  for (long e = 0; e < nj; ++e) bad_idx += rows[e] < 1 || rows[e] > dim || cols[e] < 1 || cols[e] > (long)D * K;
  CHECK(bad_idx == 0, "%ld structure entries out of range", bad_idx);

  double *delta = (double*)malloc(sizeof(double) * dim), *vals = (double*)malloc(sizeof(double) * nj);
  double *dref = (double*)malloc(sizeof(double) * dim), *vref = (double*)malloc(sizeof(double) * nj);
  CHECK(pb2_residual_jacobian(h, Z, delta, vals, PB2_HOST) == PB2_OK, "residual_jacobian: %s", pb2_last_error());
  CHECK(pbo_residual(b, n_b, m, K, D, 0, n_x, n_x + 2, G0, G1, Z, dref, 1) == 0, "oracle residual");
  CHECK(pbo_jacobian(b, n_b, m, K, D, 0, n_x, n_x + 2, G0, G1, Z, vref, 1) == 0, "oracle jacobian");
  CHECK(maxdiff(delta, dref, dim) < 1e-12, "residual differs from the reference algorithm by %.3e", maxdiff(delta, dref, dim));
  CHECK(maxdiff(vals, vref, nj) < 1e-11, "Jacobian differs from the reference algorithm by %.3e", maxdiff(vals, vref, nj));

  double* d2 = (double*)malloc(sizeof(double) * dim);
  CHECK(pb2_residual(h, Z, d2, PB2_HOST) == PB2_OK && maxdiff(d2, delta, dim) < 1e-13, "pb2_residual vs fused call");
  double* v2 = (double*)malloc(sizeof(double) * nj);
  CHECK(pb2_jacobian(h, Z, v2, PB2_HOST) == PB2_OK && maxdiff(v2, vals, nj) < 1e-13, "pb2_jacobian vs fused call");

  double *mu = (double*)malloc(sizeof(double) * dim), *hv = (double*)malloc(sizeof(double) * nh), *href = (double*)malloc(sizeof(double) * nh);
  for (long i = 0; i < dim; ++i) mu[i] = 2.0 * urand() - 1.0;
  CHECK(pb2_hess_lagrangian(h, Z, mu, hv, PB2_HOST) == PB2_OK, "hess_lagrangian: %s", pb2_last_error());
  CHECK(pbo_hessian(b, n_b, m, K, D, 0, n_x, n_x + 2, G0, G1, Z, mu, href, 1) == 0, "oracle hessian");
  double hmax = 0.0;
  for (long i = 0; i < nh; ++i) hmax = fmax(hmax, fabs(href[i]));
  CHECK(maxdiff(hv, href, nh) < 1e-9 * fmax(1.0, hmax), "Hessian differs by %.3e (max |H| %.3e)", maxdiff(hv, href, nh), hmax);

  /* rollout of the controls from the first state column; the divergence numbers are finite and consistent */
  double* states = (double*)malloc(sizeof(double) * n_x * K);
  double out3[3] = {0, 0, 0};
  CHECK(pb2_rollout(h, Z, NULL, states, out3, PB2_HOST) == PB2_OK, "rollout: %s", pb2_last_error());
  CHECK(maxdiff(states, Z, n_x) == 0.0, "rollout starts from the first state column");
  CHECK(out3[0] == out3[1] / fmax(out3[2], 1.0) && out3[1] == out3[1], "rollout_divergence = %.3e / max(%.3e, 1)", out3[1], out3[2]);

  CHECK(pb2_set_option(h, PB2_OPT_EARLY_Z, 1) == PB2_OK && pb2_set_option(h, 12345, 1) == PB2_EINVAL, "set_option");
  CHECK(pb2_launch_count(h) > 0, "launch accounting");
  pb2_destroy(h);
  printf(fails ? "test_capi: %d check(s) FAILED\n" : "test_capi: all checks passed\n", fails);
  return fails ? 1 : 0;
}
