// Tensor-core knot kernel (residual + Jacobian): one warp per knot, everything in registers.
//
// Per knot k the reference's BilinearIntegrator evaluates (constructed at
// /root/reference/src/control/integrators.jl:35-95; constraint docs/src/concepts/index.md:21,62)
//     delta_k = x_{k+1} - expv(dt_k, Ghat(u_k), x_k)
// and differentiates it with forward duals over u.  ExponentialAction.expv is the
// Al-Mohy--Higham truncated-Taylor *action*: n_sub sub-steps of a degree-M Horner recurrence
// applied to the vectors.  This kernel runs that same recurrence -- on the state columns, on
// their first-order jets in every drive direction, and on the columns of the identity (which
// yields the propagator E itself for the d/dx_k block) -- with FP64 tensor-core MMAs:
//
//     T   <-  base   + (dt'/q) G(u) T                    q = M .. 1        (Horner)
//     T_j <-  base_j + (dt'/q) (G(u) T_j + G_j Y)        (jet of drive j; G_j sparse, ELL)
//
// Layout trick that keeps the whole recurrence in registers: the warp holds the TRANSPOSED
// iterate, T^T (state columns x b), as mma.m8n8k4 accumulator tiles and computes
// T^T <- T^T G^T.  The accumulator fragment of one step (lane (g,q) holds rows 2q,2q+1 of
// column g) is exactly the A-operand fragment of the next step if the contraction index is
// enumerated in the order pi(4 kt + q) = 8 (kt/2) + 2q + (kt%2); G's B-fragments are stored
// in that permuted order once at setup.  So a Horner step is NMT*NT*KT DMMAs and nothing else:
// no shared-memory round trip, no shuffles, no block barrier.  Shared memory is used only to
// exchange the current Y columns for the sparse G_j Y coupling (1 KB per knot).
//
// Generators with the real-isomorphism structure G = [[S, R], [-R, S]] (every ket / unitary
// generator, isomorphisms.jl:350,359) need only the first b/2 columns of E: E = [[P,-Q],[Q,P]].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <vector>

#include "knot_generic.cuh"

namespace pb2 {

struct __align__(16) EllEntry {
  double val;
  int idx;
  int pad;
};

struct DmmaParams {
  int b, n_b, m, K, D, x_off, dt_off, u_off, nnz_jac;
  int ncT;       // identity columns carried (b/2 for iso generators, b otherwise, 0 = residual only)
  int m_jets;    // drive jets carried (m, or 0 for residual only)
  int W;         // ELL width of the drive generators
  int iso;       // 1: E is rebuilt from its first b/2 columns
  int max_sub;   // bound on the number of Taylor sub-steps
  const double* Gfrag;   // (m+1) * KT*NT*32 doubles, B-fragment order
  const EllEntry* ell;   // (m+1) * Bp * W   (drive m = all-zero dummy)
  const double* Z;
  double* delta;
  double* jac;
};

constexpr int kDmmaMaxTiles = 8;
constexpr int kDmmaMaxW = 4;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// max that propagates NaN (fmax would drop it)
__device__ __forceinline__ double nan_max(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : fmax(a, b));
}

template <bool VEC>
__device__ __forceinline__ void store_pair(double* ptr, double v0, double v1, bool second_ok) {
  if (VEC) {
    *reinterpret_cast<double2*>(ptr) = make_double2(v0, v1);
  } else {
    ptr[0] = v0;
    if (second_ok) ptr[1] = v1;
  }
}

__host__ __device__ inline int dmma_ldy(int n_b) { return ((n_b + 5) / 8) * 8 + 2; }

// NT = padded generator size / 8 (1 or 2); NMT = 8-column tiles of T^T carried by the warp.
template <int NT, int NMT, bool VEC>
__global__ void __launch_bounds__(32) knot_dmma_kernel(DmmaParams p) {
  constexpr int KT = 2 * NT, Bp = 8 * NT;
  extern __shared__ __align__(16) unsigned char dmma_smem[];
  const int lane = threadIdx.x;
  const int g = lane >> 2, q = lane & 3;
  const int k = blockIdx.x;
  const int b = p.b, n_b = p.n_b, m = p.m, ncT = p.ncT, mj = p.m_jets;
  const int ldy = dmma_ldy(n_b);

  EllEntry* ell = reinterpret_cast<EllEntry*>(dmma_smem);
  const int n_ell = (mj > 0) ? (m + 1) * Bp * p.W : 0;
  double* ys = reinterpret_cast<double*>(dmma_smem + sizeof(EllEntry) * (size_t)n_ell);
  for (int e = lane; e < n_ell; e += 32) ell[e] = p.ell[e];

  const double* z = p.Z + (size_t)k * p.D;
  double dt = z[p.dt_off];

  // ---- G(u) = G0 + sum_j u_j G_j in permuted B-fragment order ---------------------------
  double Gu[KT][NT];
  {
    const double* gf = p.Gfrag + lane;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) Gu[kt][nt] = __ldg(gf + (kt * NT + nt) * 32);
    for (int j = 0; j < m; ++j) {
      const double uj = z[p.u_off + j];
      const double* gj = gf + (size_t)(1 + j) * KT * NT * 32;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) Gu[kt][nt] = fma(uj, __ldg(gj + (kt * NT + nt) * 32), Gu[kt][nt]);
    }
  }
  // ---- ||dt G||_1 -> Taylor degree M and number of sub-steps (warp-uniform) --------------
  double nrm = 0.0;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    double cs = 0.0;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) cs += fabs(Gu[kt][nt]);
    cs += __shfl_xor_sync(0xffffffffu, cs, 4);
    cs += __shfl_xor_sync(0xffffffffu, cs, 8);
    cs += __shfl_xor_sync(0xffffffffu, cs, 16);   // column sum of |G| for column pi(4 kt + q)
    nrm = nan_max(nrm, cs);
  }
  nrm = nan_max(nrm, __shfl_xor_sync(0xffffffffu, nrm, 1));
  nrm = nan_max(nrm, __shfl_xor_sync(0xffffffffu, nrm, 2));
  nrm *= fabs(dt);
  int n_sub = 1, M = 1;
  if (nrm > c_theta[kMaxDeg]) {
    const double ns = ceil(nrm / c_theta[kMaxDeg]);
    if (ns <= (double)p.max_sub) n_sub = (int)ns;
    else dt = __longlong_as_double(0x7ff8000000000000LL);  // norm beyond the supported range: NaN out
  }
  {
    const double per = nrm / (double)n_sub;
    while (M < kMaxDeg && c_theta[M] < per) ++M;
  }
  const double dts = dt / (double)n_sub;

  // ---- column bookkeeping: tile mt, lane column c = 8 mt + g -------------------------------
  // kind 0: identity column (propagator E)   1: state column (Y = E x)   2: jet of drive jd
  // 3: padding
  int kind[NMT], ccs[NMT], jd[NMT];
  bool tile_cpl[NMT], tile_y[NMT], tile_e[NMT];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt) {
    const int c = 8 * mt + g;
    kind[mt] = 3; ccs[mt] = 0; jd[mt] = m;
    if (c < ncT) {
      kind[mt] = 0; ccs[mt] = c;
    } else {
      const int rel = c - ncT, slab = rel / n_b;
      ccs[mt] = rel - slab * n_b;
      if (slab == 0) kind[mt] = 1;
      else if (slab <= mj) { kind[mt] = 2; jd[mt] = slab - 1; }
      else ccs[mt] = 0;
    }
    tile_e[mt] = (8 * mt < ncT);
    tile_y[mt] = (8 * mt + 7 >= ncT) && (8 * mt < ncT + n_b);
    tile_cpl[mt] = (mj > 0) && (8 * mt + 7 >= ncT + n_b) && (8 * mt < ncT + (1 + mj) * n_b);
  }

  // ---- base = [I | X | 0 ...] -------------------------------------------------------------------
  double t[NMT][KT], base[NMT][KT];   // index 2 nt + s  <->  row 8 nt + 2 q + s
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int r = 8 * (i >> 1) + 2 * q + (i & 1);
      double v = 0.0;
      if (kind[mt] == 0) v = (r == ccs[mt]) ? 1.0 : 0.0;
      else if (kind[mt] == 1 && r < b) v = z[p.x_off + ccs[mt] * b + r];
      base[mt][i] = v;
      t[mt][i] = v;
    }
  __syncwarp();  // ELL table visible

  // ---- Taylor action: n_sub sub-steps of a degree-M Horner recurrence ----------------------
  int par = 0;
  for (int sub = 0; sub < n_sub; ++sub) {
    for (int qd = M; qd >= 1; --qd) {
      const double coef = dts / (double)qd;
      double Gq[KT][NT];
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) Gq[kt][nt] = Gu[kt][nt] * coef;

      double* buf = ys + par * Bp * ldy;
      if (mj > 0) {
        // publish the current state columns Y for the sparse G_j Y terms
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
          if (tile_y[mt] && kind[mt] == 1) {
#pragma unroll
            for (int i = 0; i < KT; ++i)
              buf[(8 * (i >> 1) + 2 * q + (i & 1)) * ldy + ccs[mt]] = t[mt][i];
          }
        __syncwarp();
        par ^= 1;
      }
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt) {
        double d[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { d[nt][0] = base[mt][2 * nt]; d[nt][1] = base[mt][2 * nt + 1]; }
        if (tile_cpl[mt]) {
          const EllEntry* row = ell + (size_t)(jd[mt] * Bp + 2 * q) * p.W;
          const double* ycol = buf + (kind[mt] == 0 ? 0 : ccs[mt]);  // always an initialised column
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const EllEntry* e = row + (size_t)(8 * nt + s) * p.W;
              double acc = 0.0;
              for (int w = 0; w < p.W; ++w) {
                const EllEntry en = e[w];
                acc = fma(en.val, ycol[en.idx * ldy], acc);
              }
              d[nt][s] = fma(coef, acc, d[nt][s]);
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) dmma884(d[nt], t[mt][kt], Gq[kt][nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { t[mt][2 * nt] = d[nt][0]; t[mt][2 * nt + 1] = d[nt][1]; }
      }
    }
    if (sub + 1 < n_sub) {
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
        for (int i = 0; i < KT; ++i) base[mt][i] = t[mt][i];
    }
  }

  // ---- outputs straight from the accumulator fragments -----------------------------------------
  const int n_x = b * n_b;
  const int half = b >> 1;
  double* jac = p.jac ? p.jac + (size_t)k * p.nnz_jac : nullptr;
  double* dl = p.delta ? p.delta + (size_t)k * n_x : nullptr;
  const double* zn = z + p.D + p.x_off;

#pragma unroll
  for (int mt = 0; mt < NMT; ++mt) {
    if (tile_e[mt] && jac && kind[mt] == 0) {
      // -E, replicated for each of the n_b state columns (the d/dx_k block is I (x) E)
      const int c = ccs[mt];
      for (int copy = 0; copy < n_b; ++copy) {
        double* blk = jac + (size_t)copy * b * b;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int r = 8 * nt + 2 * q;
          if (r < b) {
            const double v0 = -t[mt][2 * nt], v1 = -t[mt][2 * nt + 1];
            store_pair<VEC>(blk + c * b + r, v0, v1, r + 1 < b);
            if (p.iso) {
              // E = [[P, -Q], [Q, P]] : column half + c from column c
              if (VEC) {   // half is even here, so the pair stays together
                const bool top = r < half;
                const int rr = top ? r + half : r - half;
                store_pair<true>(blk + (c + half) * b + rr, top ? v0 : -v0, top ? v1 : -v1, true);
              } else {
                double* col = blk + (c + half) * b;
                col[r < half ? r + half : r - half] = r < half ? v0 : -v0;
                if (r + 1 < b) col[r + 1 < half ? r + 1 + half : r + 1 - half] = r + 1 < half ? v1 : -v1;
              }
            }
          }
        }
      }
    }
    if (tile_y[mt] || tile_cpl[mt]) {
      if (kind[mt] == 1) {
        const int off = ccs[mt] * b;
        if (dl) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int r = 8 * nt + 2 * q;
            if (r < b) {
              const double x0 = zn[off + r], x1 = (r + 1 < b) ? zn[off + r + 1] : 0.0;
              store_pair<VEC>(dl + off + r, x0 - t[mt][2 * nt], x1 - t[mt][2 * nt + 1], r + 1 < b);
            }
          }
        }
      }
      if (jac && mj > 0) {
        if (kind[mt] == 2) {
          double* out = jac + (size_t)n_b * b * b + (size_t)jd[mt] * n_x + ccs[mt] * b;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int r = 8 * nt + 2 * q;
            if (r < b) store_pair<VEC>(out + r, -t[mt][2 * nt], -t[mt][2 * nt + 1], r + 1 < b);
          }
        }
      }
    }
    if (tile_y[mt] && jac) {
      // d/d dt = -G(u) E x  : one more (unscaled) generator product on the state tile
      double d[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { d[nt][0] = 0.0; d[nt][1] = 0.0; }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) dmma884(d[nt], t[mt][kt], Gu[kt][nt]);
      if (kind[mt] == 1) {
        double* out = jac + (size_t)n_b * b * b + (size_t)m * n_x + ccs[mt] * b;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int r = 8 * nt + 2 * q;
          if (r < b) {
            store_pair<VEC>(out + r, -d[nt][0], -d[nt][1], r + 1 < b);
            store_pair<VEC>(out + n_x + r, 1.0, 1.0, r + 1 < b);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side: fragment / ELL tables, support test, launch
// ------------------------------------------------------------------------------------------
struct DmmaPlan {
  bool ok = false;
  int NT = 0, Bp = 0, W = 1, iso = 0, ncT = 0;
  int tiles_full = 0, tiles_res = 0;
  std::vector<double> gfrag;
  std::vector<EllEntry> ell;
};

inline int dmma_perm(int kt, int q) { return 8 * (kt / 2) + 2 * q + (kt % 2); }

// G0, Gj: host, column-major b x b.  kind_iso: caller says the state uses the half layout.
inline DmmaPlan dmma_plan(int b, int n_b, int m, bool allow_iso, const double* G0, const double* Gj) {
  DmmaPlan pl;
  if (b < 1 || b > 16) return pl;
  pl.NT = b <= 8 ? 1 : 2;
  pl.Bp = 8 * pl.NT;
  const int KT = 2 * pl.NT, NT = pl.NT, Bp = pl.Bp;
  auto at = [&](int mat, int r, int c) -> double {
    if (r >= b || c >= b) return 0.0;
    const double* A = mat == 0 ? G0 : Gj + (size_t)(mat - 1) * b * b;
    return A[r + (size_t)c * b];
  };
  // real-isomorphism structure [[S, R], [-R, S]] of every matrix?
  bool iso = allow_iso && (b % 2 == 0);
  if (iso) {
    const int h = b / 2;
    for (int mat = 0; mat <= m && iso; ++mat)
      for (int i = 0; i < h && iso; ++i)
        for (int j = 0; j < h; ++j)
          if (at(mat, i, j) != at(mat, i + h, j + h) || at(mat, i, j + h) != -at(mat, i + h, j)) {
            iso = false;
            break;
          }
  }
  pl.iso = iso ? 1 : 0;
  pl.ncT = iso ? b / 2 : b;
  const int cols_full = pl.ncT + (1 + m) * n_b, cols_res = n_b;
  pl.tiles_full = (cols_full + 7) / 8;
  pl.tiles_res = (cols_res + 7) / 8;
  if (pl.tiles_full > kDmmaMaxTiles) return pl;
  // ELL width of the drive generators
  int W = 1;
  for (int j = 0; j < m; ++j)
    for (int r = 0; r < b; ++r) {
      int cnt = 0;
      for (int c = 0; c < b; ++c) cnt += at(1 + j, r, c) != 0.0;
      W = cnt > W ? cnt : W;
    }
  if (W > kDmmaMaxW) return pl;
  pl.W = W;
  pl.gfrag.assign((size_t)(m + 1) * KT * NT * 32, 0.0);
  for (int mat = 0; mat <= m; ++mat)
    for (int kt = 0; kt < KT; ++kt)
      for (int nt = 0; nt < NT; ++nt)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, q = lane & 3;
          pl.gfrag[((size_t)mat * KT * NT + kt * NT + nt) * 32 + lane] = at(mat, 8 * nt + g, dmma_perm(kt, q));
        }
  pl.ell.assign((size_t)(m + 1) * Bp * W, EllEntry{0.0, 0, 0});
  for (int j = 0; j < m; ++j)
    for (int r = 0; r < b; ++r) {
      int w = 0;
      for (int c = 0; c < b; ++c)
        if (at(1 + j, r, c) != 0.0) pl.ell[((size_t)j * Bp + r) * W + w++] = EllEntry{at(1 + j, r, c), c, 0};
    }
  pl.ok = true;
  return pl;
}

inline size_t dmma_smem_bytes(const DmmaPlan& pl, int n_b, int m, bool jets) {
  const size_t n_ell = jets ? (size_t)(m + 1) * pl.Bp * pl.W : 0;
  return sizeof(EllEntry) * n_ell + 2 * (size_t)pl.Bp * dmma_ldy(n_b) * sizeof(double);
}

template <int NT, bool VEC>
inline cudaError_t dmma_launch_nt(int tiles, const DmmaParams& p, int blocks, size_t smem, cudaStream_t st) {
  switch (tiles) {
#define PB2_CASE(N) \
  case N: knot_dmma_kernel<NT, N, VEC><<<blocks, 32, smem, st>>>(p); break;
    PB2_CASE(1) PB2_CASE(2) PB2_CASE(3) PB2_CASE(4) PB2_CASE(5) PB2_CASE(6) PB2_CASE(7) PB2_CASE(8)
#undef PB2_CASE
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

inline cudaError_t dmma_launch(int NT, int tiles, bool vec, const DmmaParams& p, int blocks, size_t smem,
                               cudaStream_t st) {
  if (NT == 1) return vec ? dmma_launch_nt<1, true>(tiles, p, blocks, smem, st)
                          : dmma_launch_nt<1, false>(tiles, p, blocks, smem, st);
  return vec ? dmma_launch_nt<2, true>(tiles, p, blocks, smem, st)
             : dmma_launch_nt<2, false>(tiles, p, blocks, smem, st);
}

}  // namespace pb2
