// Register-pressure probe for the single-round kernel of notes/next_round/v3_single_round_kernel.md:
// the Horner step loop of a compute warp that holds THREE tiles, in the accumulator-preload form, with
// the same launch bounds the real kernel would have (512 threads, one CTA per SM => 128 registers).
// Compile only (no GPU needed):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -c tools/v3_step_probe.cu -o /dev/null
// Run (next round): cycles per step with 14 such warps per SM against the 1 344 / 1 536-cycle pipe floor.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int OFF>
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void mma_acc(double (&d)[2][2], const double (&t)[4], const double (&A)[4][2]) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    dmma884(d[0], t[kt], A[kt][0]);
    dmma884(d[1], t[kt], A[kt][1]);
  }
}

struct ProbeParams {
  int M, n_slots;
  const double* Z;      // fake inputs
  double* out;
  long long* cycles;
};

// one jet tile: coupling term (ELL value re-read from shared memory, W = 1) into the accumulator, product
template <int PAR>
__device__ __forceinline__ void jet_tile(double (&t)[4], const double (&A)[4][2], const uint32_t (&yad)[4], uint32_t evad) {
  double d[2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i >> 1][i & 1] = lds_f64<0>(evad + 8 * i) * lds_f64<PAR * 1024>(yad[i]);
  mma_acc(d, t, A);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = d[i >> 1][i & 1];
}

// warp A: tiles (E, X, J1).  warp B: tiles (J2, J3, J4).
template <int PAR>
__device__ __forceinline__ void step_A(double (&tE)[4], double (&tX)[4], double (&tJ)[4], const double (&bX)[4],
                                       const double (&A)[4][2], int iE, uint32_t ypub, uint32_t ck_next, int bar,
                                       double (&accX)[4], double& cE, const uint32_t (&yad)[4], uint32_t evad) {
#pragma unroll
  for (int i = 0; i < 4; ++i) sts_f64<PAR * 1024>(ypub + i * 256, tX[i]);
  bar_sync(bar, 64);
  {
    double d[2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i >> 1][i & 1] = accX[i];
    mma_acc(d, tX, A);
#pragma unroll
    for (int i = 0; i < 4; ++i) tX[i] = d[i >> 1][i & 1];
  }
  {
    double d[2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i >> 1][i & 1] = (i == iE) ? cE : 0.0;
    mma_acc(d, tE, A);
#pragma unroll
    for (int i = 0; i < 4; ++i) tE[i] = d[i >> 1][i & 1];
  }
  const double ckn = lds_f64<0>(ck_next);
#pragma unroll
  for (int i = 0; i < 4; ++i) accX[i] = ckn * bX[i];
  cE = ckn;
  jet_tile<PAR>(tJ, A, yad, evad);
}

template <int PAR>
__device__ __forceinline__ void step_B(double (&t)[3][4], const double (&A)[4][2], int bar, const uint32_t (&yad)[3][4],
                                       uint32_t evad) {
  bar_sync(bar, 64);
#pragma unroll
  for (int a = 0; a < 3; ++a) jet_tile<PAR>(t[a], A, yad[a], evad + 32 * a);
}

__global__ void __launch_bounds__(512, 1) v3_step_probe(const __grid_constant__ ProbeParams p) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int slot = w >> 1, role = w & 1, g = lane >> 2, q = lane & 3;
  if (slot >= p.n_slots) return;
  const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(smem) + 8u * (uint32_t)(slot * 1024);
  for (int e = lane; e < 1024; e += 32) smem[slot * 1024 + e] = p.Z[(blockIdx.x * 16 + slot) * 1024 + e];
  __syncwarp();
  bar_sync(1 + slot, 64);
  const uint32_t a_prep = a0, a_c = a0 + 8u * 256u, a_y = a0 + 8u * 320u, a_ev = a0 + 8u * 600u;
  double A[4][2];
#pragma unroll
  for (int s = 0; s < 8; ++s) A[s >> 1][s & 1] = lds_f64<0>(a_prep + 8u * (uint32_t)(s * 32 + lane));
  const int M = p.M;
  long long t0 = clock64();
  double res = 0.0;
  if (role == 0) {
    const int iE = (g == 2 * q) ? 0 : ((g == 2 * q + 1) ? 1 : -1);
    double bX[4], tE[4], tX[4], tJ[4], accX[4], cE;
    uint32_t yad[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bX[i] = lds_f64<0>(a0 + 8u * (uint32_t)(700 + 4 * lane + i));
      tX[i] = bX[i];
      tE[i] = (i == iE) ? 1.0 : 0.0;
      tJ[i] = 0.0;
      yad[i] = a_y + 8u * (uint32_t)(((lane * 7 + i * 3) & 31) + 32 * i);
    }
    cE = lds_f64<0>(a_c + 8u * (uint32_t)(M - 1));
#pragma unroll
    for (int i = 0; i < 4; ++i) accX[i] = cE * bX[i];
    const uint32_t ypub = a_y + 8u * (uint32_t)(g * 4 + q);
    int kq = M - 1;
    for (; kq >= 1; kq -= 2) {
      step_A<0>(tE, tX, tJ, bX, A, iE, ypub, a_c + 8u * kq - 8u, 1 + slot, accX, cE, yad, a_ev + 8u * (uint32_t)(4 * lane));
      step_A<1>(tE, tX, tJ, bX, A, iE, ypub, a_c + 8u * (uint32_t)(kq >= 2 ? kq - 2 : 0), 1 + slot, accX, cE, yad,
                a_ev + 8u * (uint32_t)(4 * lane));
    }
    if (kq == 0) step_A<0>(tE, tX, tJ, bX, A, iE, ypub, a_c, 1 + slot, accX, cE, yad, a_ev + 8u * (uint32_t)(4 * lane));
#pragma unroll
    for (int i = 0; i < 4; ++i) res += tE[i] + tX[i] + tJ[i];
  } else {
    double t[3][4];
    uint32_t yad[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        t[a][i] = 0.0;
        yad[a][i] = a_y + 8u * (uint32_t)(((lane * 5 + i * 3 + a) & 31) + 32 * i);
      }
    int kq = M - 1;
    for (; kq >= 1; kq -= 2) {
      step_B<0>(t, A, 1 + slot, yad, a_ev + 8u * (uint32_t)(128 + 12 * lane));
      step_B<1>(t, A, 1 + slot, yad, a_ev + 8u * (uint32_t)(128 + 12 * lane));
    }
    if (kq == 0) step_B<0>(t, A, 1 + slot, yad, a_ev + 8u * (uint32_t)(128 + 12 * lane));
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 4; ++i) res += t[a][i];
  }
  long long t1 = clock64();
  p.out[blockIdx.x * 512 + threadIdx.x] = res;
  if (blockIdx.x == 0 && lane == 0) p.cycles[w] = t1 - t0;
}

int main() {
  const int M = 8, slots = 7, blocks = 148;
  ProbeParams p{};
  p.M = M; p.n_slots = slots;
  double* Z; cudaMalloc(&Z, sizeof(double) * blocks * 16 * 1024);
  cudaMemset(Z, 0, sizeof(double) * blocks * 16 * 1024);
  cudaMalloc(&p.out, sizeof(double) * blocks * 512);
  cudaMallocManaged(&p.cycles, sizeof(long long) * 16);
  p.Z = Z;
  cudaFuncSetAttribute(v3_step_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 8);
  for (int rep = 0; rep < 3; ++rep) {
    v3_step_probe<<<blocks, 512, 16 * 1024 * 8>>>(p);
    cudaDeviceSynchronize();
  }
  printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int w = 0; w < 2 * slots; ++w) printf("warp %2d (%s): %lld cycles for %d steps = %.0f per step\n", w, (w & 1) ? "J2-J4" : "E,X,J1",
                                               p.cycles[w], M, (double)p.cycles[w] / M);
  printf("pipe floor per step: 7 slots x 48 DMMA x 16 / 4 = 1344 cycles (1536 on a sub-partition with 4 compute warps)\n");
  return 0;
}
