#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
// CH independent chains, each N deep, accumulate-dependent (C chain)
template <int CH>
__global__ void k_c(double* out, long long* cyc, double a, double b, int N) {
  double c[CH][2];
  for (int j = 0; j < CH; ++j) c[j][0] = c[j][1] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < CH; ++j) dmma884(c[j][0], c[j][1], a, b);
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < CH; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// A-dependent chain: result feeds next A operand (as in Horner feed)
__global__ void k_a(double* out, long long* cyc, double a, double b, int N) {
  double c0 = threadIdx.x, c1 = 1;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) { double d0 = 0, d1 = 0; dmma884(d0, d1, c0, b); c0 = d0; c1 = d1; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void k_big(double* out, long long* cyc, double a, double b, int N) {
  double c[CH][4], av[8] = {a, b, a, b, a, b, a, b}, bv[4] = {b, a, b, a};
  for (int j = 0; j < CH; ++j) for (int q = 0; q < 4; ++q) c[j][q] = threadIdx.x + j + q;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < CH; ++j) dmma16816(c[j], av, bv);
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < CH; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 1024); cudaMallocManaged(&cyc, 8);
  const int N = 1000;
  auto rep = [&](const char* nm, int warps, int ch, double per) {
    cudaDeviceSynchronize();
    printf("%-28s warps/SM=%2d chains=%d : %8.2f cyc per dmma-slot ; SM rate %.3f cyc/inst\n", nm, warps, ch, (double)*cyc / N, (double)*cyc / N / (warps * ch) );
  };
  for (int w : {1, 4, 8, 12, 16, 32}) {
    k_c<1><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("884 C-chain", w, 1, 0);
    k_c<2><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("884 C-chain", w, 2, 0);
    k_c<4><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("884 C-chain", w, 4, 0);
    k_c<8><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("884 C-chain", w, 8, 0);
  }
  k_a<<<1, 32>>>(out, cyc, 1.0, 1e-9, N); rep("884 A-chain", 1, 1, 0);
  for (int w : {1, 4, 8, 12}) {
    k_big<1><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("16816 C-chain", w, 1, 0);
    k_big<2><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("16816 C-chain", w, 2, 0);
    k_big<4><<<1, 32 * w>>>(out, cyc, 1.0, 1e-9, N); rep("16816 C-chain", w, 4, 0);
  }
  return 0;
}
