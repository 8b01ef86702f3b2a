// FP64 pipe microbenchmark for B200 (sm_100a): DFMA, DMMA m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16,
// and DFMA+DMMA interleaved.  Used only to establish the FP64 ceiling the knot kernels are
// measured against (the driver's MEASURED_PEAKS.json has no FP64 figure).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_dmma884(double* out, double a, double b) {
  double c[8][2];
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = threadIdx.x + j;
#pragma unroll 2
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1684(double* out, double a, double b) {
  double c[4][4];
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[j][q] = threadIdx.x + j + q;
#pragma unroll 2
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dmma1684(c[j], a, b, a);
  double s = 0;
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1688(double* out, double a, double b) {
  double c[4][4], av[4] = {a, b, a, b};
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[j][q] = threadIdx.x + j + q;
#pragma unroll 2
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dmma1688(c[j], av, a, b);
  double s = 0;
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, double a, double b) {
  double c[4][4], av[8] = {a, b, a, b, a, b, a, b}, bv[4] = {b, a, b, a};
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) c[j][q] = threadIdx.x + j + q;
#pragma unroll 2
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dmma16816(c[j], av, bv);
  double s = 0;
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// interleave: per iteration 4 DMMA m8n8k4 (1024 FMA) + 8 DFMA/thread (256 FMA per warp)
__global__ void k_mixed(double* out, double a, double b) {
  double c[4][2];
  for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = threadIdx.x + j;
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 2
  for (int i = 0; i < ITERS; ++i) {
    dmma884(c[0][0], c[0][1], a, b); x0 = fma(x0, a, b); x1 = fma(x1, a, b);
    dmma884(c[1][0], c[1][1], a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    dmma884(c[2][0], c[2][1], a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b);
    dmma884(c[3][0], c[3][1], a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount, blocks = sms * 4, threads = 512;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const double warps = (double)blocks * threads / 32;
  struct { const char* name; double fma_per_warp_iter; double ms; } r[7];
  int n = 0;
  r[n++] = {"dfma", 8.0 * 32, time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); })};
  r[n++] = {"dmma.m8n8k4", 8.0 * 256, time_ms([&] { k_dmma884<<<blocks, threads>>>(out, 1.0000001, 1e-9); })};
  r[n++] = {"dmma.m16n8k4", 4.0 * 512, time_ms([&] { k_dmma1684<<<blocks, threads>>>(out, 1.0000001, 1e-9); })};
  r[n++] = {"dmma.m16n8k8", 4.0 * 1024, time_ms([&] { k_dmma1688<<<blocks, threads>>>(out, 1.0000001, 1e-9); })};
  r[n++] = {"dmma.m16n8k16", 4.0 * 2048, time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, 1.0000001, 1e-9); })};
  r[n++] = {"mixed(4 dmma884 + 8 dfma)", 4.0 * 256 + 8.0 * 32, time_ms([&] { k_mixed<<<blocks, threads>>>(out, 1.0000001, 1e-9); })};
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [", pr.name, sms);
  for (int i = 0; i < n; ++i) {
    double fl = 2.0 * r[i].fma_per_warp_iter * ITERS * warps;
    printf("%s{\"kernel\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f}", i ? ", " : "", r[i].name, r[i].ms, fl / r[i].ms / 1e9);
  }
  printf("]}\n");
  return 0;
}
