// Dependent-issue latency of FP64 CUDA-core ops on sm_100a, alone and while other warps of the same
// sub-partition keep the FP64 tensor pipe busy (what a Horner step's epilogue sees).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_lat tools/dfma_lat.cu && tools/dfma_lat
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// warp 0 (and every warp with id % 4 == 0 .. `probes`) runs a dependent chain of `kind`;
// the other warps issue independent DMMAs back to back (load on the pipe).
// kind 0: DFMA chain, 1: DADD chain, 2: DMUL chain, 3: LDS -> DFMA chain, 4: DMMA accumulate chain -> DFMA
__global__ void k_lat(double* out, long long* cyc, double a, double b, int N, int kind, int loaders) {
  __shared__ double sm[64];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) sm[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
  __syncthreads();
  double v = 1.0 + lane * 1e-3, c0 = lane, c1 = 1.0;
  if (w == 0) {
    long long t0 = clock64();
    if (kind == 0) for (int i = 0; i < N; ++i) v = fma(v, a, b);
    if (kind == 1) for (int i = 0; i < N; ++i) v = v + b;
    if (kind == 2) for (int i = 0; i < N; ++i) v = v * a;
    if (kind == 3) for (int i = 0; i < N; ++i) { int idx = (int)v & 31; v = fma(sm[idx], a, b); }
    if (kind == 4) for (int i = 0; i < N; ++i) { dmma884(c0, c1, a, b); c0 = fma(c0, a, b); }
    long long t1 = clock64();
    if (lane == 0) *cyc = t1 - t0;
  } else if (w <= loaders) {
    double d[4][2] = {{1, 2}, {3, 4}, {5, 6}, {7, 8}};
    for (int i = 0; i < 4 * N; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(d[j][0], d[j][1], a, b);
    c0 = d[0][0] + d[1][1] + d[2][0] + d[3][1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + c0 + c1;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 4096); cudaMallocManaged(&cyc, 8);
  const int N = 2000;
  const char* names[] = {"DFMA chain", "DADD chain", "DMUL chain", "LDS->DFMA chain", "DMMA->DFMA chain"};
  // loaders: 0 = idle SM; 4 = one DMMA-issuing warp per sub-partition *other* than... warps 1..4 land on
  // sub-partitions 1,2,3,0: with 4 loaders the probe's own sub-partition (0) hosts warp 4; 8 / 12 = 2 / 3 per sub-partition
  for (int loaders : {0, 3, 4, 8, 12}) {
    for (int kind = 0; kind < 5; ++kind) {
      k_lat<<<1, 32 * (1 + (loaders ? loaders : 0))>>>(out, cyc, 1.0000001, 1e-9, N, kind, loaders);
      cudaDeviceSynchronize();
      printf("loaders=%2d  %-18s %8.1f cycles per link\n", loaders, names[kind], (double)*cyc / N);
    }
  }
  return 0;
}
