// Hermitian (ket / unitary) fast path -- placeholder until the complex kernels land.
#pragma once
#include <cuda_runtime.h>
#include "knot_generic.cuh"

namespace pb2 {
struct HermitianConsts { int unused; };
inline bool hermitian_supported(int, int, int, const double*, const double*) { return false; }
inline cudaError_t hermitian_setup(int, int, int, const double*, const double*, HermitianConsts**) {
  return cudaErrorNotSupported;
}
inline cudaError_t launch_hermitian_resjac(const KnotParams&, const HermitianConsts*, cudaStream_t) {
  return cudaErrorNotSupported;
}
inline cudaError_t launch_hermitian_hess(const KnotParams&, const HermitianConsts*, cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace pb2
