// Internal declarations of the dense helpers shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200lm {
// csrc/dense_linalg.cu
cudaError_t potrf(int n, const double* A, int lda, double shift, double* L, int ldl, double* d_linv, int* d_info,
                  cudaStream_t s);
cudaError_t trsm(int n, int nrhs, const double* L, int ldl, const double* d_linv, int trans, double* B, int ldb,
                 double* X, int ldx, cudaStream_t s);
// csrc/bootstrap.cu
cudaError_t normals(long long g0, long long count, unsigned long long seed, double* d_z, uint32_t* d_raw, cudaStream_t s);
}  // namespace b200lm
