// fp32 products of the training backward (motion_train.cu owns the kernels; nets_train.cu shares them).
#pragma once
#include "common.cuh"

namespace dyn {

// C[m, n] (+)= sum_k A(m, k) B(k, n) s(k);  A(m, k) = A[m * sam + k * sak],
// B(k, n) = B[(k / bdiv) * sbk + n * sbn] (bdiv > 1: one B row serves `bdiv` consecutive k -- a per-point
// tensor against per-(point, view) rows), s(k) = kscale ? kscale[k] : 1 (the row scale of a layer input).
// gridDim.z splits K; with more than one split (or accumulate) the tile is added with atomicAdd.
struct GemmArgs {
  const float *A, *B;
  float* C;
  long long M, N, K;
  long long sam, sak, sbk, sbn;
  long long ldc;
  int accumulate;
  long long k_per_split;
  long long bdiv = 1;
  const float* kscale = nullptr;
};

int launch_gemm(GemmArgs a, bool split_k, cudaStream_t st);
// db[c] += sum_r g[r * ldg + c], c < width
int launch_colsum(const float* g, long long ldg, int width, long long N, float* db, cudaStream_t st);

}  // namespace dyn
