#include "linear_f32.cuh"

namespace dyn {

namespace {
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_ELU: return elu_f(v);
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(256) linear_f32_kernel(const __grid_constant__ LinArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += BK) {
    // A tile: element e -> (m = e / BK, k = e % BK)
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      int e = tid + i * 256;
      int k = e % BK, m = e / BK;
      long long row = m0 + m;
      int col = k0 + k;
      float v = 0.f;
      if (row < a.M && col < a.K) {
        int c = col;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s < a.nseg) {
            if (c >= 0 && c < a.seg[s].width)
              v = a.seg[s].p[(row / a.seg[s].div) * a.seg[s].ld + c];
            c -= a.seg[s].width;
          }
        }
        if (a.row_scale != nullptr) v *= a.row_scale[row];
      }
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      int e = tid + i * 256;
      int k = e % BK, n = e / BK;
      int col = k0 + k, on = n0 + n;
      Ws[k][n] = (on < a.N && col < a.K) ? a.W[(long long)on * a.K + col] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Ws[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long row = m0 + ty * TM + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int col = n0 + tx * TN + j;
      if (col >= a.N) continue;
      float v = acc[i][j] + (a.b ? a.b[col] : 0.f);
      a.Y[row * a.ldy + col] = apply_act(v, a.act);
    }
  }
}
}  // namespace

int launch_linear(const LinArgs& a, cudaStream_t st) {
  if (a.M == 0) return DYN_OK;
  int ksum = 0;
  for (int s = 0; s < a.nseg; ++s) ksum += a.seg[s].width;
  if (ksum != a.K) return fail(DYN_E_INVALID, "linear: segment widths %d != K %d", ksum, a.K);
  dim3 grid(cdiv(a.M, BM), cdiv(a.N, BN));
  linear_f32_kernel<<<grid, 256, 0, st>>>(a);
  DYN_LAUNCH_CHECK();
  return DYN_OK;
}

}  // namespace dyn
