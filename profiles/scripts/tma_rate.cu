// Micro-benchmark: sustained cp.async.bulk (global/L2 -> shared) rate per SM as a function of the
// number of 16 KB copies in flight and of the CTAs per SM, all SMs streaming the same 384 KB
// L2-resident buffer (the weight-streaming pattern of the fused kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dynibar_b200/csrc -o profiles/scripts/tma_rate profiles/scripts/tma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace dyn::tc;

// `nprod` producer threads (lane 0 of warps 0 .. nprod-1), each with its own `slots` ring slots
__global__ void __launch_bounds__(128) rate_kernel(const uint8_t* src, int src_chunks, int slots, int chunk_bytes,
                                                   int n_copies, long long* out, int nprod) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < slots * nprod; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if ((tid & 31) == 0 && (tid >> 5) < nprod) {
    const int p = tid >> 5;
    uint64_t* mybars = bars + p * slots;
    uint8_t* mysmem = smem + (size_t)p * slots * chunk_bytes;
    long long t0 = clock64();
    for (int c = 0; c < n_copies + slots; ++c) {
      const int s = c % slots;
      if (c >= slots) mbar_wait(smem_u32(&mybars[s]), ((c / slots) - 1) & 1);  // copy c - slots has landed
      if (c < n_copies) {
        mbar_arrive_expect_tx(smem_u32(&mybars[s]), chunk_bytes);
        bulk_g2s(smem_u32(mysmem + (size_t)s * chunk_bytes),
                 src + (size_t)((c + blockIdx.x + 7 * p) % src_chunks) * chunk_bytes, chunk_bytes, smem_u32(&mybars[s]));
      }
    }
    if (p == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int chunk = 16384, src_chunks = 24;
  uint8_t* src; cudaMalloc(&src, (size_t)chunk * src_chunks); cudaMemset(src, 1, (size_t)chunk * src_chunks);
  long long* out; cudaMalloc(&out, sizeof(long long) * 296);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * chunk);
  const int n = 400;
  for (int cps : {1, 2}) {
    for (int slots : {1, 2, 3, 4, 6, 8}) {
      if (cps == 2 && slots > 6) continue;
      const int grid = 148 * cps;
      const int smem = cps == 2 ? 6 * chunk : 8 * chunk;  // pins the CTAs-per-SM count
      for (int rep = 0; rep < 2; ++rep) rate_kernel<<<grid, 128, smem>>>(src, src_chunks, slots, chunk, n, out, 1);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[296]; cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
      const double cyc = s / grid;
      printf("CTAs/SM %d  copies in flight/CTA %d  : %6.1f B/clk per CTA, %6.1f B/clk per SM, %6.0f cycles per 16 KB copy round trip  %s\n",
             cps, slots, (double)n * chunk / cyc, (double)n * chunk / cyc * cps, cyc / n * slots, cudaGetErrorString(e));
    }
  }
  // one CTA per SM, several producer threads (different warps), 2-4 copies in flight each
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * chunk);
  for (int nprod : {1, 2, 3, 4}) {
    for (int slots : {2, 3}) {
      for (int rep = 0; rep < 2; ++rep) rate_kernel<<<148, 128, 12 * chunk>>>(src, src_chunks, slots, chunk, n, out, nprod);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, out, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      const double cyc = s / 148;
      printf("1 CTA/SM, %d producer threads x %d copies in flight: %6.1f B/clk per SM  %s\n", nprod, slots,
             (double)n * chunk * nprod / cyc, cudaGetErrorString(e));
    }
  }
  // smaller copies from one producer: 8 x 2 KB pieces per 16 KB stage (same bytes, more requests)
  return 0;
}
