// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) for the operand
// layouts / N values the fused kernels use.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dynibar_b200/csrc -o profiles/scripts/mma_rate profiles/scripts/mma_rate.cu
//   profiles/scripts/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace dyn::tc;

struct Res { long long issue, done; };

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;            // LBO (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8-row group stride
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

template <int N, int MODE, int CE>  // MODE 0: no swizzle (kernel layout), 1: 128B swizzle; CE: commit every CE mmas
__global__ void __launch_bounds__(128, 1) rate_kernel(Res* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 131072) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1 << 20); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tslot;
  if (warp == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    const uint32_t idesc = idesc_bf16_f32(128, N);
    long long t0 = clock64(), t1 = 0;
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
          uint64_t da, db;
          if (MODE == 0) {
            da = smem_desc(a0 + ks * 4096u, 2048u, 128u);
            db = smem_desc(b0 + ks * 2u * (N * 16u), N * 16u, 128u);
          } else {
            da = desc_sw128(a0 + (ks >> 2) * 16384u + (ks & 3) * 32u);
            db = desc_sw128(b0 + (ks >> 2) * (N * 128u) + (ks & 3) * 32u);
          }
          mma_bf16_ss(tm + (r & 1) * 256, da, db, idesc, ks ? 1u : 0u);
          if (CE && (ks % (CE ? CE : 1)) == CE - 1) mma_commit(smem_u32(&bar2));
        }
      }
      mma_commit(smem_u32(&bar));
    }
    __syncwarp();
    t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (tid == 0) { out[blockIdx.x].issue = t1 - t0; out[blockIdx.x].done = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tm, 512); }
}

template <int N, int MODE, int CE = 0>
void run(const char* name, int grid, int reps) {
  Res* d; cudaMalloc(&d, sizeof(Res) * grid);
  const int smem = 65536 + 131072;
  cudaFuncSetAttribute(rate_kernel<N, MODE, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int i = 0; i < 2; ++i) rate_kernel<N, MODE, CE><<<grid, 128, smem>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  Res h[148]; cudaMemcpy(h, d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
  double si = 0, sd = 0; for (int i = 0; i < grid; ++i) { si += h[i].issue; sd += h[i].done; }
  const int n = reps * 16;
  printf("%-28s grid %3d  mmas %4d  issue %7.1f cyc/mma  complete %7.1f cyc/mma  (floor %d)  %s\n", name, grid, n,
         si / grid / n, sd / grid / n, N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {148}) {
    run<256, 0>("N=256 no-swizzle", grid, 16);
    run<256, 1>("N=256 swizzle-128B", grid, 16);
    run<128, 0>("N=128 no-swizzle", grid, 16);
    run<128, 1>("N=128 swizzle-128B", grid, 16);
    run<48, 0>("N=48 no-swizzle", grid, 16);
    run<48, 1>("N=48 swizzle-128B", grid, 16);
    run<256, 0>("N=256 no-swizzle x64", grid, 64);
    run<256, 0, 2>("N=256 commit every 2", grid, 16);
    run<128, 0, 4>("N=128 commit every 4", grid, 16);
    run<128, 0, 1>("N=128 commit every 1", grid, 16);
  }
  return 0;
}
