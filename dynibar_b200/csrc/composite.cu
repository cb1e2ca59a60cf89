// a12 alpha compositing (raw2outputs / raw2outputs_vanilla, render_ray.py:134-330)
// and a13 hierarchical resampling (sample_pdf + merge, render_ray.py:19-64,
// :790-819).  One warp per ray; the transmittance product is a warp-level
// inclusive scan.  All fp32.
#include "common.cuh"

namespace dyn {

// inclusive multiplicative scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// alpha = 1 - exp(-softplus(sigma) * delta), delta = 1 except last = 1e10
// (render_ray.py:154-184, USE_DISTANCE=False)
__device__ __forceinline__ float alpha_of(float sigma, bool last) {
  return 1.f - expf(-softplus_f(sigma) * (last ? 1e10f : 1.f));
}

// number of views that see a sample
__device__ __forceinline__ float views_seen(const float* mask, long long p, int V) {
  float s = 0.f;
  for (int v = 0; v < V; ++v) s += mask[p * V + v];
  return s;
}

// composite == true : raw2outputs (dynamic + static)
// composite == false: raw2outputs_vanilla on raw_a alone
template <bool kComposite>
__global__ void composite_kernel(const float* __restrict__ raw_a, const float* __restrict__ raw_b,
                                 const float* __restrict__ z_vals, const float* __restrict__ mask_a,
                                 int V_a, int min_a, const float* __restrict__ mask_b, int V_b,
                                 int min_b, int R, int S, float* __restrict__ out_rays,
                                 float* __restrict__ out_samples) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const long long RS = (long long)R * S;
  float carry = 1.f;  // T at the start of this 32-sample segment
  float acc_a[3] = {0.f, 0.f, 0.f}, acc_b[3] = {0.f, 0.f, 0.f};
  float depth = 0.f, cnt_a = 0.f, cnt_b = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool ok = s < S;
    const long long p = (long long)r * S + (ok ? s : 0);
    float4 ra = ok ? reinterpret_cast<const float4*>(raw_a)[p] : make_float4(0, 0, 0, 0);
    float al_a = ok ? alpha_of(ra.w, s == S - 1) : 0.f;
    float al_b = 0.f;
    float4 rb = make_float4(0, 0, 0, 0);
    float al = al_a;
    if (kComposite) {
      rb = ok ? reinterpret_cast<const float4*>(raw_b)[p] : make_float4(0, 0, 0, 0);
      al_b = ok ? alpha_of(rb.w, s == S - 1) : 0.f;
      al = 1.f - (1.f - al_b) * (1.f - al_a);  // render_ray.py:287
    }
    float f = ok ? (1.f - al + 1e-10f) : 1.f;
    float inc = warp_scan_mul(f, lane);
    float excl = __shfl_up_sync(0xffffffffu, inc, 1);
    float T = carry * (lane == 0 ? 1.f : excl);
    carry *= __shfl_sync(0xffffffffu, inc, 31);
    if (ok) {
      float z = z_vals[p];
      float w = al * T;
      depth += w * z;
      if (kComposite) {
        float w_a = al_a * T, w_b = al_b * T;
        acc_a[0] += w_a * ra.x; acc_a[1] += w_a * ra.y; acc_a[2] += w_a * ra.z;
        acc_b[0] += w_b * rb.x; acc_b[1] += w_b * rb.y; acc_b[2] += w_b * rb.z;
        out_samples[0 * RS + p] = al_a;
        out_samples[1 * RS + p] = w_a;
        out_samples[2 * RS + p] = w_b;
        out_samples[3 * RS + p] = al;
        out_samples[4 * RS + p] = w;
        cnt_b += views_seen(mask_b, p, V_b) > (float)min_b ? 1.f : 0.f;
      } else {
        acc_a[0] += w * ra.x; acc_a[1] += w * ra.y; acc_a[2] += w * ra.z;
        out_samples[0 * RS + p] = w;
        out_samples[1 * RS + p] = al;
      }
      cnt_a += views_seen(mask_a, p, V_a) > (float)min_a ? 1.f : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) { acc_a[i] = warp_sum(acc_a[i]); acc_b[i] = warp_sum(acc_b[i]); }
  depth = warp_sum(depth);
  cnt_a = warp_sum(cnt_a);
  cnt_b = warp_sum(cnt_b);
  if (lane == 0) {
    if (kComposite) {
      float* o = out_rays + (long long)r * 11;
      for (int i = 0; i < 3; ++i) {
        o[i] = acc_a[i] + acc_b[i];  // rgb = rgb_dy + rgb_static (render_ray.py:306)
        o[3 + i] = acc_b[i];
        o[6 + i] = acc_a[i];
      }
      o[9] = depth;
      o[10] = (cnt_a > 8.f || cnt_b > 8.f) ? 1.f : 0.f;  // :311-313
    } else {
      float* o = out_rays + (long long)r * 5;
      for (int i = 0; i < 3; ++i) o[i] = acc_a[i];
      o[3] = depth;
      o[4] = cnt_a > 8.f ? 1.f : 0.f;  // :197-199
    }
  }
}

// ---------------------------------------------------------------------------
// a13: one warp per ray.  Everything stays in shared memory: pdf/cdf of the
// S-2 interior weights, inverse-CDF lookup by counting, then a rank-merge of
// the (sorted) coarse depths with the (monotone-in-u only when det) fine
// depths -- general u needs a real sort, so the S+Ni values are sorted with a
// warp-cooperative rank sort (stable w.r.t. value, ties by index).
// ---------------------------------------------------------------------------
__global__ void resample_kernel(const float* __restrict__ z_vals, const float* __restrict__ weights,
                                const float* __restrict__ u_in, int R, int S, int Ni, int inv_uniform,
                                float* __restrict__ z_out) {
  extern __shared__ float smem[];
  const int wpb = blockDim.x >> 5;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * wpb + wid;
  const int M = S - 2;           // number of pdf bins
  const int stride = 2 * (S) + (S + Ni);
  float* bins = smem + wid * stride;  // [M+1]
  float* cdf = bins + S;              // [M+1]
  float* all = cdf + S;               // [S+Ni]
  if (r >= R) return;
  const float* z = z_vals + (long long)r * S;
  const float* w = weights + (long long)r * S;
  // bins: mid points (of 1/z, flipped, when inv_uniform) -- render_ray.py:793-803
  for (int i = lane; i < M + 1; i += 32) {
    if (inv_uniform) {
      int j = M - i;  // flip
      bins[i] = 0.5f * (1.f / z[j + 1] + 1.f / z[j]);
    } else {
      bins[i] = 0.5f * (z[i + 1] + z[i]);
    }
  }
  // pdf = (w + 1e-5) / sum ; cdf = [0, cumsum(pdf)]
  float part = 0.f;
  for (int i = lane; i < M; i += 32) {
    int j = inv_uniform ? (M - 1 - i) : i;
    part += w[1 + j] + 1e-5f;
  }
  float tot = warp_sum(part);
  __syncwarp();
  if (lane == 0) {
    // sequential cumsum in the reference's order (torch.cumsum over M <= 126 entries)
    float c = 0.f;
    cdf[0] = 0.f;
    for (int i = 0; i < M; ++i) {
      int j = inv_uniform ? (M - 1 - i) : i;
      c += (w[1 + j] + 1e-5f) / tot;
      cdf[i + 1] = c;
    }
  }
  __syncwarp();
  for (int k = lane; k < Ni; k += 32) {
    float u;
    if (u_in != nullptr) u = u_in[(long long)r * Ni + k];
    else {
      // torch.linspace(0,1,Ni): first half from start, second half from end
      float step = 1.f / (float)(Ni - 1);
      u = k < Ni / 2 ? step * (float)k : 1.f - step * (float)(Ni - 1 - k);
    }
    int above = 0;
    for (int i = 0; i < M; ++i) above += (u >= cdf[i]) ? 1 : 0;  // first M entries only (:38-39)
    int below = above - 1 < 0 ? 0 : above - 1;
    float c0 = cdf[below], c1 = cdf[above];
    float b0 = bins[below], b1 = bins[above];
    float den = c1 - c0;
    if (den < 1e-5f) den = 1.f;
    float t = (u - c0) / den;
    float smp = b0 + t * (b1 - b0);
    all[S + k] = inv_uniform ? 1.f / smp : smp;
  }
  for (int i = lane; i < S; i += 32) all[i] = z[i];
  __syncwarp();
  // rank sort of S+Ni values (<= 256): rank = #smaller + #equal with lower index
  const int n = S + Ni;
  for (int i = lane; i < n; i += 32) {
    float v = all[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      float o = all[j];
      rank += (o < v || (o == v && j < i)) ? 1 : 0;
    }
    z_out[(long long)r * n + rank] = v;
  }
}

}  // namespace dyn

using namespace dyn;

extern "C" {

int dyn_composite(const float* raw_dy, const float* raw_st, const float* z_vals, const float* mask_dy,
                  int V_dy, int min_views_dy, const float* mask_st, int V_st, int min_views_st, int R,
                  int S, float* out_rays, float* out_samples, void* stream) {
  if (R == 0) return DYN_OK;  // empty batch: nothing to do (pointers may be null)
  DYN_CHECK_ARG(raw_dy && raw_st && z_vals && mask_dy && mask_st && out_rays && out_samples);
  DYN_CHECK_ARG(R >= 0 && S >= 1 && V_dy >= 1 && V_st >= 1);
  if (R == 0) return DYN_OK;
  composite_kernel<true><<<cdiv((long long)R * 32, 128), 128, 0, (cudaStream_t)stream>>>(
      raw_dy, raw_st, z_vals, mask_dy, V_dy, min_views_dy, mask_st, V_st, min_views_st, R, S, out_rays,
      out_samples);
  DYN_LAUNCH_CHECK();
  return DYN_OK;
}

int dyn_composite_vanilla(const float* raw, const float* z_vals, const float* mask, int V,
                          int min_views, int R, int S, float* out_rays, float* out_samples,
                          void* stream) {
  if (R == 0) return DYN_OK;  // empty batch: nothing to do (pointers may be null)
  DYN_CHECK_ARG(raw && z_vals && mask && out_rays && out_samples && R >= 0 && S >= 1 && V >= 1);
  if (R == 0) return DYN_OK;
  composite_kernel<false><<<cdiv((long long)R * 32, 128), 128, 0, (cudaStream_t)stream>>>(
      raw, nullptr, z_vals, mask, V, min_views, nullptr, 0, 0, R, S, out_rays, out_samples);
  DYN_LAUNCH_CHECK();
  return DYN_OK;
}

int dyn_resample(const float* z_vals, const float* weights, const float* u, int R, int S, int Ni,
                 int inv_uniform, float* z_out, void* stream) {
  if (R == 0) return DYN_OK;  // empty batch: nothing to do (pointers may be null)
  DYN_CHECK_ARG(z_vals && weights && z_out && R >= 0 && S >= 3 && Ni >= 2);
  DYN_CHECK_ARG(S + Ni <= 1024);
  if (R == 0) return DYN_OK;
  const int wpb = 4;
  size_t smem = (size_t)wpb * (2 * S + S + Ni) * sizeof(float);
  if (smem > 48 * 1024)
    DYN_CUDA(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resample_kernel<<<cdiv(R, wpb), wpb * 32, smem, (cudaStream_t)stream>>>(z_vals, weights, u, R, S, Ni,
                                                                          inv_uniform, z_out);
  DYN_LAUNCH_CHECK();
  return DYN_OK;
}

}  // extern "C"
