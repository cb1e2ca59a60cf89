// Fused per-(point, view) stage of the two aggregation networks on tcgen05.
//
// One persistent CTA per SM.  Per iteration it owns 256 (point, view) rows
// (= two M=128 UMMA tiles; VP = 8 or 16 view slots per point so a point's views
// are adjacent lanes of one warp) and runs, without leaving the SM:
//
//   projection + in-front/in-bounds masks + view-angle difference   (a4, a6)
//   bilinear gather of source RGB + features from L2-resident maps  (a5)
//   [static]  Plucker coords, positional encodings, ray_dir_fc     (a7, a8, a10)
//   pooling weights, weighted mean/var over views (warp shuffles)  (a9/a10)
//   base_fc -> vis_fc -> vis_fc2 (tensor cores, fp32 accum in TMEM)
//   visibility re-weighting and the second mean/var pooling -> G[257] per point
//
// Activations never touch HBM: each layer's epilogue (TMEM -> registers ->
// bias/ELU -> bf16) writes the next layer's A operand straight into the
// canonical shared-memory tile; `x` (128 fp32 per row) is parked in the unused
// TMEM columns.  Weights stream from L2 as pre-packed UMMA images through a
// 4-stage cp.async.bulk ring shared by both row tiles.
//
//   warps 0-7 : one row each: operand construction + all epilogues
//   warp 8    : lane 0 issues every tcgen05.mma (follows the chunk table)
//   warp 9    : lane 0 is the weight producer (bulk copies, runs ahead)
//
// Reference semantics: ibrnet/projection.py:103-176, ibrnet/mlp_network.py:236-284
// (dynamic) and :423-497 (static).
#include <stdlib.h>

#include "fused_engine.cuh"
#include "geometry.cuh"
#include "nets.cuh"

namespace dyn {

using namespace tc;
using namespace fe;

namespace {

// offsets (floats) inside the smem constant block
constexpr int C_B1 = 0;      // [256] ray_dir_fc.0 bias (static) / unused
constexpr int C_B2 = 256;    // [48]  ray_dir_fc.2 bias (static)
constexpr int C_B3 = 304;    // [256] base_fc.0 bias
constexpr int C_B4 = 560;    // [128] base_fc.2 bias
constexpr int C_B5 = 688;    // [128] vis_fc.0 bias
constexpr int C_B6 = 816;    // [128] vis_fc.2 bias rows 0..127
constexpr int C_W6V = 944;   // [128] vis_fc.2 weight row 128 (the visibility logit)
constexpr int C_B7 = 1072;   // [128] vis_fc2.0 bias
constexpr int C_W8 = 1200;   // [128] vis_fc2.2 weight
constexpr int C_MISC = 1328; // [0] vis_fc.2 bias[128], [1] vis_fc2.2 bias, [2] |s|
constexpr int C_DFEAT = 1344;  // [40] dynamic time feature

template <int VP, bool ST>
__global__ void __launch_bounds__(320, 1) view_fused_kernel(const __grid_constant__ ViewFusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + 2 * kATileBytes;
  float* cst = reinterpret_cast<float*>(ring + kRing * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(cst + kConstFloats);
  // bars: see fused_engine.cuh (ping-pong schedule: per-tile a_ready / acc_full)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar0 = smem_u32(bars);
  __shared__ __align__(16) FusedChunk s_tab[kMaxChunks];
  stage_chunks(s_tab, a.chunks, a.nchunks);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (tid == 0) init_barriers(bar0, /*pp=*/true);
  // constants -> smem
  {
    const float* prm = a.params;
    for (int i = tid; i < 256; i += blockDim.x) {
      if (ST) cst[C_B1 + i] = prm[a.o_b1 + i];
      cst[C_B3 + i] = prm[a.o_b3 + i];
    }
    for (int i = tid; i < 128; i += blockDim.x) {
      cst[C_B4 + i] = prm[a.o_b4 + i];
      cst[C_B5 + i] = prm[a.o_b5 + i];
      cst[C_B6 + i] = prm[a.o_b6 + i];
      cst[C_W6V + i] = prm[a.o_w6 + 128 * 128 + i];
      cst[C_B7 + i] = prm[a.o_b7 + i];
      cst[C_W8 + i] = prm[a.o_w8 + i];
    }
    if (tid < 48) cst[C_B2 + tid] = (ST && tid < kF) ? prm[a.o_b2 + tid] : 0.f;
    if (tid < 40) cst[C_DFEAT + tid] = (!ST && tid < kF) ? a.dfeat[tid] : 0.f;
    if (tid == 0) {
      cst[C_MISC + 0] = prm[a.o_b6 + 128];
      cst[C_MISC + 1] = prm[a.o_b8];
      cst[C_MISC + 2] = (ST && a.o_s >= 0) ? fabsf(prm[a.o_s]) : 0.f;
    }
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_rows = a.P * VP;
  const int n_iter = (int)((n_rows + 255) / 256);
  const FusedChunk* chunks = s_tab;
  const int nchunks = a.nchunks;

  if (warp == 9) {
    if ((tid & 31) == 0) producer_loop<true>(chunks, nchunks, a.wimg, n_iter, ring, bar0);
  } else if (warp == 8) {
    issuer_loop<true>(chunks, nchunks, n_iter, smem, ring, bar0, tmem_base);
  } else {
    // ------------------------- row threads -------------------------
    const int tile = tid >> 7, r = tid & 127;
    uint8_t* arow = smem + tile * kATileBytes + (r >> 3) * 128 + (r & 7) * 16;
    const uint32_t tacc = tmem_addr(tmem_base, (uint32_t)((warp & 3) * 32), (uint32_t)(tile * 256));
    const int v = tid % VP;
    const int gl = tid & (VP - 1);  // == v; lane bits used by the reduce-scatter
    uint32_t acc_cnt = 0;
    const float wh = a.w_img, hh = a.h_img;

    for (int it = blockIdx.x; it < n_iter; it += gridDim.x) {
      const long long pl = ((long long)it * 256 + tid) / VP;
      const bool pt_ok = pl < a.P;
      const bool valid = pt_ok && v < a.V;
      const long long m = pl * a.V + v;  // compact (point, view) index for outputs
      const long long ray = pt_ok ? pl / a.S : 0;

      // ---- geometry -------------------------------------------------------
      float p3[3] = {0.f, 0.f, 0.f}, q3[3];
      if (pt_ok) { p3[0] = a.pts[pl * 3]; p3[1] = a.pts[pl * 3 + 1]; p3[2] = a.pts[pl * 3 + 2]; }
      q3[0] = p3[0]; q3[1] = p3[1]; q3[2] = p3[2];
      if (!ST && valid) {
        const float* q = a.pts_seq + ((long long)v * a.seq_stride + pl) * 3;
        q3[0] = q[0]; q3[1] = q[1]; q3[2] = q[2];
      }
      const int vc = valid ? v : 0;
      float pu, pv;
      bool front;
      project_point(a.cams.P[vc], q3[0], q3[1], q3[2], pu, pv, front);
      const bool inb = (pu <= wh - 1.f) && (pu >= 0.f) && (pv <= hh - 1.f) && (pv >= 0.f);
      const float mask_proj = (valid && inb && front) ? 1.f : 0.f;
      float rd[4];
      {
        float a0 = a.cams.tgt[0] - p3[0], a1 = a.cams.tgt[1] - p3[1], a2 = a.cams.tgt[2] - p3[2];
        normalize3(a0, a1, a2);
        float b0 = a.cams.center[vc][0] - q3[0], b1 = a.cams.center[vc][1] - q3[1],
              b2 = a.cams.center[vc][2] - q3[2];
        normalize3(b0, b1, b2);
        rd[0] = a0 - b0; rd[1] = a1 - b1; rd[2] = a2 - b2;
        rd[3] = a0 * b0 + a1 * b1 + a2 * b2;
        normalize3(rd[0], rd[1], rd[2]);
      }

      if (ST) {
        // ---- static layer-1 operand: [PE(pts) 33 | PE(src plucker) 66 | ray_diff 4] ----
        float xin[112];
        pe_pow2<3, 5>(p3, xin);
        float pl6[6];
        {
          const float ox = a.cams.center[vc][0], oy = a.cams.center[vc][1], oz = a.cams.center[vc][2];
          float dx = p3[0] - ox, dy = p3[1] - oy, dz = p3[2] - oz;
          normalize3(dx, dy, dz);
          pl6[0] = dx; pl6[1] = dy; pl6[2] = dz;
          pl6[3] = oy * dz - oz * dy;
          pl6[4] = oz * dx - ox * dz;
          pl6[5] = ox * dy - oy * dx;
        }
        pe_pow2<6, 5>(pl6, xin + 33);
#pragma unroll
        for (int i = 0; i < 4; ++i) xin[99 + i] = rd[i];
#pragma unroll
        for (int i = 103; i < 112; ++i) xin[i] = 0.f;
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 112; ++i) xin[i] = 0.f;
        }
#pragma unroll
        for (int g = 0; g < 14; ++g) store8(arow, 8 * g, xin + 8 * g);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(bar_aready(bar0, tile));
      }

      // ---- bilinear gather (fp32 maps, L2 resident); overlaps the first MMA ----
      float feat[kF];
#pragma unroll
      for (int i = 0; i < kF; ++i) feat[i] = 0.f;
      if (valid) {
        const float gx = 2.f * pu / (wh - 1.f) - 1.f, gy = 2.f * pv / (hh - 1.f) - 1.f;
        {
          const float fx = (gx + 1.f) * 0.5f * (float)(a.w - 1), fy = (gy + 1.f) * 0.5f * (float)(a.h - 1);
          const float x0f = floorf(fx), y0f = floorf(fy);
          const int x0 = (int)x0f, y0 = (int)y0f;
          const float ax = fx - x0f, ay = fy - y0f, bx = (x0f + 1.f) - fx, by = (y0f + 1.f) - fy;
          const float* base = a.feat_cl + (long long)v * a.h * a.w * kC;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int xi = x0 + dx, yi = y0 + dy;
              const float wgt = (dx ? ax : bx) * (dy ? ay : by);
              if (xi >= 0 && xi < a.w && yi >= 0 && yi < a.h) {
                const float4* t = reinterpret_cast<const float4*>(base + ((long long)yi * a.w + xi) * kC);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 q = __ldg(t + j);
                  feat[3 + 4 * j] += q.x * wgt; feat[4 + 4 * j] += q.y * wgt;
                  feat[5 + 4 * j] += q.z * wgt; feat[6 + 4 * j] += q.w * wgt;
                }
              }
            }
        }
        {
          const float fx = (gx + 1.f) * 0.5f * (float)(a.W - 1), fy = (gy + 1.f) * 0.5f * (float)(a.H - 1);
          const float x0f = floorf(fx), y0f = floorf(fy);
          const int x0 = (int)x0f, y0 = (int)y0f;
          const float ax = fx - x0f, ay = fy - y0f, bx = (x0f + 1.f) - fx, by = (y0f + 1.f) - fy;
          const float* base = a.rgbs + (long long)v * a.H * a.W * 3;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const int xi = x0 + dx, yi = y0 + dy;
              const float wgt = (dx ? ax : bx) * (dy ? ay : by);
              if (xi >= 0 && xi < a.W && yi >= 0 && yi < a.H) {
                const float* t = base + ((long long)yi * a.W + xi) * 3;
                feat[0] += __ldg(t) * wgt; feat[1] += __ldg(t + 1) * wgt; feat[2] += __ldg(t + 2) * wgt;
              }
            }
        }
      }
      float mask = mask_proj;
      if (ST && a.mask_rgb) mask *= ((feat[0] + feat[1] + feat[2]) > 1e-3f) ? 1.f : 0.f;
      if (valid) {
        a.mask_proj[m] = mask_proj;
        if (ST) {
          a.mask_eff[m] = mask;
          reinterpret_cast<float4*>(a.ray_diff)[m] = make_float4(rd[0], rd[1], rd[2], rd[3]);
          a.rgb_in[m * 3] = feat[0]; a.rgb_in[m * 3 + 1] = feat[1]; a.rgb_in[m * 3 + 2] = feat[2];
        }
      }

      float acc[32];
      float f2[40];  // second half of the static per-view feature (src_feat * ref_feat)
      if (ST) {
        // ---- F1 epilogue: ELU(ray_dir_fc.0) -> A[256] ----
        mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
        tc_fence_after_sync();
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
          tmem_ld32(tacc + cb * 32, acc);
          tmem_wait_ld();
          epi32_to_A<true>(arow, cb * 32, acc, cst + C_B1, 1.f);
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(bar_aready(bar0, tile));
        // ---- F2 epilogue: src_feat = ray_dir_fc.2 (35 of 48 cols), times ref_feat ----
        mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
        tc_fence_after_sync();
        float t16[16];
        tmem_ld32(tacc, acc);
        tmem_ld16(tacc + 32, t16);
        tmem_wait_ld();
        const float* rf = a.ref_feat + ray * kF;
#pragma unroll
        for (int i = 0; i < 40; ++i) {
          const float sv = i < 32 ? acc[i < 32 ? i : 0] : t16[i >= 32 ? i - 32 : 0];
          f2[i] = i < kF ? (sv + cst[C_B2 + i]) * __ldg(rf + (i < kF ? i : 0)) : 0.f;
        }
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 40; ++i) f2[i] = 0.f;
        }
      } else {
        // dynamic: feat += time feature (mlp_network.py:244-247)
#pragma unroll
        for (int i = 0; i < kF; ++i) feat[i] = valid ? feat[i] + cst[C_DFEAT + i] : 0.f;
      }

      // ---- pooling weights (mlp_network.py:249-257 / :461-469) ----
      float w1;
      if (ST && a.anti_alias) {
        float e = ex2f(cst[C_MISC + 2] * (rd[3] - 1.f) * 1.4426950408889634f);
        float emin = group_min<VP>(valid ? e : INFINITY);
        w1 = valid ? (e - emin) * mask : 0.f;
      } else {
        w1 = mask;
      }
      w1 = w1 / (group_sum<VP>(w1) + 1e-8f);

      // ---- weighted mean / var over views, written channel-group-wise as
      //      [mean8 | var8 | feat8] (weight columns are permuted to match) ----
      {
        constexpr int NG = ST ? 9 : 5;  // 72 / 40 channels
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          float o[24];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = 8 * g + j;
            float fv;
            if (ST) fv = c < kF ? feat[c < kF ? c : 0] : (c < 2 * kF ? f2[(c - kF) < 40 && c >= kF ? c - kF : 0] : 0.f);
            else fv = c < kF ? feat[c < kF ? c : 0] : 0.f;
            const float s1 = group_sum<VP>(w1 * fv);
            const float d = fv - s1;
            const float s2 = group_sum<VP>(w1 * d * d);
            o[j] = s1; o[8 + j] = s2; o[16 + j] = fv;
          }
          store8(arow, 24 * g, o);
          store8(arow, 24 * g + 8, o + 8);
          store8(arow, 24 * g + 16, o + 16);
        }
        // zero the K padding up to the layer's padded width
        float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store8(arow, 24 * NG, z);
        if (!ST) { /* 120 -> 128 */ } else { /* 216 -> 224 */ }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_aready(bar0, tile));

      // ---- F3: ELU(base_fc.0) -> A[256] ----
      mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
      tc_fence_after_sync();
#pragma unroll 1
      for (int cb = 0; cb < 8; ++cb) {
        tmem_ld32(tacc + cb * 32, acc);
        tmem_wait_ld();
        epi32_to_A<true>(arow, cb * 32, acc, cst + C_B3, 1.f);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_aready(bar0, tile));

      // ---- F4: x = ELU(base_fc.2); park x in TMEM cols [128,256); A = x * w1 ----
      mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
      tc_fence_after_sync();
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        tmem_ld32(tacc + cb * 32, acc);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = elu_fast(acc[i] + cst[C_B4 + cb * 32 + i]);
        tmem_st32(tacc + 128 + cb * 32, acc);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] *= w1;
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(arow, cb * 32 + 8 * g, acc + 8 * g);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_aready(bar0, tile));

      // ---- F5: h = ELU(vis_fc.0); A = h; visibility logit = ELU(w_128 . h + b) ----
      mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
      tc_fence_after_sync();
      float vlogit = cst[C_MISC + 0];
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        tmem_ld32(tacc + cb * 32, acc);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[i] = elu_fast(acc[i] + cst[C_B5 + cb * 32 + i]);
          vlogit = fmaf(acc[i], cst[C_W6V + cb * 32 + i], vlogit);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(arow, cb * 32 + 8 * g, acc + 8 * g);
      }
      const float vis1 = sigmoid_fast(elu_fast(vlogit)) * mask;  // mlp_network.py:272-274
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_aready(bar0, tile));

      // ---- F6: x += ELU(vis_fc.2[:128]); A = x * vis1 ----
      mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
      tc_fence_after_sync();
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        float xs[32];
        tmem_ld32(tacc + cb * 32, acc);
        tmem_ld32(tacc + 128 + cb * 32, xs);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) xs[i] += elu_fast(acc[i] + cst[C_B6 + cb * 32 + i]);
        tmem_st32(tacc + 128 + cb * 32, xs);
        if (ST && valid) {
          // spilled as bf16: its only consumer is the blending head's A operand
          uint4* xo = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.X) + m * 128 + cb * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            xo[i] = make_uint4(pack_bf16x2(xs[8 * i], xs[8 * i + 1]), pack_bf16x2(xs[8 * i + 2], xs[8 * i + 3]),
                               pack_bf16x2(xs[8 * i + 4], xs[8 * i + 5]), pack_bf16x2(xs[8 * i + 6], xs[8 * i + 7]));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) xs[i] *= vis1;
#pragma unroll
        for (int g = 0; g < 4; ++g) store8(arow, cb * 32 + 8 * g, xs + 8 * g);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_aready(bar0, tile));

      // ---- F7: vis2 = sigmoid(vis_fc2.2 . ELU(vis_fc2.0)) * mask ----
      mbar_wait(bar_acc(bar0, tile), acc_cnt & 1); ++acc_cnt;
      tc_fence_after_sync();
      float v2 = cst[C_MISC + 1];
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        tmem_ld32(tacc + cb * 32, acc);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v2 = fmaf(elu_fast(acc[i] + cst[C_B7 + cb * 32 + i]), cst[C_W8 + cb * 32 + i], v2);
      }
      const float vis2 = sigmoid_fast(v2) * mask;
      if (ST && valid) a.vis2[m] = vis2;
      const float vsum = group_sum<VP>(vis2);
      const float w2 = vis2 / (vsum + 1e-8f);
      const float W = group_sum<VP>(w2);
      const float nval = group_sum<VP>(mask);

      // ---- second pooling: reduce-scatter of sum(w x) and sum(w x^2) over the
      //      views; lane (b0,b1,b2[,b3]) ends up owning 16 (8) channels ----
      {
        const bool b0 = gl & 1, b1 = gl & 2, b2 = gl & 4, b3 = gl & 8;
        constexpr int NO = VP == 16 ? 8 : 16;
        const int cbase = (b0 ? 64 : 0) + (b1 ? 32 : 0) + (b2 ? 16 : 0) + ((VP == 16 && b3) ? 8 : 0);
        float mean[NO], sq[NO];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float s1[64];
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {  // channels [32h, 32h+32) pair with [64+32h, ...)
            float lo[32], hi[32];
            tmem_ld32(tacc + 128 + hlf * 32, lo);
            tmem_ld32(tacc + 128 + 64 + hlf * 32, hi);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float l = q ? w2 * lo[i] * lo[i] : w2 * lo[i];
              const float h = q ? w2 * hi[i] * hi[i] : w2 * hi[i];
              const float send = b0 ? l : h, keep = b0 ? h : l;
              s1[hlf * 32 + i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
          }
          float s2[32], s3[16];
          rs_step<64>(s1, s2, b1, 2);
          rs_step<32>(s2, s3, b2, 4);
          float* dst = q ? sq : mean;
          if (VP == 16) {
            float s4[8];
            rs_step<16>(s3, s4, b3, 8);
#pragma unroll
            for (int i = 0; i < NO; ++i) dst[i] = s4[i];
          } else {
#pragma unroll
            for (int i = 0; i < NO; ++i) dst[i] = s3[i < 16 ? i : 0];
          }
        }
        if (pt_ok) {
          float* g = a.G + pl * kGStride;
#pragma unroll
          for (int i = 0; i < NO; ++i) {
            const float mu = mean[i];
            g[cbase + i] = mu;
            g[128 + cbase + i] = sq[i] - mu * mu * (2.f - W);  // sum w (x-mu)^2 with sum w = W
          }
          if (gl == 0) {
            g[256] = W / (float)a.V;  // weight.mean(dim=2), mlp_network.py:281
            a.nvalid[pl] = nval;
          }
        }
      }
      // (next iteration overwrites A and the accumulators only after the reads above:
      //  the MMA issuer is gated by this thread's next arrive on a_ready)
      tc_fence_before_sync();
      if (!ST) {
        // dynamic has no layer-1: the pooled operand above IS the first operand of the
        // next iteration, nothing else to do here
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// host side: weight images + chunk table
// ---------------------------------------------------------------------------
size_t fused_view_bytes(int kind) {
  // upper bound: images + table (static is the larger one)
  (void)kind;
  return (size_t)(512 * 1024);
}

int fused_view_build(dyn_net* n, const float* host_params, void* dst_dev, size_t dst_bytes,
                     cudaStream_t st) {
  std::vector<uint8_t> img;
  std::vector<FusedChunk> tab;
  const float* P = host_params;
  auto add = [&](const LinearP& l, int N, int Npad, int Kpad, std::vector<int> map) {
    HostLayer L;
    L.W = P + l.w; L.N = N; L.Kw = l.in; L.Npad = Npad; L.Kpad = Kpad; L.colmap = std::move(map);
    append_layer(L, img, tab);
  };
  if (n->kind == DYN_NET_STATIC) {
    const StaticLayout& L = n->sl;
    add(L.ray_dir0, 256, 256, 112, identity_map(103, 112));
    add(L.ray_dir2, kF, 48, 256, identity_map(256, 256));
    add(L.base0, 256, 256, 224, pooled_map(2 * kF, 9, 224));
    add(L.base2, 128, 128, 256, identity_map(256, 256));
    add(L.vis0, 128, 128, 128, identity_map(128, 128));
    add(L.vis2, 128, 128, 128, identity_map(128, 128));
    add(L.vis2_0, 128, 128, 128, identity_map(128, 128));
  } else {
    const DynamicLayout& L = n->dl;
    add(L.base0, 256, 256, 128, pooled_map(kF, 5, 128));
    add(L.base2, 128, 128, 256, identity_map(256, 256));
    add(L.vis0, 128, 128, 128, identity_map(128, 128));
    add(L.vis2, 128, 128, 128, identity_map(128, 128));
    add(L.vis2_0, 128, 128, 128, identity_map(128, 128));
  }
  const size_t img_bytes = (img.size() + 255) & ~(size_t)255;
  const size_t need = img_bytes + tab.size() * sizeof(FusedChunk);
  if (need > dst_bytes) return fail(DYN_E_INVALID, "fused images need %zu bytes, have %zu", need, dst_bytes);
  DYN_CUDA(cudaMemcpyAsync(dst_dev, img.data(), img.size(), cudaMemcpyHostToDevice, st));
  DYN_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(dst_dev) + img_bytes, tab.data(),
                           tab.size() * sizeof(FusedChunk), cudaMemcpyHostToDevice, st));
  DYN_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
  n->fused_img = dst_dev;
  n->fused_tab = reinterpret_cast<char*>(dst_dev) + img_bytes;
  n->fused_nchunks = (int)tab.size();
  if (tab.size() > (size_t)kMaxChunks) return fail(DYN_E_INVALID, "chunk table too long (%zu)", tab.size());
  return DYN_OK;
}

int launch_view_fused(const dyn_net* n, ViewFusedArgs& a, int V, cudaStream_t st) {
  if (n->fused_img == nullptr) return fail(DYN_E_INVALID, "net has no fused tensor-core images");
  if (V > 16) return fail(DYN_E_INVALID, "fused per-view kernel supports V <= 16 (got %d)", V);
  a.wimg = n->fused_img;
  a.chunks = reinterpret_cast<const FusedChunk*>(n->fused_tab);
  a.nchunks = n->fused_nchunks;
  a.params = n->params;
  const bool st_net = n->kind == DYN_NET_STATIC;
  if (st_net) {
    const StaticLayout& L = n->sl;
    a.o_b1 = L.ray_dir0.b; a.o_b2 = L.ray_dir2.b; a.o_b3 = L.base0.b; a.o_b4 = L.base2.b;
    a.o_b5 = L.vis0.b; a.o_b6 = L.vis2.b; a.o_w6 = L.vis2.w; a.o_b7 = L.vis2_0.b;
    a.o_w8 = L.vis2_2.w; a.o_b8 = L.vis2_2.b; a.o_s = L.s;
    a.anti_alias = n->anti_alias; a.mask_rgb = n->mask_rgb;
  } else {
    const DynamicLayout& L = n->dl;
    a.o_b1 = 0; a.o_b2 = 0; a.o_b3 = L.base0.b; a.o_b4 = L.base2.b;
    a.o_b5 = L.vis0.b; a.o_b6 = L.vis2.b; a.o_w6 = L.vis2.w; a.o_b7 = L.vis2_0.b;
    a.o_w8 = L.vis2_2.w; a.o_b8 = L.vis2_2.b; a.o_s = -1;
    a.anti_alias = 0; a.mask_rgb = 0;
  }
  {
    // default: twin-warp kernel (16 row warps per SM); DYN_VIEW_TWIN=0 selects the
    // one-thread-per-row kernel below (kept for A/B measurements)
    static int use_twin = -1;
    if (use_twin < 0) {
      const char* e = getenv("DYN_VIEW_TWIN");
      use_twin = (e && e[0] == '0') ? 0 : 1;
    }
    if (use_twin && n->twin.img != nullptr) return launch_view_twin(n, a, V, st);
  }
  int dev = 0, sms = 148;
  DYN_CUDA(cudaGetDevice(&dev));
  DYN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int VP = V <= 8 ? 8 : 16;
  const long long n_iter = (a.P * VP + 255) / 256;
  const int grid = (int)(n_iter < sms ? n_iter : sms);
  if (grid == 0) return DYN_OK;
  ProfScope prof(st_net ? PROF_VIEW_ST : PROF_VIEW_DY, st);
#define LAUNCH_VF(VPV, STV)                                                                   \
  do {                                                                                        \
    DYN_CUDA(cudaFuncSetAttribute(view_fused_kernel<VPV, STV>,                                \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFused));  \
    view_fused_kernel<VPV, STV><<<grid, 320, kSmemFused, st>>>(a);                            \
  } while (0)
  if (st_net) { if (VP == 8) LAUNCH_VF(8, true); else LAUNCH_VF(16, true); }
  else { if (VP == 8) LAUNCH_VF(8, false); else LAUNCH_VF(16, false); }
#undef LAUNCH_VF
  DYN_LAUNCH_CHECK();
  return DYN_OK;
}

}  // namespace dyn
