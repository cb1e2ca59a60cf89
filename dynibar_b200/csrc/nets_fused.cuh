// Fused per-view tensor-core stage (nets_fused.cu): argument block + host API.
#pragma once
#include "geometry.cuh"

namespace dyn {

constexpr int kGStride = 272;  // row stride of the pooled per-point feature G (257 used)

struct FusedChunk {
  uint32_t off, bytes;   // weight chunk inside the image
  uint16_t npad;         // UMMA N of the layer
  uint8_t ksteps;        // K = 16 * ksteps in this chunk
  uint8_t flags;         // 1: first chunk of a layer (wait for the operand, overwrite D)
                         // 2: last chunk of a layer (signal the epilogue)
  uint16_t a_kgroup;     // first 8-column group of the A tile this chunk consumes
  uint16_t pad;
};

struct ViewFusedArgs {
  // inputs
  const float* pts;       // [P,3] reference-time sample points
  const float* pts_seq;   // motion-displaced points, view v at pts_seq + v*seq_stride*3 (dynamic) or null
  long long seq_stride;
  const float* rgbs;      // [V,H,W,3]
  const float* feat_cl;   // [V,h,w,32] channels-last feature maps
  const float* ref_feat;  // [R,35] static: ref_feature_fc(PE(ref plucker)) per ray
  const float* dfeat;     // [35] dynamic: time feature
  const float* params;    // fp32 parameter blob (biases, small heads)
  const void* wimg;       // packed weight chunks
  const FusedChunk* chunks;
  int nchunks;
  ViewCams cams;
  float h_img, w_img;
  int H, W, h, w, V, S;
  long long P;
  int anti_alias, mask_rgb;
  int o_b1, o_b2, o_b3, o_b4, o_b5, o_b6, o_w6, o_b7, o_w8, o_b8, o_s;
  // outputs
  float* G;          // [P, kGStride]: mean128 | var128 | mean weight
  float* nvalid;     // [P]
  float* mask_proj;  // [P*V] projector mask (in front & in bounds)
  float* mask_eff;   // static [P*V]: mask after mask_rgb gating
  float* X;          // static [P*V,128]: per-view feature after the visibility residual
  float* vis2;       // static [P*V]
  float* ray_diff;   // static [P*V,4]
  float* rgb_in;     // static [P*V,3] gathered source colours
};

size_t fused_view_bytes(int kind);
int fused_view_build(dyn_net* n, const float* host_params, void* dst_dev, size_t dst_bytes,
                     cudaStream_t st);
int launch_view_fused(const dyn_net* n, ViewFusedArgs& a, int V, cudaStream_t st);

}  // namespace dyn
