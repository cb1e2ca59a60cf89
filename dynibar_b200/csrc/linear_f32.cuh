// Generic fp32 SIMT linear layer  Y = act(concat(segments) * W^T + b)
// used by the DYN_PREC_FP32 (parity) mode and for the small per-ray layers.
#pragma once
#include "common.cuh"

namespace dyn {

enum Act { ACT_NONE = 0, ACT_ELU = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };

// One input segment: logical columns [start, start+width) of the layer input
// come from p[(row / div) * ld + col].  `div` > 1 broadcasts a per-point (or
// per-ray) tensor over the rows of a finer-grained matrix without
// materialising the reference's expand()+cat() (mlp_network.py:267-269 etc).
struct Seg {
  const float* p;
  int width, ld, div;
};

struct LinArgs {
  Seg seg[4];
  int nseg;
  const float* row_scale;  // optional per-row multiplier on the input (x * weight)
  const float* W;          // [N,K] row-major (nn.Linear.weight)
  const float* b;          // [N] or null
  float* Y;
  int ldy;
  long long M;
  int N, K;
  int act;
};

int launch_linear(const LinArgs& a, cudaStream_t st);

// convenience: single dense input
static inline LinArgs lin1(const float* X, int ldx, const float* W, const float* b, float* Y,
                           int ldy, long long M, int N, int K, int act) {
  LinArgs a;
  memset(&a, 0, sizeof(a));
  a.seg[0] = Seg{X, K, ldx, 1};
  a.nseg = 1;
  a.W = W; a.b = b; a.Y = Y; a.ldy = ldy; a.M = M; a.N = N; a.K = K; a.act = act;
  return a;
}

}  // namespace dyn
