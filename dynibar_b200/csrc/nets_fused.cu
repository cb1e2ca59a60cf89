// Host-side dispatch of the fused per-(point, view) stage of the two aggregation networks.
// The kernel itself is the twin-warp one in view_twin.cu (the earlier one-thread-per-row
// kernel was removed once the twin-warp version superseded it; see profiles/r01_kernels.md).
//
// Reference semantics: ibrnet/projection.py:103-176, ibrnet/mlp_network.py:236-284
// (dynamic) and :423-497 (static).
#include <stdlib.h>

#include "fused_engine.cuh"
#include "geometry.cuh"
#include "nets.cuh"

namespace dyn {

// no separate image for the removed kernel (kept so dyn_net_packed_bytes' layout code is unchanged)
size_t fused_view_bytes(int kind) { (void)kind; return 0; }
int fused_view_build(dyn_net* n, const float* host_params, void* dst_dev, size_t dst_bytes,
                     cudaStream_t st) {
  (void)host_params; (void)dst_dev; (void)dst_bytes; (void)st;
  n->fused_img = nullptr; n->fused_tab = nullptr; n->fused_nchunks = 0;
  return DYN_OK;
}

int launch_view_fused(const dyn_net* n, ViewFusedArgs& a, int V, cudaStream_t st) {
  if (V > 16) return fail(DYN_E_INVALID, "fused per-view kernel supports V <= 16 (got %d)", V);
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
  return launch_view_twin(n, a, V, st);
}

}  // namespace dyn
