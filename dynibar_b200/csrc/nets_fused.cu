// Host-side dispatch of the fused per-(point, view) stage of the two aggregation networks: argument
// block + choice between the twin-warp kernel (view_twin.cu, default) and the quad-schedule kernel
// (view_quad.cu, kept for comparison).
//
// Reference semantics: ibrnet/projection.py:103-176, ibrnet/mlp_network.py:236-284
// (dynamic) and :423-497 (static).
#include <stdlib.h>

#include "fused_engine.cuh"
#include "geometry.cuh"
#include "nets.cuh"

namespace dyn {

// Two schedules of the same per-tile work: the twin-warp kernel (view_twin.cu: two independent CTAs per SM,
// default -- it is the faster one, profiles/r02_view_kernels.md) and the quad kernel (view_quad.cu: one CTA
// per SM alternating between two tiles; DYN_VIEW_KERNEL=quad or dyn_debug_set_view_kernel(1)).
static int g_view_kernel = -1;  // 0 twin, 1 quad
void set_view_kernel(int quad) { g_view_kernel = quad ? 1 : 0; }
static bool use_twin_kernel() {
  if (g_view_kernel < 0) {
    const char* e = getenv("DYN_VIEW_KERNEL");
    g_view_kernel = (e != nullptr && e[0] == 'q') ? 1 : 0;
  }
  return g_view_kernel == 0;
}

bool use_twin_chains() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DYN_CHAINS");
    v = (e != nullptr && e[0] == 'f') ? 0 : 1;
  }
  return v == 1;
}

int producer_lanes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DYN_PRODUCERS");
    v = (e != nullptr && e[0] == '2') ? 2 : 1;  // two lanes of ONE warp do not help (divergence): r02_view_kernels.md
  }
  return v;
}

int launch_view_fused(const dyn_net* n, ViewFusedArgs& a, int V, cudaStream_t st) {
  if (V > 16) return fail(DYN_E_INVALID, "fused per-view kernel supports V <= 16 (got %d)", V);
  a.params = n->params;
  a.producers = producer_lanes();
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
  if (use_twin_kernel()) return launch_view_twin(n, a, V, st);
  return launch_view_quad(n, a, V, st);
}

}  // namespace dyn
