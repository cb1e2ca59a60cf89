// Workspace layout of the staged fp32 evaluation of the two aggregation nets (nets_f32.cu).  The training
// path (nets_train.cu) runs the same forward with `train = true` -- nothing is overwritten in place, one
// internal chunk -- and its backward finds every activation again by replaying the same bump allocation.
#pragma once
#include "common.cuh"

namespace dyn {

// workspace bump allocator (also used to SIZE the workspace with base == null)
struct Bump {
  char* base;
  size_t off;
  float* f(size_t n) {
    off = (off + 255) & ~(size_t)255;
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += n * sizeof(float);
    return p;
  }
};

// Shared trunk after the first pooling: base_fc ... ray transformer.
//   H1 = ELU(base_fc.0)   X = ELU(base_fc.2)   H2 = ELU(vis_fc.0(X w1))   XV = ELU(vis_fc.2)  [M,129]
//   X2 = X + XV[:, :128]  vis1 = sigmoid(XV[:,128]) mask      H3 = ELU(vis_fc2.0(X2 vis1))
//   vis2 = sigmoid(vis_fc2.2) mask       G = [mean, var, mean_v(w2)]  nvalid
//   GH, G2 (+ sinusoid) = geometry_fc    Q, K, V, O, O2 = fc(O), G3 = LayerNorm(O2 + G2)
// Inference aliases X2 = X and H3 = H2 (updated in place).
struct TrunkBufs {
  float *H1, *X, *H2, *XV, *X2, *H3, *vis1, *vis2, *G, *nvalid, *GH, *G2, *Q, *K, *V, *O, *O2, *G3;
};

inline void trunk_alloc(Bump& b, long long M, long long P, TrunkBufs* t, bool train) {
  t->H1 = b.f(M * 256); t->X = b.f(M * 128); t->H2 = b.f(M * 128); t->XV = b.f(M * 129);
  t->X2 = train ? b.f(M * 128) : t->X;
  t->H3 = train ? b.f(M * 128) : t->H2;
  t->vis1 = b.f(M); t->vis2 = b.f(M); t->G = b.f(P * 257); t->nvalid = b.f(P);
  // G2, Q, K, V, O double as the fused path's tile-layout buffers: whole 256-row iterations
  const long long Pt = ((P + 255) / 256) * 256;
  t->GH = b.f(P * 256); t->G2 = b.f(Pt * 128); t->Q = b.f(Pt * 128); t->K = b.f(Pt * 128);
  t->V = b.f(Pt * 128); t->O = b.f(Pt * 128); t->O2 = b.f(P * 128); t->G3 = b.f(P * 128);
}

struct DynBufs {
  float *dfeat, *feat, *mv, *w1, *ptspe, *dirpe, *G4h, *G4, *sh, *sig, *ch, *ch2, *rgb;
  TrunkBufs t;
};

inline size_t dyn_alloc(Bump& b, int R, int S, int V, DynBufs* d, bool train = false) {
  long long P = (long long)R * S, M = P * V;
  d->dfeat = b.f(train ? 384 : 64);  // training keeps [dfeat 35 | pad | hidden 256 | PE(t) 21] of ray_dir_fc
  d->feat = b.f(M * kF); d->mv = b.f(P * 2 * kF); d->w1 = b.f(M);
  trunk_alloc(b, M, P, &d->t, train);
  d->ptspe = b.f(P * 33); d->dirpe = b.f((long long)R * 27);
  d->G4h = b.f(P * 256); d->G4 = b.f(P * 128); d->sh = b.f(P * 128); d->sig = b.f(P);
  d->ch = b.f(P * 128); d->ch2 = b.f(P * 64); d->rgb = b.f(P * 3);
  return b.off;
}

struct StBufs {
  float *ptspe, *srcpe, *refpe, *H0, *SF, *reff, *feat70, *mv, *w1, *meff, *sh, *sig, *ch, *ch2, *logit;
  TrunkBufs t;
};

inline size_t st_alloc(Bump& b, int R, int S, int V, StBufs* d, bool train = false) {
  long long P = (long long)R * S, M = P * V;
  d->ptspe = b.f(P * 33); d->srcpe = b.f(M * 66); d->refpe = b.f((long long)R * 66);
  d->H0 = b.f(M * 256); d->SF = b.f(M * kF); d->reff = b.f((long long)R * kF);
  d->feat70 = b.f(M * 2 * kF); d->mv = b.f(P * 4 * kF); d->w1 = b.f(M); d->meff = b.f(M);
  trunk_alloc(b, M, P, &d->t, train);
  d->sh = b.f(P * 128); d->sig = b.f(P);
  d->ch = b.f(M * 128); d->ch2 = b.f(M * 64); d->logit = b.f(M);
  return b.off;
}

}  // namespace dyn
