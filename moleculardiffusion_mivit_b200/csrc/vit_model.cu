// Orchestration of the motion-informed ViT regressor (reference helpers/models.py:278-361
// GeneralTransformer) forward / backward / training step on the kernels of this library.
// Parameters live in ONE flat fp32 buffer in the canonical order produced by ParamLayout (the
// Python module mirrors the order by state_dict key); gradients and AdamW moments use the same
// offsets, so the optimiser and the data-parallel all-reduce are single passes over 2 MB.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "vit.h"
#include "../../include/mivit.h"

#define CK(expr)              \
  do {                        \
    int _rc = (expr);         \
    if (_rc) return _rc;      \
  } while (0)

namespace {

constexpr int kGuard = 128;
constexpr int kMaxLayers = 32;

struct ParamLayout {
  long off = 0;
  int count = 0;
  long sizes[24 + 16 * kMaxLayers + 16];
  long add(long n) {
    const long o = off;
    sizes[count++] = n;
    off += n;
    return o;
  }
  // embedding
  long conv0_w = -1, bn0_g = -1, bn0_b = -1;
  struct RB { long c1_w, bn1_g, bn1_b, c2_w, bn2_g, bn2_b, sk_w, bns_g, bns_b; } rb[2];
  long fc_w = -1, fc_b = -1, proj_w = -1, proj_b = -1;
  long norm_g, norm_b, reg = -1, pos = -1;
  struct Lyr { long q_w, q_b, k_w, k_b, v_w, v_b, o_w, o_b, n1_g, n1_b, f1_w, f1_b, f2_w, f2_b, n2_g, n2_b; } lyr[kMaxLayers];
  long tn_g, tn_b, fp0_w = -1, fp0_b = -1, fp2_w = -1, fp2_b = -1, h0_w, h0_b, h3_w, h3_b;
  int head_in = 0;
  // ModularTransformer (helpers/models.py:366-593)
  long fe0_w = -1, fe0_b = -1, feln_g = -1, feln_b = -1, fe3_w = -1, fe3_b = -1, fu_w = -1, fu_b = -1;
  int E_img = 0;         // output width of the image embedding (embed_dim - features_dim for 'concat_features')
  bool has_img = true;   // an image embedding exists (everything but 'features_only')
  bool has_femb = false; // a feature embedding exists ('features_only' / 'both' unless 'concat_features')
};

// width of the image embedding output
inline int image_embed_dim(const mivit_vit_config* c) {
  return (c->modular && c->mod_mode == 2 && c->mod_fusion == 2) ? c->E - c->feat_dim : c->E;
}

int build_layout(const mivit_vit_config* c, ParamLayout& L) {
  MIVIT_CHECK_ARG(c != nullptr, "config is NULL");
  MIVIT_CHECK_ARG(c->embedding >= 0 && c->embedding <= 2, "embedding must be 0 (linear), 1 (cnn) or 2 (deepresnet)");
  MIVIT_CHECK_ARG(c->L >= 1 && c->L <= kMaxLayers, "num_layers out of range");
  MIVIT_CHECK_ARG(c->E >= 8 && c->E <= 256 && c->E % c->H == 0, "embed_dim must be in [8,256] and divisible by num_heads");
  MIVIT_CHECK_ARG(c->P >= 1 && c->P <= 100 && c->F >= 1, "bad patch size / frame count");
  MIVIT_CHECK_ARG(c->F + (c->use_reg ? 1 : 0) <= 128, "more than MAX_TOKENS = 128 tokens");
  const int E = c->E, HD = c->HD, PP = c->P * c->P;
  if (c->modular) {
    MIVIT_CHECK_ARG(c->mod_mode >= 0 && c->mod_mode <= 2, "mode must be one of: 'images_only', 'features_only', 'both'");
    MIVIT_CHECK_ARG(c->mod_fusion >= 0 && c->mod_fusion <= 2, "fusion_method must be one of: 'add', 'concat_proj', 'concat_features'");
    MIVIT_CHECK_ARG(c->mod_fembed >= 0 && c->mod_fembed <= 1, "Unknown feature_embedding_type");
    MIVIT_CHECK_ARG(!c->use_feat, "global features (use_feat) belong to GeneralTransformer, not to ModularTransformer");
    MIVIT_CHECK_ARG(c->mod_mode == 0 || c->feat_dim >= 1, "features_dim must be provided when using features");
    MIVIT_CHECK_ARG(image_embed_dim(c) > 0, "embed_dim (%d) must be greater than features_dim (%d) when using 'concat_features' fusion",
                    c->E, c->feat_dim);
    MIVIT_CHECK_ARG(!(c->mod_mode != 0 && c->mod_fembed == 1) || 2 * E <= 256, "'mlp' feature embedding needs embed_dim <= 128");
  }
  MIVIT_CHECK_ARG(!c->per_frame || (!c->use_reg && !c->use_feat),
                  "per-frame outputs (single_prediction=False) need use_regression_token=False and no global features");
  L.E_img = image_embed_dim(c);
  L.has_img = !c->modular || c->mod_mode != 1;
  L.has_femb = c->modular && c->mod_mode != 0 && !(c->mod_mode == 2 && c->mod_fusion == 2);
  const int Ei = L.E_img;
  if (!L.has_img) {
    // 'features_only': no image embedding
  } else if (c->embedding == 2) {
    L.conv0_w = L.add(32 * 9); L.bn0_g = L.add(32); L.bn0_b = L.add(32);
    const int ci[2] = {32, 64}, co[2] = {64, 128};
    for (int b = 0; b < 2; ++b) {
      L.rb[b].c1_w = L.add((long)co[b] * ci[b] * 9); L.rb[b].bn1_g = L.add(co[b]); L.rb[b].bn1_b = L.add(co[b]);
      L.rb[b].c2_w = L.add((long)co[b] * co[b] * 9); L.rb[b].bn2_g = L.add(co[b]); L.rb[b].bn2_b = L.add(co[b]);
      L.rb[b].sk_w = L.add((long)co[b] * ci[b]); L.rb[b].bns_g = L.add(co[b]); L.rb[b].bns_b = L.add(co[b]);
    }
    L.fc_w = L.add((long)Ei * 128); L.fc_b = L.add(Ei);
  } else {
    L.proj_w = L.add((long)Ei * PP); L.proj_b = L.add(Ei);
  }
  if (L.has_femb) {
    const int fd = c->feat_dim;
    if (c->mod_fembed == 0) {
      L.fe0_w = L.add((long)E * fd); L.fe0_b = L.add(E);
    } else {
      L.fe0_w = L.add(2L * E * fd); L.fe0_b = L.add(2 * E); L.feln_g = L.add(2 * E); L.feln_b = L.add(2 * E);
      L.fe3_w = L.add((long)E * 2 * E); L.fe3_b = L.add(E);
    }
  }
  if (c->modular && c->mod_mode == 2 && c->mod_fusion == 1) { L.fu_w = L.add((long)E * 2 * E); L.fu_b = L.add(E); }
  L.norm_g = L.add(E); L.norm_b = L.add(E);
  if (c->use_reg) L.reg = L.add(E);
  if (c->use_pos) L.pos = L.add(128L * E);
  for (int l = 0; l < c->L; ++l) {
    auto& y = L.lyr[l];
    y.q_w = L.add((long)E * E); y.q_b = L.add(E); y.k_w = L.add((long)E * E); y.k_b = L.add(E);
    y.v_w = L.add((long)E * E); y.v_b = L.add(E); y.o_w = L.add((long)E * E); y.o_b = L.add(E);
    y.n1_g = L.add(E); y.n1_b = L.add(E);
    y.f1_w = L.add((long)HD * E); y.f1_b = L.add(HD); y.f2_w = L.add((long)E * HD); y.f2_b = L.add(E);
    y.n2_g = L.add(E); y.n2_b = L.add(E);
  }
  L.tn_g = L.add(E); L.tn_b = L.add(E);
  if (c->use_feat) {
    MIVIT_CHECK_ARG(c->feat_dim >= 1, "global_feature_dim must be given with use_global_features");
    L.fp0_w = L.add((long)E * c->feat_dim); L.fp0_b = L.add(E); L.fp2_w = L.add((long)E * E); L.fp2_b = L.add(E);
  }
  L.head_in = (c->use_feat && c->fusion == 1) ? 2 * E : E;
  L.h0_w = L.add((long)c->head_hidden * L.head_in); L.h0_b = L.add(c->head_hidden);
  L.h3_w = L.add(c->head_hidden); L.h3_b = L.add(1);
  return MIVIT_OK;
}

// ------------------------------------------------------------------- workspace ------------
struct Bump {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct RowsT {  // one pitched-rows bf16 tensor with guards
  __nv_bfloat16* buf = nullptr;  // allocation start
  __nv_bfloat16* row0 = nullptr;
  int C = 0;
};

struct BnScratch {
  float *stats, *mi, *ss;  // [2C] each: (sum,sumsq) | (mean,invstd) | (scale,shift)
  int C;
};

// image source of the from-trajectories entry points: the frames are rendered inside the call (Linear / CNN embedding: fused
// with the embedding, they never exist in HBM; DeepResNet: rendered into `frames`, [B,F,P,P])
struct TrajSrc {
  const double* traj;
  int T;
  const mivit_render_params* prm;
  uint64_t seed, seq_offset;
  const uint64_t* seq_off_dev;
  float* frames;
};

struct LayerWS {
  float *q, *k, *v, *ctx, *probs, *ao, *z1, *m1, *r1, *x1, *hpre, *hact, *ff, *z2, *m2, *r2, *x2;
};

struct Workspace {
  // deepresnet embedding
  RowsT raw0, act0, raw1, act1, raw2, raws1, act2, raw3, act3, raw4, raws2;   // the last block output is pooled on the fly
  RowsT draw4, draws2, dact3, draw3, dact2m, dact2s, draw2, draws1, dact1, draw1, dact0m, dact0s, draw0;
  __nv_bfloat16 *wp_c1[2], *wp_c2[2], *wp_sk[2], *wd_c1[2], *wd_c2[2], *wd_sk[2];
  BnScratch bn[7];
  float* bn_sums;
  float *pooled, *dpooled, *fsums;
  // tokens
  float *emb, *m0, *r0, *tok;
  float *fp_pre, *fp_h, *fp_out;
  // ModularTransformer: image embedding before fusion, cleaned features, feature embedding (+ its MLP internals), concat buffer
  float *emb_img, *demb_img, *fclean, *femb, *dfemb, *fe_h, *fe_m, *fe_r, *fe_ln, *fe_act, *dfe_a, *dfe_b, *cat, *dcat;
  LayerWS lyr[kMaxLayers];
  float *mf, *rf, *xf, *headin, *hh, *pred_dummy;
  // backward temporaries
  float *dxa, *dxb, *dq, *dk, *dv, *dctx, *dh, *dh2, *dheadin, *dhh, *dfp, *dfp_h, *demb;
  float* enc_wt;   // transposed encoder weights of the fused layer backward (encoder_fused_bwd.cu)
  long long rows = 0, rows_pad = 0;
  size_t bytes = 0;
};

void take_rows(Bump& b, RowsT& t, int C, long long rows_pad) {
  t.C = C;
  t.buf = b.take<__nv_bfloat16>((size_t)(rows_pad + 2 * kGuard) * C);
  t.row0 = t.buf ? t.buf + (size_t)kGuard * C : nullptr;
}

void carve(const mivit_vit_config* c, int B, void* base, Workspace& w) {
  Bump b{reinterpret_cast<uint8_t*>(base)};
  const int E = c->E, HD = c->HD, F = c->F, S = F + (c->use_reg ? 1 : 0), H = c->H;
  const long long NF = (long long)B * F, T = (long long)B * S;
  const bool has_img = !c->modular || c->mod_mode != 1;
  if (c->embedding == 2 && has_img) {
    w.rows = NF * (c->P + 1) * (c->P + 1);
    w.rows_pad = (w.rows + 127) / 128 * 128;
    const long long rp = w.rows_pad;
    take_rows(b, w.raw0, 32, rp); take_rows(b, w.act0, 32, rp);
    take_rows(b, w.raw1, 64, rp); take_rows(b, w.act1, 64, rp); take_rows(b, w.raw2, 64, rp);
    take_rows(b, w.raws1, 64, rp); take_rows(b, w.act2, 64, rp);
    take_rows(b, w.raw3, 128, rp); take_rows(b, w.act3, 128, rp); take_rows(b, w.raw4, 128, rp);
    take_rows(b, w.raws2, 128, rp);
    // The 13 gradient tensors of the embedding backward live in THREE slots (two of 128 channels, one of 64): a tensor moves
    // into a slot when the previous occupant has been consumed (order of vit_backward_impl; 31 -> 10 units of 0.39 GB at
    // B = 1024).  Guards are zeroed right before each tensor's producer runs (zero_guards): they overlay the dead occupant's rows.
    RowsT slotA, slotB, slotD;
    take_rows(b, slotA, 128, rp); take_rows(b, slotB, 128, rp); take_rows(b, slotD, 64, rp);
    auto sub = [&](RowsT& t, const RowsT& sl, int C, int part) {
      t.C = C;
      t.buf = sl.buf ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(sl.buf) + (size_t)part * (rp + 2 * kGuard) * C * 2) : nullptr;
      t.row0 = t.buf ? t.buf + (size_t)kGuard * C : nullptr;
    };
    sub(w.draw4, slotA, 128, 0); sub(w.draws2, slotB, 128, 0);
    sub(w.dact2s, slotD, 64, 0);                                  // block 2's skip dgrad runs first ...
    sub(w.dact3, slotB, 128, 0);                                  // ... so draws2 is consumed when conv2's dgrad writes dact3
    sub(w.draw3, slotA, 128, 0);                                  // draw4: consumed by conv2's wgrad / dgrad
    sub(w.dact2m, slotB, 64, 0);                                  // dact3: consumed by bn1's backward
    sub(w.draw2, slotA, 64, 0); sub(w.draws1, slotA, 64, 1);      // draw3: consumed by conv1's wgrad / dgrad
    sub(w.dact1, slotB, 64, 0); sub(w.draw1, slotB, 64, 1);       // dact2m (and dact2s): consumed by block 1's bn2 / skip-bn backward
    sub(w.dact0m, slotD, 32, 0); sub(w.dact0s, slotD, 32, 1);
    sub(w.draw0, slotA, 32, 0);                                   // draw2: consumed by block 1's conv2 wgrad / dgrad
    const int ci[2] = {32, 64}, co[2] = {64, 128};
    for (int i = 0; i < 2; ++i) {
      w.wp_c1[i] = b.take<__nv_bfloat16>((size_t)9 * ci[i] * co[i]); w.wd_c1[i] = b.take<__nv_bfloat16>((size_t)9 * ci[i] * co[i]);
      w.wp_c2[i] = b.take<__nv_bfloat16>((size_t)9 * co[i] * co[i]); w.wd_c2[i] = b.take<__nv_bfloat16>((size_t)9 * co[i] * co[i]);
      w.wp_sk[i] = b.take<__nv_bfloat16>((size_t)ci[i] * co[i]); w.wd_sk[i] = b.take<__nv_bfloat16>((size_t)ci[i] * co[i]);
    }
    const int bc[7] = {32, 64, 64, 64, 128, 128, 128};
    for (int i = 0; i < 7; ++i) {
      w.bn[i].C = bc[i];
      w.bn[i].stats = b.take<float>(2 * bc[i]); w.bn[i].mi = b.take<float>(2 * bc[i]); w.bn[i].ss = b.take<float>(2 * bc[i]);
    }
    w.bn_sums = b.take<float>(12 * 128);
    w.pooled = b.take<float>(NF * 128); w.dpooled = b.take<float>(NF * 128); w.fsums = b.take<float>(NF * 3 * 128);
  }
  w.emb = b.take<float>(NF * E); w.m0 = b.take<float>(NF); w.r0 = b.take<float>(NF);
  w.tok = b.take<float>(T * E);
  w.emb_img = w.emb;
  if (c->modular && c->mod_mode != 0) {
    const int Ei = image_embed_dim(c), fd = c->feat_dim;
    if (c->mod_mode == 2) { w.emb_img = b.take<float>(NF * Ei); w.demb_img = b.take<float>(NF * Ei); }
    w.fclean = b.take<float>(NF * fd); w.femb = b.take<float>(NF * E); w.dfemb = b.take<float>(NF * E);
    w.fe_h = b.take<float>(NF * 2 * E); w.fe_ln = b.take<float>(NF * 2 * E); w.fe_act = b.take<float>(NF * 2 * E);
    w.dfe_a = b.take<float>(NF * 2 * E); w.dfe_b = b.take<float>(NF * 2 * E); w.fe_m = b.take<float>(NF); w.fe_r = b.take<float>(NF);
    w.cat = b.take<float>(NF * 2 * E); w.dcat = b.take<float>(NF * 2 * E);
  }
  if (c->use_feat) { w.fp_pre = b.take<float>((size_t)B * E); w.fp_h = b.take<float>((size_t)B * E); w.fp_out = b.take<float>((size_t)B * E); }
  for (int l = 0; l < c->L; ++l) {
    LayerWS& y = w.lyr[l];
    y.q = b.take<float>(T * E); y.k = b.take<float>(T * E); y.v = b.take<float>(T * E); y.ctx = b.take<float>(T * E);
    y.probs = b.take<float>(attention_probs_floats(B, S, E, H)); y.ao = b.take<float>(T * E);
    y.z1 = b.take<float>(T * E); y.m1 = b.take<float>(T); y.r1 = b.take<float>(T); y.x1 = b.take<float>(T * E);
    y.hpre = b.take<float>(T * HD); y.hact = b.take<float>(T * HD); y.ff = b.take<float>(T * E);
    y.z2 = b.take<float>(T * E); y.m2 = b.take<float>(T); y.r2 = b.take<float>(T); y.x2 = b.take<float>(T * E);
  }
  const int hin = (c->use_feat && c->fusion == 1) ? 2 * E : E;
  const long long HR = c->per_frame ? T : B;   // rows the MLP head runs on: every token (ModularTransformer, single_prediction=False) or one per sequence
  w.mf = b.take<float>(T); w.rf = b.take<float>(T); w.xf = b.take<float>(T * E);
  w.headin = b.take<float>((size_t)B * hin); w.hh = b.take<float>((size_t)HR * c->head_hidden);
  w.dxa = b.take<float>(T * E); w.dxb = b.take<float>(T * E); w.dq = b.take<float>(T * E); w.dk = b.take<float>(T * E);
  w.dv = b.take<float>(T * E); w.dctx = b.take<float>(T * E); w.dh = b.take<float>(T * HD); w.dh2 = b.take<float>(T * HD);
  w.dheadin = b.take<float>((size_t)B * hin); w.dhh = b.take<float>((size_t)HR * c->head_hidden);
  w.dfp = b.take<float>((size_t)B * E); w.dfp_h = b.take<float>((size_t)B * E); w.demb = b.take<float>(NF * E);
  w.enc_wt = b.take<float>((size_t)c->L * encoder_bwd_packed_floats(E, HD));
  w.bytes = (b.off + 255) & ~(size_t)255;
}

// tensor-core (tf32) nn.Linear kernels are used when the product path is selected (conv_impl == 1) and the
// shape / alignment allows it; the fp32 SIMT GEMM otherwise (and always for the SIMT cross-check path).
static thread_local int g_linear_tc = 0;
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Y[M,N] = X[M,K] W[N,K]^T + b
int linear_fwd(const float* X, const float* W, const float* bias, float* Y, int M, int N, int K, int relu, cudaStream_t st) {
  if (g_linear_tc && linear_tc_supported(M, K, N) && al16(X) && al16(W) && al16(Y) && (bias == nullptr || al16(bias)))
    return linear_tc(X, W, bias, Y, M, K, N, 0, relu, 0, st);
  return gemm_f32(X, K, 1, W, 1, K, Y, N, M, N, K, bias, relu, 0, 1, st);
}
// dW[N,K] += dY[M,N]^T X[M,K];  db[N] += colsum(dY);  dX[M,K] (=|+=) dY W
int linear_bwd(const float* X, const float* W, const float* dY, float* dW, float* db, float* dX, int M, int N, int K,
               int accumulate_dx, cudaStream_t st) {
  if (g_linear_tc && linear_wgrad_tc_supported(M, N, K) && al16(X) && al16(dY) && al16(dW)) {
    CK(linear_wgrad_tc(dY, X, dW, db, M, N, K, st));   // bias gradient rides along as an extra accumulator column
    db = nullptr;
  } else {
    int split = M / 128;  // token-dimension split-K: enough CTAs to fill 148 SMs even for 64x64 outputs
    if (split < 1) split = 1;
    if (split > 296) split = 296;
    if (split == 1) {
      CK(gemm_f32(dY, 1, N, X, K, 1, dW, K, N, K, M, nullptr, 0, 1, 1, st));
    } else {
      CK(gemm_f32(dY, 1, N, X, K, 1, dW, K, N, K, M, nullptr, 0, 0, split, st));
    }
  }
  if (db) CK(colsum_f32(dY, N, M, N, db, st));
  if (dX) {
    // dX[M,K] = dY[M,N] W[N,K]: reduction dim N, output cols K, W is the MN-major operand [N][K]
    if (g_linear_tc && linear_tc_supported(M, N, K) && al16(dY) && al16(W) && al16(dX)) {
      CK(linear_tc(dY, W, nullptr, dX, M, N, K, 1, 0, accumulate_dx, st));
    } else {
      CK(gemm_f32(dY, N, 1, W, K, 1, dX, K, M, K, N, nullptr, 0, accumulate_dx, 1, st));
    }
  }
  return MIVIT_OK;
}

int zero_guards(RowsT& t, long long rows_pad, cudaStream_t st) {
  MIVIT_CUDA_CHECK(cudaMemsetAsync(t.buf, 0, (size_t)kGuard * t.C * 2, st));
  MIVIT_CUDA_CHECK(cudaMemsetAsync(t.row0 + (size_t)rows_pad * t.C, 0, (size_t)kGuard * t.C * 2, st));
  return MIVIT_OK;
}

// ---- synchronised BatchNorm (data-parallel parity mode) --------------------------------------------------------------
// The host side (training.py) registers a SUM all-reduce over the data-parallel group; with it, every BatchNorm of a
// TRAINING forward/backward uses statistics of the GLOBAL batch (torch.nn.SyncBatchNorm semantics): forward all-reduces
// (sum x, sum x^2) per layer, backward all-reduces (sum g, sum g*xhat); counts are multiplied by the world size (equal
// per-rank batches).  Parameter gradients keep the LOCAL sums -- the gradient all-reduce adds the ranks up afterwards.
mivit_allreduce_fn g_ar_fn = nullptr;
void* g_ar_user = nullptr;
int g_ar_world = 1;

int run_bn_finalize(const mivit_vit_config* c, Workspace& w, int i, const float* g, const float* bta, float* bn_running,
                    long long* nbt, double count, int training, cudaStream_t st) {
  static const int rm_off[7] = {0, 64, 192, 320, 448, 704, 960};  // [mean C | var C] per layer: 32,64,64,64,128,128,128
  BnScratch& s = w.bn[i];
  if (training && mivit_bn_sync_world() > 1) {
    CK(mivit_bn_sync(s.stats, 2 * s.C, st));
    count *= mivit_bn_sync_world();
  }
  float* rm = bn_running ? bn_running + rm_off[i] : nullptr;
  float* rv = rm ? rm + s.C : nullptr;
  return bn_finalize(s.stats, g, bta, rm, rv, nbt ? nbt + i : nullptr, s.mi, s.mi + s.C, s.ss, s.ss + s.C, s.C, count, c->bn_eps,
                     c->bn_momentum, training, st);
}

}  // namespace

int mivit_bn_sync_world() { return mivit_bn_peer_active() ? mivit_bn_peer_world() : g_ar_fn != nullptr ? g_ar_world : 1; }
int mivit_bn_sync(float* buf, long long n, cudaStream_t st) {
  if (mivit_bn_peer_active()) return mivit_bn_peer_sync(buf, n, st);
  if (g_ar_fn == nullptr || g_ar_world <= 1) return MIVIT_OK;
  const int rc = g_ar_fn(buf, (int64_t)n, (void*)st, g_ar_user);
  if (rc != 0) {
    mivit_set_error("the registered all-reduce hook failed (rc = %d)", rc);
    return MIVIT_ERR_INVALID;
  }
  return MIVIT_OK;
}
extern "C" void mivit_set_allreduce_hook(mivit_allreduce_fn fn, void* user, int32_t world_size) {
  g_ar_fn = fn;
  g_ar_user = user;
  g_ar_world = fn != nullptr && world_size > 1 ? world_size : 1;
}

extern "C" int mivit_vit_param_sizes(const mivit_vit_config* cfg, int64_t* sizes, int32_t max_count) {
  ParamLayout L;
  CK(build_layout(cfg, L));
  if (sizes != nullptr) {
    MIVIT_CHECK_ARG(max_count >= L.count, "sizes array too small (%d < %d)", max_count, L.count);
    for (int i = 0; i < L.count; ++i) sizes[i] = L.sizes[i];
  }
  return MIVIT_OK;
}

extern "C" int32_t mivit_vit_param_count(const mivit_vit_config* cfg) {
  ParamLayout L;
  if (build_layout(cfg, L)) return -1;
  return L.count;
}

extern "C" int64_t mivit_vit_workspace_bytes(const mivit_vit_config* cfg, int32_t B) {
  ParamLayout L;
  if (build_layout(cfg, L) || B < 1) return -1;
  Workspace w;
  carve(cfg, B, nullptr, w);
  return (int64_t)w.bytes;
}

static int check_src(const mivit_vit_config* c, const ParamLayout& L, const TrajSrc* src) {
  MIVIT_CHECK_ARG(L.has_img, "a 'features_only' ModularTransformer has no image path to render for");
  MIVIT_CHECK_ARG(src->traj && src->prm, "NULL trajectory / render parameters");
  MIVIT_CHECK_ARG(src->prm->P == c->P, "output_size %d does not match the model's patch_size %d", src->prm->P, c->P);
  MIVIT_CHECK_ARG(src->prm->n >= 1 && src->T % src->prm->n == 0, "T is not divisble by posPerFrame");
  MIVIT_CHECK_ARG(src->T / src->prm->n == c->F, "the trajectories give %d frames, the configuration says %d", src->T / src->prm->n, c->F);
  MIVIT_CHECK_ARG(c->embedding != 2 || src->frames, "the DeepResNet embedding needs a [B,F,P,P] frame buffer (batch statistics)");
  return MIVIT_OK;
}

static int vit_forward_impl(const mivit_vit_config* c, int32_t B, const float* x, const float* features,
                            const float* params, float* bn_running, int64_t* bn_num_batches, void* workspace,
                            float* pred, int32_t training, void* stream, const TrajSrc* src) {
  ParamLayout L;
  CK(build_layout(c, L));
  MIVIT_CHECK_ARG(B >= 1 && params && workspace && pred, "bad arguments");
  if (src != nullptr) {
    CK(check_src(c, L, src));
    if (c->embedding == 2) {   // batch statistics need every frame of the batch: render to HBM (20 KB per sequence), then as usual
      CK(render_v1_launch(src->traj, B, src->T, src->prm, src->seed, src->seq_offset, src->seq_off_dev, src->frames,
                          (long long)c->F * c->P * c->P, (cudaStream_t)stream));
      x = src->frames;
    } else {
      x = reinterpret_cast<const float*>(src->traj);   // non-NULL marker; the fused kernel reads the trajectories
    }
  }
  MIVIT_CHECK_ARG(x || !L.has_img, c->modular ? (c->mod_mode == 2 ? "Both images and features are required for 'both' mode"
                                                                   : "Images are required for 'images_only' mode") : "bad arguments");
  MIVIT_CHECK_ARG(!c->use_feat || features, "Global features required for %s fusion", c->fusion ? "late" : "early");
  MIVIT_CHECK_ARG(!(c->modular && c->mod_mode != 0) || features,
                  c->mod_mode == 2 ? "Both images and features are required for 'both' mode" : "Features are required for 'features_only' mode");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  carve(c, B, workspace, w);
  g_linear_tc = c->conv_impl == 1;
  if (training) mivit_bn_peer_begin_step();   // small-exchange call indices restart with every training forward
  const int E = c->E, HD = c->HD, F = c->F, S = F + (c->use_reg ? 1 : 0), H = c->H, P = c->P;
  const int NF = B * F, T = B * S;
  const float* p = params;
  const int Ei = L.E_img;
  float* emb_img = w.emb_img;   // == w.emb unless ModularTransformer fuses it with a feature embedding
  if (!L.has_img) {
    // 'features_only'
  } else if (c->embedding == 2) {
    MIVIT_CHECK_ARG(training || bn_running, "eval-mode BatchNorm needs the running statistics");
    const long long rows = w.rows, rp = w.rows_pad;
    const double cnt = (double)NF * P * P;
    const int impl = c->conv_impl;
    RowsT* guarded[] = {&w.act0, &w.act1, &w.act2, &w.act3};
    for (RowsT* t : guarded) CK(zero_guards(*t, rp, st));
    for (int i = 0; i < 7; ++i) MIVIT_CUDA_CHECK(cudaMemsetAsync(w.bn[i].stats, 0, 2 * w.bn[i].C * sizeof(float), st));
    const ConvShifts s3 = make_shifts(P, 9, false);
    long long* nbt = (long long*)bn_num_batches;
    // stem
    CK(conv0_forward(x, p + L.conv0_w, w.raw0.row0, w.bn[0].stats, rows, rp, P, st));
    CK(run_bn_finalize(c, w, 0, p + L.bn0_g, p + L.bn0_b, bn_running, nbt, cnt, training, st));
    CK(bn_apply(w.raw0.row0, w.bn[0].ss, nullptr, nullptr, w.act0.row0, rows, rp, P, 32, st));
    const int ci[2] = {32, 64}, co[2] = {64, 128};
    RowsT* in[2] = {&w.act0, &w.act2};
    RowsT* r1[2] = {&w.raw1, &w.raw3};
    RowsT* a1[2] = {&w.act1, &w.act3};
    RowsT* r2[2] = {&w.raw2, &w.raw4};
    RowsT* rs[2] = {&w.raws1, &w.raws2};
    RowsT* out[2] = {&w.act2, nullptr};
    for (int b = 0; b < 2; ++b) {
      const auto& R = L.rb[b];
      const int i1 = 1 + 3 * b, i2 = 2 + 3 * b, is = 3 + 3 * b;
      CK(pack_conv_weights(p + R.c1_w, w.wp_c1[b], co[b], ci[b], 9, 0, st));
      CK(pack_conv_weights(p + R.c2_w, w.wp_c2[b], co[b], co[b], 9, 0, st));
      CK(pack_conv_weights(p + R.sk_w, w.wp_sk[b], co[b], ci[b], 1, 0, st));
      // conv1 (3x3) and the 1x1 skip convolution read the same input: one fused launch
      CK(conv_rows_forward_fused(in[b]->row0, w.wp_c1[b], w.wp_sk[b], r1[b]->row0, rs[b]->row0, w.bn[i1].stats, w.bn[is].stats,
                                 rows, P, ci[b], co[b], 9, s3, impl, st));
      CK(run_bn_finalize(c, w, i1, p + R.bn1_g, p + R.bn1_b, bn_running, nbt, cnt, training, st));
      CK(bn_apply(r1[b]->row0, w.bn[i1].ss, nullptr, nullptr, a1[b]->row0, rows, rp, P, co[b], st));
      CK(conv_rows_forward(a1[b]->row0, w.wp_c2[b], r2[b]->row0, w.bn[i2].stats, rows, P, co[b], co[b], 9, s3, impl, st));
      CK(run_bn_finalize(c, w, i2, p + R.bn2_g, p + R.bn2_b, bn_running, nbt, cnt, training, st));
      CK(run_bn_finalize(c, w, is, p + R.bns_g, p + R.bns_b, bn_running, nbt, cnt, training, st));
      if (b == 0) {
        CK(bn_apply(r2[b]->row0, w.bn[i2].ss, rs[b]->row0, w.bn[is].ss, out[b]->row0, rows, rp, P, co[b], st));
      } else {  // last block: its output only feeds the average pool -> fused, the activation is never materialised
        CK(bn_apply_pool(r2[b]->row0, w.bn[i2].ss, rs[b]->row0, w.bn[is].ss, w.pooled, training ? w.fsums : nullptr, NF, P, co[b], st));
      }
    }
    CK(linear_fwd(w.pooled, p + L.fc_w, p + L.fc_b, emb_img, NF, Ei, 128, 0, st));
  } else if (src != nullptr) {
    // renderer fused with the frame embedding: emb = W . frame + b straight from the trajectories, frames never reach HBM
    CK(render_embed_linear_launch(src->traj, B, src->T, src->prm, src->seed, src->seq_offset, src->seq_off_dev, p + L.proj_w, 0,
                                  p + L.proj_b, Ei, emb_img, nullptr, 0, st));
  } else {
    CK(linear_fwd(x, p + L.proj_w, p + L.proj_b, emb_img, NF, Ei, P * P, 0, st));
  }
  if (c->modular && c->mod_mode != 0) {   // ModularTransformer.forward, helpers/models.py:528-570
    const int fd = c->feat_dim;
    CK(nan_to_num_f32(features, w.fclean, (long long)NF * fd, st));
    float* femb = c->mod_mode == 1 ? w.emb : w.femb;
    if (L.has_femb) {
      if (c->mod_fembed == 0) {
        CK(linear_fwd(w.fclean, p + L.fe0_w, p + L.fe0_b, femb, NF, E, fd, 0, st));
      } else {   // Linear(fd, 2E) -> LayerNorm(2E) -> GELU -> Linear(2E, E)
        CK(linear_fwd(w.fclean, p + L.fe0_w, p + L.fe0_b, w.fe_h, NF, 2 * E, fd, 0, st));
        CK(layernorm_fwd(w.fe_h, nullptr, p + L.feln_g, p + L.feln_b, nullptr, w.fe_ln, w.fe_m, w.fe_r, NF, 2 * E, c->ln_eps, 0, 0, 0, st));
        CK(act_fwd(w.fe_ln, w.fe_act, (long long)NF * 2 * E, 1, st));
        CK(linear_fwd(w.fe_act, p + L.fe3_w, p + L.fe3_b, femb, NF, E, 2 * E, 0, st));
      }
    }
    if (c->mod_mode == 2) {
      if (c->mod_fusion == 0) {          // 'add'
        CK(add_f32(emb_img, w.femb, w.emb, (long long)NF * E, st));
      } else if (c->mod_fusion == 1) {   // 'concat_proj': fusion_layer(cat([image, feature], -1))
        MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.cat, (size_t)2 * E * 4, emb_img, (size_t)E * 4, (size_t)E * 4, NF, cudaMemcpyDeviceToDevice, st));
        MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.cat + E, (size_t)2 * E * 4, w.femb, (size_t)E * 4, (size_t)E * 4, NF, cudaMemcpyDeviceToDevice, st));
        CK(linear_fwd(w.cat, p + L.fu_w, p + L.fu_b, w.emb, NF, E, 2 * E, 0, st));
      } else {                           // 'concat_features': cat([image (E - fd wide), raw features], -1)
        MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.emb, (size_t)E * 4, emb_img, (size_t)Ei * 4, (size_t)Ei * 4, NF, cudaMemcpyDeviceToDevice, st));
        MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.emb + Ei, (size_t)E * 4, w.fclean, (size_t)fd * 4, (size_t)fd * 4, NF, cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  // x = self.norm(x), written straight into the token slots behind the regression token
  CK(layernorm_fwd(w.emb, nullptr, p + L.norm_g, p + L.norm_b, nullptr, w.tok, w.m0, w.r0, NF, E, c->ln_eps, F, S,
                   c->use_reg ? 1 : 0, st));
  if (c->use_feat) {
    CK(linear_fwd(features, p + L.fp0_w, p + L.fp0_b, w.fp_h, B, E, c->feat_dim, 1, st));
    CK(linear_fwd(w.fp_h, p + L.fp2_w, p + L.fp2_b, w.fp_out, B, E, E, 0, st));
  }
  CK(tokens_finish(w.tok, c->use_reg ? p + L.reg : nullptr, (c->use_feat && c->fusion == 0 && c->use_reg) ? w.fp_out : nullptr,
                   c->use_pos ? p + L.pos : nullptr, B, S, E, st));
  const float* xin = w.tok;
  bool fused_layers = g_linear_tc && c->activation == 0 && encoder_fused_supported(B, S, E, HD, H) && !getenv("MIVIT_NO_FUSED_ENCODER");
  for (int l = 0; fused_layers && l < c->L; ++l) {   // TMA reads the weights in place: 16-byte aligned matrices
    const auto& Y = L.lyr[l];
    fused_layers = al16(p + Y.q_w) && al16(p + Y.k_w) && al16(p + Y.v_w) && al16(p + Y.o_w) && al16(p + Y.f1_w) && al16(p + Y.f2_w);
  }
  for (int l = 0; l < c->L; ++l) {
    const auto& Y = L.lyr[l];
    LayerWS& y = w.lyr[l];
    if (fused_layers) {   // the whole layer in one persistent kernel (encoder_fused.cu); writes what the backward below reads
      EncoderLayerIO io;
      io.x = xin;
      io.wq = p + Y.q_w; io.bq = p + Y.q_b; io.wk = p + Y.k_w; io.bk = p + Y.k_b; io.wv = p + Y.v_w; io.bv = p + Y.v_b;
      io.wo = p + Y.o_w; io.bo = p + Y.o_b; io.g1 = p + Y.n1_g; io.be1 = p + Y.n1_b; io.w1 = p + Y.f1_w; io.bf1 = p + Y.f1_b;
      io.w2 = p + Y.f2_w; io.bf2 = p + Y.f2_b; io.g2 = p + Y.n2_g; io.be2 = p + Y.n2_b;
      io.q = y.q; io.k = y.k; io.v = y.v; io.ctx = y.ctx; io.lse = y.probs; io.z1 = y.z1; io.m1 = y.m1; io.r1 = y.r1; io.x1 = y.x1;
      io.hact = y.hact; io.z2 = y.z2; io.m2 = y.m2; io.r2 = y.r2; io.x2 = y.x2;
      CK(encoder_layer_fwd(io, B, S, E, HD, H, c->ln_eps, st));
      xin = y.x2;
      continue;
    }
    if (g_linear_tc && linear_tc_supported(T, E, E) && al16(xin) && al16(p + Y.q_w) && al16(p + Y.k_w) && al16(p + Y.v_w) &&
        al16(p + Y.q_b) && al16(p + Y.k_b) && al16(p + Y.v_b) && al16(y.q) && al16(y.k) && al16(y.v)) {
      // the three projections read the same tokens and are independent: one launch, three problems (linear_tc.cu)
      const float* As[3] = {xin, xin, xin};
      const float* Ws[3] = {p + Y.q_w, p + Y.k_w, p + Y.v_w};
      const float* bs[3] = {p + Y.q_b, p + Y.k_b, p + Y.v_b};
      float* Ys[3] = {y.q, y.k, y.v};
      CK(linear_tc_batched(3, As, Ws, bs, Ys, T, E, E, 0, 0, 0, st));
    } else {
      CK(linear_fwd(xin, p + Y.q_w, p + Y.q_b, y.q, T, E, E, 0, st));
      CK(linear_fwd(xin, p + Y.k_w, p + Y.k_b, y.k, T, E, E, 0, st));
      CK(linear_fwd(xin, p + Y.v_w, p + Y.v_b, y.v, T, E, E, 0, st));
    }
    CK(attention_fwd(y.q, y.k, y.v, y.ctx, y.probs, B, S, E, H, st));
    CK(linear_fwd(y.ctx, p + Y.o_w, p + Y.o_b, y.ao, T, E, E, 0, st));
    CK(layernorm_fwd(y.ao, xin, p + Y.n1_g, p + Y.n1_b, y.z1, y.x1, y.m1, y.r1, T, E, c->ln_eps, 0, 0, 0, st));
    if (c->activation == 0) {   // F.relu: fused into the fc1 epilogue (its gradient only needs the sign, i.e. hact > 0)
      CK(linear_fwd(y.x1, p + Y.f1_w, p + Y.f1_b, y.hact, T, HD, E, 1, st));
    } else {
      CK(linear_fwd(y.x1, p + Y.f1_w, p + Y.f1_b, y.hpre, T, HD, E, 0, st));
      CK(act_fwd(y.hpre, y.hact, (long long)T * HD, c->activation, st));
    }
    CK(linear_fwd(y.hact, p + Y.f2_w, p + Y.f2_b, y.ff, T, E, HD, 0, st));
    CK(layernorm_fwd(y.ff, y.x1, p + Y.n2_g, p + Y.n2_b, y.z2, y.x2, y.m2, y.r2, T, E, c->ln_eps, 0, 0, 0, st));
    xin = y.x2;
  }
  CK(layernorm_fwd(xin, nullptr, p + L.tn_g, p + L.tn_b, nullptr, w.xf, w.mf, w.rf, T, E, c->ln_eps, 0, 0, 0, st));
  if (c->per_frame) {   // ModularTransformer, no regression token, single_prediction=False: the head sees every token (:585-593)
    CK(linear_fwd(w.xf, p + L.h0_w, p + L.h0_b, w.hh, T, c->head_hidden, E, 1, st));
    CK(linear_fwd(w.hh, p + L.h3_w, p + L.h3_b, pred, T, 1, c->head_hidden, 0, st));
    return MIVIT_OK;
  }
  CK(pool_tokens(w.xf, w.headin, B, S, E, L.head_in, c->use_reg, st));
  if (c->use_feat && c->fusion == 1)
    MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.headin + E, (size_t)L.head_in * 4, w.fp_out, (size_t)E * 4, (size_t)E * 4, B,
                                       cudaMemcpyDeviceToDevice, st));
  CK(linear_fwd(w.headin, p + L.h0_w, p + L.h0_b, w.hh, B, c->head_hidden, L.head_in, 1, st));
  CK(linear_fwd(w.hh, p + L.h3_w, p + L.h3_b, pred, B, 1, c->head_hidden, 0, st));
  return MIVIT_OK;
}

extern "C" int mivit_vit_forward(const mivit_vit_config* c, int32_t B, const float* x, const float* features,
                                 const float* params, float* bn_running, int64_t* bn_num_batches, void* workspace,
                                 float* pred, int32_t training, void* stream) {
  return vit_forward_impl(c, B, x, features, params, bn_running, bn_num_batches, workspace, pred, training, stream, nullptr);
}

extern "C" int mivit_vit_forward_traj(const mivit_vit_config* c, int32_t B, const double* traj, int32_t T,
                                      const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset,
                                      const uint64_t* seq_offset_dev, float* frames,
                                      const float* features, const float* params, float* bn_running, int64_t* bn_num_batches,
                                      void* workspace, float* pred, int32_t training, void* stream) {
  const TrajSrc src{traj, T, prm, seed, seq_offset, seq_offset_dev, frames};
  return vit_forward_impl(c, B, nullptr, features, params, bn_running, bn_num_batches, workspace, pred, training, stream, &src);
}

// part: 0 = whole backward; 1 = everything up to (not including) the image embedding -- head, encoder layers, tokens, feature
// paths, embedding LayerNorm: all gradients behind mivit_vit_embedding_param_count() are final afterwards; 2 = image embedding only.
static int vit_backward_impl(const mivit_vit_config* c, int32_t B, const float* x, const float* features, const float* dpred,
                             const float* params, float* grads, void* workspace, int part, void* stream, const TrajSrc* src = nullptr) {
  ParamLayout L;
  CK(build_layout(c, L));
  if (src != nullptr) {
    CK(check_src(c, L, src));
    x = c->embedding == 2 ? src->frames : reinterpret_cast<const float*>(src->traj);
  }
  MIVIT_CHECK_ARG(B >= 1 && (x || !L.has_img) && dpred && params && grads && workspace, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  carve(c, B, workspace, w);
  const int E = c->E, HD = c->HD, F = c->F, S = F + (c->use_reg ? 1 : 0), H = c->H, P = c->P;
  const int Ei = L.E_img;
  const int NF = B * F, T = B * S, hin = L.head_in, hhid = c->head_hidden;
  const float* p = params;
  float* g = grads;
  g_linear_tc = c->conv_impl == 1;
  // gradient of the image embedding output [NF, Ei]: the embedding LayerNorm's input gradient, or its image slice / share
  const float* d_img = (c->modular && c->mod_mode == 2 && c->mod_fusion != 0) ? w.demb_img : w.demb;
  auto tokens_part = [&]() -> int {
  MIVIT_CUDA_CHECK(cudaMemsetAsync(g, 0, (size_t)L.off * sizeof(float), st));
  // head
  const int HR = c->per_frame ? T : B;
  CK(linear_bwd(w.hh, p + L.h3_w, dpred, g + L.h3_w, g + L.h3_b, w.dhh, HR, 1, hhid, 0, st));
  CK(act_bwd(w.dhh, w.hh, w.dhh, (long long)HR * hhid, 0, st));
  const bool late = c->use_feat && c->fusion == 1, early = c->use_feat && c->fusion == 0 && c->use_reg;
  if (c->per_frame) {   // the head read every token of the final LayerNorm's output: its input gradient IS d(xf)
    CK(linear_bwd(w.xf, p + L.h0_w, w.dhh, g + L.h0_w, g + L.h0_b, w.dxa, T, hhid, E, 0, st));
  } else {
    CK(linear_bwd(w.headin, p + L.h0_w, w.dhh, g + L.h0_w, g + L.h0_b, w.dheadin, B, hhid, hin, 0, st));
    if (late)
      MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.dfp, (size_t)E * 4, w.dheadin + E, (size_t)hin * 4, (size_t)E * 4, B,
                                         cudaMemcpyDeviceToDevice, st));
    CK(pool_tokens_bwd(w.dheadin, w.dxa, B, S, E, hin, c->use_reg, st));
  }
  const float* xlast = w.lyr[c->L - 1].x2;
  CK(layernorm_bwd(w.dxa, xlast, w.mf, w.rf, p + L.tn_g, w.dxb, g + L.tn_g, g + L.tn_b, T, E, 0, 0, 0, st));
  float* dx = w.dxb;    // gradient w.r.t. the current layer output
  float* tmp = w.dxa;
  // one persistent kernel per layer for the whole input-gradient chain (encoder_fused_bwd.cu) + three weight-gradient launches
  bool fused_bwd = g_linear_tc && c->activation == 0 && encoder_fused_supported(B, S, E, HD, H) && c->L <= kEncMaxLayers &&
                   linear_wgrad_tc_supported(T, E, E) && linear_wgrad_tc_supported(T, E, HD) && linear_wgrad_tc_supported(T, HD, E) &&
                   !getenv("MIVIT_NO_FUSED_ENCODER_BWD");
  for (int l = 0; fused_bwd && l < c->L; ++l) {   // the weight-gradient flushes are 16-byte vector reductions
    const auto& Y = L.lyr[l];
    fused_bwd = al16(g + Y.q_w) && al16(g + Y.k_w) && al16(g + Y.v_w) && al16(g + Y.o_w) && al16(g + Y.f1_w) && al16(g + Y.f2_w);
  }
  if (fused_bwd) {
    long long offs[kEncMaxLayers][6];
    for (int l = 0; l < c->L; ++l) {
      const auto& Y = L.lyr[l];
      const long long o[6] = {Y.q_w, Y.k_w, Y.v_w, Y.o_w, Y.f1_w, Y.f2_w};
      for (int i = 0; i < 6; ++i) offs[l][i] = o[i];
    }
    CK(encoder_pack_bwd_weights(p, offs, c->L, E, HD, w.enc_wt, st));
  }
  for (int l = c->L - 1; l >= 0; --l) {
    const auto& Y = L.lyr[l];
    LayerWS& y = w.lyr[l];
    const float* xin = l == 0 ? w.tok : w.lyr[l - 1].x2;
    if (fused_bwd) {
      EncoderLayerBwdIO io;
      io.dy = dx; io.dx = dx;
      io.z2 = y.z2; io.m2 = y.m2; io.r2 = y.r2; io.hact = y.hact; io.z1 = y.z1; io.m1 = y.m1; io.r1 = y.r1;
      io.q = y.q; io.k = y.k; io.v = y.v; io.lse = y.probs; io.ctx = y.ctx;
      io.g2 = p + Y.n2_g; io.g1 = p + Y.n1_g;
      io.packed = w.enc_wt + (size_t)l * encoder_bwd_packed_floats(E, HD);
      io.dz2 = tmp; io.dh2 = w.dh2; io.dz1 = w.dctx; io.dq = w.dq; io.dk = w.dk; io.dv = w.dv;
      io.dg2 = g + Y.n2_g; io.db2 = g + Y.n2_b; io.dg1 = g + Y.n1_g; io.db1 = g + Y.n1_b;
      CK(encoder_layer_bwd(io, B, S, E, HD, H, st));
      // the six weight (and bias) gradients of the layer: one launch
      const float* dYs[6] = {w.dq, w.dk, w.dv, w.dctx, tmp, w.dh2};
      const float* Xs[6] = {xin, xin, xin, y.ctx, y.hact, y.x1};
      float* dWs[6] = {g + Y.q_w, g + Y.k_w, g + Y.v_w, g + Y.o_w, g + Y.f2_w, g + Y.f1_w};
      float* dbs[6] = {g + Y.q_b, g + Y.k_b, g + Y.v_b, g + Y.o_b, g + Y.f2_b, g + Y.f1_b};
      const int Ns[6] = {E, E, E, E, E, HD}, Ks[6] = {E, E, E, E, HD, E};
      if (linear_wgrad_tc_multi_supported(6, T, Ns, Ks) && !getenv("MIVIT_NO_WGRAD_MULTI")) {
        CK(linear_wgrad_tc_multi(6, dYs, Xs, dWs, dbs, T, Ns, Ks, st));
      } else {
        CK(linear_wgrad_tc_batched(4, dYs, Xs, dWs, dbs, T, E, E, st));
        CK(linear_wgrad_tc(tmp, y.hact, g + Y.f2_w, g + Y.f2_b, T, E, HD, st));
        CK(linear_wgrad_tc(w.dh2, y.x1, g + Y.f1_w, g + Y.f1_b, T, HD, E, st));
      }
      continue;
    }
    // x2 = LN2(x1 + ff)
    CK(layernorm_bwd(dx, y.z2, y.m2, y.r2, p + Y.n2_g, tmp, g + Y.n2_g, g + Y.n2_b, T, E, 0, 0, 0, st));  // tmp = dz2 = dff = dx1
    if (c->activation == 0 && g_linear_tc && linear_tc_supported(T, E, HD) && al16(tmp) && al16(p + Y.f2_w) && al16(w.dh2) &&
        al16(y.hact)) {
      // F.relu: its backward (dh2 = dh * 1[hact > 0]) is a gate in the epilogue of the fc2 input gradient
      CK(linear_bwd(y.hact, p + Y.f2_w, tmp, g + Y.f2_w, g + Y.f2_b, nullptr, T, E, HD, 0, st));
      const float* A1 = tmp; const float* W1 = p + Y.f2_w; float* Y1 = w.dh2;
      CK(linear_tc_batched(1, &A1, &W1, nullptr, &Y1, T, E, HD, 1, 0, 0, st, y.hact));
    } else {
      CK(linear_bwd(y.hact, p + Y.f2_w, tmp, g + Y.f2_w, g + Y.f2_b, w.dh, T, E, HD, 0, st));
      CK(act_bwd(w.dh, c->activation == 0 ? y.hact : y.hpre, w.dh2, (long long)T * HD, c->activation, st));
    }
    CK(linear_bwd(y.x1, p + Y.f1_w, w.dh2, g + Y.f1_w, g + Y.f1_b, tmp, T, HD, E, 1, st));               // tmp += dhpre W1
    // x1 = LN1(xin + ao)
    CK(layernorm_bwd(tmp, y.z1, y.m1, y.r1, p + Y.n1_g, dx, g + Y.n1_g, g + Y.n1_b, T, E, 0, 0, 0, st));   // dx = dz1 = dao = dxin
    const bool tc_attn = g_linear_tc && linear_wgrad_tc_supported(T, E, E) && linear_tc_supported(T, E, E) && al16(xin) && al16(w.dq) &&
                         al16(w.dk) && al16(w.dv) && al16(g + Y.q_w) && al16(g + Y.k_w) && al16(g + Y.v_w) && al16(g + Y.o_w) &&
                         al16(p + Y.q_w) && al16(p + Y.k_w) && al16(p + Y.v_w) && al16(p + Y.o_w) && al16(dx) && al16(y.ctx) &&
                         al16(w.dctx);
    if (tc_attn) {
      CK(linear_tc(dx, p + Y.o_w, nullptr, w.dctx, T, E, E, 1, 0, 0, st));                 // dctx = dz1 W_o
      CK(attention_bwd(y.q, y.k, y.v, y.probs, y.ctx, w.dctx, w.dq, w.dk, w.dv, B, S, E, H, st));
      // the four E x E weight gradients of the attention block (q, k, v, out) in one launch ...
      const float* dYs[4] = {w.dq, w.dk, w.dv, dx};
      const float* Xs[4] = {xin, xin, xin, y.ctx};
      float* dWs[4] = {g + Y.q_w, g + Y.k_w, g + Y.v_w, g + Y.o_w};
      float* dbs[4] = {g + Y.q_b, g + Y.k_b, g + Y.v_b, g + Y.o_b};
      CK(linear_wgrad_tc_batched(4, dYs, Xs, dWs, dbs, T, E, E, st));
      // ... and the three input gradients in one launch, added onto dx (= dz1, the residual path) with vector atomics
      const float* As[3] = {w.dq, w.dk, w.dv};
      const float* Ws[3] = {p + Y.q_w, p + Y.k_w, p + Y.v_w};
      float* Ys[3] = {dx, dx, dx};
      CK(linear_tc_batched(3, As, Ws, nullptr, Ys, T, E, E, 1, 0, 2, st));
    } else {
      CK(linear_bwd(y.ctx, p + Y.o_w, dx, g + Y.o_w, g + Y.o_b, w.dctx, T, E, E, 0, st));
      CK(attention_bwd(y.q, y.k, y.v, y.probs, y.ctx, w.dctx, w.dq, w.dk, w.dv, B, S, E, H, st));
      CK(linear_bwd(xin, p + Y.q_w, w.dq, g + Y.q_w, g + Y.q_b, dx, T, E, E, 1, st));
      CK(linear_bwd(xin, p + Y.k_w, w.dk, g + Y.k_w, g + Y.k_b, dx, T, E, E, 1, st));
      CK(linear_bwd(xin, p + Y.v_w, w.dv, g + Y.v_w, g + Y.v_b, dx, T, E, E, 1, st));
    }
  }
  // tokens: regression token, positional embedding, early-fusion projection
  CK(tokens_finish_bwd(dx, c->use_reg ? g + L.reg : nullptr, early ? w.dfp : nullptr, c->use_pos ? g + L.pos : nullptr, B, S, E, st));
  if (c->use_feat && (late || early)) {
    CK(linear_bwd(w.fp_h, p + L.fp2_w, w.dfp, g + L.fp2_w, g + L.fp2_b, w.dfp_h, B, E, E, 0, st));
    CK(act_bwd(w.dfp_h, w.fp_h, w.dfp_h, (long long)B * E, 0, st));
    CK(linear_bwd(features, p + L.fp0_w, w.dfp_h, g + L.fp0_w, g + L.fp0_b, nullptr, B, E, c->feat_dim, 0, st));
  }
  // embedding LayerNorm (dy is read from the token slots)
  CK(layernorm_bwd(dx, w.emb, w.m0, w.r0, p + L.norm_g, w.demb, g + L.norm_g, g + L.norm_b, NF, E, F, S, c->use_reg ? 1 : 0, st));
  if (c->modular && c->mod_mode != 0) {   // backward of the ModularTransformer fusion (forward: helpers/models.py:528-570)
    const int fd = c->feat_dim;
    const float* d_femb = w.demb;         // 'features_only' and 'add': the fused gradient itself
    if (c->mod_mode == 2 && c->mod_fusion == 1) {
      CK(linear_bwd(w.cat, p + L.fu_w, w.demb, g + L.fu_w, g + L.fu_b, w.dcat, NF, E, 2 * E, 0, st));
      MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.demb_img, (size_t)E * 4, w.dcat, (size_t)2 * E * 4, (size_t)E * 4, NF, cudaMemcpyDeviceToDevice, st));
      MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.dfemb, (size_t)E * 4, w.dcat + E, (size_t)2 * E * 4, (size_t)E * 4, NF, cudaMemcpyDeviceToDevice, st));
      d_img = w.demb_img;
      d_femb = w.dfemb;
    } else if (c->mod_mode == 2 && c->mod_fusion == 2) {
      MIVIT_CUDA_CHECK(cudaMemcpy2DAsync(w.demb_img, (size_t)Ei * 4, w.demb, (size_t)E * 4, (size_t)Ei * 4, NF, cudaMemcpyDeviceToDevice, st));
      d_img = w.demb_img;
    }
    if (L.has_femb) {
      if (c->mod_fembed == 0) {
        CK(linear_bwd(w.fclean, p + L.fe0_w, d_femb, g + L.fe0_w, g + L.fe0_b, nullptr, NF, E, fd, 0, st));
      } else {
        CK(linear_bwd(w.fe_act, p + L.fe3_w, d_femb, g + L.fe3_w, g + L.fe3_b, w.dfe_a, NF, E, 2 * E, 0, st));
        CK(act_bwd(w.dfe_a, w.fe_ln, w.dfe_b, (long long)NF * 2 * E, 1, st));
        CK(layernorm_bwd(w.dfe_b, w.fe_h, w.fe_m, w.fe_r, p + L.feln_g, w.dfe_a, g + L.feln_g, g + L.feln_b, NF, 2 * E, 0, 0, 0, st));
        CK(linear_bwd(w.fclean, p + L.fe0_w, w.dfe_a, g + L.fe0_w, g + L.fe0_b, nullptr, NF, 2 * E, fd, 0, st));
      }
    }
  }
  return MIVIT_OK;
  };
  if (part != 2) CK(tokens_part());
  if (part == 1 || !L.has_img) return MIVIT_OK;
  if (c->embedding != 2) {
    if (src != nullptr)   // frames re-rendered from the trajectories inside the weight-gradient kernel (bit-identical to the forward's)
      CK(render_embed_wgrad_launch(src->traj, B, src->T, src->prm, src->seed, src->seq_offset, src->seq_off_dev, d_img, Ei,
                                   g + L.proj_w, g + L.proj_b, st));
    else
      CK(linear_bwd(x, p + L.proj_w, d_img, g + L.proj_w, g + L.proj_b, nullptr, NF, Ei, P * P, 0, st));
    return MIVIT_OK;
  }
  CK(linear_bwd(w.pooled, p + L.fc_w, d_img, g + L.fc_w, g + L.fc_b, w.dpooled, NF, Ei, 128, 0, st));
  const long long rows = w.rows, rp = w.rows_pad;
  const double cnt = (double)NF * P * P;
  const int impl = c->conv_impl;
  const ConvShifts s3 = make_shifts(P, 9, false), s3m = make_shifts(P, 9, true), s1 = make_shifts(P, 1, false);
  const int ci[2] = {32, 64}, co[2] = {64, 128};
  RowsT* in[2] = {&w.act0, &w.act2};
  RowsT* r1[2] = {&w.raw1, &w.raw3};
  RowsT* a1[2] = {&w.act1, &w.act3};
  RowsT* r2[2] = {&w.raw2, &w.raw4};
  RowsT* rs[2] = {&w.raws1, &w.raws2};
  RowsT* d2[2] = {&w.draw2, &w.draw4};      // grad of raw (conv2 output)
  RowsT* ds[2] = {&w.draws1, &w.draws2};    // grad of raw skip
  RowsT* da1[2] = {&w.dact1, &w.dact3};     // grad of act1 (conv2 input)
  RowsT* d1[2] = {&w.draw1, &w.draw3};      // grad of raw1
  RowsT* dinm[2] = {&w.dact0m, &w.dact2m};  // grad of block input through conv1
  RowsT* dins[2] = {&w.dact0s, &w.dact2s};  // grad of block input through the skip
  bool fused_in[2] = {false, false};
  for (int b = 1; b >= 0; --b) {
    const auto& R = L.rb[b];
    const int i1 = 1 + 3 * b, i2 = 2 + 3 * b, is = 3 + 3 * b;
    CK(pack_conv_weights(p + R.c1_w, w.wd_c1[b], co[b], ci[b], 9, 1, st));
    CK(pack_conv_weights(p + R.c2_w, w.wd_c2[b], co[b], co[b], 9, 1, st));
    CK(pack_conv_weights(p + R.sk_w, w.wd_sk[b], co[b], ci[b], 1, 1, st));
    // out = relu(bn2(raw2) + bns(raws))
    // (the gradient tensors share three slots, see carve(): guards are zeroed when a tensor moves in, i.e. before its producer)
    CK(zero_guards(*d2[b], rp, st));
    CK(zero_guards(*ds[b], rp, st));
    if (b == 1) {
      CK(bn_backward(nullptr, nullptr, w.dpooled, r2[b]->row0, w.bn[i2].ss, w.bn[i2].mi, p + R.bn2_g, d2[b]->row0, g + R.bn2_g,
                     g + R.bn2_b, rs[b]->row0, w.bn[is].ss, w.bn[is].mi, p + R.bns_g, ds[b]->row0, g + R.bns_g, g + R.bns_b,
                     w.bn_sums, rows, rp, P, co[b], cnt, w.fsums, 0, st));
    } else {
      CK(bn_backward(w.dact2m.row0, fused_in[1] ? nullptr : w.dact2s.row0, nullptr, r2[b]->row0, w.bn[i2].ss, w.bn[i2].mi, p + R.bn2_g, d2[b]->row0,
                     g + R.bn2_g, g + R.bn2_b, rs[b]->row0, w.bn[is].ss, w.bn[is].mi, p + R.bns_g, ds[b]->row0, g + R.bns_g,
                     g + R.bns_b, w.bn_sums, rows, rp, P, co[b], cnt, nullptr, 0, st));
    }
    // The skip path first: its gradient tensor leaves its slot before conv2's input gradient needs one (carve()).  Block 1 on
    // the product path keeps it for the ONE kernel that accumulates the conv1 (3x3) and skip (1x1) input gradients into one
    // accumulator and one output tensor (dinm); otherwise two convolutions, and BatchNorm's backward sums the two tensors.
    const bool try_dual = impl == 1 && b == 0;   // (block 2: the resident weights would force 32-column slices, measured slower)
    // Block 2 on the product path: the skip convolution's weight gradient AND input gradient from one pass over its output
    // gradient (conv_skip_bwd.cu); otherwise the two stream kernels.
    bool skip_fused = false;
    if (impl == 1 && !try_dual)
      CK(conv_skip_backward_fused(in[b]->row0, ds[b]->row0, w.wd_sk[b], g + R.sk_w, dins[b]->row0, rows, P, ci[b], co[b], st, &skip_fused));
    // Block 1 on the product path: the skip convolution's weight gradient rides along with conv1's (same input, conv_wgrad3.cu);
    // its output gradient stays in its slot until then (carve())
    const bool skip_with_c1 = impl == 1 && try_dual;
    if (!skip_fused) {
      if (!skip_with_c1) CK(conv_rows_wgrad(in[b]->row0, ds[b]->row0, g + R.sk_w, rows, P, ci[b], co[b], 1, s1, impl, st));
      if (!try_dual) CK(conv_rows_forward(ds[b]->row0, w.wd_sk[b], dins[b]->row0, nullptr, rows, P, co[b], ci[b], 1, s1, impl, st));
    }
    // conv2: block 1 (64 -> 64) on the product path computes weight gradient and input gradient in one pass over the output
    // gradient (conv_layer64_bwd.cu); block 2 (128 -> 128) keeps the clustered weight gradient and the CTA-pair dgrad with the
    // BatchNorm-backward sums in its epilogue.
    bool c2_fused = false;
    if (impl == 1 && b == 0)
      CK(conv_layer64_backward_fused(a1[b]->row0, d2[b]->row0, w.wd_c2[b], g + R.c2_w, da1[b]->row0, rows, P, co[b], co[b], s3, s3m, st,
                                     &c2_fused));
    if (!c2_fused) CK(conv_rows_wgrad(a1[b]->row0, d2[b]->row0, g + R.c2_w, rows, P, co[b], co[b], 9, s3, impl, st));
    // conv2 input gradient; on the product path its epilogue also accumulates the backward sums of bn1 (no reduction pass)
    bool bn1_summed = false;
    if (impl == 1 && !c2_fused) {
      MIVIT_CUDA_CHECK(cudaMemsetAsync(w.bn_sums, 0, 3 * co[b] * sizeof(float), st));
      CK(conv_rows_dgrad_bnsums(d2[b]->row0, w.wd_c2[b], da1[b]->row0, r1[b]->row0, w.bn[i1].ss, w.bn_sums, rows, P, co[b], co[b], s3m,
                                st, &bn1_summed));
    }
    if (!bn1_summed && !c2_fused) CK(conv_rows_forward(d2[b]->row0, w.wd_c2[b], da1[b]->row0, nullptr, rows, P, co[b], co[b], 9, s3m, impl, st));
    // act1 = relu(bn1(raw1))
    CK(zero_guards(*d1[b], rp, st));
    CK(bn_backward(da1[b]->row0, nullptr, nullptr, r1[b]->row0, w.bn[i1].ss, w.bn[i1].mi, p + R.bn1_g, d1[b]->row0, g + R.bn1_g,
                   g + R.bn1_b, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, w.bn_sums, rows, rp, P, co[b], cnt, nullptr, bn1_summed ? 1 : 0, st));
    // Block 1 on the product path: both weight gradients AND the dual-input dgrad from one pass over the two output gradients
    // (conv_block1_bwd.cu); otherwise the ride-along weight gradient + the dual-input dgrad, or the separate kernels.
    bool blk1 = false;
    if (skip_with_c1)
      CK(conv_block1_backward_fused(in[b]->row0, d1[b]->row0, ds[b]->row0, w.wd_c1[b], w.wd_sk[b], g + R.c1_w, g + R.sk_w, dinm[b]->row0,
                                    rows, P, ci[b], co[b], s3, s3m, st, &blk1));
    if (!blk1) {
      {
        bool both = false;
        if (skip_with_c1)
          CK(conv_rows_wgrad_skip_v3(in[b]->row0, d1[b]->row0, ds[b]->row0, g + R.c1_w, g + R.sk_w, rows, P, ci[b], co[b], s3, st, &both));
        if (!both) {
          if (skip_with_c1) CK(conv_rows_wgrad(in[b]->row0, ds[b]->row0, g + R.sk_w, rows, P, ci[b], co[b], 1, s1, impl, st));
          CK(conv_rows_wgrad(in[b]->row0, d1[b]->row0, g + R.c1_w, rows, P, ci[b], co[b], 9, s3, impl, st));
        }
      }
      fused_in[b] = false;
      if (try_dual) {
        bool handled = false;
        CK(conv_rows_forward_dual_v3(d1[b]->row0, w.wd_c1[b], ds[b]->row0, w.wd_sk[b], dinm[b]->row0, rows, P, co[b], co[b], ci[b], s3m,
                                     st, &handled));
        fused_in[b] = handled;
        if (!handled) CK(conv_rows_forward(ds[b]->row0, w.wd_sk[b], dins[b]->row0, nullptr, rows, P, co[b], ci[b], 1, s1, impl, st));
      }
    }
    if (blk1) fused_in[b] = true;
    if (!fused_in[b]) CK(conv_rows_forward(d1[b]->row0, w.wd_c1[b], dinm[b]->row0, nullptr, rows, P, co[b], ci[b], 9, s3m, impl, st));
  }
  // act0 = relu(bn0(raw0)); upstream = conv1 path + skip path of block 1
  CK(zero_guards(w.draw0, rp, st));
  CK(bn_backward(w.dact0m.row0, fused_in[0] ? nullptr : w.dact0s.row0, nullptr, w.raw0.row0, w.bn[0].ss, w.bn[0].mi, p + L.bn0_g, w.draw0.row0, g + L.bn0_g,
                 g + L.bn0_b, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, w.bn_sums, rows, rp, P, 32, cnt, nullptr, 0, st));
  CK(conv0_wgrad(x, w.draw0.row0, g + L.conv0_w, rows, P, st));
  return MIVIT_OK;
}

extern "C" int mivit_vit_backward(const mivit_vit_config* c, int32_t B, const float* x, const float* features,
                                  const float* dpred, const float* params, float* grads, void* workspace, void* stream) {
  return vit_backward_impl(c, B, x, features, dpred, params, grads, workspace, 0, stream);
}
extern "C" int mivit_vit_backward_part(const mivit_vit_config* c, int32_t B, const float* x, const float* features,
                                       const float* dpred, const float* params, float* grads, void* workspace, int32_t part,
                                       void* stream) {
  MIVIT_CHECK_ARG(part >= 0 && part <= 2, "part must be 0 (all), 1 (tokens) or 2 (image embedding)");
  return vit_backward_impl(c, B, x, features, dpred, params, grads, workspace, part, stream);
}
extern "C" int mivit_vit_backward_traj(const mivit_vit_config* c, int32_t B, const double* traj, int32_t T,
                                       const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset,
                                       const uint64_t* seq_offset_dev, float* frames,
                                       const float* features, const float* dpred, const float* params, float* grads,
                                       void* workspace, int32_t part, void* stream) {
  MIVIT_CHECK_ARG(part >= 0 && part <= 2, "part must be 0 (all), 1 (tokens) or 2 (image embedding)");
  const TrajSrc src{traj, T, prm, seed, seq_offset, seq_offset_dev, frames};
  return vit_backward_impl(c, B, nullptr, features, dpred, params, grads, workspace, part, stream, &src);
}
extern "C" int64_t mivit_vit_embedding_param_count(const mivit_vit_config* c) {
  ParamLayout L;
  if (build_layout(c, L)) return -1;
  // the image-embedding block is the head of the flat buffer; the feature embedding / fusion layer / norm.weight follow it
  return (int64_t)(L.has_femb ? L.fe0_w : L.fu_w >= 0 ? L.fu_w : L.norm_g);
}

extern "C" int mivit_vit_train_step(const mivit_vit_config* c, int32_t B, const float* x, const float* features,
                                    const float* target, float* params, float* grads, float* adam_m, float* adam_v,
                                    float* bn_running, int64_t* bn_num_batches, void* workspace, float* pred, float* loss,
                                    float* dpred, float lr, float beta1, float beta2, float eps, float weight_decay,
                                    int64_t step, int32_t apply_update, void* stream) {
  ParamLayout L;
  CK(build_layout(c, L));
  MIVIT_CHECK_ARG(target && pred && loss && dpred, "bad arguments");
  CK(mivit_vit_forward(c, B, x, features, params, bn_running, bn_num_batches, workspace, pred, 1, stream));
  CK(mse_loss(pred, target, mivit_vit_pred_rows(c, B), loss, dpred, (cudaStream_t)stream));
  CK(mivit_vit_backward(c, B, x, features, dpred, params, grads, workspace, stream));
  if (apply_update) {
    MIVIT_CHECK_ARG(adam_m && adam_v && step >= 1, "optimizer state missing");
    CK(adamw_flat(params, grads, adam_m, adam_v, L.off, lr, beta1, beta2, eps, weight_decay, step, 1.0f, (cudaStream_t)stream));
  }
  return MIVIT_OK;
}

extern "C" int32_t mivit_vit_pred_rows(const mivit_vit_config* c, int32_t B) {
  return c->per_frame ? B * (c->F + (c->use_reg ? 1 : 0)) : B;
}

// The reference loop body starting from the TRAJECTORIES: render (helpers/helpersGeneration.py:128-278 + normalize_images
// :356-400), model forward, MSE, backward, AdamW -- for the Linear / CNN embeddings the frames exist only inside the fused
// render->embedding kernel (forward) and its re-rendering weight-gradient twin (backward).
extern "C" int mivit_vit_train_step_traj(const mivit_vit_config* c, int32_t B, const double* traj, int32_t T,
                                         const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset,
                                         const uint64_t* seq_offset_dev, float* frames,
                                         const float* features, const float* target, float* params, float* grads, float* adam_m,
                                         float* adam_v, float* bn_running, int64_t* bn_num_batches, void* workspace, float* pred,
                                         float* loss, float* dpred, float lr, float beta1, float beta2, float eps,
                                         float weight_decay, int64_t step, int32_t apply_update, void* stream) {
  ParamLayout L;
  CK(build_layout(c, L));
  MIVIT_CHECK_ARG(target && pred && loss && dpred, "bad arguments");
  CK(mivit_vit_forward_traj(c, B, traj, T, prm, seed, seq_offset, seq_offset_dev, frames, features, params, bn_running,
                            bn_num_batches, workspace, pred, 1, stream));
  CK(mse_loss(pred, target, mivit_vit_pred_rows(c, B), loss, dpred, (cudaStream_t)stream));
  CK(mivit_vit_backward_traj(c, B, traj, T, prm, seed, seq_offset, seq_offset_dev, frames, features, dpred, params, grads, workspace,
                             0, stream));
  if (apply_update) {
    MIVIT_CHECK_ARG(adam_m && adam_v && step >= 1, "optimizer state missing");
    CK(adamw_flat(params, grads, adam_m, adam_v, L.off, lr, beta1, beta2, eps, weight_decay, step, 1.0f, (cudaStream_t)stream));
  }
  return MIVIT_OK;
}
