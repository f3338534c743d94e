// Internal (non-ABI) declarations shared by the ViT kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct ConvShifts {
  int d[9];
};

// Row geometry of the pitched-rows activation layout (see conv_tc.cu).
static inline long long rows_per_frame(int P) { return (long long)(P + 1) * (P + 1); }
static inline ConvShifts make_shifts(int P, int taps, bool mirrored) {
  ConvShifts s;
  for (int i = 0; i < 9; ++i) s.d[i] = 0;
  if (taps == 9)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int dlt = (kh - 1) * (P + 1) + (kw - 1);
        s.d[kh * 3 + kw] = mirrored ? -dlt : dlt;
      }
  return s;
}

// Y[rows,cout] = sum_tap X[rows+shift,cin] * Wp[tap]; optional per-channel (sum, sumsq) atomics.
// impl: 1 = tcgen05 tensor-core kernel (product path), 0 = SIMT cross-check.
int conv_rows_forward(const __nv_bfloat16* X, const __nv_bfloat16* Wp, __nv_bfloat16* Y, float* stats, long long rows,
                      int P, int cin, int cout, int taps, const ConvShifts& sh, int impl, cudaStream_t st);

// pipelined / fused variants (conv_tc3.cu: TMA ring; conv_tc4.cu: CTA pairs); impl: 1 = pipelined tcgen05, 2 = serial tcgen05, 0 = SIMT
int conv_rows_forward_v3(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                         __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout, int taps,
                         const ConvShifts& sh, cudaStream_t st, bool* handled);
int conv_rows_forward_v4(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                         __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout, int taps,
                         const ConvShifts& sh, cudaStream_t st, bool* handled);
int conv_rows_dgrad_bnsums(const __nv_bfloat16* X, const __nv_bfloat16* Wp, __nv_bfloat16* Y, const __nv_bfloat16* raw,
                           const float* ss, float* sums, long long rows, int P, int cin, int cout, const ConvShifts& sh,
                           cudaStream_t st, bool* handled);
int conv_rows_forward_dual_v3(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* X2, const __nv_bfloat16* Wp2,
                              __nv_bfloat16* Y, long long rows, int P, int cin, int cin2, int cout, const ConvShifts& sh,
                              cudaStream_t st, bool* handled);
int conv_rows_forward_fused(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                            __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout,
                            int taps, const ConvShifts& sh, int impl, cudaStream_t st);

// W fp32 [cout][cin][k][k] -> bf16 core-matrix packs (see conv_aux.cu)
int pack_conv_weights(const float* W, __nv_bfloat16* out, int cout, int cin, int taps, int dgrad, cudaStream_t st);

// dW[cout][cin][k][k] (fp32, pre-zeroed) += sum_r dY[r,cout]^T X[r+shift,cin]
int conv_rows_wgrad(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int cin, int cout,
                    int taps, const ConvShifts& sh, int impl, cudaStream_t st);

int conv_layer64_backward_fused(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* Wd, float* dW, __nv_bfloat16* dX,
                                long long rows, int P, int cin, int cout, const ConvShifts& sh_w, const ConvShifts& sh_d,
                                cudaStream_t st, bool* handled);
int conv_block1_backward_fused(const __nv_bfloat16* X, const __nv_bfloat16* dY1, const __nv_bfloat16* dY2, const __nv_bfloat16* Wd1,
                               const __nv_bfloat16* Wd2, float* dW1, float* dW2, __nv_bfloat16* dX, long long rows, int P, int cin,
                               int cout, const ConvShifts& sh_w, const ConvShifts& sh_d, cudaStream_t st, bool* handled);
int conv_rows_wgrad_skip_v3(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* dYskip, float* dW, float* dWskip,
                            long long rows, int P, int cin, int cout, const ConvShifts& sh, cudaStream_t st, bool* handled);
int conv_skip_backward_fused(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* Wp, float* dW, __nv_bfloat16* dX,
                             long long rows, int P, int cin, int cout, cudaStream_t st, bool* handled);
int conv_rows_wgrad_v3(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int cin, int cout,
                       int taps, const ConvShifts& sh, cudaStream_t st, bool* handled);

// gemm_simt.cu
int gemm_f32(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
             long long scm, int M, int N, int K, const float* bias, int relu, int accumulate, int split_k, cudaStream_t st);
int colsum_f32(const float* X, long long ld, int M, int N, float* out, cudaStream_t st);

// transformer.cu
int layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* z_out, float* y,
                  float* mean, float* rstd, int rows, int E, float eps, int group, int out_group, int out_off, cudaStream_t st);
int layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma, float* dz,
                  float* dgamma, float* dbeta, int rows, int E, int group, int out_group, int out_off, cudaStream_t st);
int attention_fwd(const float* q, const float* k, const float* v, float* ctx, float* probs, int B, int S, int E, int H,
                  cudaStream_t st);
int attention_bwd(const float* q, const float* k, const float* v, const float* probs, const float* ctx, const float* dctx,
                  float* dq, float* dk, float* dv, int B, int S, int E, int H, cudaStream_t st);
size_t attention_probs_floats(int B, int S, int E, int H);   // size of the per-layer `probs` buffer (log-sum-exp rows, or S x S)
int act_fwd(const float* pre, float* post, long long n, int mode, cudaStream_t st);
int nan_to_num_f32(const float* in, float* out, long long n, cudaStream_t st);
int add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st);
int act_bwd(const float* dpost, const float* pre, float* dpre, long long n, int mode, cudaStream_t st);
int tokens_finish(float* tok, const float* reg, const float* proj, const float* pos, int B, int S, int E, cudaStream_t st);
int tokens_finish_bwd(const float* dtok, float* dreg, float* dproj, float* dpos, int B, int S, int E, cudaStream_t st);
int pool_tokens(const float* x, float* out, int B, int S, int E, int ld, int use_reg, cudaStream_t st);
int pool_tokens_bwd(const float* dout, float* dx, int B, int S, int E, int ld, int use_reg, cudaStream_t st);

// encoder_fused.cu -- one encoder layer (q/k/v, attention, out-proj + LN, FFN + LN) as one persistent tcgen05 kernel
struct EncoderLayerIO {
  const float* x;   // [B*S, E] layer input
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *g1, *be1, *w1, *bf1, *w2, *bf2, *g2, *be2;
  // everything the backward reads: [B*S, E] tensors, hact [B*S, HD], lse [B, H, S], LayerNorm row statistics [B*S]
  float *q, *k, *v, *ctx, *lse, *z1, *m1, *r1, *x1, *hact, *z2, *m2, *r2, *x2;
};
bool encoder_fused_supported(int B, int S, int E, int HD, int H);
int encoder_layer_fwd(const EncoderLayerIO& io, int B, int S, int E, int HD, int H, float ln_eps, cudaStream_t st);

// encoder_fused_bwd.cu -- the input-gradient half of a layer's backward as one persistent kernel (the weight gradients reduce
// over all tokens and stay batched launches of linear_tc.cu)
constexpr int kEncMaxLayers = 16;
struct EncoderLayerBwdIO {
  const float* dy;   // [B*S, E] gradient of the layer output
  float* dx;         // [B*S, E] gradient of the layer input (may alias dy)
  const float *z2, *m2, *r2, *hact, *z1, *m1, *r1, *q, *k, *v, *lse, *ctx;   // written by the forward
  const float *g2, *g1;          // LayerNorm weights (norm2, norm1)
  const float* packed;           // this layer's block of encoder_pack_bwd_weights
  float *dz2, *dh2, *dz1, *dq, *dk, *dv;   // [B*S, E] ([B*S, HD] for dh2): the dY operands of the weight-gradient launches
  float *dg2, *db2, *dg1, *db1;  // LayerNorm parameter gradients, accumulated
};
size_t encoder_bwd_packed_floats(int E, int HD);   // floats per layer of the packed (transposed) weights
int encoder_pack_bwd_weights(const float* params, const long long (*offs)[6] /* q, k, v, o, f1, f2 weight offsets per layer */, int L,
                             int E, int HD, float* out, cudaStream_t st);
int encoder_layer_bwd(const EncoderLayerBwdIO& io, int B, int S, int E, int HD, int H, cudaStream_t st);

// synchronised BatchNorm hook (vit_model.cu): SUM all-reduce of a small device buffer over the data-parallel group
int mivit_bn_sync_world();                                      // 1 when no hook is registered
int mivit_bn_sync(float* buf, long long n, cudaStream_t st);

// peer_comm.cu: synchronised BatchNorm through the peer segments (plain kernels instead of the host hook)
bool mivit_bn_peer_active();
int mivit_bn_peer_world();
void mivit_bn_peer_begin_step();
int mivit_bn_peer_sync(float* buf, long long n, cudaStream_t st);

// bn.cu
int bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                long long* num_batches, float* mean, float* invstd, float* scale, float* shift, int C, double count, float eps,
                float momentum, int training, cudaStream_t st);
int bn_apply(const __nv_bfloat16* raw_a, const float* ss_a, const __nv_bfloat16* raw_b, const float* ss_b,
             __nv_bfloat16* act, long long rows, long long rows_pad, int P, int C, cudaStream_t st);
int bn_apply_pool(const __nv_bfloat16* raw_a, const float* ss_a, const __nv_bfloat16* raw_b, const float* ss_b, float* pooled,
                  float* fsums /* [frames][3][C] masked per-frame sums for bn_backward, or NULL */, long long n_frames, int P, int C,
                  cudaStream_t st);
int pool_rows(const __nv_bfloat16* act, float* pooled, long long n_frames, int P, int C, cudaStream_t st);
int bn_backward(const __nv_bfloat16* up_a, const __nv_bfloat16* up_b, const float* dpooled, const __nv_bfloat16* raw_a,
                const float* ss_a, const float* mi_a, const float* gamma_a, __nv_bfloat16* draw_a, float* dgamma_a, float* dbeta_a,
                const __nv_bfloat16* raw_b, const float* ss_b, const float* mi_b, const float* gamma_b, __nv_bfloat16* draw_b,
                float* dgamma_b, float* dbeta_b, float* sums, long long rows, long long rows_pad, int P, int C, double count,
                const float* fsums /* with dpooled: per-frame sums of bn_apply_pool replace the reduction pass */,
                int presummed /* sums already holds (sum g, sum g*raw_a) from conv_rows_dgrad_bnsums: no reduction pass */,
                cudaStream_t st);
// bn_frames.cu: the pooled-gradient BatchNorm pair one frame at a time (TMA-fed); same outputs as bn_backward's streaming pass
bool bn_frames_supported(int P, int C);
int bn_backward_pooled_frames(const float* dpooled, const __nv_bfloat16* raw_a, const float* ss_a, const float* coef_a,
                              __nv_bfloat16* draw_a, const __nv_bfloat16* raw_b, const float* ss_b, const float* coef_b,
                              __nv_bfloat16* draw_b, long long rows, long long rows_pad, int P, int C, cudaStream_t st);
bool bn_reduce_frames_supported(int P, int C, int n_streams);
int bn_backward_reduce_frames(const __nv_bfloat16* up_a, const __nv_bfloat16* up_b, const __nv_bfloat16* raw_a, const float* ss_a,
                              const float* mi_a, const __nv_bfloat16* raw_b, const float* ss_b, const float* mi_b, float* sums,
                              long long rows, int P, int C, cudaStream_t st);
int bn_apply_pool_frames(const __nv_bfloat16* raw_a, const float* ss_a, const __nv_bfloat16* raw_b, const float* ss_b, float* pooled,
                         float* fsums, long long n_frames, int P, int C, cudaStream_t st);
int conv0_forward(const float* frames, const float* W0, __nv_bfloat16* raw0, float* stats, long long rows, long long rows_pad,
                  int P, cudaStream_t st);
int conv0_wgrad(const float* frames, const __nv_bfloat16* draw0, float* dW0, long long rows, int P, cudaStream_t st);

// render.cu -- renderer fused with the Linear / CNN frame embedding and its re-rendering weight gradient
struct mivit_render_params;
// seq_off_dev: optional DEVICE counter added to seq_offset (a captured CUDA graph advances the global sequence ids through it)
int render_v1_launch(const double* traj, long long N, int T, const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset,
                     const uint64_t* seq_off_dev, float* out, long long out_seq_stride, cudaStream_t stream);
int render_embed_linear_launch(const double* traj, long long N, int T, const mivit_render_params* prm, uint64_t seed,
                               uint64_t seq_offset, const uint64_t* seq_off_dev, const float* W, int w_transposed, const float* bias,
                               int E, float* emb, float* frames_out, long long frames_seq_stride, cudaStream_t st);
int render_embed_wgrad_launch(const double* traj, long long N, int T, const mivit_render_params* prm, uint64_t seed,
                              uint64_t seq_offset, const uint64_t* seq_off_dev, const float* demb, int E, float* dW, float* db,
                              cudaStream_t st);

// optim.cu
int mse_loss(const float* pred, const float* target, int n, float* loss, float* dpred, cudaStream_t st);
int adamw_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
               long long step, float grad_scale, cudaStream_t st);

// linear_tc.cu -- tcgen05 kind::tf32 nn.Linear kernels
bool linear_tc_supported(int M, int K, int N);
int linear_tc(const float* A, const float* W, const float* bias, float* Y, int M, int K, int N, int mode, int relu, int accumulate,
              cudaStream_t st);
int linear_tc_batched(int nb, const float* const* A, const float* const* W, const float* const* bias, float* const* Y, int M, int K,
                      int N, int mode, int relu, int accumulate, cudaStream_t st,
                      const float* gate = nullptr);   // nb <= 4 problems of one shape, one launch; gate: Y = gate > 0 ? Y : 0; accumulate 2: atomic add (problems may share Y)
int linear_wgrad_tc_batched(int nb, const float* const* dY, const float* const* X, float* const* dW, float* const* db, int M, int N,
                            int K, cudaStream_t st);
bool linear_wgrad_tc_supported(int M, int N, int K);
// nb <= 6 weight gradients of different shapes over the same M rows in one launch (CTAs dealt to the problems by bytes per row)
bool linear_wgrad_tc_multi_supported(int nb, int M, const int* N, const int* K);
int linear_wgrad_tc_multi(int nb, const float* const* dY, const float* const* X, float* const* dW, float* const* db, int M, const int* N,
                          const int* K, cudaStream_t st);
int linear_wgrad_tc(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, cudaStream_t st);
