/* mivit_b200 -- C ABI of the B200-native MiViT hot path (render + ViT train step).
 *
 * The reference (Biomedical-Imaging-Group/MolecularDiffusion_MiViT) is pure Python and has
 * no FFI; the boundary it exposes for this path is the Python call surface listed in
 * SURVEY.md section 8b.  Each entry point below names the reference function whose work it
 * replaces (paths relative to the reference root).  The Python mirrors in
 * moleculardiffusion_mivit_b200/{helpersGeneration,experiments,models}.py bind these symbols
 * with ctypes.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  All data pointers are DEVICE pointers
 *    unless the name ends in _host.  Nothing is allocated inside; the caller owns buffers.
 *  - `stream` is a cudaStream_t passed as void*.  Calls enqueue work and return without
 *    synchronising.
 *  - return value: 0 = ok, non-zero = error; mivit_last_error() gives the thread-local text.
 *  - sm_100a only; there is no CPU fallback.
 */
#ifndef MIVIT_H_
#define MIVIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIVIT_ABI_VERSION 3

int mivit_abi_version(void);
const char* mivit_last_error(void);
/* number of kernels launched by this library since the last reset (bench.py: gpu_launches) */
int64_t mivit_launch_count(void);
void mivit_reset_launch_count(void);
void mivit_add_launch_count(int64_t n); /* kernels launched as nodes of a replayed CUDA graph captured from this library */

/* Optional device timing of the tagged hot kernels (CUDA events on the launching stream), used by
 * bench.py for the roofline object.  total_work = algorithmic FLOPs (convolutions) or bytes
 * (renderer) summed over the launches. */
typedef struct mivit_kernel_time {
  char name[48];
  int64_t launches;
  double total_ms;
  double total_work;
} mivit_kernel_time;
void mivit_profile_enable(int32_t on);
int32_t mivit_profile_read(mivit_kernel_time* out, int32_t max_entries); /* after a stream sync */

/* ------------------------------------------------------------------ renderer ---------- */

/* Scalar set-up that helpers/helpersGeneration.py:225-247 derives from `image_props`. */
typedef struct mivit_render_params {
  double scale;      /* trajectory -> pixel factor: unit/(resolution*1e9)  (:231), 1 if unit == -1 */
  double sigma_hr;   /* PSF sigma in high-res pixels: U/resolution*fwhm/2.355 (:242)               */
  int32_t P;         /* output_size                                                                */
  int32_t U;         /* upsampling_factor                                                          */
  int32_t n;         /* nPosPerFrame                                                               */
  int32_t center;    /* subtract the frame's mean sub-position (:291)                              */
  int32_t flip_y;    /* render with -y (the effect of the in-place flip :197)                      */
  int32_t draw;      /* particle_mean > 1e-4 && particle_std > 1e-4 (:299); 0 -> blank particle    */
  float part_mean;   /* particle_intensity[0]                                                      */
  float part_std;    /* particle_intensity[1]                                                      */
  float bg_mean;     /* background_intensity[0]                                                    */
  float bg_std;      /* background_intensity[1]                                                    */
  float poisson;     /* poisson_noise; -1 disables (:316)                                          */
  int32_t normalize; /* fuse normalize_images (:389-395): (v - norm_sub) / norm_div                */
  float norm_sub;    /* background_mean - background_sigma                                         */
  float norm_div;    /* theoretical_max - (background_mean - background_sigma)                     */
  int32_t mean_noise;/* test hook: every normal draw returns its mean, Poisson(lam) -> lam         */
} mivit_render_params;

/* Replaces helpers/helpersGeneration.py:128-278 trajectories_to_video + :283-319
 * trajectory_to_video (+ optional :356-400 normalize_images).
 * traj: [N,T,2] float64, read only (the caller applies the reference's in-place y flip).
 * out : float32; frame f of sequence s is written at out + s*out_seq_stride + f*P*P
 *       (out_seq_stride = F*P*P for the plain (N,F,P,P) result).
 * seq_offset: global id of sequence 0 (noise streams are keyed by global id). */
int mivit_render_v1(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm,
                    uint64_t seed, uint64_t seq_offset, float* out, int64_t out_seq_stride,
                    void* stream);

/* Renderer fused with the frame embedding of LinearProjectionEmbedding / CNNEmbedding
 * (reference helpers/models.py:146-167 and :170-199: both are emb[f,:] = W[E,P*P] . frame[f] + b applied to the output of
 * helpers/helpersGeneration.py:128-278 + normalize_images :356-400).  Same trajectory / parameter / RNG contract as
 * mivit_render_v1; the frame is never written to HBM unless frames_out != NULL (needed only for the weight gradient).
 * Wt: device fp32 [P*P][E] = the nn.Linear weight transposed (Conv2d weight [E,1,P,P] is the same matrix);
 * bias: device fp32 [E]; emb: device fp32 [N][F][E]; frames_out (optional): [N][..frames_seq_stride..] like render_v1. */
int mivit_render_embed_linear(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm, uint64_t seed,
                              uint64_t seq_offset, const float* Wt, const float* bias, int32_t E, float* emb,
                              float* frames_out, int64_t frames_seq_stride, void* stream);

/* Weight gradient of that fused layer with the frames RE-RENDERED from the trajectories (same parameters / seed / seq_offset
 * as the forward call, hence bit-identical frames): dW[E][P*P] += sum_frames demb[frame][:] (x) frame, db[E] += sum_frames
 * demb[frame][:] (the backward of helpers/models.py:160-164 / :188-197 w.r.t. proj.weight / conv.weight and the bias).
 * demb: device fp32 [N][F][E]; dW, db: device fp32, ACCUMULATED into (zero them first); db may be NULL.  P <= 16, E <= 256. */
int mivit_render_embed_linear_wgrad(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm, uint64_t seed,
                                    uint64_t seq_offset, const float* demb, int32_t E, float* dW, float* db, void* stream);

/* Host-only helper (no GPU needed): the 256-entry alias table the V1 renderer draws its multiplicative Poisson(pn) factor from
 * (helpers/helpersGeneration.py:316-317; layout in csrc/philox.cuh): entries_host[j] = alias << 24 | threshold(24 bit),
 * k = *k0_host + (accepted ? j : alias).  Returns non-zero when lam is outside (0, 380] (the renderer then uses PTRS). */
int mivit_poisson_alias_table(double lam, uint32_t* entries_host, int32_t* k0_host);

/* Replaces helpers/helpersGeneration.py:422-540 trajectories_to_video_multiple_settings / trajectory_to_mult_settings (the
 * Denoising experiments' generator): one intensity per frame, four float32 [N,F,P,P] outputs -- noise free, + clipped Gaussian
 * background, + Poisson(x*pn)/pn, + Gaussian filter (sigma 0.5, 'nearest' borders) of the Poisson frame.  prm as for
 * mivit_render_v1 (scale = trajectory_unit*1e-9/resolution, :464; flip_y honoured; normalize ignored). */
int mivit_render_multi(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm, uint64_t seed,
                       uint64_t seq_offset, float* out_no_noise, float* out_gauss, float* out_poisson,
                       float* out_filter, void* stream);

/* Replaces helpers/helpersGeneration.py:557-587 richardson_lucy_tv / richardson_lucy_tv_iter_list (with tv_gradient :542-555)
 * for a batch of P x P frames (P <= 16): images [n_images][P][P] float32 (device), psf_dev [K][K] float64 (device, K odd),
 * iterations_host[n_iterations] = the reference's `iterations_list` (0-based indices of the iterations whose estimate is kept,
 * ascending; iterations_list[-1] + 1 iterations run); out [n_images][n_iterations][P][P] float32. */
int mivit_rl_tv(const float* images, int64_t n_images, int32_t P, const double* psf_dev, int32_t K,
                const int32_t* iterations_host, int32_t n_iterations, float tv_weight, float* out, void* stream);

/* Replaces Experiments/PSFNoise/trainSettingsPSFNoise.py:196-309 trajs_to_vid_psf_noise.
 * psf_div[n_psf], noise_frac[n_noise] are HOST arrays (PSF_Settings, Noise_Settings);
 * part_mean_global is the module-level `part_mean` used for the background sigma (:302).
 * out: [N, n_psf, n_noise, F, P, P] float32.  prm->flip_y and prm->normalize are ignored. */
int mivit_render_psfnoise(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm,
                          const float* psf_div_host, int32_t n_psf, const float* noise_frac_host,
                          int32_t n_noise, float part_mean_global, uint64_t seed,
                          uint64_t seq_offset, float* out, void* stream);

/* Trajectory source.  Replaces the call sites of andi_datasets models_phenom().single_state
 * (Experiments/PSFNoise/trainModelsPSFNoise.py:128-132) with the in-repo equivalent
 * helpers/helpersGeneration.py:9-45 brownian_motion: D ~ N(mean, sqrt(var)) redrawn until
 * positive, steps ~ N(0, 2 D) per axis, positions = cumulative sum starting at 0, divided
 * by `div` (traj_div_factor).  Sequence with global id g uses group g % n_groups.
 * traj: [N,T,2] float64 out;  D_out: [N] float32 out (the label before /D_max). */
int mivit_brownian(int64_t N, int32_t T, const float* group_mean_host, const float* group_var_host,
                   int32_t n_groups, double div, uint64_t seed, uint64_t seq_offset, double* traj,
                   float* D_out, void* stream);

/* ------------------------------------------------- ViT building blocks (tests / ViT) ---- */

/* Frame-averaged positions (reference helpers/helpersGeneration.py:48-74 average_trajectories_frames):
 * out[N][T/n][2] = mean over groups of n consecutive sub-positions of traj[N][T][2] (device float64). */
int mivit_average_frames(const double* traj, int64_t N, int32_t T, int32_t n, double* out, void* stream);

/* The 25 trajectory features the ViT takes as `features` (reference helpers/helpersFeatures.py:448-520
 * compute_diffusion_features, order of feature_names :7-34; msd :102-132, power-law fit :135-191, efficiency :194-218,
 * fractal_dim :221-247, gaussianity :250-284, kurtosis :287-324, msd_ratio :327-347, trappedness :350-378,
 * convex_hull_area :381-402).  traj: device float64 [N][L][2] (already frame-averaged), out: device float64 [N][25],
 * raw values like the reference's function (NaN / -inf where it returns them; L < 3 gives 25 NaNs).  3 <= L <= 512. */
int mivit_diffusion_features(const double* traj, int64_t N, int32_t L, double dt, double* out, void* stream);

/* Activation layout of the DeepResNetEmbedding kernels ("pitched rows", bf16, channels last):
 * frame f, pixel (y,x) -> row f*(P+1)^2 + y*(P+1) + x of a [rows, C] matrix; column P of every
 * line and line P of every frame are zero; >= 128 zero guard rows precede row 0 and follow the
 * last 128-row tile.  `*_row0` pointers address row 0. */

/* nn.Conv2d weight fp32 [cout][cin][k][k] (reference helpers/models.py:206,209,216,233) ->
 * bf16 operand pack.  dgrad = 0: [k*k][cin/8][cout][8];  dgrad = 1: [k*k][cout/8][cin][8]. */
int mivit_conv_pack_weights(const float* W, void* out_bf16, int32_t cout, int32_t cin, int32_t ksize,
                            int32_t dgrad, void* stream);

/* Stride-1 / pad-(k/2) convolution on pitched rows: Y[r,:] = sum_tap X[r + delta_tap,:] * W_tap^T
 * (F.conv2d of helpers/models.py:222,225,247 without bias).  mirrored = 1 negates the tap
 * shifts (input gradient with a dgrad weight pack; cin/cout are then those of the GEMM, i.e.
 * swapped).  stats (optional): [2][cout] fp32, per-channel sum and sum of squares of the
 * stored outputs are ADDED (BatchNorm2d batch statistics).  impl: 1 = pipelined tcgen05 kernel
 * (product path), 2 = serial tcgen05 kernel, 0 = SIMT cross-check kernel. */
int mivit_conv_rows(const void* X_row0, const void* Wp, void* Y_row0, float* stats, int64_t rows,
                    int32_t P, int32_t cin, int32_t cout, int32_t ksize, int32_t mirrored, int32_t impl,
                    void* stream);

/* 3x3 convolution + the 1x1 skip convolution of the SAME input in one launch
 * (ResidualBlock.conv1 and ResidualBlock.skip[0], helpers/models.py:206,216,221-222). */
int mivit_conv_rows_fused(const void* X_row0, const void* Wp, const void* Wskip, void* Y_row0, void* Yskip_row0,
                          float* stats, float* stats_skip, int64_t rows, int32_t P, int32_t cin, int32_t cout,
                          int32_t impl, void* stream);

/* dW (fp32 [cout][cin][k][k], pre-zeroed by the caller) += sum_r dY[r,:]^T X[r + delta_tap,:]
 * (weight gradient of the same convolution).  impl as above. */
int mivit_conv_rows_wgrad(const void* X_row0, const void* dY_row0, float* dW, int64_t rows, int32_t P,
                          int32_t cin, int32_t cout, int32_t ksize, int32_t impl, void* stream);

/* nn.Linear (helpers/models.py:20-23,64-65,241,268-273) on tcgen05 kind::tf32, fp32 in / fp32 out.
 *   mode 0: out[M,out_f] = A[M,in_f] W[out_f,in_f]^T + bias (relu)
 *   mode 1: out[M,in_f] (+)= A[M,out_f] W[out_f,in_f]                (input gradient)
 *   mode 2: out[out_f,in_f] += A[M,out_f]^T W'[M,in_f]               (weight gradient; W' = layer input;
 *           when `bias` is not NULL it is the bias GRADIENT: bias[out_f] += column sums of A)
 * Fails for shapes outside the tensor-core kernels' range (M >= 512, features multiples of 32, <= 256). */
int mivit_linear_tf32(int32_t mode, const float* A, const float* W, const float* bias, float* out, int32_t M,
                      int32_t in_features, int32_t out_features, int32_t relu, int32_t accumulate, void* stream);

/* ---------------------------------------------------------------- ViT -------------------- */

/* Replaces helpers/models.py:278-361 GeneralTransformer (+ its embedding classes :146-257,
 * Transformer :111-141, MLPHead :260-276).  Parameters are ONE flat fp32 buffer in the canonical
 * order below; gradients / AdamW moments share the offsets.
 *   deepresnet: embedding.initial_conv.weight, embedding.bn1.{weight,bias},
 *               embedding.res_block{1,2}.{conv1.weight, bn1.weight, bn1.bias, conv2.weight, bn2.weight,
 *               bn2.bias, skip.0.weight, skip.1.weight, skip.1.bias}, embedding.fc.{weight,bias}
 *   linear/cnn: embedding.proj.{weight,bias}  /  embedding.conv.{weight,bias}
 *   norm.{weight,bias}, [reg_token], [transformer.pos_embedding],
 *   transformer.encoder_layers.i.{self_attn.{q,k,v,out}_proj.{weight,bias}, norm1.{weight,bias},
 *               feed_forward.fc1.{weight,bias}, feed_forward.fc2.{weight,bias}, norm2.{weight,bias}},
 *   transformer.norm.{weight,bias}, [feature_projector.{0,2}.{weight,bias}],
 *   mlp_head.mlp.{0,3}.{weight,bias}
 * BatchNorm running statistics: flat fp32 [mean C | var C] for the 7 BatchNorm2d layers in the
 * order bn1, res_block1.{bn1,bn2,skip.1}, res_block2.{bn1,bn2,skip.1} (1216 floats) and int64[7]
 * num_batches_tracked. */
typedef struct mivit_vit_config {
  int32_t embedding;   /* 0 LinearProjectionEmbedding, 1 CNNEmbedding, 2 DeepResNetEmbedding       */
  int32_t P, F;        /* patch (= image) size, frames per sequence                                 */
  int32_t E, H, HD, L; /* embed_dim, num_heads, hidden_dim, num_layers                              */
  int32_t activation;  /* tr_activation_fct: 0 F.relu, 1 F.gelu, 2 F.leaky_relu                     */
  int32_t use_pos, use_reg, use_feat;
  int32_t fusion;      /* 0 'early', 1 'late'                                                       */
  int32_t feat_dim;    /* global_feature_dim                                                        */
  int32_t head_hidden; /* MLPHead hidden_dim (128)                                                  */
  int32_t conv_impl;   /* 1 = tcgen05 convolutions (product path), 0 = SIMT cross-check             */
  float bn_eps, bn_momentum, ln_eps;
  /* ModularTransformer (helpers/models.py:366-593); all four 0 for GeneralTransformer.  With modular != 0 the `features`
   * argument of forward / backward / train_step is PER FRAME, [B,F,feat_dim] (feat_dim = features_dim), use_feat must be 0,
   * and the flat parameter order is: image_embedding.* (as `embedding.*` above, with embed_dim - features_dim outputs for
   * 'concat_features'), feature_embedding.{weight,bias} ('linear') or feature_embedding.{0.weight,0.bias,1.weight,1.bias,
   * 3.weight,3.bias} ('mlp': Linear, LayerNorm, GELU, Linear), fusion_layer.{weight,bias} ('concat_proj'), then norm.* ... as above. */
  int32_t modular;     /* 1 = ModularTransformer.forward                                             */
  int32_t mod_mode;    /* mode: 0 'images_only', 1 'features_only' (x may be NULL), 2 'both'         */
  int32_t mod_fembed;  /* feature_embedding_type: 0 'linear', 1 'mlp'                                */
  int32_t mod_fusion;  /* fusion_method: 0 'add', 1 'concat_proj', 2 'concat_features'               */
  int32_t per_frame;   /* ModularTransformer(use_regression_token=False, single_prediction=False): the MLP head runs on every
                        * token and pred / dpred / target are [B*F,1] (helpers/models.py:585-593); needs use_reg = use_feat = 0 */
} mivit_vit_config;

/* Synchronised BatchNorm for data-parallel training (torch.nn.SyncBatchNorm semantics): the host registers
 * a SUM all-reduce over its process group (training.py wraps torch.distributed.all_reduce).  While a hook with
 * world_size > 1 is registered, every TRAINING forward / backward all-reduces the per-channel BatchNorm sums
 * (<= 384 floats per call, 12 calls per step) through it and uses world_size x the local element count, so the
 * normalisation, the running statistics and the gradients equal the single-process reference at the same global batch
 * (per-rank batches must be equal).  fn is called on the calling thread with a DEVICE buffer; it must enqueue the
 * reduction in order with `stream` and return 0.  fn == NULL unregisters. */
typedef int (*mivit_allreduce_fn)(void* device_buf, int64_t n_floats, void* stream, void* user);
void mivit_set_allreduce_hook(mivit_allreduce_fn fn, void* user, int32_t world_size);

/* -------------------------------------------------- data-parallel exchange over NVLink peer memory (csrc/peer_comm.cu) ----
 * The reference is single-process; these entry points implement SURVEY.md section 8e without NCCL on the data path: every rank
 * owns a SEGMENT -- header (flags, counters, AdamW step, lr) | small-exchange slots | the model's flat gradient buffer --
 * allocated by mivit_comm_alloc, exported as a 64-byte CUDA IPC handle, mapped by its peers with mivit_comm_ipc_open (the host
 * exchanges the handles, e.g. with torch.distributed.all_gather_object).  All ranks must enqueue the same exchanges in the same
 * order.  Up to 8 ranks (one NVSwitch domain). */
typedef struct mivit_peer_comm {
  int32_t rank, world;
  void* segment[8];     /* segment base pointers AS MAPPED IN THIS PROCESS; segment[rank] is the rank's own */
} mivit_peer_comm;
int64_t mivit_comm_segment_bytes(int64_t n_grad_floats);     /* segment size for a flat gradient of n floats */
int64_t mivit_comm_grad_offset_bytes(void);                  /* the gradient buffer starts this far into the segment */
int mivit_comm_alloc(int64_t bytes, void** segment);         /* cudaMalloc + zero (the one allocation this library makes) */
int mivit_comm_free(void* segment);
int mivit_comm_ipc_handle(void* segment, uint8_t* handle64);
int mivit_comm_ipc_open(const uint8_t* handle64, void** segment);
int mivit_comm_ipc_close(void* segment);
/* AdamW's learning rate and 1-based step count live in the segment (a replayed CUDA graph reads them there) */
int mivit_comm_set_lr(const mivit_peer_comm* comm, float lr, void* stream);
int mivit_comm_set_step(const mivit_peer_comm* comm, int64_t steps_done, void* stream);
/* Gradient all-reduce FUSED with torch.optim.AdamW (Experiments/PSFNoise/trainSettingsPSFNoise.py:119) on flat offsets
 * [lo, hi) (lo % 4 == 0): g = sum over ranks (rank order) of the peers' gradient replicas, read straight from peer memory;
 * p, m, v updated with grad_scale = 1 / world, lr and step from the segment.  advance_step != 0 on the LAST bucket of a step.
 * grad_sum (optional, local, same offsets): receives the reduced gradient.  bucket in [0,4): exchanges that may be in flight
 * concurrently need different buckets.  max_ctas: 0 = one CTA per SM; an exchange that overlaps other kernels should ask for
 * a few dozen (it is latency bound).  One kernel, no host synchronisation, capturable into a CUDA graph. */
int mivit_allreduce_adamw(const mivit_peer_comm* comm, int32_t bucket, int64_t lo, int64_t hi, float* p, float* m, float* v,
                          float beta1, float beta2, float eps, float weight_decay, int32_t advance_step, float* grad_sum,
                          int32_t max_ctas, void* stream);
/* In-place SUM all-reduce of n <= 512 floats at `buf` (local); `call` in [0,32) names the exchange within a step. */
int mivit_allreduce_small(const mivit_peer_comm* comm, int32_t call, float* buf, int32_t n, void* stream);
/* While set (non-NULL, world > 1) the synchronised-BatchNorm reductions of mivit_vit_forward / backward go through the peer
 * segments (mivit_allreduce_small) instead of the host hook of mivit_set_allreduce_hook: plain kernels, capturable. */
int mivit_set_bn_sync_comm(const mivit_peer_comm* comm);

int32_t mivit_vit_param_count(const mivit_vit_config* cfg);                 /* -1 on a bad config */
int mivit_vit_param_sizes(const mivit_vit_config* cfg, int64_t* sizes, int32_t max_count);
int64_t mivit_vit_workspace_bytes(const mivit_vit_config* cfg, int32_t B);  /* -1 on a bad config */

/* GeneralTransformer.forward (:328-361).  x: [B,F,P,P] fp32; features: [B,feat_dim] or NULL;
 * pred: [B,1].  training != 0: BatchNorm uses batch statistics and updates bn_running /
 * bn_num_batches (either may be NULL to skip the update); the workspace then holds everything
 * mivit_vit_backward needs. */
int mivit_vit_forward(const mivit_vit_config* cfg, int32_t B, const float* x, const float* features,
                      const float* params, float* bn_running, int64_t* bn_num_batches, void* workspace,
                      float* pred, int32_t training, void* stream);

/* Backward of the forward call that last used `workspace`.  grads (flat, same layout as params)
 * is OVERWRITTEN with dLoss/dparams given dpred = dLoss/dpred [B,1]. */
int mivit_vit_backward(const mivit_vit_config* cfg, int32_t B, const float* x, const float* features,
                       const float* dpred, const float* params, float* grads, void* workspace,
                       void* stream);

/* The same backward in two parts, for a data-parallel trainer that overlaps the gradient all-reduce with the backward
 * (SURVEY.md 8e): part 1 = head, encoder layers, tokens, feature paths and the embedding LayerNorm -- afterwards every gradient
 * at flat offset >= mivit_vit_embedding_param_count(cfg) is final and can be reduced while part 2 = the image embedding
 * (DeepResNet: 99 % of the FLOPs, ~45 % of the step) runs; part 0 = both (= mivit_vit_backward).  Part 1 zeroes `grads`. */
int mivit_vit_backward_part(const mivit_vit_config* cfg, int32_t B, const float* x, const float* features,
                            const float* dpred, const float* params, float* grads, void* workspace,
                            int32_t part, void* stream);
int64_t mivit_vit_embedding_param_count(const mivit_vit_config* cfg);   /* floats of the image-embedding block; -1: bad config */

/* rows of pred / dpred / target: B, or B*F for per_frame configurations */
int32_t mivit_vit_pred_rows(const mivit_vit_config* cfg, int32_t B);

/* The same forward / backward / training step starting from the TRAJECTORIES (BASELINE north_star (1): "fused with the
 * patch/temporal embedding so frames never round-trip HBM").  Replaces `model(torch.Tensor(normalize_images(
 * trajectories_to_video(trajs, ...))))` of the training loops (helpers/helpersGeneration.py:128-278,356-400 feeding
 * helpers/models.py:146-199,328-361).  traj: device float64 [B,T,2] (the caller applies the in-place y flip of :197), prm /
 * seed / seq_offset as for mivit_render_v1 (prm->normalize selects the fused normalize_images), T / prm->n == cfg->F.
 *   LinearProjectionEmbedding / CNNEmbedding: the forward runs mivit_render_embed_linear on the model's own weights and the
 *     backward mivit_render_embed_linear_wgrad (frames re-rendered, bit-identical); `frames` may be NULL, no frame reaches HBM.
 *   DeepResNetEmbedding: BatchNorm needs the statistics of all frames before any can be consumed, so they are rendered into
 *     `frames` (device fp32 [B,F,P,P], required) and the usual path runs; the backward reads them again for initial_conv.
 * seq_offset_dev (optional, may be NULL): DEVICE uint64 added to seq_offset when the kernels run -- a captured CUDA graph
 * advances the global sequence ids by updating it between replays.
 * The backward entry must be given the same traj / prm / seed / seq_offset as the forward that filled `workspace`. */
int mivit_vit_forward_traj(const mivit_vit_config* cfg, int32_t B, const double* traj, int32_t T,
                           const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset, const uint64_t* seq_offset_dev,
                           float* frames,
                           const float* features, const float* params, float* bn_running, int64_t* bn_num_batches,
                           void* workspace, float* pred, int32_t training, void* stream);
int mivit_vit_backward_traj(const mivit_vit_config* cfg, int32_t B, const double* traj, int32_t T,
                            const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset, const uint64_t* seq_offset_dev,
                            float* frames,
                            const float* features, const float* dpred, const float* params, float* grads, void* workspace,
                            int32_t part, void* stream);
int mivit_vit_train_step_traj(const mivit_vit_config* cfg, int32_t B, const double* traj, int32_t T,
                              const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset, const uint64_t* seq_offset_dev,
                              float* frames,
                              const float* features, const float* target, float* params, float* grads, float* adam_m,
                              float* adam_v, float* bn_running, int64_t* bn_num_batches, void* workspace, float* pred,
                              float* loss, float* dpred, float lr, float beta1, float beta2, float eps, float weight_decay,
                              int64_t step, int32_t apply_update, void* stream);

/* ------------------------------------------------- CNN baselines ----------------
 * Replaces helpers/models.py:600-772: BasicBlock, LightResNet, MultiImageResNet (ext_dim == 0) and LightImagesFeaturesResNet /
 * MultiImageFeatureResNet (ext_dim > 0) with num_blocks = [1,1,1] -- the ResNet every reference experiment trains beside its
 * ViTs (Experiments/PSFNoise/trainSettingsPSFNoise.py:114, trainSettingsImagesFeatures.py:170-173).  fp32, train-mode BatchNorm.
 * Flat parameter order (= state_dict order): resnet.conv1.weight, resnet.bn1.{weight,bias}, resnet.layer1.0.{conv1.weight,
 * bn1.weight, bn1.bias, conv2.weight, bn2.weight, bn2.bias}, resnet.layer{2,3}.0.{the same six, shortcut.0.weight,
 * shortcut.1.weight, shortcut.1.bias}, resnet.fc1.{weight,bias}, then resnet.fc2.{weight,bias} (MultiImageResNet) or
 * mlp.0.{weight,bias}, mlp.2.{weight,bias} (MultiImageFeatureResNet).  BatchNorm running statistics: flat [mean C | var C] for
 * bn1, layer1.0.{bn1,bn2}, layer2.0.{bn1,bn2,shortcut.1}, layer3.0.{bn1,bn2,shortcut.1} (1344 floats) + int64[9] counters. */
typedef struct mivit_resnet_config {
  int32_t P, F;              /* image size, frames per sequence                                                   */
  int32_t feature_size;      /* LightResNet feature_size (64)                                                     */
  int32_t ext_dim;           /* 0: MultiImageResNet; > 0: MultiImageFeatureResNet with external_dim features      */
  int32_t hidden;            /* MultiImageFeatureResNet hidden_size (128); ignored when ext_dim == 0              */
  int32_t single_prediction; /* MultiImageResNet: mean of the per-frame predictions [B,1] (1) or all of them [B,F,1] */
  int32_t activation;        /* 0 = nn.ReLU (the only one on the CUDA path)                                       */
  float bn_eps, bn_momentum;
} mivit_resnet_config;
int32_t mivit_resnet_param_count(const mivit_resnet_config* cfg);
int mivit_resnet_param_sizes(const mivit_resnet_config* cfg, int64_t* sizes, int32_t max_count);
int64_t mivit_resnet_workspace_bytes(const mivit_resnet_config* cfg, int32_t B);
int32_t mivit_resnet_pred_rows(const mivit_resnet_config* cfg, int32_t B);
/* x: [B,F,P,P] fp32; ext: [B,ext_dim] or NULL; pred: [pred_rows,1].  Same conventions as mivit_vit_forward / backward /
 * train_step (grads is overwritten; the backward follows the forward that last used `workspace`). */
int mivit_resnet_forward(const mivit_resnet_config* cfg, int32_t B, const float* x, const float* ext, const float* params,
                         float* bn_running, int64_t* bn_num_batches, void* workspace, float* pred, int32_t training,
                         void* stream);
int mivit_resnet_backward(const mivit_resnet_config* cfg, int32_t B, const float* x, const float* ext, const float* dpred,
                          const float* params, float* grads, void* workspace, void* stream);
int mivit_resnet_train_step(const mivit_resnet_config* cfg, int32_t B, const float* x, const float* ext, const float* target,
                            float* params, float* grads, float* adam_m, float* adam_v, float* bn_running,
                            int64_t* bn_num_batches, void* workspace, float* pred, float* loss, float* dpred, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int64_t step, int32_t apply_update,
                            void* stream);

/* nn.MSELoss() (mean) and its gradient w.r.t. pred (Experiments/PSFNoise/trainSettingsPSFNoise.py:31). */
int mivit_mse_loss(const float* pred, const float* target, int32_t n, float* loss, float* dpred,
                   void* stream);

/* torch.optim.AdamW step on a flat buffer (trainSettingsPSFNoise.py:119; decoupled weight decay,
 * bias correction; step is the 1-based step count).  grad_scale multiplies g first (1/world_size
 * after a sum all-reduce). */
int mivit_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                     void* stream);

/* One training step of the reference loop body (trainModelsPSFNoise.py:187-193):
 * forward, MSE, backward and (apply_update != 0) AdamW, enqueued on `stream` without host sync.
 * loss: [1], pred/dpred: [B,1] device scratch/outputs. */
int mivit_vit_train_step(const mivit_vit_config* cfg, int32_t B, const float* x, const float* features,
                         const float* target, float* params, float* grads, float* adam_m, float* adam_v,
                         float* bn_running, int64_t* bn_num_batches, void* workspace, float* pred,
                         float* loss, float* dpred, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int64_t step, int32_t apply_update, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIVIT_H_ */
