// Weight packing for the shifted-row implicit-GEMM convolutions + C-ABI entry points that
// expose the convolution on the pitched-rows layout (used by the ViT and by the parity tests).
#include "common.cuh"
#include "vit.h"
#include "../../include/mivit.h"

namespace {

// W fp32 [cout][cin][k][k]  ->  forward pack  bf16 [taps][cin/8][cout][8]   (B operand, K = cin)
//                               dgrad pack    bf16 [taps][cout/8][cin][8]   (B operand, K = cout)
__global__ void pack_weights_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int cout, int cin,
                                    int taps, int dgrad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = taps * cout * cin;
  if (idx >= total) return;
  const int t = idx / (cout * cin);
  const int rem = idx - t * cout * cin;
  int co, ci;
  size_t dst;
  if (!dgrad) {
    const int chunk = rem / (cout * 8);
    const int r2 = rem - chunk * cout * 8;
    co = r2 / 8;
    ci = chunk * 8 + (r2 & 7);
    dst = (size_t)idx;
  } else {
    const int chunk = rem / (cin * 8);
    const int r2 = rem - chunk * cin * 8;
    ci = r2 / 8;
    co = chunk * 8 + (r2 & 7);
    dst = (size_t)idx;
  }
  out[dst] = __float2bfloat16_rn(W[((size_t)co * cin + ci) * taps + t]);
}

}  // namespace

int pack_conv_weights(const float* W, __nv_bfloat16* out, int cout, int cin, int taps, int dgrad, cudaStream_t st) {
  const int total = taps * cout * cin;
  pack_weights_kernel<<<mivit_ceil_div(total, 256), 256, 0, st>>>(W, out, cout, cin, taps, dgrad);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_conv_pack_weights(const float* W, void* out_bf16, int32_t cout, int32_t cin, int32_t ksize,
                                       int32_t dgrad, void* stream) {
  MIVIT_CHECK_ARG(W && out_bf16, "NULL pointer");
  MIVIT_CHECK_ARG(ksize == 1 || ksize == 3, "kernel size must be 1 or 3");
  MIVIT_CHECK_ARG(cin % 8 == 0 && cout % 8 == 0, "channels must be multiples of 8");
  return pack_conv_weights(W, (__nv_bfloat16*)out_bf16, cout, cin, ksize * ksize, dgrad, (cudaStream_t)stream);
}

extern "C" int mivit_conv_rows(const void* X_row0, const void* Wp, void* Y_row0, float* stats, int64_t rows, int32_t P,
                               int32_t cin, int32_t cout, int32_t ksize, int32_t mirrored, int32_t impl, void* stream) {
  MIVIT_CHECK_ARG(X_row0 && Wp && Y_row0, "NULL pointer");
  MIVIT_CHECK_ARG(ksize == 1 || ksize == 3, "kernel size must be 1 or 3");
  const ConvShifts sh = make_shifts(P, ksize * ksize, mirrored != 0);
  return conv_rows_forward((const __nv_bfloat16*)X_row0, (const __nv_bfloat16*)Wp, (__nv_bfloat16*)Y_row0, stats, rows, P,
                           cin, cout, ksize * ksize, sh, impl, (cudaStream_t)stream);
}

extern "C" int mivit_conv_rows_wgrad(const void* X_row0, const void* dY_row0, float* dW, int64_t rows, int32_t P, int32_t cin,
                                     int32_t cout, int32_t ksize, int32_t impl, void* stream) {
  MIVIT_CHECK_ARG(X_row0 && dY_row0 && dW, "NULL pointer");
  MIVIT_CHECK_ARG(ksize == 1 || ksize == 3, "kernel size must be 1 or 3");
  const ConvShifts sh = make_shifts(P, ksize * ksize, false);
  return conv_rows_wgrad((const __nv_bfloat16*)X_row0, (const __nv_bfloat16*)dY_row0, dW, rows, P, cin, cout, ksize * ksize, sh,
                         impl, (cudaStream_t)stream);
}

extern "C" int mivit_conv_rows_fused(const void* X_row0, const void* Wp, const void* Wskip, void* Y_row0, void* Yskip_row0,
                                     float* stats, float* stats_skip, int64_t rows, int32_t P, int32_t cin, int32_t cout,
                                     int32_t impl, void* stream) {
  MIVIT_CHECK_ARG(X_row0 && Wp && Wskip && Y_row0 && Yskip_row0, "NULL pointer");
  const ConvShifts sh = make_shifts(P, 9, false);
  return conv_rows_forward_fused((const __nv_bfloat16*)X_row0, (const __nv_bfloat16*)Wp, (const __nv_bfloat16*)Wskip,
                                 (__nv_bfloat16*)Y_row0, (__nv_bfloat16*)Yskip_row0, stats, stats_skip, rows, P, cin, cout, 9, sh,
                                 impl, (cudaStream_t)stream);
}

// nn.Linear building blocks on the tf32 tensor path (tests / ViT).  mode 0: Y = X W^T + b (relu);
// mode 1: dX (+)= dY W;  mode 2: dW += dY^T X (and, when `bias` is given, bias[out] += column sums of dY).  Returns an error when the shape is not supported by the
// tensor-core kernels (the ViT then uses its fp32 SIMT GEMM).
extern "C" int mivit_linear_tf32(int32_t mode, const float* A, const float* W, const float* bias, float* out, int32_t M,
                                 int32_t in_features, int32_t out_features, int32_t relu, int32_t accumulate, void* stream) {
  MIVIT_CHECK_ARG(A && W && out, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) {
    MIVIT_CHECK_ARG(linear_tc_supported(M, in_features, out_features), "shape not supported by the tf32 linear kernel");
    return linear_tc(A, W, bias, out, M, in_features, out_features, 0, relu, 0, st);
  }
  if (mode == 1) {
    MIVIT_CHECK_ARG(linear_tc_supported(M, out_features, in_features), "shape not supported by the tf32 linear kernel");
    return linear_tc(A, W, nullptr, out, M, out_features, in_features, 1, 0, accumulate, st);
  }
  MIVIT_CHECK_ARG(mode == 2 && linear_wgrad_tc_supported(M, out_features, in_features), "shape not supported by the tf32 wgrad kernel");
  return linear_wgrad_tc(A, W, out, const_cast<float*>(bias), M, out_features, in_features, st);  // A = dY [M,out], W = X [M,in], out = dW [out,in]
}
