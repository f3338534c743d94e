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

// W fp32 [cout][cin][k][k] -> bf16 core-matrix packs (see conv_aux.cu)
int pack_conv_weights(const float* W, __nv_bfloat16* out, int cout, int cin, int taps, int dgrad, cudaStream_t st);
