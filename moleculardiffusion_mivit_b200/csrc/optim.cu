// Loss and optimiser kernels of the training step
// (reference Experiments/PSFNoise/trainSettingsPSFNoise.py:31 nn.MSELoss(), :119 optim.AdamW(lr=1e-4);
//  torch defaults betas=(0.9,0.999), eps=1e-8, weight_decay=0.01).
#include "common.cuh"
#include "vit.h"
#include "../../include/mivit.h"

namespace {

// loss = mean((pred-target)^2);  dpred = 2 (pred-target) / n      (single block; n = batch size)
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int n,
                                                  float* __restrict__ loss, float* __restrict__ dpred) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = pred[i] - target[i];
    s = fmaf(d, d, s);
    if (dpred) dpred[i] = 2.0f * d / (float)n;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    if (loss) *loss = t / (float)n;
  }
}

// torch.optim.AdamW single step on a flat parameter buffer (decoupled weight decay).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                    float wd, float step_size, float inv_sqrt_bc2, float grad_scale) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = ga[k] * grad_scale;
      pa[k] *= (1.0f - lr * wd);
      ma[k] = ma[k] + (1.0f - b1) * (gr - ma[k]);          // lerp
      va[k] = b2 * va[k] + (1.0f - b2) * gr * gr;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    *reinterpret_cast<float4*>(p + i4) = pp;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (long long i = i4; i < n; ++i) {
      const float gr = g[i] * grad_scale;
      float pv = p[i] * (1.0f - lr * wd);
      const float mv = m[i] + (1.0f - b1) * (gr - m[i]);
      const float vv = b2 * v[i] + (1.0f - b2) * gr * gr;
      pv -= step_size * (mv / (sqrtf(vv) * inv_sqrt_bc2 + eps));
      p[i] = pv; m[i] = mv; v[i] = vv;
    }
  }
}

}  // namespace

int mse_loss(const float* pred, const float* target, int n, float* loss, float* dpred, cudaStream_t st) {
  mse_kernel<<<1, 256, 0, st>>>(pred, target, n, loss, dpred);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int adamw_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
               long long step, float grad_scale, cudaStream_t st) {
  if (n <= 0) return MIVIT_OK;
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  adamw_kernel<<<mivit_ceil_div((n + 3) / 4, 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, (float)(lr / bc1),
                                                                 (float)(1.0 / sqrt(bc2)), grad_scale);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_mse_loss(const float* pred, const float* target, int32_t n, float* loss, float* dpred, void* stream) {
  MIVIT_CHECK_ARG(pred && target && n > 0, "bad arguments");
  return mse_loss(pred, target, n, loss, dpred, (cudaStream_t)stream);
}

extern "C" int mivit_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  MIVIT_CHECK_ARG(p && g && m && v, "NULL pointer");
  MIVIT_CHECK_ARG(step >= 1, "step is 1-based");
  MIVIT_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "buffers must be 16-byte aligned");
  return adamw_flat(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream);
}
