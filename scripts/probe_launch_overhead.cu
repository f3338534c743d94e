// Development probe: fixed cost of a small tcgen05 kernel launch (empty / +TMEM alloc / +16 KB weight fill / +mbarrier init),
// 248 CTAs x 256 threads, 96 KB dynamic shared memory, back to back in one stream.
#include <cstdio>
#include <cuda_runtime.h>
#include "../moleculardiffusion_mivit_b200/csrc/umma.cuh"
__global__ void __launch_bounds__(256) k(int mode, const float* W, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bars[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (mode >= 3 && tid == 0) { for (int i = 0; i < 8; ++i) umma::mbar_init(bars + i, 1); umma::mbar_fence_init(); }
  if (mode >= 1 && warp == 0) umma::tmem_alloc<128>(&slot);
  if (mode >= 2) for (int i = tid; i < 1024; i += 256) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(W) + i);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (mode >= 2 && tid == 0 && smem[17] == 77) out[blockIdx.x] = 1.f;
  __syncthreads();
  if (mode >= 1 && warp == 0) umma::tmem_dealloc<128>(slot);
}
int main() {
  float *W, *out; cudaMalloc(&W, 1 << 20); cudaMalloc(&out, 1 << 20); cudaMemset(W, 0, 1 << 20);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int smem : {96 * 1024}) for (int mode = 0; mode < 4; ++mode) {
    for (int i = 0; i < 10; ++i) k<<<248, 256, smem>>>(mode, W, out);
    cudaEventRecord(e0);
    for (int i = 0; i < 200; ++i) k<<<248, 256, smem>>>(mode, W, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("smem %6d mode %d: %.2f us per launch (%s)\n", smem, mode, ms * 1000 / 200, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
