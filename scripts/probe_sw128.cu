// Hardware probe (development tool): tcgen05.mma kind::f16 reading a [row][128 B] SWIZZLE_128B shared-memory
// tile (the layout a TMA box {64 ch, R rows} with CU_TENSOR_MAP_SWIZZLE_128B produces) whose descriptor start
// address is shifted by an arbitrary number of rows (the convolution taps), as
//   test 0: K-major  A operand (rows = M, 16 channels = K)      -> forward / dgrad
//   test 1: MN-major B operand (rows = K, 64 or 128 channels = N) -> weight gradient
// The other operand is a one-hot selector, so D spells out which logical element the hardware fetched.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../moleculardiffusion_mivit_b200/csrc/umma.cuh"

constexpr int R = 176;   // rows per 64-channel region
#ifndef PITCH
#define PITCH 128
#endif
constexpr uint32_t kPitch = PITCH, kSwzMask = PITCH == 128 ? 7u : 3u, kLayout = PITCH == 128 ? 2u : 4u, kSbo = PITCH * 8, kChRow = PITCH / 2;

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)kLayout << 61;   // SWIZZLE_128B / SWIZZLE_64B
  return d;
}


// what: 0 -> values = row, 1 -> values = channel
__global__ void probe(float* out, int test, int N, int delta, int kblock, int use_base_off, int what) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* onehot = reinterpret_cast<__half*>(smem);                  // 128 x 16 K-major no-swizzle: [2 chunks][128 rows][8]
  uint8_t* tile = smem + 8192;                                        // 2 regions x R rows x 128 B, swizzled
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 8; i += blockDim.x) {
    const int c = i / 1024, r = (i / 8) % 128, e = i % 8;
    onehot[i] = __float2half((c * 8 + e) == (r % 16) ? 1.f : 0.f);
  }
  for (int i = tid; i < 2 * R * (int)kChRow; i += blockDim.x) {
    const int reg = i / (R * kChRow), row = (i / kChRow) % R, ch = i % kChRow;
    uint32_t off = (uint32_t)reg * R * kPitch + row * kPitch + ch * 2;
    off ^= ((off >> 7) & kSwzMask) << 4;                               // Swizzle<3|2,4,3> on the byte offset (tile is 1024-aligned)
    *reinterpret_cast<__half*>(tile + off) = __float2half(what == 0 ? (float)row : (float)(reg * kChRow + ch));
  }
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  if (warp == 0) umma::tmem_alloc<128>(&slot);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint64_t d_onehot = umma::make_desc(umma::smem_u32(onehot), 128u * 16u, 128u);
    if (test == 0) {
      // A = tile rows [delta, delta+128), channels [16*kblock, +16) (kblock < 4: region 0)
      const uint32_t addr = umma::smem_u32(tile) + (uint32_t)delta * kPitch + (uint32_t)kblock * 32u;
      const uint64_t da = desc_sw128(addr, 0, kSbo, use_base_off ? (addr >> 7) & 7u : 0u);
      // B one-hot: B[n][k] = (k == n % 16), N = 16
      umma::mma_bf16(tmem, da, d_onehot, idesc_f16(128, 16, 0, 0), 0);
    } else {
      // B = tile rows [delta + 16*kblock, +16) = K, channels [0, N) = MN;  LBO = next 64 channels, SBO = 8 rows
      const uint32_t addr = umma::smem_u32(tile) + (uint32_t)(delta + 16 * kblock) * kPitch;
      const uint64_t db = desc_sw128(addr, R * kPitch, kSbo, use_base_off ? (addr >> 7) & 7u : 0u);
      umma::mma_bf16(tmem, d_onehot, db, idesc_f16(128, N, 0, 1), 0);
    }
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  if (warp < 4) {
    float v[32];
    const int ncols = test == 0 ? 16 : N;
    for (int g = 0; g < (ncols + 31) / 32; ++g) {
      umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, v);
      for (int i = 0; i < 32 && g * 32 + i < ncols; ++i) out[(size_t)tid * 128 + g * 32 + i] = v[i];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<128>(tmem);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 128 * 4);
  float* h0 = (float*)malloc(128 * 128 * 4);
  float* h1 = (float*)malloc(128 * 128 * 4);
  const int smem = 8192 + 2 * R * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int deltas[] = {0, 1, 3, 7, 8, 13, 15, 29};
  for (int test = 0; test < 2; ++test)
    for (int use_bo = 0; use_bo < 2; ++use_bo)
      for (int N : {(int)kChRow, 2 * (int)kChRow})
        for (int kblock : {0, 1, PITCH == 128 ? 3 : 1})
          for (int delta : deltas) {
            if (test == 0 && N == 2 * (int)kChRow) continue;
            if (use_bo) continue;
            probe<<<1, 128, smem>>>(d, test, N, delta, kblock, use_bo, 0);
            cudaMemcpy(h0, d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            probe<<<1, 128, smem>>>(d, test, N, delta, kblock, use_bo, 1);
            cudaError_t e = cudaMemcpy(h1, d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            int bad = 0, first_m = -1, first_n = -1;
            if (test == 0) {
              for (int m = 0; m < 128; ++m)
                for (int k = 0; k < 16; ++k) {
                  const int row = (int)h0[m * 128 + k], ch = (int)h1[m * 128 + k];
                  if (row != delta + m || ch != 16 * kblock + k) { if (!bad) { first_m = m; first_n = k; } ++bad; }
                }
            } else {
              for (int m = 0; m < 16; ++m)   // D[m][n] = B[k = m][n]
                for (int n = 0; n < N; ++n) {
                  const int row = (int)h0[m * 128 + n], ch = (int)h1[m * 128 + n];
                  if (row != delta + 16 * kblock + m || ch != n) { if (!bad) { first_m = m; first_n = n; } ++bad; }
                }
            }
            printf("test %d base_off %d N %3d kblock %d delta %2d : %s bad=%d", test, use_bo, N, kblock, delta, cudaGetErrorString(e), bad);
            if (bad) printf("  first (m=%d,n=%d): got row %d ch %d", first_m, first_n, (int)h0[first_m * 128 + first_n], (int)h1[first_m * 128 + first_n]);
            printf("\n");
          }
  return 0;
}
