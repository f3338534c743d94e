// Hardware probe (development tool, not product): which shared-memory word does tcgen05.mma kind::tf32 read
// for element (k, n) of an MN-major operand in the no-swizzle layout?  A one-hot K-major operand selects
// row k; the probed operand's smem holds its own word index, so D spells out the address map.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/probe scripts/probe_tf32_mn.cu && gpurun_out/probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../moleculardiffusion_mivit_b200/csrc/umma.cuh"

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// which = 0: probe B (MN-major), A one-hot K-major.  which = 1: probe A (MN-major), B one-hot K-major.
// shift: probed words hold (index >> shift) & 1023
__global__ void probe(float* out, int which, int N, uint32_t lbo, uint32_t sbo, int shift, int swz) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* onehot = reinterpret_cast<float*>(smem);              // K-major [2 chunks][128 rows][4]
  float* probed = reinterpret_cast<float*>(smem + 8192);       // 64 KB of indices
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 4; i += blockDim.x) {
    const int c = i / 512, r = (i / 4) % 128, e = i % 4;
    const int k = c * 4 + e;
    onehot[i] = (k == (r % 8)) ? 1.f : 0.f;
  }
  for (int i = tid; i < 16384; i += blockDim.x) probed[i] = (float)((i >> shift) & 1023);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  if (warp == 0) umma::tmem_alloc<64>(&slot);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint64_t d_onehot = umma::make_desc(umma::smem_u32(onehot), 128u * 16u, 128u);
    const uint64_t d_probed = umma::make_desc(umma::smem_u32(probed), lbo, sbo) | ((uint64_t)swz << 61);
    if (which == 0) mma_tf32(tmem, d_onehot, d_probed, idesc_tf32(128, N, 0, 1), 0);
    else if (which == 1) mma_tf32(tmem, d_probed, d_onehot, idesc_tf32(128, N, 1, 0), 0);
    else if (which == 2) mma_tf32(tmem, d_onehot, d_probed, idesc_tf32(128, N, 0, 0), 0);   // control: B K-major
    else mma_tf32(tmem, d_probed, d_onehot, idesc_tf32(128, N, 0, 0), 0);                    // control: A K-major
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  if (warp < 4) {
    float v[32];
    for (int g = 0; g < N / 32; ++g) {
      umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, v);
      for (int i = 0; i < 32; ++i) out[(size_t)tid * N + g * 32 + i] = v[i];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<64>(tmem);
}

int main() {
  float* d;
  const int N = 32;
  cudaMalloc(&d, 128 * 64 * 4);
  float* h = (float*)malloc(128 * 64 * 4);
  float* h2 = (float*)malloc(128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 65536);
  const uint32_t cfg[][2] = {{128, 1024}, {1024, 128}, {2048, 512}, {512, 2048}};
  const int swzs[] = {0, 1, 2, 4, 6};
  for (int swz : swzs)
  for (int which = 0; which < 2; ++which)
    for (auto& c : cfg) {
      printf("#### swizzle code %d\n", swz);
      probe<<<1, 128, 8192 + 65536>>>(d, which, N, c[0], c[1], 0, swz);
      cudaMemcpy(h, d, 128 * N * 4, cudaMemcpyDeviceToHost);
      probe<<<1, 128, 8192 + 65536>>>(d, which, N, c[0], c[1], 10, swz);
      cudaError_t e = cudaMemcpy(h2, d, 128 * N * 4, cudaMemcpyDeviceToHost);
      printf("== probe %s which=%d, LBO=%u SBO=%u  (%s)\n", (which & 1) ? "A" : "B", which, c[0], c[1], cudaGetErrorString(e));
      if ((which & 1) == 0) {
        // D[m][n] = B[k = m%8][n] -> word index
        for (int k = 0; k < 8; ++k) {
          printf("k=%d:", k);
          for (int n = 0; n < N; ++n) printf(" %d", (int)h[k * N + n] + 1024 * (int)h2[k * N + n]);
          printf("\n");
        }
      } else {
        // D[m][n] = A[m][k = n%8]
        for (int k = 0; k < 8; ++k) {
          printf("k=%d:", k);
          for (int m = 0; m < 40; ++m) printf(" %d", (int)h[m * N + k] + 1024 * (int)h2[m * N + k]);
          printf(" ... m=64: %d m=127: %d\n", (int)h[64 * N + k] + 1024 * (int)h2[64 * N + k],
                 (int)h[127 * N + k] + 1024 * (int)h2[127 * N + k]);
        }
      }
    }
  return 0;
}
