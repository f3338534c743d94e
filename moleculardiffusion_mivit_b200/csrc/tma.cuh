// TMA (cp.async.bulk.tensor) helpers for the pitched-rows activation tensors.
// A [rows][C] bf16 matrix is described to the TMA unit as the 3-D tensor
//     d0 = 8 channels (16 B, contiguous)   d1 = rows (stride C*2 B)   d2 = C/8 channel chunks (stride 16 B)
// so that ONE box {8, R, C/8} lands in shared memory as [chunk][row][16 B] -- exactly the no-swizzle UMMA
// core-matrix order the convolution kernels use (K-major for forward/dgrad, MN-major for wgrad).  The
// global->shared transposition that 3-7 producer warps did with 16-byte cp.async is done by the copy engine.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "umma.cuh"

// cuTensorMapEncodeTiled is the only driver-API entry point this library needs.  It is resolved through the runtime
// (cudaGetDriverEntryPoint, capi.cu) instead of linking libcuda, so libmivit_b200.so loads -- and its ABI version / symbol table
// can be checked -- on a machine without a driver (the build container); it fails loudly at the first launch there.
typedef CUresult (*mivit_tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
mivit_tensor_map_encode_fn mivit_tensor_map_encoder();   // nullptr (and mivit_last_error set) when the driver is not available

// base = first guard row (row0 - guard_rows*C); total_rows includes both guards.  box = box_chunks x box_rows (<= 256) x 16 B.
static inline int make_rows_tensor_map(CUtensorMap* tm, const void* base, int C, long long total_rows, int box_rows,
                                       int box_chunks) {
  const cuuint64_t gdim[3] = {8, (cuuint64_t)total_rows, (cuuint64_t)(C / 8)};
  const cuuint64_t gstride[2] = {(cuuint64_t)C * 2, 16};   // bytes, dims 1..2
  const cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_chunks};
  const cuuint32_t estr[3] = {1, 1, 1};
  mivit_tensor_map_encode_fn encode = mivit_tensor_map_encoder();
  if (encode == nullptr) return MIVIT_ERR_CUDA;
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mivit_set_error("cuTensorMapEncodeTiled failed (%d) for C=%d rows=%lld box=%d", (int)r, C, total_rows, box_rows);
    return MIVIT_ERR_CUDA;
  }
  return MIVIT_OK;
}

// Row tile for the swizzled UMMA layouts: 2-D view {C channels, rows}; one box = {min(C,64) channels, box_rows rows}
// lands as [row][128 B] (C >= 64, SWIZZLE_128B) or [row][64 B] (C = 32, SWIZZLE_64B).  Probed on B200
// (scripts/probe_sw128.cu): both tcgen05 K-major (rows = M) and MN-major (rows = K) descriptors read such a tile
// correctly from a start address shifted by ANY number of rows, with base_offset = 0 (the XOR uses absolute
// shared-memory address bits), so the 9 convolution taps are 9 start addresses into one tile.
static inline int make_rows_tensor_map_sw(CUtensorMap* tm, const void* base, int C, long long total_rows, int box_rows,
                                          int box_cols = 0) {
  if (box_cols <= 0) box_cols = C >= 64 ? 64 : C;   // 64 channels = one 128-byte swizzled row; 32 channels -> SWIZZLE_64B
  const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)total_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  mivit_tensor_map_encode_fn encode = mivit_tensor_map_encoder();
  if (encode == nullptr) return MIVIT_ERR_CUDA;
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mivit_set_error("cuTensorMapEncodeTiled failed (%d) for C=%d rows=%lld box=%d", (int)r, C, total_rows, box_rows);
    return MIVIT_ERR_CUDA;
  }
  return MIVIT_OK;
}

// fp32 row-major matrix [rows][cols] for the tf32 kernels: one box = {32 floats = 128 B, box_rows rows} lands as a [row][128 B]
// SWIZZLE_128B tile, which tcgen05 kind::tf32 reads as a K-major operand (8 reduction elements = 32 B per MMA: the descriptor
// start address advances by 32 B inside the 128-byte atom, then to the next box).  Rows past `rows` read as zeros.
static inline int make_f32_tensor_map_sw(CUtensorMap* tm, const void* base, int cols, long long rows, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  mivit_tensor_map_encode_fn encode = mivit_tensor_map_encoder();
  if (encode == nullptr) return MIVIT_ERR_CUDA;
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mivit_set_error("cuTensorMapEncodeTiled failed (%d) for fp32 cols=%d rows=%lld box=%d", (int)r, cols, rows, box_rows);
    return MIVIT_ERR_CUDA;
  }
  return MIVIT_OK;
}

// The same for MN-major tf32 operands (rows = reduction index, 32 floats of the M / N dimension per 128-byte line):
// SWIZZLE_128B with 32-byte atoms = the UMMA layout SWIZZLE_128B_BASE32B, whose 32-byte chunks are XOR-ed with the row (the
// address map linear_tc.cu's mn_off() reproduces when it fills shared memory with cp.async).
static inline int make_f32_tensor_map_sw32(CUtensorMap* tm, const void* base, int cols, long long rows, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  mivit_tensor_map_encode_fn encode = mivit_tensor_map_encoder();
  if (encode == nullptr) return MIVIT_ERR_CUDA;
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mivit_set_error("cuTensorMapEncodeTiled (128B swizzle, 32-byte atoms) failed (%d) for fp32 cols=%d rows=%lld box=%d", (int)r, cols, rows, box_rows);
    return MIVIT_ERR_CUDA;
  }
  return MIVIT_OK;
}

namespace tma {

// UMMA shared-memory descriptor for a swizzled row tile (pitch 128 -> SWIZZLE_128B, pitch 64 -> SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t pitch) {
  return umma::make_desc(smem_addr, lbo_bytes, 8u * pitch) | ((uint64_t)(pitch == 128u ? 2u : 4u) << 61);
}
__device__ __forceinline__ void load_tile(void* smem_dst, const CUtensorMap* tm, int ch0, int row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   umma::smem_u32(smem_dst)),
               "l"(tm), "r"(ch0), "r"(row), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void load_tile_multicast(void* smem_dst, const CUtensorMap* tm, int ch0, int row, uint64_t* bar,
                                                    uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], "
      "%5;" ::"r"(umma::smem_u32(smem_dst)),
      "l"(tm), "r"(ch0), "r"(row), "r"(umma::smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
// box at (0, row, chunk) -> smem; completes `bytes` on the CTA-local mbarrier
__device__ __forceinline__ void load_rows(void* smem_dst, const CUtensorMap* tm, int row, int chunk, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tm), "r"(0), "r"(row), "r"(chunk), "r"(umma::smem_u32(bar))
      : "memory");
}
// same box delivered to the same shared-memory offset of every CTA in cta_mask (and its mbarrier there)
__device__ __forceinline__ void load_rows_multicast(void* smem_dst, const CUtensorMap* tm, int row, int chunk, uint64_t* bar,
                                                    uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], "
      "[%5], %6;" ::"r"(umma::smem_u32(smem_dst)),
      "l"(tm), "r"(0), "r"(row), "r"(chunk), "r"(umma::smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit arriving on the same mbarrier offset in every CTA of cta_mask
__device__ __forceinline__ void commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   umma::smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace tma
