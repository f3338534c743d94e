// nn.Linear forward / input-gradient / weight-gradient of the transformer stack and heads
// (reference helpers/models.py:20-23,64-65,241,268-273) on tcgen05 with kind::tf32: the fp32
// activations and weights are consumed directly as TF32 operands (no conversion pass, no packed
// copies), accumulators are fp32 in TMEM.  Same machinery as the convolutions (conv_tc2.cu):
// 128-token row slabs in the no-swizzle core-matrix order [chunk of 4 floats][row][16 B] filled with
// cp.async, double-buffered slabs and accumulators, resident weights, one elected issuing thread.
//   forward  Y[M,N]  = X[M,K] W[N,K]^T (+bias)(relu)    W is the K-major  B operand
//   dgrad    dX[M,N] = dY[M,K] W[K,N]  (+= optional)    W is the MN-major B operand
//   wgrad    dW[N,K] += dY[M,N]^T X[M,K]                both operands MN-major, rows = reduction
// MN-major TF32 operands: probed on B200 (scripts/probe_tf32_mn.cu), tcgen05.mma kind::tf32 returns zeros for
// MN-major operands in every descriptor layout except SWIZZLE_128B_BASE32B (layout type 1), whose address map is
//   byte(k, mn) = (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + (((mn%32)/8) ^ (k%4))*32 + (mn%8)*4
// (atoms of 4 reduction rows x 32 MN elements, 32-byte chunks XOR-swizzled by the row).  mn_off() below is that map.
// These GEMMs are HBM-bound (K, N <= 256 on 30k+ tokens): the point of the tensor path is to get the
// math out of the way so the kernel streams at memory speed.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kRows = 128;

__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}
// byte offset of the 16-byte chunk holding MN elements [4j, 4j+4) of reduction row k (SBO = 512: row groups contiguous)
__device__ __forceinline__ uint32_t mn_off(int k, int j, uint32_t lbo) {
  const int mn = 4 * j;
  return (uint32_t)(mn >> 5) * lbo + (uint32_t)(k >> 2) * 512u + (uint32_t)(k & 3) * 128u +
         (uint32_t)((((mn & 31) >> 3) ^ (k & 3)) << 5) + (uint32_t)((j & 1) << 4);
}
__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes) {
  return umma::make_desc(smem_addr, lbo_bytes, 512u) | ((uint64_t)1 << 61);   // SWIZZLE_128B_BASE32B
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM columns are allocated by need (power of two >= 32), not 512: with <= 256 columns and <= 113 KB of shared memory
// two CTAs share an SM, so the second tile of an SM no longer waits for the first one's load-MMA-store chain (these
// launches are latency-bound: 12.5 us fixed + 2.5 us per 1024 sequences).
__device__ __forceinline__ void tmem_alloc_n(uint32_t* slot, int cols) {
  if (cols <= 64) umma::tmem_alloc<64>(slot);
  else if (cols <= 128) umma::tmem_alloc<128>(slot);
  else if (cols <= 256) umma::tmem_alloc<256>(slot);
  else umma::tmem_alloc<512>(slot);
}
__device__ __forceinline__ void tmem_dealloc_n(uint32_t taddr, int cols) {
  if (cols <= 64) umma::tmem_dealloc<64>(taddr);
  else if (cols <= 128) umma::tmem_dealloc<128>(taddr);
  else if (cols <= 256) umma::tmem_dealloc<256>(taddr);
  else umma::tmem_dealloc<512>(taddr);
}
__host__ __device__ inline int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

// Up to four independent problems of the same shape in one launch (blockIdx.y): the q / k / v projections of an encoder
// layer (forward) and their weight gradients.  Each of these launches is a latency chain (load -> MMA -> store) over a single
// wave of CTAs; three problems in one grid let the load phase of one CTA overlap the epilogue of another on the same SM.
struct LinBatch {
  const float* A[4];
  const float* W[4];
  const float* bias[4];
  float* Y[4];
  const float* gate;   // optional [M,N]: Y = gate > 0 ? Y : 0  (backward of a ReLU fused into the dgrad that feeds it)
};
template <typename T>
__device__ __forceinline__ T pick3(T const (&a)[4], int z) { return z == 0 ? a[0] : z == 1 ? a[1] : z == 2 ? a[2] : a[3]; }

// ------------------------------------------------------------------ forward / dgrad -------------
// warps 0-3 epilogue, warp 4 MMA issuer, warps 5-7 producers.
template <bool BMN>
__global__ void __launch_bounds__(256, 1)
linear_tc_kernel(const __grid_constant__ LinBatch batch, int M, int K, int N, int n_tiles, int relu, int accumulate) {
  const float* __restrict__ A = pick3(batch.A, (int)blockIdx.y);
  const float* __restrict__ W = pick3(batch.W, (int)blockIdx.y);
  const float* __restrict__ bias = pick3(batch.bias, (int)blockIdx.y);
  float* __restrict__ Y = pick3(batch.Y, (int)blockIdx.y);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int kch = K / 4;                       // 16-byte chunks along the reduction dim of A
  const int slab_bytes = kch * kRows * 16;
  uint8_t* wsm = smem;                         // K-major: [K/4][N][16B]   MN-major: mn_off atoms
  uint8_t* slab0 = wsm + (size_t)K * N * 4;
  uint8_t* slab1 = slab0 + slab_bytes;
  uint8_t* stage = slab1 + slab_bytes;         // 4 epilogue warps x 4 KB transposition blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 4 * 4096);
  uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, 3);
      umma::mbar_init(empty + i, 1);
      umma::mbar_init(tfull + i, 1);
      umma::mbar_init(tempty + i, 4);
    }
    umma::mbar_fence_init();
  }
  const int acc_cols = pow2_cols(N);            // two accumulators, acc_cols columns apart
  if (warp == 0) tmem_alloc_n(tmem_slot, 2 * acc_cols);
  // The first tile's cp.async copies go out BEFORE the weights are fetched and the CTA synchronises: at B = 1024 every CTA
  // has exactly one tile, so the kernel is one latency chain and the two global round trips (weights, tile) used to be serial.
  // (Both slabs are free at start: no barrier is needed for this copy; the producer loop below only waits for it.)
  if (warp >= 5 && (int)blockIdx.x < n_tiles) {
    const int pt = (warp - 5) * 32 + lane;
    const long long m0 = (long long)blockIdx.x * kRows;
    const int rows_here = min(kRows, M - (int)m0);
    const uint4* src = reinterpret_cast<const uint4*>(A + m0 * K);
    for (int i = pt; i < rows_here * kch; i += 96) {
      const int r = i / kch, c = i - r * kch;
      cp16(slab0 + ((size_t)c * kRows + r) * 16, src + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (!BMN) {  // W[N][K] -> [K/4][N][16B]
    for (int i = tid; i < kch * N; i += 256) {
      const int n = i / kch, c = i - n * kch;
      *reinterpret_cast<uint4*>(wsm + ((size_t)c * N + n) * 16) = __ldg(reinterpret_cast<const uint4*>(W) + i);
    }
  } else {     // W[K][N] -> MN-major atoms (mn_off), 32-column groups K*128 bytes apart
    const int nch = N / 4;
    for (int i = tid; i < nch * K; i += 256) {
      const int k = i / nch, c = i - k * nch;
      *reinterpret_cast<uint4*>(wsm + mn_off(k, c, (uint32_t)K * 128u)) = __ldg(reinterpret_cast<const uint4*>(W) + i);
    }
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    const int pt = (warp - 5) * 32 + lane;
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int buf = k & 1;
      if (k > 0) {   // tile 0 was issued in the prologue
        umma::mbar_wait(empty + buf, ((k >> 1) & 1) ^ 1);
        uint8_t* slab = buf ? slab1 : slab0;
        const long long m0 = (long long)tile * kRows;
        const int rows_here = min(kRows, M - (int)m0);
        const uint4* src = reinterpret_cast<const uint4*>(A + m0 * K);
        for (int i = pt; i < rows_here * kch; i += 96) {
          const int r = i / kch, c = i - r * kch;
          cp16(slab + ((size_t)c * kRows + r) * 16, src + i);
        }
      }
      cp_wait_all();
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive(full + buf);
    }
  } else if (warp == 4) {
    const uint32_t idesc = idesc_tf32(kRows, N, 0, BMN ? 1 : 0);
    const uint64_t da_b[2] = {umma::make_desc(umma::smem_u32(slab0), (uint32_t)kRows * 16u, 128u),
                              umma::make_desc(umma::smem_u32(slab1), (uint32_t)kRows * 16u, 128u)};
    // K-major B: LBO = chunk stride (N*16), SBO = 128.   MN-major B: SW128_BASE32B atoms, LBO = K*128 (32-column groups)
    const uint64_t db0 = BMN ? make_desc_mn32(umma::smem_u32(wsm), (uint32_t)K * 128u)
                             : umma::make_desc(umma::smem_u32(wsm), (uint32_t)N * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da_b[0] >> 32), b_hi = (uint32_t)(db0 >> 32), b_lo0 = (uint32_t)db0;
    const uint32_t a_step = 2u * kRows;                           // two 4-float chunks per K = 8 step
    const uint32_t b_step = BMN ? 64u : 2u * (uint32_t)N;         // MN-major: 8 k-rows = 1024 B; K-major: 2 chunks
    const int ksteps = K / 8;
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int buf = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      umma::mbar_wait(full + buf, ph);
      umma::mbar_wait(tempty + buf, ph ^ 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        uint32_t a_lo = (uint32_t)da_b[buf], b_lo = b_lo0;
        const uint32_t acc = tmem + (uint32_t)(buf * acc_cols);
        for (int j = 0; j < ksteps; ++j) {
          mma_tf32(acc, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, j > 0 ? 1u : 0u);
          a_lo += a_step;
          b_lo += b_step;
        }
        umma::commit(empty + buf);
        umma::commit(tfull + buf);
      }
      __syncwarp();
    }
  } else {
    int k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      const long long row0 = (long long)tile * kRows + warp * 32;
      const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * acc_cols);
      const int groups = N / 32;          // N % 32 == 0 on this path
      // Each warp transposes its 32 rows x 32 columns through a private, XOR-swizzled 4 KB staging block so that the
      // global stores are 128 contiguous bytes per row (a thread owns a ROW of the accumulator: storing straight from
      // registers wrote 16-byte pieces 4*N bytes apart, 32 lines per instruction, and capped the kernel at ~1.1 TB/s).
      float4* stg = reinterpret_cast<float4*>(stage + warp * 4096);
      const int c = lane & 7, rsub = lane >> 3;
      for (int g = 0; g < groups; ++g) {
        float v[32];
        umma::tmem_ld32(acc + (uint32_t)(g * 32), v);
        if (g == groups - 1) {
          umma::fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive(tempty + buf);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) stg[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) bb = __ldg(reinterpret_cast<const float4*>(bias + g * 32) + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 4 + rsub;
          if (row0 + row < M) {
            float4 o = stg[row * 8 + (c ^ (row & 7))];
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            float4* p = reinterpret_cast<float4*>(Y + (row0 + row) * N + g * 32) + c;
            if (accumulate == 2) {   // several problems of the launch add into the SAME output (dx += dq Wq + dk Wk + dv Wv)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
              continue;
            }
            if (accumulate) {
              const float4 old = *p;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (batch.gate != nullptr) {
              const float4 gt = __ldg(reinterpret_cast<const float4*>(batch.gate + (row0 + row) * N + g * 32) + c);
              o.x = gt.x > 0.f ? o.x : 0.f; o.y = gt.y > 0.f ? o.y : 0.f; o.z = gt.z > 0.f ? o.z : 0.f; o.w = gt.w > 0.f ? o.w : 0.f;
            }
            *p = o;
          }
        }
        __syncwarp();
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc_n(tmem, 2 * acc_cols);
}

// ------------------------------------------------------------------ weight gradient -------------
// warps 0-3 producers + final flush, warp 4 MMA issuer, warps 5-7 producers.
// tmA / tmB (optional): fp32 tensor maps of dY / X with 32-byte-atom 128-byte swizzle (tma.cuh: make_f32_tensor_map_sw32) -- one
// elected thread then stages both operands with TMA (N / 32 + K / 32 boxes of 128 rows per stage) instead of seven warps of cp.async.
__device__ __forceinline__ void wgrad_body(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ dW,
                                           float* __restrict__ db, int M, int N, int K, int s_begin, int s_end, int buf_bytes,
                                           const CUtensorMap* tmA = nullptr, const CUtensorMap* tmB = nullptr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int nch = N / 4, kch = K / 4;
  const int a_bytes = nch * kRows * 16;
  uint8_t* buf0 = smem;
  uint8_t* buf1 = smem + buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * buf_bytes);
  uint64_t *full = bars, *empty = bars + 2, *done = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, tmA != nullptr ? 1 : 7);
      umma::mbar_init(empty + i, 1);
    }
    umma::mbar_init(done, 1);
    umma::mbar_fence_init();
    if (tmA != nullptr) { tma::prefetch_map(tmA); tma::prefetch_map(tmB); }
  }
  if (warp == 0) tmem_alloc_n(tmem_slot, K + 32);
  // bias gradient for free: a 33rd..-th column group of the B operand whose first column is all ones makes accumulator
  // column K the column sum of dY (db).  The group lives behind the X slab of both stage buffers and is written once.
  if (db != nullptr) {
    const int b_bytes = kch * kRows * 16;
    for (int i = tid; i < 2 * kRows * 8; i += 256) {
      const int b = i / (kRows * 8), r = (i >> 3) & (kRows - 1), j = i & 7;
      uint8_t* g = (b ? smem + buf_bytes : smem) + a_bytes + b_bytes;     // one more 32-column group: LBO = kRows*128 after the last
      *reinterpret_cast<float4*>(g + mn_off(r, j, kRows * 128u)) = make_float4(j == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
    }
    umma::fence_proxy_async();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (tmA != nullptr) {
    if (tid == 0) {   // TMA producer: rows past M read as zeros
      int k = 0;
      for (int s = s_begin; s < s_end; ++s, ++k) {
        const int b = k & 1;
        umma::mbar_wait(empty + b, ((k >> 1) & 1) ^ 1);
        uint8_t* aslab = b ? buf1 : buf0;
        uint8_t* bslab = aslab + a_bytes;
        tma::expect_tx(full + b, (uint32_t)((N + K) * kRows * 4));
        for (int g = 0; g < N / 32; ++g) tma::load_tile(aslab + (size_t)g * kRows * 128, tmA, g * 32, s * kRows, full + b);
        for (int g = 0; g < K / 32; ++g) tma::load_tile(bslab + (size_t)g * kRows * 128, tmB, g * 32, s * kRows, full + b);
      }
    }
  }
  if (warp != 4 && tmA == nullptr) {
    const int pt = (warp < 4 ? warp : warp - 1) * 32 + lane;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int b = k & 1;
      umma::mbar_wait(empty + b, ((k >> 1) & 1) ^ 1);
      uint8_t* aslab = b ? buf1 : buf0;
      uint8_t* bslab = aslab + a_bytes;
      const long long r0 = (long long)s * kRows;
      const int rows_here = min(kRows, M - (int)r0);
      const uint4* srca = reinterpret_cast<const uint4*>(dY + r0 * N);
      const uint4* srcb = reinterpret_cast<const uint4*>(X + r0 * K);
      for (int i = pt; i < kRows * nch; i += 224) {
        const int r = i / nch, c = i - r * nch;
        uint8_t* d = aslab + mn_off(r, c, kRows * 128u);
        if (r < rows_here) cp16(d, srca + i); else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
      }
      for (int i = pt; i < kRows * kch; i += 224) {
        const int r = i / kch, c = i - r * kch;
        uint8_t* d = bslab + mn_off(r, c, kRows * 128u);
        if (r < rows_here) cp16(d, srcb + i); else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
      }
      cp_wait_all();
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive(full + b);
    }
  }
  if (warp == 4) {
    const uint32_t idesc = idesc_tf32(128, db != nullptr ? K + 32 : K, 1, 1);
    // MN-major operands (SW128_BASE32B atoms): 32-channel groups kRows*128 bytes apart, 4-row groups 512 bytes apart
    const uint64_t da_b[2] = {make_desc_mn32(umma::smem_u32(buf0), kRows * 128u), make_desc_mn32(umma::smem_u32(buf1), kRows * 128u)};
    const uint64_t db_b[2] = {make_desc_mn32(umma::smem_u32(buf0) + (uint32_t)a_bytes, kRows * 128u),
                              make_desc_mn32(umma::smem_u32(buf1) + (uint32_t)a_bytes, kRows * 128u)};
    const uint32_t a_hi = (uint32_t)(da_b[0] >> 32), b_hi = (uint32_t)(db_b[0] >> 32);
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int b = k & 1;
      umma::mbar_wait(full + b, (k >> 1) & 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t a_lo0 = (uint32_t)da_b[b], b_lo0 = (uint32_t)db_b[b];
#pragma unroll
        for (int kk = 0; kk < kRows / 8; ++kk)   // K = 8 rows per tf32 MMA
          mma_tf32(tmem, ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(kk * 64)), ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)(kk * 64)),
                   idesc, (k > 0 || kk > 0) ? 1u : 0u);
        umma::commit(empty + b);
        if (s == s_end - 1) umma::commit(done);
      }
      __syncwarp();
    }
  }
  if (warp < 4 && s_end > s_begin) {
    umma::mbar_wait(done, 0);
    umma::fence_after_sync();
    if (warp * 32 < N) {
      const int n = warp * 32 + lane;
      for (int cg = 0; cg < K / 32; ++cg) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cg * 32), v);
        if (n < N) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)   // 16-byte vector reductions: 4x fewer L2 atomic operations
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dW + (size_t)n * K + cg * 32 + i), "f"(v[i]), "f"(v[i + 1]),
                         "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        }
      }
      if (db != nullptr) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)K, v);
        if (n < N) atomicAdd(db + n, v[0]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc_n(tmem, K + 32);
}

struct LinMaps {
  CUtensorMap a[4], b[4];   // dY / X of each problem (make_f32_tensor_map_sw32, 128-row boxes)
};
__global__ void __launch_bounds__(256, 1)
linear_wgrad_tc_kernel(const __grid_constant__ LinBatch batch /* A = dY, W = X, Y = dW, bias = db */, const __grid_constant__ LinMaps tm,
                       int M, int N, int K, int n_stages, int stages_per_cta, int buf_bytes) {
  const int s_begin = blockIdx.x * stages_per_cta;
  wgrad_body(pick3(batch.A, (int)blockIdx.y), pick3(batch.W, (int)blockIdx.y), pick3(batch.Y, (int)blockIdx.y),
             const_cast<float*>(pick3(batch.bias, (int)blockIdx.y)), M, N, K, s_begin, min(n_stages, s_begin + stages_per_cta), buf_bytes,
             &tm.a[blockIdx.y], &tm.b[blockIdx.y]);
}

// Several weight gradients of DIFFERENT shapes over the same M rows in one launch (the six nn.Linear layers of an encoder layer):
// the CTAs of the grid are dealt to the problems in proportion to their bytes per row, each CTA owns a contiguous row range of
// ONE problem and flushes its accumulator once -- ~148 flushes per layer instead of one per 256 rows and problem, and one
// launch instead of three.
constexpr int kMaxMulti = 6;
struct WgradMulti {
  const float* dY[kMaxMulti];
  const float* X[kMaxMulti];
  float* dW[kMaxMulti];
  float* db[kMaxMulti];
  int N[kMaxMulti], K[kMaxMulti], spc[kMaxMulti];
  int cta0[kMaxMulti + 1];
  int nb;
};
struct WgradMaps {
  CUtensorMap a[kMaxMulti], b[kMaxMulti];   // dY / X of each problem (make_f32_tensor_map_sw32, 128-row boxes)
};
__global__ void __launch_bounds__(256, 1) linear_wgrad_multi_kernel(const __grid_constant__ WgradMulti mp, const __grid_constant__ WgradMaps tm,
                                                                    int M, int n_stages, int buf_bytes) {
  int pidx = 0;
#pragma unroll
  for (int i = 1; i < kMaxMulti; ++i)
    if (i < mp.nb && (int)blockIdx.x >= mp.cta0[i]) pidx = i;
  const int local = (int)blockIdx.x - mp.cta0[pidx];
  const int s_begin = local * mp.spc[pidx];
  wgrad_body(mp.dY[pidx], mp.X[pidx], mp.dW[pidx], mp.db[pidx], M, mp.N[pidx], mp.K[pidx], s_begin, min(n_stages, s_begin + mp.spc[pidx]), buf_bytes,
             &tm.a[pidx], &tm.b[pidx]);
}

int sm_count() {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  return sms ? sms : 148;
}

}  // namespace

bool linear_tc_supported(int M, int K, int N) {
  return M >= 512 && K % 8 == 0 && N % 32 == 0 && K >= 8 && N >= 32 && N <= 256 && K <= 256 &&
         (size_t)K * N * 4 + 2 * (size_t)(K / 4) * kRows * 16 + 4 * 4096 + 256 <= 227 * 1024;
}

// mode 0: Y = A W^T (+bias)(relu), W [N][K].   mode 1: Y (+)= A W, W [K][N].   nb <= 4 problems of the same shape; accumulate = 2 adds with vector atomics (problems may share Y).
int linear_tc_batched(int nb, const float* const* A, const float* const* W, const float* const* bias, float* const* Y, int M, int K,
                      int N, int mode, int relu, int accumulate, cudaStream_t st, const float* gate) {
  LinBatch b = {};
  b.gate = gate;
  for (int i = 0; i < nb; ++i) { b.A[i] = A[i]; b.W[i] = W[i]; b.bias[i] = bias ? bias[i] : nullptr; b.Y[i] = Y[i]; }
  int smem = K * N * 4 + 2 * (K / 4) * kRows * 16 + 4 * 4096 + 256;
  if (2 * pow2_cols(N) > 256 && smem < 120 * 1024) smem = 120 * 1024;   // 512 TMEM columns: one CTA per SM
  const int n_tiles = (M + kRows - 1) / kRows;
  const int per_sm = (smem <= 113 * 1024 && 2 * pow2_cols(N) <= 256) ? 2 : 1;
  const int grid = n_tiles < per_sm * sm_count() ? n_tiles : per_sm * sm_count();
  MivitProfScope prof(mode ? "linear_tc_dgrad" : "linear_tc_fwd", 2.0 * M * K * N * nb, st);
  if (mode == 0) {
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    linear_tc_kernel<false><<<dim3(grid, nb), 256, smem, st>>>(b, M, K, N, n_tiles, relu, accumulate);
  } else {
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    linear_tc_kernel<true><<<dim3(grid, nb), 256, smem, st>>>(b, M, K, N, n_tiles, relu, accumulate);
  }
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
int linear_tc(const float* A, const float* W, const float* bias, float* Y, int M, int K, int N, int mode, int relu, int accumulate,
              cudaStream_t st) {
  return linear_tc_batched(1, &A, &W, &bias, &Y, M, K, N, mode, relu, accumulate, st, nullptr);
}

bool linear_wgrad_tc_supported(int M, int N, int K) {
  // N = out_features (rows of dW, <= 128 lanes), K = in_features (TMEM columns); + one 32-column group for the bias gradient
  size_t buf = (size_t)(N / 4 + K / 4 + 8) * kRows * 16;
  if (buf < (size_t)32 * kRows * 16) buf = (size_t)32 * kRows * 16;
  return M >= 512 && N % 32 == 0 && K % 32 == 0 && N >= 32 && N <= 128 && K >= 32 && K <= 256 && 2 * buf + 256 <= 227 * 1024;
}

// dW[N][K] += dY[M,N]^T X[M,K]   (dW pre-zeroed / holds the value to accumulate onto)
// db (optional, [N]) += column sums of dY
int linear_wgrad_tc_batched(int nb, const float* const* dY, const float* const* X, float* const* dW, float* const* db, int M, int N,
                            int K, cudaStream_t st) {
  LinBatch b = {};
  for (int i = 0; i < nb; ++i) { b.A[i] = dY[i]; b.W[i] = X[i]; b.Y[i] = dW[i]; b.bias[i] = db ? db[i] : nullptr; }
  // the M = 128 MMA over-reads the A slab up to 32 chunks: keep that inside the stage buffer
  int buf_bytes = (N / 4 + K / 4 + 8) * kRows * 16;
  const int need = 32 * kRows * 16;
  if (buf_bytes < need) buf_bytes = need;
  int smem = 2 * buf_bytes + 256;
  if (K + 32 > 256 && smem < 120 * 1024) smem = 120 * 1024;   // 512 TMEM columns: one CTA per SM
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(linear_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_stages = (M + kRows - 1) / kRows;
  int ctas = n_stages < sm_count() ? n_stages : sm_count();
  const int spc = (n_stages + ctas - 1) / ctas;
  ctas = (n_stages + spc - 1) / spc;
  LinMaps tm;
  for (int i = 0; i < nb; ++i) {
    int rc = make_f32_tensor_map_sw32(&tm.a[i], dY[i], N, M, kRows);
    if (!rc) rc = make_f32_tensor_map_sw32(&tm.b[i], X[i], K, M, kRows);
    if (rc) return rc;
  }
  for (int i = nb; i < 4; ++i) { tm.a[i] = tm.a[0]; tm.b[i] = tm.b[0]; }
  MivitProfScope prof("linear_tc_wgrad", 2.0 * M * K * N * nb, st);
  linear_wgrad_tc_kernel<<<dim3(ctas, nb), 256, smem, st>>>(b, tm, M, N, K, n_stages, spc, buf_bytes);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
int linear_wgrad_tc(const float* dY, const float* X, float* dW, float* db, int M, int N, int K, cudaStream_t st) {
  return linear_wgrad_tc_batched(1, &dY, &X, &dW, &db, M, N, K, st);
}

// nb <= 6 weight gradients dW_i[N_i][K_i] += dY_i[M,N_i]^T X_i[M,K_i] (+ bias gradients) over the same M rows in ONE launch
bool linear_wgrad_tc_multi_supported(int nb, int M, const int* N, const int* K) {
  if (nb < 1 || nb > kMaxMulti) return false;
  size_t buf = 0;
  for (int i = 0; i < nb; ++i) {
    if (!linear_wgrad_tc_supported(M, N[i], K[i])) return false;
    size_t b = (size_t)(N[i] / 4 + K[i] / 4 + 8) * kRows * 16;
    if (b < (size_t)32 * kRows * 16) b = (size_t)32 * kRows * 16;
    if (b > buf) buf = b;
  }
  return 2 * buf + 256 <= 227 * 1024;
}
int linear_wgrad_tc_multi(int nb, const float* const* dY, const float* const* X, float* const* dW, float* const* db, int M, const int* N,
                          const int* K, cudaStream_t st) {
  MIVIT_CHECK_ARG(linear_wgrad_tc_multi_supported(nb, M, N, K), "batched weight gradient: unsupported shapes");
  WgradMulti mp = {};
  mp.nb = nb;
  int buf_bytes = 32 * kRows * 16;
  double bytes_total = 0, flops = 0;
  for (int i = 0; i < nb; ++i) {
    mp.dY[i] = dY[i]; mp.X[i] = X[i]; mp.dW[i] = dW[i]; mp.db[i] = db ? db[i] : nullptr; mp.N[i] = N[i]; mp.K[i] = K[i];
    const int b = (N[i] / 4 + K[i] / 4 + 8) * kRows * 16;
    if (b > buf_bytes) buf_bytes = b;
    bytes_total += N[i] + K[i];
    flops += 2.0 * M * N[i] * K[i];
  }
  const int n_stages = (M + kRows - 1) / kRows;
  const int sms = sm_count();
  int used = 0;
  for (int i = 0; i < nb; ++i) {
    // CTAs in proportion to the bytes a row of the problem moves, at least one, at most one per stage
    int c = (int)((double)sms * (N[i] + K[i]) / bytes_total);
    if (c < 1) c = 1;
    if (c > n_stages) c = n_stages;
    const int spc = (n_stages + c - 1) / c;
    c = (n_stages + spc - 1) / spc;
    mp.spc[i] = spc;
    mp.cta0[i] = used;
    used += c;
  }
  mp.cta0[nb] = used;
  for (int i = nb + 1; i <= kMaxMulti; ++i) mp.cta0[i] = used;
  const int smem = 2 * buf_bytes + 256;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(linear_wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  WgradMaps tm;
  for (int i = 0; i < nb; ++i) {
    int rc = make_f32_tensor_map_sw32(&tm.a[i], dY[i], N[i], M, kRows);
    if (!rc) rc = make_f32_tensor_map_sw32(&tm.b[i], X[i], K[i], M, kRows);
    if (rc) return rc;
  }
  for (int i = nb; i < kMaxMulti; ++i) { tm.a[i] = tm.a[0]; tm.b[i] = tm.b[0]; }
  MivitProfScope prof("linear_tc_wgrad", flops, st);
  linear_wgrad_multi_kernel<<<used, 256, smem, st>>>(mp, tm, M, n_stages, buf_bytes);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
