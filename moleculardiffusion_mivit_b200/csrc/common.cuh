// Shared helpers for the mivit_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define MIVIT_OK 0
#define MIVIT_ERR_INVALID 1
#define MIVIT_ERR_CUDA 2

// thread-local message returned by mivit_last_error()
void mivit_set_error(const char* fmt, ...);

#define MIVIT_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      mivit_set_error(__VA_ARGS__);           \
      return MIVIT_ERR_INVALID;               \
    }                                         \
  } while (0)

#define MIVIT_CUDA_CHECK(expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      mivit_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,            \
                      cudaGetErrorString(_e));                                        \
      return MIVIT_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define MIVIT_LAUNCH_CHECK() MIVIT_CUDA_CHECK(cudaGetLastError())

// launch counter (bench.py reports gpu_launches from it)
void mivit_count_launch(int n = 1);

static inline int mivit_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Validity (not a pad row / column) of the rows r0, r0 + stride, r0 + 2*stride, ... of the pitched-rows activation layout
// (conv_tc.cu) without a division per row: r mod (P+1)^2 is carried incrementally and the split into (y, x) is a 32-bit
// multiply-high (exact: q < 2^16, (P+1) <= 101).  The convolution epilogues call this once per 128-row tile and thread; the
// 64-bit '%' and the int division they used before were ~45 dependent instructions with two MUFU.RCP on the epilogue's critical path.
struct RowWalker {
  long long r, stride;
  uint32_t q, step, rpf, pitch, magic;
  __device__ __forceinline__ void init(long long r0, long long stride_, int P) {
    pitch = (uint32_t)(P + 1);
    rpf = pitch * pitch;
    magic = (uint32_t)((0x100000000ull + pitch - 1) / pitch);
    r = r0;
    stride = stride_;
    q = (uint32_t)(r0 % (long long)rpf);
    step = (uint32_t)(stride_ % (long long)rpf);
  }
  __device__ __forceinline__ bool valid(long long rows, int P) const {
    const uint32_t y = __umulhi(q, magic), x = q - y * pitch;
    return r < rows && y < (uint32_t)P && x < (uint32_t)P;
  }
  __device__ __forceinline__ void next() {
    r += stride;
    q += step;
    if (q >= rpf) q -= rpf;
  }
};

// optional device timing of tagged launches (capi.cu); work = algorithmic FLOPs or bytes of the launch
int mivit_prof_tag(const char* name);
bool mivit_prof_enabled();
void mivit_prof_begin(int tag, double work, cudaStream_t st);
void mivit_prof_end(cudaStream_t st);
struct MivitProfScope {
  cudaStream_t st;
  bool on;
  MivitProfScope(const char* name, double work, cudaStream_t s) : st(s), on(mivit_prof_enabled()) {
    if (on) mivit_prof_begin(mivit_prof_tag(name), work, st);
  }
  ~MivitProfScope() {
    if (on) mivit_prof_end(st);
  }
};
