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
