// Data-parallel exchange steps of the training loop as ONE kernel each, over NVLink / NVSwitch peer memory (SURVEY.md section 8e;
// nothing of this exists in the single-process reference):
//
//   mivit_allreduce_adamw   gradient all-reduce FUSED with torch.optim.AdamW (reference Experiments/PSFNoise/
//                           trainSettingsPSFNoise.py:119): every rank reads the W gradient replicas of its parameter range
//                           straight from its peers' memory (one-shot all-reduce: 2 MB per replica, latency bound, so every
//                           rank sums everything itself -- in rank order, hence bit-identical sums and bit-identical weights on
//                           all ranks), and applies the optimizer update in the same pass.  No NCCL call, no host round trip:
//                           the kernel is an ordinary graph node, so a whole data-parallel step replays as one CUDA graph.
//   mivit_allreduce_small   SUM all-reduce of <= 512 floats (synchronised-BatchNorm statistics, 12 per step), one CTA.
//
// Every rank owns a SEGMENT (cudaMalloc'ed here, exported with cudaIpcGetMemHandle, mapped by the peers):
//   header: flags + counters | small-exchange slots | the flat gradient buffer of the model.
// Protocol of one exchange with sequence number c (kept in the segment, advanced by the kernel itself, so graph replays need
// no host-provided value):
//   start:  rank r stores c into ready[r] of every peer (release, system scope) -- its gradients were written by earlier kernels
//           of the same stream -- and every CTA spins (acquire) until all W entries of its own ready[] hold c;
//   body :  volatile 16-byte loads of the peers' replicas (peer memory is never cached in the reader's L1);
//   end  :  the LAST CTA to finish (grid-wide completion counter) stores c into done[r] of every peer and waits until its own
//           done[] holds c everywhere: the kernel does not retire -- and the next backward does not overwrite this rank's
//           gradients -- before every peer has finished reading them.
// Spins are bounded by %globaltimer (20 s): a missing rank traps instead of hanging the GPU.
#include <string.h>

#include "common.cuh"
#include "vit.h"
#include "../../include/mivit.h"

namespace {

constexpr int kMaxRanks = 8;
constexpr int kBuckets = 4;        // independent gradient exchanges in flight (the trainer uses 2)
constexpr int kSmallCalls = 32;    // distinct small exchanges per step (synchronised BatchNorm: 12)
constexpr int kSmallFloats = 512;

struct SegHeader {
  unsigned long long ready[kBuckets][kMaxRanks];
  unsigned long long done[kBuckets][kMaxRanks];
  unsigned long long seq[kBuckets];
  unsigned int finished[kBuckets];
  unsigned int pad0[kBuckets];
  unsigned long long small_ready[kSmallCalls][kMaxRanks];
  unsigned long long small_seq[kSmallCalls];
  long long step;            // AdamW step count (1-based after the first update)
  float lr;
  float pad1[3];
};
constexpr size_t kHeaderBytes = (sizeof(SegHeader) + 1023) / 1024 * 1024;
constexpr size_t kSmallBytes = (size_t)kSmallCalls * 2 * kSmallFloats * sizeof(float);   // two parities per call (see below)

struct Peers {
  uint8_t* seg[kMaxRanks];
  int rank, world;
};

__device__ __forceinline__ SegHeader* hdr(uint8_t* seg) { return reinterpret_cast<SegHeader*>(seg); }
__device__ __forceinline__ float* small_slot(uint8_t* seg, int call, unsigned long long c) {
  return reinterpret_cast<float*>(seg + kHeaderBytes) + ((size_t)call * 2 + (size_t)(c & 1ull)) * kSmallFloats;
}
__device__ __forceinline__ float* grad_of(uint8_t* seg) { return reinterpret_cast<float*>(seg + kHeaderBytes + kSmallBytes); }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_all(const unsigned long long* flags, int world, unsigned long long c) {
  const unsigned long long t0 = globaltimer();
  for (int r = 0; r < world; ++r) {
    while (ld_acquire_sys(flags + r) < c) {
      if (globaltimer() - t0 > 20000000000ull) __trap();   // 20 s: a rank is missing
      __nanosleep(64);
    }
  }
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// ---- gradient all-reduce fused with AdamW ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) allreduce_adamw_kernel(Peers pr, int bucket, long long lo, long long hi, float* __restrict__ p,
                                                              float* __restrict__ m, float* __restrict__ v, float b1, float b2,
                                                              float eps, float wd, int advance_step, float* __restrict__ gsum) {
  uint8_t* mine = pr.seg[pr.rank];
  SegHeader* h = hdr(mine);
  __shared__ unsigned long long c_s;
  __shared__ float bc_s[3];
  if (threadIdx.x == 0) {
    const unsigned long long c = h->seq[bucket] + 1;          // unchanged until the last CTA of THIS launch advances it
    c_s = c;
    if (blockIdx.x == 0) {
      __threadfence_system();
      for (int r = 0; r < pr.world; ++r) st_release_sys(&hdr(pr.seg[r])->ready[bucket][pr.rank], c);
    }
    wait_all(h->ready[bucket], pr.world, c);
    const double step = (double)(h->step + 1);
    const float lr = h->lr;
    bc_s[0] = lr;
    bc_s[1] = (float)((double)lr / (1.0 - pow((double)b1, step)));           // lr / bias_correction1
    bc_s[2] = (float)(1.0 / sqrt(1.0 - pow((double)b2, step)));              // 1 / sqrt(bias_correction2)
  }
  __syncthreads();
  const float lr = bc_s[0], step_size = bc_s[1], inv_sqrt_bc2 = bc_s[2];
  const float scale = 1.0f / (float)pr.world;
  const long long n4 = (hi - lo) >> 2;                         // the launcher aligns lo / hi to 4 floats
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long o = lo + 4 * i;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < pr.world; ++r) {                       // rank order: identical sums on every rank
      const float4 t = ld_volatile4(grad_of(pr.seg[r]) + o);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    if (gsum != nullptr) *reinterpret_cast<float4*>(gsum + o) = g;
    float4 pp = *reinterpret_cast<float4*>(p + o), mm = *reinterpret_cast<float4*>(m + o), vv = *reinterpret_cast<float4*>(v + o);
    float* pa = &pp.x; const float* ga = &g.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {                              // same arithmetic as optim.cu adamw_kernel
      const float gr = ga[k] * scale;
      pa[k] *= (1.0f - lr * wd);
      ma[k] = ma[k] + (1.0f - b1) * (gr - ma[k]);
      va[k] = b2 * va[k] + (1.0f - b2) * gr * gr;
      pa[k] -= step_size * (ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps));
    }
    *reinterpret_cast<float4*>(p + o) = pp;
    *reinterpret_cast<float4*>(m + o) = mm;
    *reinterpret_cast<float4*>(v + o) = vv;
  }
  // ---- end barrier by the last CTA
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&h->finished[bucket], 1u);
    if (prev == gridDim.x - 1) {
      const unsigned long long c = c_s;
      for (int r = 0; r < pr.world; ++r) st_release_sys(&hdr(pr.seg[r])->done[bucket][pr.rank], c);
      wait_all(h->done[bucket], pr.world, c);
      h->finished[bucket] = 0;
      h->seq[bucket] = c;
      if (advance_step) h->step += 1;
      __threadfence();
    }
  }
}

// ---- small SUM all-reduce (synchronised BatchNorm statistics) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) allreduce_small_kernel(Peers pr, int call, float* __restrict__ buf, int n) {
  uint8_t* mine = pr.seg[pr.rank];
  SegHeader* h = hdr(mine);
  const unsigned long long c = h->small_seq[call] + 1;       // every thread reads it; thread 0 advances it after the last barrier
  // Slots alternate with the parity of c and there is no end barrier: exchange c + 2 of this call reuses the slot of c, and a
  // rank can only start it after every peer ARRIVED at exchange c + 1, i.e. after each of them finished reading exchange c.
  float* slot = small_slot(mine, call, c);
  for (int i = threadIdx.x; i < n; i += blockDim.x) slot[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int r = 0; r < pr.world; ++r) st_release_sys(&hdr(pr.seg[r])->small_ready[call][pr.rank], c);
    wait_all(h->small_ready[call], pr.world, c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < pr.world; ++r) {                       // rank order: identical sums on every rank
      float t;
      asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(t) : "l"(small_slot(pr.seg[r], call, c) + i) : "memory");
      s += t;
    }
    buf[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) h->small_seq[call] = c;
}

Peers g_bn_peers;
bool g_bn_peers_on = false;
int g_bn_call = 0;

int to_peers(const mivit_peer_comm* c, Peers& pr) {
  MIVIT_CHECK_ARG(c != nullptr && c->world >= 1 && c->world <= kMaxRanks && c->rank >= 0 && c->rank < c->world, "bad peer communicator");
  for (int r = 0; r < kMaxRanks; ++r) pr.seg[r] = r < c->world ? reinterpret_cast<uint8_t*>(c->segment[r]) : nullptr;
  for (int r = 0; r < c->world; ++r) MIVIT_CHECK_ARG(pr.seg[r] != nullptr, "peer segment %d is not mapped", r);
  pr.rank = c->rank;
  pr.world = c->world;
  return MIVIT_OK;
}

}  // namespace

// synchronised BatchNorm through the peer segments (vit_model.cu calls these in place of the host all-reduce hook)
bool mivit_bn_peer_active() { return g_bn_peers_on; }
int mivit_bn_peer_world() { return g_bn_peers_on ? g_bn_peers.world : 1; }
void mivit_bn_peer_begin_step() { g_bn_call = 0; }
int mivit_bn_peer_sync(float* buf, long long n, cudaStream_t st) {
  MIVIT_CHECK_ARG(n >= 1 && n <= kSmallFloats, "small all-reduce of %lld floats (max %d)", n, kSmallFloats);
  MIVIT_CHECK_ARG(g_bn_call < kSmallCalls, "more than %d small all-reduces in one step", kSmallCalls);
  allreduce_small_kernel<<<1, 256, 0, st>>>(g_bn_peers, g_bn_call++, buf, (int)n);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int64_t mivit_comm_segment_bytes(int64_t n_grad_floats) {
  if (n_grad_floats < 0) return -1;
  return (int64_t)(kHeaderBytes + kSmallBytes + ((size_t)n_grad_floats + 63) / 64 * 64 * sizeof(float));
}
extern "C" int64_t mivit_comm_grad_offset_bytes(void) { return (int64_t)(kHeaderBytes + kSmallBytes); }

extern "C" int mivit_comm_alloc(int64_t bytes, void** segment) {
  MIVIT_CHECK_ARG(segment != nullptr && bytes >= (int64_t)(kHeaderBytes + kSmallBytes), "bad segment size");
  void* p = nullptr;
  MIVIT_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  MIVIT_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  MIVIT_CUDA_CHECK(cudaDeviceSynchronize());
  *segment = p;
  return MIVIT_OK;
}
extern "C" int mivit_comm_free(void* segment) {
  if (segment != nullptr) MIVIT_CUDA_CHECK(cudaFree(segment));
  return MIVIT_OK;
}
extern "C" int mivit_comm_ipc_handle(void* segment, uint8_t* handle64) {
  MIVIT_CHECK_ARG(segment && handle64, "NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  MIVIT_CUDA_CHECK(cudaIpcGetMemHandle(&h, segment));
  memcpy(handle64, &h, 64);
  return MIVIT_OK;
}
extern "C" int mivit_comm_ipc_open(const uint8_t* handle64, void** segment) {
  MIVIT_CHECK_ARG(segment && handle64, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  MIVIT_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *segment = p;
  return MIVIT_OK;
}
extern "C" int mivit_comm_ipc_close(void* segment) {
  if (segment != nullptr) MIVIT_CUDA_CHECK(cudaIpcCloseMemHandle(segment));
  return MIVIT_OK;
}

extern "C" int mivit_comm_set_lr(const mivit_peer_comm* c, float lr, void* stream) {
  Peers pr;
  int rc = to_peers(c, pr);
  if (rc) return rc;
  // 4-byte host -> device copy ordered with the stream (pageable source: staged before the call returns).  The value lives in
  // the segment so that a replayed graph picks it up.
  MIVIT_CUDA_CHECK(cudaMemcpyAsync(pr.seg[pr.rank] + offsetof(SegHeader, lr), &lr, sizeof(float), cudaMemcpyHostToDevice,
                                   (cudaStream_t)stream));
  return MIVIT_OK;
}
extern "C" int mivit_comm_set_step(const mivit_peer_comm* c, int64_t step, void* stream) {
  Peers pr;
  int rc = to_peers(c, pr);
  if (rc) return rc;
  const long long v = (long long)step;
  MIVIT_CUDA_CHECK(cudaMemcpyAsync(pr.seg[pr.rank] + offsetof(SegHeader, step), &v, sizeof(long long), cudaMemcpyHostToDevice,
                                   (cudaStream_t)stream));
  return MIVIT_OK;
}

extern "C" int mivit_allreduce_adamw(const mivit_peer_comm* c, int32_t bucket, int64_t lo, int64_t hi, float* p, float* m, float* v,
                                     float beta1, float beta2, float eps, float weight_decay, int32_t advance_step, float* grad_sum,
                                     int32_t max_ctas, void* stream) {
  Peers pr;
  int rc = to_peers(c, pr);
  if (rc) return rc;
  MIVIT_CHECK_ARG(bucket >= 0 && bucket < kBuckets, "bucket out of range");
  MIVIT_CHECK_ARG(p && m && v && lo >= 0 && hi >= lo && lo % 4 == 0, "bad range / NULL pointer");
  MIVIT_CHECK_ARG((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad_sum) & 15) == 0, "buffers must be 16-byte aligned");
  const long long hi4 = (hi + 3) / 4 * 4;     // the flat buffers are padded by >= 4 floats; the pad's gradient is zero
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long blocks = ((hi4 - lo) / 4 + 255) / 256;
  if (blocks > sms) blocks = sms;             // at most one CTA per SM: every CTA spins on the start barrier
  if (max_ctas > 0 && blocks > max_ctas) blocks = max_ctas;   // an exchange overlapped with compute leaves the SMs to the compute
  if (blocks < 1) blocks = 1;
  MivitProfScope prof("allreduce_adamw", (double)(hi4 - lo) * 4.0 * (pr.world + 6), (cudaStream_t)stream);
  allreduce_adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pr, bucket, lo, hi4, p, m, v, beta1, beta2, eps,
                                                                             weight_decay, advance_step, grad_sum);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_allreduce_small(const mivit_peer_comm* c, int32_t call, float* buf, int32_t n, void* stream) {
  Peers pr;
  int rc = to_peers(c, pr);
  if (rc) return rc;
  MIVIT_CHECK_ARG(call >= 0 && call < kSmallCalls && buf && n >= 1 && n <= kSmallFloats, "bad small all-reduce arguments");
  allreduce_small_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pr, call, buf, n);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

// synchronised BatchNorm inside the ViT forward / backward goes through the peer segments while a communicator is set (the
// collectives are then plain kernels: no host callback, capturable into the step's CUDA graph); NULL unsets
extern "C" int mivit_set_bn_sync_comm(const mivit_peer_comm* c) {
  if (c == nullptr) {
    g_bn_peers_on = false;
    return MIVIT_OK;
  }
  int rc = to_peers(c, g_bn_peers);
  if (rc) return rc;
  g_bn_peers_on = g_bn_peers.world > 1;
  return MIVIT_OK;
}
