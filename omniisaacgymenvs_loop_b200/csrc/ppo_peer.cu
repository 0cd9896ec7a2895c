// One-shot all-reduce of the PPO gradient span over NVLink peer memory (the only data-path collective of the
// whole loop is the 75 KB gradient + statistics vector, once per minibatch -- latency-bound, ~256 per epoch).
//
// Every rank owns a "window" allocated by this library with cudaMalloc and exported through CUDA IPC:
//     [ 2 x cap floats : double-buffered payload | 64 x uint32 : signal pad, word r = last sequence number announced by rank r ]
// and maps the windows of its peers (cudaIpcOpenMemHandle -> P2P loads / stores over NVLink / NVSwitch).  One kernel per
// all-reduce (a few co-resident CTAs): copy the local vector into the local window, release-store the sequence number into every
// peer's pad, acquire-spin (bounded) until all peers have announced it, then sum the peers' payloads in rank order (identical bits
// on every rank).
// No NCCL call, no host in the loop, so the whole PPO update phase stays capturable in one CUDA graph on every rank.
// Double buffering by sequence parity makes the next iteration's payload write safe while a slow peer still reads this one.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "usv_common.cuh"

namespace usv {

constexpr int kPeerMax = 16;
constexpr int kPeerThreads = 256;
constexpr uint32_t kSpinLimit = 1u << 24;   // ~seconds; on expiry the kernel flags an error instead of hanging the GPU

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Pad layout (uint32 words after the two payload buffers): [0..15] sequence number last announced by rank r, [32] arrival counter
// of this rank's own CTAs.  The grid is a handful of co-resident CTAs (<= kPeerMaxCtas), one float4 per thread per trip:
//   1. every CTA copies its slice of the local vector into the local window, fences, and arrives on the local counter;
//   2. the LAST CTA to arrive release-stores the sequence number into every peer's pad (and bumps the device-side counter);
//   3. every CTA acquire-spins on the LOCAL pad until all peers have announced this sequence number (bounded);
//   4. every CTA sums its slice over the peers' windows (all remote float4 loads in flight together, rank order of the adds).
constexpr int kPeerMaxCtas = 32;

__global__ void __launch_bounds__(kPeerThreads, 1) peer_allreduce_kernel(PpoPeerComm c, const float* __restrict__ src, float* __restrict__ dst,
                                                                        int64_t count, uint32_t* __restrict__ seq_dev,
                                                                        uint32_t* __restrict__ err_flag) {
  __shared__ int s_bad;
  const uint32_t seq = *seq_dev + 1u;      // written back only after every CTA has arrived (step 2), i.e. after every CTA read it
  if (threadIdx.x == 0) s_bad = 0;
  const int64_t off = (int64_t)(seq & 1u) * c.cap;
  float* mine = c.windows[c.rank] + off;
  uint32_t* mypad = reinterpret_cast<uint32_t*>(c.windows[c.rank] + 2 * c.cap);
  const int64_t n4 = count >> 2;
  const int64_t gtid = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x, gstride = (int64_t)gridDim.x * kPeerThreads;
  // 1. local vector -> local window
  for (int64_t q = gtid; q < n4; q += gstride) reinterpret_cast<float4*>(mine)[q] = reinterpret_cast<const float4*>(src)[q];
  for (int64_t i = (n4 << 2) + gtid; i < count; i += gstride) mine[i] = src[i];
  __threadfence_system();
  __syncthreads();
  // 2. the last CTA of this rank to arrive announces: one fence, then `world` threads store to the `world` pads in parallel
  //    (a chain of release stores issued by one thread cost ~2 us per peer on the 8-GPU box)
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(mypad + 32, 1u);
    s_last = (old == gridDim.x - 1) ? 1 : 0;
    if (s_last) {
      mypad[32] = 0;
      *seq_dev = seq;
    }
  }
  __syncthreads();
  if (s_last && threadIdx.x < c.world) {
    __threadfence_system();
    st_relaxed_sys(reinterpret_cast<uint32_t*>(c.windows[threadIdx.x] + 2 * c.cap) + c.rank, seq);
  }
  // 3. wait for every peer's announcement (local memory, written remotely)
  if (threadIdx.x < c.world) {
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mypad + threadIdx.x) - seq) < 0) {
      if (++spins > kSpinLimit) { s_bad = 1; break; }
      __nanosleep(32);
    }
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0 && err_flag) atomicOr(err_flag, 1u);
    return;
  }
  // 4. reduce this CTA's slice in rank order
  for (int64_t q = gtid; q < n4; q += gstride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < c.world; r0 += 8) {     // 8 remote float4 loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < c.world) v[r] = __ldcv(reinterpret_cast<const float4*>(c.windows[r0 + r] + off) + q);
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < c.world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    }
    reinterpret_cast<float4*>(dst)[q] = acc;
  }
  for (int64_t i = (n4 << 2) + gtid; i < count; i += gstride) {
    float acc = 0.0f;
    for (int r = 0; r < c.world; ++r) acc += __ldcv(c.windows[r] + off + i);
    dst[i] = acc;
  }
}

}  // namespace usv

using namespace usv;

extern "C" {

int64_t ppo_peer_window_bytes(int64_t cap) { return (cap < 0 ? 0 : 2 * cap) * (int64_t)sizeof(float) + 64 * (int64_t)sizeof(uint32_t); }

int ppo_peer_window_alloc(int64_t cap, void** window_out, unsigned char* handle64_out) {
  if (!window_out || !handle64_out || cap <= 0) return USV_E_NULL;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)ppo_peer_window_bytes(cap));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, (size_t)ppo_peer_window_bytes(cap));
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return (int)e; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64_out, &h, 64);
  *window_out = p;
  return USV_OK;
}

int ppo_peer_window_open(const unsigned char* handle64, void** window_out) {
  if (!handle64 || !window_out) return USV_E_NULL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(window_out, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? USV_OK : (int)e;
}

int ppo_peer_window_close(void* window, int32_t owned) {
  if (!window) return USV_OK;
  cudaError_t e = owned ? cudaFree(window) : cudaIpcCloseMemHandle(window);
  return e == cudaSuccess ? USV_OK : (int)e;
}

int ppo_peer_allreduce_f32(const PpoPeerComm* c, const float* src, float* dst, int64_t count, uint32_t* seq_dev, uint32_t* err_flag,
                           void* stream) {
  if (!c || !src || !dst || !seq_dev) return USV_E_NULL;
  if (c->world < 1 || c->world > kPeerMax || c->rank < 0 || c->rank >= c->world) return USV_E_PARAM;
  if (count < 0 || count > c->cap) return USV_E_SIZE;
  for (int r = 0; r < c->world; ++r)
    if (!c->windows[r]) return USV_E_NULL;
  if ((c->cap & 3) || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return USV_E_ALIGN;
  int64_t ctas = ((count >> 2) + kPeerThreads - 1) / kPeerThreads;
  ctas = ctas < 1 ? 1 : (ctas > kPeerMaxCtas ? kPeerMaxCtas : ctas);   // all CTAs spin on flags: they must be co-resident
  peer_allreduce_kernel<<<(int)ctas, kPeerThreads, 0, (cudaStream_t)stream>>>(*c, src, dst, count, seq_dev, err_flag);
  return finish_launch();
}

}  // extern "C"
