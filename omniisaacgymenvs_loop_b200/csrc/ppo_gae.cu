// P1: GAE(lambda) reverse scan + returns, one thread per (vector of) env(s).
// [ref: RLG/common/a2c_common.py:525-540 (discount_values), :761-763 (returns = advs + values)]
// Layout (T, n) row-major: for each t a warp reads one contiguous line of envs -> fully coalesced;
// the scan runs along t in registers.  Pure streaming: 149 B read + 128 B written per env per
// rollout at T=16 -> HBM-bound.  VEC=4 envs per thread (float4 / uchar4) keeps
// >= 48 x 16 B requests in flight per thread.
#include "usv_common.cuh"

namespace usv {

template <int VEC> struct VecT;
template <> struct VecT<1> { using F = float; using B = uint8_t; };
template <> struct VecT<4> { using F = float4; using B = uchar4; };

__device__ __forceinline__ void unpack(float v, float* o) { o[0] = v; }
__device__ __forceinline__ void unpack(float4 v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack(uint8_t v, float* o) { o[0] = (float)v; }
__device__ __forceinline__ void unpack(uchar4 v, float* o) { o[0] = (float)v.x; o[1] = (float)v.y; o[2] = (float)v.z; o[3] = (float)v.w; }
__device__ __forceinline__ void pack(const float* o, float& v) { v = o[0]; }
__device__ __forceinline__ void pack(const float* o, float4& v) { v = make_float4(o[0], o[1], o[2], o[3]); }

template <int VEC>
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ dones,
                                                  const float* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, float gamma, float gamma_tau,
                                                  float* __restrict__ adv, float* __restrict__ ret, int T, int64_t n) {
  using F = typename VecT<VEC>::F;
  using B = typename VecT<VEC>::B;
  const int64_t nv = n / VEC;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  float nextv[VEC], nnt[VEC], last[VEC];
  {
    float d[VEC];
    unpack(reinterpret_cast<const F*>(last_values)[j], nextv);
    unpack(reinterpret_cast<const B*>(last_dones)[j], d);
#pragma unroll
    for (int c = 0; c < VEC; ++c) { nnt[c] = 1.0f - d[c]; last[c] = 0.0f; }
  }
#pragma unroll 4
  for (int t = T - 1; t >= 0; --t) {
    float r[VEC], v[VEC], d[VEC], a[VEC], rt[VEC];
    unpack(reinterpret_cast<const F*>(rewards + (int64_t)t * n)[j], r);
    unpack(reinterpret_cast<const F*>(values + (int64_t)t * n)[j], v);
    unpack(reinterpret_cast<const B*>(dones + (int64_t)t * n)[j], d);
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      // delta = r + gamma*nextvalues*nextnonterminal - v ; A = delta + gamma*tau*nextnonterminal*A'
      const float delta = __fsub_rn(__fadd_rn(r[c], __fmul_rn(__fmul_rn(gamma, nextv[c]), nnt[c])), v[c]);
      last[c] = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_tau, nnt[c]), last[c]));
      a[c] = last[c];
      rt[c] = __fadd_rn(last[c], v[c]);
      nextv[c] = v[c];
      nnt[c] = 1.0f - d[c];  // dones[t] is the flag ENTERING step t = "next" for step t-1
    }
    F av;
    pack(a, av);
    reinterpret_cast<F*>(adv + (int64_t)t * n)[j] = av;
    if (ret) {
      F rv;
      pack(rt, rv);
      reinterpret_cast<F*>(ret + (int64_t)t * n)[j] = rv;
    }
  }
}

// play_steps bookkeeping of one control step, fused (the eager loop spends ~40 tiny elementwise / reduction launches on it per step,
// more than the env step and the policy inference together): reward shaping, uint8 dones, per-env running return / length, the sums
// over the envs that finished, the fp64 epoch accumulators and the two windowed AverageMeters.  One CTA, fixed-order tree reduction:
// deterministic.   [ref: RLG/common/a2c_common.py:708-747 ; RLG/algos_torch/torch_ext.py:281-307]
constexpr int kBookThreads = 1024;
__device__ __forceinline__ void meter_update(float* mean, float* size, float max_size, float value_sum, float count) {
  if (count > 0.0f) {   // size = clip(count, 0, max); old = min(max - size, current); mean = (mean*old + new_mean*size) / (old + size)
    const float new_mean = value_sum / fmaxf(count, 1.0f);
    const float sz = fminf(fmaxf(count, 0.0f), max_size);
    const float old = fminf(max_size - sz, *size);
    const float tot = old + sz;
    *mean = (*mean * old + new_mean * sz) / fmaxf(tot, 1.0f);
    *size = tot;
  }
}
__global__ void __launch_bounds__(kBookThreads) rollout_bookkeep_kernel(const float* __restrict__ rew, const int64_t* __restrict__ dones,
                                                                        float scale, float* __restrict__ rewards_out,
                                                                        uint8_t* __restrict__ dones_out, float* __restrict__ cur_rew,
                                                                        float* __restrict__ cur_len, double* __restrict__ episode_acc,
                                                                        float* __restrict__ meter_r, float* __restrict__ meter_r_size,
                                                                        float* __restrict__ meter_l, float* __restrict__ meter_l_size,
                                                                        float max_size, int64_t n) {
  __shared__ float sh[3][kBookThreads / 32];
  float rs = 0.f, ls = 0.f, cnt = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += kBookThreads) {
    const float r = rew[i];
    const float d = dones[i] != 0 ? 1.0f : 0.0f;
    rewards_out[i] = r * scale;                  // DefaultRewardsShaper
    dones_out[i] = (uint8_t)(d != 0.0f);
    const float cr = cur_rew[i] + r, cl = cur_len[i] + 1.0f;
    rs += cr * d; ls += cl * d; cnt += d;
    cur_rew[i] = cr * (1.0f - d);
    cur_len[i] = cl * (1.0f - d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rs += __shfl_xor_sync(0xffffffffu, rs, o); ls += __shfl_xor_sync(0xffffffffu, ls, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = rs; sh[1][threadIdx.x >> 5] = ls; sh[2][threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x < 32) {
    rs = sh[0][threadIdx.x]; ls = sh[1][threadIdx.x]; cnt = sh[2][threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rs += __shfl_xor_sync(0xffffffffu, rs, o); ls += __shfl_xor_sync(0xffffffffu, ls, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (threadIdx.x == 0) {
      episode_acc[0] += (double)rs; episode_acc[1] += (double)ls; episode_acc[2] += (double)cnt;
      meter_update(meter_r, meter_r_size, max_size, rs, cnt);
      meter_update(meter_l, meter_l_size, max_size, ls, cnt);
    }
  }
}

}  // namespace usv

using namespace usv;

extern "C" int ppo_rollout_bookkeep_f32(const float* rew, const int64_t* dones, float scale, float* rewards_out, uint8_t* dones_out,
                                        float* cur_rew, float* cur_len, double* episode_acc, float* meter_r, float* meter_r_size,
                                        float* meter_l, float* meter_l_size, float max_size, int64_t n, void* stream) {
  if (n < 0) return USV_E_SIZE;
  if (n == 0) return USV_OK;
  if (!rew || !dones || !rewards_out || !dones_out || !cur_rew || !cur_len || !episode_acc || !meter_r || !meter_r_size || !meter_l ||
      !meter_l_size)
    return USV_E_NULL;
  rollout_bookkeep_kernel<<<1, kBookThreads, 0, (cudaStream_t)stream>>>(rew, dones, scale, rewards_out, dones_out, cur_rew, cur_len,
                                                                         episode_acc, meter_r, meter_r_size, meter_l, meter_l_size,
                                                                         max_size, n);
  return finish_launch();
}

extern "C" int ppo_gae_f32(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                           const uint8_t* last_dones, float gamma, float tau, float* advantages, float* returns,
                           int32_t T, int64_t n, void* stream) {
  if (T < 0 || n < 0) return USV_E_SIZE;
  if (T == 0 || n == 0) return USV_OK;
  if (!rewards || !values || !dones || !last_values || !last_dones || !advantages) return USV_E_NULL;
  // python computes gamma*tau in double, then it meets the fp32 tensor
  const float gamma_tau = (float)((double)gamma * (double)tau);
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec_ok = (n % 4 == 0) && !(((uintptr_t)rewards | (uintptr_t)values | (uintptr_t)last_values |
                                         (uintptr_t)advantages | (uintptr_t)returns) & 15) &&
                      !(((uintptr_t)dones | (uintptr_t)last_dones) & 3);
  if (vec_ok)
    gae_kernel<4><<<grid_for(n / 4, 256), 256, 0, s>>>(rewards, values, dones, last_values, last_dones, gamma,
                                                       gamma_tau, advantages, returns, T, n);
  else
    gae_kernel<1><<<grid_for(n, 256), 256, 0, s>>>(rewards, values, dones, last_values, last_dones, gamma, gamma_tau,
                                                   advantages, returns, T, n);
  return finish_launch();
}
