// USV_PPOcontinuous_MLP on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// One CTA owns a tile of 128 samples.  The three Linear layers are tcgen05.mma.kind::tf32 instructions issued by one
// thread, with fp32 accumulators in tensor memory; the activations never leave the SM:
//
//   X[128x16] (obs, normalised, col 13 = 1 so b1 rides along as a weight column)
//     --UMMA 128x128x16-->  TMEM acc  --tcgen05.ld, tanh.approx--> H1[128x128] in smem (UMMA K-major operand layout)
//     --UMMA 128x128x128--> TMEM acc  --tcgen05.ld, +b2, tanh-->   H2[128x128] in smem
//     --UMMA 128x16x128-->  TMEM      --tcgen05.ld, +b3-->          (mu0, mu1, value)
//
// Operand tiles use the un-swizzled ("interleave") canonical layout: 8 x 16 B core matrices,
//     off(outer, inner) = (outer/8)*KC*32 + (inner/4)*32 + (outer%8)*4 + (inner%4)      [floats], KC = inner extent / 4.
// The same bytes are a K-major operand (outer = M/N index, inner = K) for the forward GEMMs and an MN-major operand
// (outer = K, inner = M/N) for the weight-gradient GEMMs of the training kernel, so the PyTorch-layout weights
// W[out][in] and the row-per-thread activations are staged exactly once.
//
// TF32 inputs (10-bit mantissa) + tanh.approx (2^-11) put this path at ~1e-3 relative of the fp32 kernels in ppo_mlp.cu,
// which remain the numerics reference.
#include <math_constants.h>
#include "philox.cuh"
#include "usv_common.cuh"

namespace ppotc {

constexpr int H = PPO_HIDDEN;   // 128
constexpr int TM = 128;         // samples per tile == UMMA M
constexpr int NT = 256;         // 8 warps: two warpgroups share the epilogue (64 columns each)
constexpr int DP = 16;          // padded obs dim (K of the first GEMM); column D carries the bias
constexpr float kHalfLog2Pi2 = 1.8378770664093453f;

struct Layout {
  int D, sigma, w1, b1, w2, b2, wv, bv, wmu, bmu, P;
  __host__ __device__ explicit Layout(int d) {
    D = d; sigma = 0; w1 = 2; b1 = w1 + H * d; w2 = b1 + H; b2 = w2 + H * H; wv = b2 + H; bv = wv + H; wmu = bv + 1;
    bmu = wmu + 2 * H; P = bmu + 2;
  }
};

// ---- canonical un-swizzled operand tile ------------------------------------------------------------------
__device__ __forceinline__ int tile_off(int outer, int inner, int kc) {
  return (outer >> 3) * (kc * 32) + (inner >> 2) * 32 + (outer & 7) * 4 + (inner & 3);
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (SM100 "version 1"), no swizzle.  lbo/sbo in bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                  // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE (0)
}
// instruction descriptor: D=F32, A=B=TF32; majors: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(mbar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive columns: thread i of the warp gets lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- staging ------------------------------------------------------------------------------------------------
// W[rows][cols] row-major in global -> operand tile (outer = row, inner = col), cols padded to `cols_pad`
__device__ inline void stage_matrix(float* tile, const float* __restrict__ w, int rows, int cols, int cols_pad) {
  const int kc = cols_pad >> 2;
  for (int e = threadIdx.x; e < rows * kc; e += NT) {
    const int r = e / kc, c4 = (e - r * kc) * 4;
    float4 v;
    v.x = (c4 + 0 < cols) ? w[r * cols + c4 + 0] : 0.f;
    v.y = (c4 + 1 < cols) ? w[r * cols + c4 + 1] : 0.f;
    v.z = (c4 + 2 < cols) ? w[r * cols + c4 + 2] : 0.f;
    v.w = (c4 + 3 < cols) ? w[r * cols + c4 + 3] : 0.f;
    *reinterpret_cast<float4*>(tile + tile_off(r, c4, kc)) = v;
  }
}

// issue a K-loop of tf32 UMMAs: D[128 x N] (+)= A[128 x K] * B[N x K]^T, both operands K-major tiles
__device__ __forceinline__ void gemm_kmajor(uint32_t tmem_d, const float* a_tile, int a_kc, const float* b_tile, int b_kc, int K, int N) {
  const uint32_t idesc = make_idesc(128, N, 0, 0);
  const uint32_t a0 = smem_u32(a_tile), b0 = smem_u32(b_tile);
  for (int k = 0; k < K; k += 8) {   // one tf32 UMMA consumes K = 8 (two 16-byte chunks)
    const uint64_t ad = make_desc(a0 + (k >> 2) * 128, 128, a_kc * 128);
    const uint64_t bd = make_desc(b0 + (k >> 2) * 128, 128, b_kc * 128);
    umma_tf32(tmem_d, ad, bd, idesc, k > 0 ? 1u : 0u);
  }
}

struct FwdSmem {
  float *w1, *w2, *w3, *x, *act, *b2, *b3;
  uint64_t* mbar;
  uint32_t* tmem_slot;
};
__host__ __device__ inline size_t fwd_smem_bytes() {
  return (size_t)(H * DP + H * H + 16 * H + TM * DP + TM * H + H + 16) * sizeof(float) + 64;
}
__device__ inline FwdSmem carve_fwd(float* base) {
  FwdSmem s;
  float* p = base;
  s.w2 = p; p += H * H;
  s.act = p; p += TM * H;
  s.w1 = p; p += H * DP;
  s.w3 = p; p += 16 * H;
  s.x = p; p += TM * DP;
  s.b2 = p; p += H;
  s.b3 = p; p += 16;
  s.mbar = reinterpret_cast<uint64_t*>(p); p += 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

// normalised, clamped observations of a tile -> X operand tile; column D is the constant 1 (bias column)
__device__ inline void stage_obs(float* xt, const float* __restrict__ obs, int D, const float* __restrict__ mean,
                                 const float* __restrict__ var, int64_t row0, int64_t M) {
  for (int e = threadIdx.x; e < TM * (DP / 4); e += NT) {
    const int r = e >> 2, c4 = (e & 3) * 4;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int d = c4 + q;
      float y = 0.f;
      if (d < D && row0 + r < M) {
        const float x = obs[(row0 + r) * D + d];
        y = fminf(fmaxf((x - mean[d]) / sqrtf(var[d] + 1e-5f), -5.0f), 5.0f);
      } else if (d == D) {
        y = 1.0f;
      }
      v[q] = y;
    }
    *reinterpret_cast<float4*>(xt + tile_off(r, c4, DP / 4)) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// epilogue of a hidden layer: TMEM accumulator -> (+bias) -> tanh -> activation operand tile.  256 threads: row = t%128,
// columns [64*(t/128), +64).  A warp may only touch TMEM lanes 32*(warp%4)..+31, which is exactly its 32 rows.
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_acc, float* act, const float* bias) {
  const int t = threadIdx.x, row = t & 127, half = t >> 7;
  const uint32_t lane_base = (uint32_t)((t >> 5) & 3) * 32u;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 32) {
    const int col = half * 64 + c0;
    float v[32];
    tmem_ld32(tmem_acc + (lane_base << 16) + (uint32_t)col, v);
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
      float4 o;
      o.x = tanh_fast(v[q + 0] + (bias ? bias[col + q + 0] : 0.f));
      o.y = tanh_fast(v[q + 1] + (bias ? bias[col + q + 1] : 0.f));
      o.z = tanh_fast(v[q + 2] + (bias ? bias[col + q + 2] : 0.f));
      o.w = tanh_fast(v[q + 3] + (bias ? bias[col + q + 3] : 0.f));
      *reinterpret_cast<float4*>(act + tile_off(row, col + q, H / 4)) = o;
    }
  }
}

__global__ void __launch_bounds__(NT, 1) forward_tc_kernel(
    const float* __restrict__ prm, const float* __restrict__ obs, int D, const float* __restrict__ omean,
    const float* __restrict__ ovar, const float* __restrict__ vmean, const float* __restrict__ vvar, uint64_t seed,
    uint64_t counter, int64_t row_offset, float* __restrict__ actions, float* __restrict__ neglogp,
    float* __restrict__ values, float* __restrict__ mus, float* __restrict__ sigmas, int64_t M) {
  extern __shared__ __align__(1024) float smem[];
  const Layout L(D);
  const FwdSmem s = carve_fwd(smem);
  const int t = threadIdx.x;
  // ---- one-time set-up: weights -> operand tiles, TMEM, mbarrier ----------------------------------------
  stage_matrix(s.w2, prm + L.w2, H, H, H);
  for (int e = t; e < H * (DP / 4); e += NT) {   // W1 [128][D] + bias column D, zero padded to 16
    const int r = e >> 2, c4 = (e & 3) * 4;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int d = c4 + q;
      v[q] = d < D ? prm[L.w1 + r * D + d] : (d == D ? prm[L.b1 + r] : 0.f);
    }
    *reinterpret_cast<float4*>(s.w1 + tile_off(r, c4, DP / 4)) = make_float4(v[0], v[1], v[2], v[3]);
  }
  for (int e = t; e < 16 * (H / 4); e += NT) {   // heads: rows 0,1 = mu, 2 = value, rest 0
    const int r = e / (H / 4), c4 = (e - r * (H / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < 2) v = make_float4(prm[L.wmu + r * H + c4], prm[L.wmu + r * H + c4 + 1], prm[L.wmu + r * H + c4 + 2], prm[L.wmu + r * H + c4 + 3]);
    else if (r == 2) v = make_float4(prm[L.wv + c4], prm[L.wv + c4 + 1], prm[L.wv + c4 + 2], prm[L.wv + c4 + 3]);
    *reinterpret_cast<float4*>(s.w3 + tile_off(r, c4, H / 4)) = v;
  }
  for (int e = t; e < H; e += NT) s.b2[e] = prm[L.b2 + e];
  if (t < 16) s.b3[t] = t < 2 ? prm[L.bmu + t] : (t == 2 ? prm[L.bv] : 0.f);
  if (t == 0) {
    mbar_init(s.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) tmem_alloc(s.tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t acc_h = tmem, acc_o = tmem + 128;
  const float ls0 = prm[L.sigma], ls1 = prm[L.sigma + 1];
  const float sg0 = expf(ls0), sg1 = expf(ls1);
  uint32_t phase = 0;
  const int64_t ntiles = (M + TM - 1) / TM;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    stage_obs(s.x, obs, D, omean, ovar, row0, M);
    fence_async_smem();            // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    if (t == 0) {                  // layer 1: [128 x 16] * [128 x 16]^T
      tc_fence_after();
      gemm_kmajor(acc_h, s.x, DP / 4, s.w1, DP / 4, DP, H);
      umma_commit(s.mbar);
    }
    mbar_wait(s.mbar, phase); phase ^= 1;
    tc_fence_after();
    hidden_epilogue(acc_h, s.act, nullptr);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {                  // layer 2: [128 x 128] * [128 x 128]^T
      tc_fence_after();
      gemm_kmajor(acc_h, s.act, H / 4, s.w2, H / 4, H, H);
      umma_commit(s.mbar);
    }
    mbar_wait(s.mbar, phase); phase ^= 1;
    tc_fence_after();
    hidden_epilogue(acc_h, s.act, s.b2);      // the MMA has finished reading H1: overwrite it with H2
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {                  // heads: [128 x 128] * [16 x 128]^T
      tc_fence_after();
      gemm_kmajor(acc_o, s.act, H / 4, s.w3, H / 4, H, 16);
      umma_commit(s.mbar);
    }
    mbar_wait(s.mbar, phase); phase ^= 1;
    tc_fence_after();
    if (t < 128) {
      const int64_t row = row0 + t;
      float o[16];
      tmem_ld16(acc_o + (((uint32_t)(t >> 5) * 32u) << 16), o);
      const float mu0 = o[0] + s.b3[0], mu1 = o[1] + s.b3[1], v = o[2] + s.b3[2];
      if (row < M) {
        if (mus) { mus[row * 2] = mu0; mus[row * 2 + 1] = mu1; }
        if (sigmas) { sigmas[row * 2] = sg0; sigmas[row * 2 + 1] = sg1; }
        if (values) {
          const float y = fminf(fmaxf(v, -5.0f), 5.0f);
          values[row] = vmean ? sqrtf(vvar[0] + 1e-5f) * y + vmean[0] : v;
        }
        if (actions) {
          const usv::Philox4 rr = usv::philox4x32_10((uint32_t)(row + row_offset), (uint32_t)counter, (uint32_t)(counter >> 32),
                                                     100u ^ ((uint32_t)((uint64_t)(row + row_offset) >> 32) << 8), (uint32_t)seed,
                                                     (uint32_t)(seed >> 32));
          const float u1 = (float)((rr.x >> 8) + 1u) * (1.0f / 16777216.0f);
          const float u2 = (float)(rr.y >> 8) * (1.0f / 16777216.0f);
          const float rad = sqrtf(-2.0f * logf(u1));
          float sn, cs;
          sincosf(6.28318530717958647692f * u2, &sn, &cs);
          const float a0 = mu0 + sg0 * (rad * cs), a1 = mu1 + sg1 * (rad * sn);
          actions[row * 2] = a0;
          actions[row * 2 + 1] = a1;
          if (neglogp) {
            const float d0 = (a0 - mu0) / sg0, d1 = (a1 - mu1) / sg1;
            neglogp[row] = 0.5f * (d0 * d0 + d1 * d1) + kHalfLog2Pi2 + (ls0 + ls1);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();               // everyone is done with TMEM / the operand tiles before the next tile reuses them
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 256);
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace ppotc

using namespace ppotc;

extern "C" int ppo_policy_forward_tc(const float* params, const float* obs, int32_t obs_dim, const float* obs_mean,
                                     const float* obs_var, const float* value_mean, const float* value_var, uint64_t seed,
                                     uint64_t counter, int64_t row_offset, float* actions, float* neglogp, float* values,
                                     float* mus, float* sigmas, int64_t M, void* stream) {
  if (M < 0 || obs_dim < 1 || obs_dim >= DP) return USV_E_SIZE;   // one padded column is reserved for the bias
  if (M == 0) return USV_OK;
  if (!params || !obs || !obs_mean || !obs_var) return USV_E_NULL;
  if ((value_mean == nullptr) != (value_var == nullptr)) return USV_E_NULL;
  const size_t smem = fwd_smem_bytes();
  cudaFuncSetAttribute(forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t ntiles = (M + TM - 1) / TM;
  const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
  forward_tc_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(params, obs, obs_dim, obs_mean, obs_var, value_mean, value_var, seed,
                                                              counter, row_offset, actions, neglogp, values, mus, sigmas, M);
  return usv::finish_launch();
}
