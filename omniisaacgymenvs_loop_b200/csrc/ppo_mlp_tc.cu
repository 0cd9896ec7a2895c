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
#include <cooperative_groups.h>
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


// ---- packed weights ---------------------------------------------------------------------------------------------
// The four weight operand tiles live pre-arranged ("packed") in global memory in exactly the byte order of the shared-
// memory tiles, so a CTA stages all of them with a few TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued by ONE
// thread instead of ~50 dependent LDG/STS rounds per thread (the first version of these kernels spent half its time there).
//   [ W2 tile 128x128 | W2^T tile 128x128 | W1 tile 128x16 (col D = b1) | heads tile 16x128 | b2[128] | b3[16] ]
constexpr int PK_W2 = 0, PK_W2T = PK_W2 + H * H, PK_W1 = PK_W2T + H * H, PK_W3 = PK_W1 + H * DP, PK_B2 = PK_W3 + 16 * H,
              PK_B3 = PK_B2 + H, PK_TOTAL = PK_B3 + 16;

__device__ __forceinline__ void tile_inv(int e, int kc, int& outer, int& inner) {
  const int og = e / (kc * 32), rem = e - og * (kc * 32);
  outer = og * 8 + ((rem & 31) >> 2);
  inner = (rem >> 5) * 4 + (rem & 3);
}
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ prm, int D, float* __restrict__ pk) {
  const Layout L(D);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= PK_TOTAL) return;
  int o, i;
  float v = 0.f;
  if (e < PK_W2T) { tile_inv(e - PK_W2, H / 4, o, i); v = prm[L.w2 + o * H + i]; }
  else if (e < PK_W1) { tile_inv(e - PK_W2T, H / 4, o, i); v = prm[L.w2 + i * H + o]; }          // outer = in-feature, inner = out-feature
  else if (e < PK_W3) { tile_inv(e - PK_W1, DP / 4, o, i); v = i < D ? prm[L.w1 + o * D + i] : (i == D ? prm[L.b1 + o] : 0.f); }
  else if (e < PK_B2) { tile_inv(e - PK_W3, H / 4, o, i); v = o < 2 ? prm[L.wmu + o * H + i] : (o == 2 ? prm[L.wv + i] : 0.f); }
  else if (e < PK_B3) { v = prm[L.b2 + (e - PK_B2)]; }
  else { const int j = e - PK_B3; v = j < 2 ? prm[L.bmu + j] : (j == 2 ? prm[L.bv] : 0.f); }
  pk[e] = v;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

struct FwdSmem {
  float *w1, *w2, *w3, *x, *act, *b2, *b3;
  uint64_t* mbar;
  uint32_t* tmem_slot;
};
__host__ __device__ inline size_t fwd_smem_bytes() {
  return (size_t)(H * DP + H * H + 16 * H + TM * DP + TM * H + H + 16) * sizeof(float) + 96;
}
__device__ inline FwdSmem carve_fwd(float* base) {
  FwdSmem s;
  float* p = base;
  s.w2 = p; p += H * H;
  s.w1 = p; p += H * DP;          // w1 | w3 | b2 | b3 contiguous: one bulk copy from the packed buffer
  s.w3 = p; p += 16 * H;
  s.b2 = p; p += H;
  s.b3 = p; p += 16;
  s.act = p; p += TM * H;
  s.x = p; p += TM * DP;
  s.mbar = reinterpret_cast<uint64_t*>(p); p += 4;   // [0] MMA completion, [1] weight bulk copies
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
    const float* __restrict__ prm, const float* __restrict__ pk, const float* __restrict__ obs, int D, const float* __restrict__ omean,
    const float* __restrict__ ovar, const float* __restrict__ vmean, const float* __restrict__ vvar, uint64_t seed,
    uint64_t counter_in, const uint64_t* __restrict__ counter_offset, int64_t row_offset, float* __restrict__ actions, float* __restrict__ neglogp,
    float* __restrict__ values, float* __restrict__ mus, float* __restrict__ sigmas, int64_t M) {
  extern __shared__ __align__(1024) float smem[];
  const Layout L(D);
  const FwdSmem s = carve_fwd(smem);
  const int t = threadIdx.x;
  const uint64_t counter = counter_in + (counter_offset ? *counter_offset : 0ull);
  // ---- one-time set-up: packed weight tiles by TMA bulk copy, TMEM, mbarriers ---------------------------------
  if (t == 0) {
    mbar_init(s.mbar, 1);
    mbar_init(s.mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t b_w2 = H * H * 4, b_rest = (H * DP + 16 * H + H + 16) * 4;
    mbar_expect_tx(s.mbar + 1, b_w2 + b_rest);
    bulk_g2s(s.w2, pk + PK_W2, b_w2, s.mbar + 1);
    bulk_g2s(s.w1, pk + PK_W1, b_rest, s.mbar + 1);
  }
  if (t < 32) tmem_alloc(s.tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  mbar_wait(s.mbar + 1, 0);        // weights have landed (async proxy writes: visible to the MMA without a proxy fence)
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


// =====================================================================================================================
// Training on the tensor cores: two kernels per minibatch.
//
// tcgen05.mma.kind::tf32 only takes K-major shared-memory operands (the transpose bits of the instruction descriptor
// return zeros for tf32 -- scripts/probes/tc_probe.cu), and the weight gradients contract over the SAMPLE index, so their
// operands must be sample-contiguous while the forward / dH operands are feature-contiguous.  Instead of transposing
// 64 KB tiles inside one CTA (shared memory cannot hold both orientations of H1, H2, dz2 plus W2 and W2^T), the work is
// split where the orientation changes:
//
//  T1 train_fwd_bwd_tc_kernel (one CTA per 128-sample tile):
//       forward (3 UMMA GEMMs) -> per-sample PPO losses -> dz3 -> dz2 = (dz3.W3)(1-H2^2) -> dH1 = dz2.W2 (UMMA, W2^T tile)
//       -> dz1 = dH1 (1-H1^2).  H1, H2, dz2, dz1, X, dz3 are written to global FEATURE-major ([feature][sample]): each
//       store instruction of a warp covers 32 consecutive samples of one feature, i.e. one full 128 B line.
//  T2 wgrad_tc_kernel (one CTA per 64-sample chunk): loads the feature-major slabs straight into K-major operand tiles
//       (float4 along the sample axis, no transposition) and runs dW2 = dz2^T.H1, [dW1|db1] = dz1^T.X, dW3^T = H2^T.dz3,
//       db2 = dz2^T.1 as UMMA GEMMs with K = samples; accumulators in TMEM; partial gradients per CTA.
//  The slabs (4 x 4 MB at the reference's 8192-sample minibatch) stay in the 126 MB L2 between the two kernels.
struct LossInTc {
  const float *actions, *old_nlp, *adv, *old_v, *ret;
  float *old_mu, *old_sigma;
};
struct TrainWs {   // feature-major workspaces, leading dimension ld (samples, multiple of 128)
  float *h1t, *h2t, *dz2t, *dz1t, *xt, *dz3t;
  int64_t ld;
};

struct TrainSmem {
  float *w1, *w2, *w2t, *w3, *x, *act, *dz3, *b2, *b3, *red;
  uint64_t* mbar;
  uint32_t* tmem_slot;
};
__host__ __device__ inline size_t train_smem_bytes() {
  return (size_t)(H * DP + 2 * H * H + 16 * H + TM * DP + TM * H + TM * 4 + H + 16 + 64) * sizeof(float) + 96;
}
__device__ inline TrainSmem carve_train(float* base) {
  TrainSmem s;
  float* p = base;
  s.w2 = p; p += H * H;           // w2 | w2t | w1 | w3 | b2 | b3: the packed-buffer order
  s.w2t = p; p += H * H;
  s.w1 = p; p += H * DP;
  s.w3 = p; p += 16 * H;
  s.b2 = p; p += H;
  s.b3 = p; p += 16;
  s.act = p; p += TM * H;
  s.x = p; p += TM * DP;
  s.dz3 = p; p += TM * 4;
  s.red = p; p += 64;
  s.mbar = reinterpret_cast<uint64_t*>(p); p += 4;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

// hidden-layer epilogue that also stores the activation feature-major to global (ht[col][sample])
__device__ __forceinline__ void hidden_epilogue_store(uint32_t tmem_acc, float* act, const float* bias, float* __restrict__ ht,
                                                     int64_t ld, int64_t sample) {
  const int t = threadIdx.x, row = t & 127, half = t >> 7;
  const uint32_t lane_base = (uint32_t)((t >> 5) & 3) * 32u;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 32) {
    const int col = half * 64 + c0;
    float v[32];
    tmem_ld32(tmem_acc + (lane_base << 16) + (uint32_t)col, v);
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = tanh_fast(v[q] + (bias ? bias[col + q] : 0.f));
#pragma unroll
    for (int q = 0; q < 32; q += 4)
      *reinterpret_cast<float4*>(act + tile_off(row, col + q, H / 4)) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
#pragma unroll
    for (int q = 0; q < 32; ++q) ht[(int64_t)(col + q) * ld + sample] = v[q];   // warp: 32 consecutive samples of one feature
  }
}

__global__ void __launch_bounds__(NT, 1) train_fwd_bwd_tc_kernel(const float* __restrict__ prm, const float* __restrict__ pk,
                                                                const float* __restrict__ obs, int D,
                                                                const float* __restrict__ omean, const float* __restrict__ ovar,
                                                                LossInTc in, PpoLossParams lp, TrainWs ws, float* __restrict__ partial,
                                                                int64_t M) {
  extern __shared__ __align__(1024) float smem[];
  const Layout L(D);
  const TrainSmem s = carve_train(smem);
  const int t = threadIdx.x, row = t & 127, half = t >> 7;
  const uint32_t lane_base = (uint32_t)((t >> 5) & 3) * 32u;
  // ---- one-time set-up: packed weight tiles by TMA bulk copy, TMEM, mbarriers ---------------------------------
  if (t == 0) {
    mbar_init(s.mbar, 1);
    mbar_init(s.mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t b_w = H * H * 4, b_rest = (H * DP + 16 * H + H + 16) * 4;
    mbar_expect_tx(s.mbar + 1, 2 * b_w + b_rest);
    bulk_g2s(s.w2, pk + PK_W2, b_w, s.mbar + 1);
    bulk_g2s(s.w2t, pk + PK_W2T, b_w, s.mbar + 1);
    bulk_g2s(s.w1, pk + PK_W1, b_rest, s.mbar + 1);
  }
  if (t < 32) tmem_alloc(s.tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  mbar_wait(s.mbar + 1, 0);
  const uint32_t acc_h = tmem, acc_o = tmem + 128;
  const float ls0 = prm[L.sigma], ls1 = prm[L.sigma + 1];
  const float sg0 = expf(ls0), sg1 = expf(ls1);
  const float invM = 1.0f / (float)M;
  float st_a = 0.f, st_c = 0.f, st_e = 0.f, st_b = 0.f, st_kl = 0.f, g_ls0 = 0.f, g_ls1 = 0.f, g_b3[3] = {0.f, 0.f, 0.f};
  uint32_t phase = 0;
  const int64_t ntiles = (M + TM - 1) / TM;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    const int64_t sample = row0 + row;          // < ws.ld always (ld is padded to a multiple of 128)
    // ---------------- forward ----------------
    stage_obs(s.x, obs, D, omean, ovar, row0, M);
    __syncthreads();
    if (t < 128) {                              // X^T to global for the weight-gradient kernel (zero rows beyond M)
#pragma unroll
      for (int d = 0; d < DP; ++d) ws.xt[(int64_t)d * ws.ld + sample] = (row0 + t < M) ? s.x[tile_off(t, d, DP / 4)] : 0.f;
    }
    fence_async_smem(); tc_fence_before(); __syncthreads();
    if (t == 0) { tc_fence_after(); gemm_kmajor(acc_h, s.x, DP / 4, s.w1, DP / 4, DP, H); umma_commit(s.mbar); }
    mbar_wait(s.mbar, phase); phase ^= 1; tc_fence_after();
    hidden_epilogue_store(acc_h, s.act, nullptr, ws.h1t, ws.ld, sample);
    fence_async_smem(); tc_fence_before(); __syncthreads();
    if (t == 0) { tc_fence_after(); gemm_kmajor(acc_h, s.act, H / 4, s.w2, H / 4, H, H); umma_commit(s.mbar); }
    mbar_wait(s.mbar, phase); phase ^= 1; tc_fence_after();
    hidden_epilogue_store(acc_h, s.act, s.b2, ws.h2t, ws.ld, sample);
    fence_async_smem(); tc_fence_before(); __syncthreads();
    if (t == 0) { tc_fence_after(); gemm_kmajor(acc_o, s.act, H / 4, s.w3, H / 4, H, 16); umma_commit(s.mbar); }
    mbar_wait(s.mbar, phase); phase ^= 1; tc_fence_after();
    // ---------------- per-sample losses -> dz3 (threads 0..127, one row each) ----------------
    if (t < 128) {
      const int64_t r = row0 + t;
      float o[16];
      tmem_ld16(acc_o + (lane_base << 16), o);
      const float mu0 = o[0] + s.b3[0], mu1 = o[1] + s.b3[1], v = o[2] + s.b3[2];
      float d_mu0 = 0.f, d_mu1 = 0.f, d_v = 0.f;
      if (r < M) {
        const float a0 = in.actions[r * 2], a1 = in.actions[r * 2 + 1];
        const float e0 = (a0 - mu0) / sg0, e1 = (a1 - mu1) / sg1;
        const float nlp = 0.5f * (e0 * e0 + e1 * e1) + kHalfLog2Pi2 + (ls0 + ls1);
        const float adv = in.adv[r];
        const float ratio = expf(in.old_nlp[r] - nlp);
        const float s1 = adv * ratio;
        const float s2 = adv * fminf(fmaxf(ratio, 1.0f - lp.e_clip), 1.0f + lp.e_clip);
        const float a_loss = fmaxf(-s1, -s2);
        const float g_nlp = (-s1 >= -s2) ? s1 : 0.f;
        const float ov = in.old_v[r], ret = in.ret[r];
        float c_loss, g_v;
        if (lp.clip_value) {
          const float dvc = fminf(fmaxf(v - ov, -lp.e_clip), lp.e_clip);
          const float vpc = ov + dvc;
          const float l1 = (v - ret) * (v - ret), l2 = (vpc - ret) * (vpc - ret);
          c_loss = fmaxf(l1, l2);
          const bool inside = fabsf(v - ov) <= lp.e_clip;
          g_v = (l1 >= l2) ? 2.0f * (v - ret) : (inside ? 2.0f * (vpc - ret) : 0.f);
        } else {
          c_loss = (ret - v) * (ret - v);
          g_v = 2.0f * (v - ret);
        }
        const float h0 = fmaxf(mu0 - lp.bound_soft, 0.f), l0 = fminf(mu0 + lp.bound_soft, 0.f);
        const float h1v = fmaxf(mu1 - lp.bound_soft, 0.f), l1v = fminf(mu1 + lp.bound_soft, 0.f);
        const float b_loss = (l0 * l0 + h0 * h0) + (l1v * l1v + h1v * h1v);
        const float ent = (0.5f + 0.5f * 1.8378770664093453f + ls0) + (0.5f + 0.5f * 1.8378770664093453f + ls1);
        const float om0 = in.old_mu[r * 2], om1 = in.old_mu[r * 2 + 1];
        const float os0 = in.old_sigma[r * 2], os1 = in.old_sigma[r * 2 + 1];
        const float kl0 = logf(os0 / sg0 + 1e-5f) + (sg0 * sg0 + (om0 - mu0) * (om0 - mu0)) / (2.0f * (os0 * os0 + 1e-5f)) - 0.5f;
        const float kl1 = logf(os1 / sg1 + 1e-5f) + (sg1 * sg1 + (om1 - mu1) * (om1 - mu1)) / (2.0f * (os1 * os1 + 1e-5f)) - 0.5f;
        st_a += a_loss; st_c += c_loss; st_e += ent; st_b += b_loss; st_kl += kl0 + kl1;
        const float wa = invM, wc = 0.5f * lp.critic_coef * invM, wb = lp.bounds_loss_coef * invM, we = lp.entropy_coef * invM;
        d_mu0 = wa * g_nlp * (-(a0 - mu0) / (sg0 * sg0)) + wb * 2.0f * (h0 + l0);
        d_mu1 = wa * g_nlp * (-(a1 - mu1) / (sg1 * sg1)) + wb * 2.0f * (h1v + l1v);
        d_v = wc * g_v;
        g_ls0 += wa * g_nlp * (1.0f - e0 * e0) - we;
        g_ls1 += wa * g_nlp * (1.0f - e1 * e1) - we;
        g_b3[0] += d_mu0; g_b3[1] += d_mu1; g_b3[2] += d_v;
        in.old_mu[r * 2] = mu0; in.old_mu[r * 2 + 1] = mu1;
        in.old_sigma[r * 2] = sg0; in.old_sigma[r * 2 + 1] = sg1;
      }
      *reinterpret_cast<float4*>(s.dz3 + t * 4) = make_float4(d_mu0, d_mu1, d_v, 0.f);
      ws.dz3t[0 * ws.ld + sample] = d_mu0;
      ws.dz3t[1 * ws.ld + sample] = d_mu1;
      ws.dz3t[2 * ws.ld + sample] = d_v;
    }
    __syncthreads();
    // ---------------- dz2 = (dz3 . W3) * (1 - H2^2), in place over H2 (operand tile) + feature-major to global -------------
    {
      const float4 g = *reinterpret_cast<const float4*>(s.dz3 + row * 4);
#pragma unroll 4
      for (int c = half * 64; c < half * 64 + 64; c += 4) {
        float* hp = s.act + tile_off(row, c, H / 4);
        const float4 h = *reinterpret_cast<const float4*>(hp);
        const float4 wa = *reinterpret_cast<const float4*>(s.w3 + tile_off(0, c, H / 4));
        const float4 wb = *reinterpret_cast<const float4*>(s.w3 + tile_off(1, c, H / 4));
        const float4 wc = *reinterpret_cast<const float4*>(s.w3 + tile_off(2, c, H / 4));
        float4 o;
        o.x = (g.x * wa.x + g.y * wb.x + g.z * wc.x) * (1.0f - h.x * h.x);
        o.y = (g.x * wa.y + g.y * wb.y + g.z * wc.y) * (1.0f - h.y * h.y);
        o.z = (g.x * wa.z + g.y * wb.z + g.z * wc.z) * (1.0f - h.z * h.z);
        o.w = (g.x * wa.w + g.y * wb.w + g.z * wc.w) * (1.0f - h.w * h.w);
        *reinterpret_cast<float4*>(hp) = o;
        ws.dz2t[(int64_t)(c + 0) * ws.ld + sample] = o.x;
        ws.dz2t[(int64_t)(c + 1) * ws.ld + sample] = o.y;
        ws.dz2t[(int64_t)(c + 2) * ws.ld + sample] = o.z;
        ws.dz2t[(int64_t)(c + 3) * ws.ld + sample] = o.w;
      }
    }
    // ---------------- dH1 = dz2 . W2   (A = dz2 tile, B = W2^T tile, both K-major) ----------------
    fence_async_smem(); tc_fence_before(); __syncthreads();
    if (t == 0) { tc_fence_after(); gemm_kmajor(acc_h, s.act, H / 4, s.w2t, H / 4, H, H); umma_commit(s.mbar); }
    mbar_wait(s.mbar, phase); phase ^= 1; tc_fence_after();
    // ---------------- dz1 = dH1 * (1 - H1^2): H1 re-read feature-major from global (L2), dz1 stored feature-major ----------
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      const int col = half * 64 + c0;
      float v[32];
      tmem_ld32(acc_h + (lane_base << 16) + (uint32_t)col, v);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float h = ws.h1t[(int64_t)(col + q) * ws.ld + sample];
        ws.dz1t[(int64_t)(col + q) * ws.ld + sample] = v[q] * (1.0f - h * h);
      }
    }
    tc_fence_before();
    __syncthreads();   // X / ACT / dz3 and the TMEM accumulators are reused by the next tile
  }
  // ---- this CTA's share of the scalar-parameter gradients and the statistics: compact 16-float slot --------------
  float* out = partial + (size_t)blockIdx.x * 16;
  float vals[10] = {st_a, st_c, st_e, st_b, st_kl, g_ls0, g_ls1, g_b3[0], g_b3[1], g_b3[2]};
  float* red = s.red;   // [10][4] (threads 0..127 hold the data: 4 warps)
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    float v = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((t & 31) == 0 && t < 128) red[q * 4 + (t >> 5)] = v;
  }
  __syncthreads();
  if (t == 0) {   // slot: [a, c, ent, b, kl (means)] [dlogstd0, dlogstd1, dbmu0, dbmu1, dbv]
    for (int q = 0; q < 10; ++q) {
      const float v = (red[q * 4] + red[q * 4 + 1]) + (red[q * 4 + 2] + red[q * 4 + 3]);
      out[q] = q < 5 ? v * invM : v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 256);
}

// second stage for the tensor-core path: matrix gradients from the wgrad slots, scalar gradients + statistics from T1's slots
__global__ void __launch_bounds__(256) reduce_tc_kernel(const float* __restrict__ scal, int g1, const float* __restrict__ mat, int g2,
                                                       int D, float* __restrict__ grads, PpoLossParams lp) {
  __shared__ float sm[4][64];
  __shared__ float sc[16];
  const Layout L(D);
  const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + col;
  float acc = 0.f;
  if (e < L.P)
    for (int c = grp; c < g2; c += 4) acc += mat[(size_t)c * L.P + e];
  sm[grp][col] = acc;
  __syncthreads();
  const bool scalar_entry = e < L.w1 || e >= L.bmu || e == L.bv;
  if (grp == 0 && e < L.P && !scalar_entry) grads[e] = (sm[0][col] + sm[1][col]) + (sm[2][col] + sm[3][col]);
  if (blockIdx.x == 0 && threadIdx.x < 10) {
    float v = 0.f;
    for (int c = 0; c < g1; ++c) v += scal[c * 16 + threadIdx.x];
    sc[threadIdx.x] = v;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float* v = sc;
    grads[L.P + PPO_STAT_A_LOSS] = v[0]; grads[L.P + PPO_STAT_C_LOSS] = v[1]; grads[L.P + PPO_STAT_ENTROPY] = v[2];
    grads[L.P + PPO_STAT_B_LOSS] = v[3]; grads[L.P + PPO_STAT_KL] = v[4];
    grads[L.P + PPO_STAT_LOSS] = v[0] + 0.5f * v[1] * lp.critic_coef - v[2] * lp.entropy_coef + v[3] * lp.bounds_loss_coef;
    grads[L.P + PPO_STAT_GRAD_NORM] = 0.f; grads[L.P + PPO_STAT_LR] = 0.f;
    grads[L.sigma] = v[5]; grads[L.sigma + 1] = v[6]; grads[L.bmu] = v[7]; grads[L.bmu + 1] = v[8]; grads[L.bv] = v[9];
  }
}

// ---- fused tail of a minibatch step (single rank): second-stage gradient reduction + clip_grad_norm_ + Adam + adaptive-KL lr +
// re-packing of the operand tiles, ONE cooperative launch instead of reduce / adam / roll / pack (4 dependent launches were
// ~1/3 of the 65 us minibatch step in the r01 launch list).  Phases are separated by grid-wide barriers:
//   A  every CTA sums its 64 gradient entries over the partial slots (as reduce_tc_kernel) and publishes their sum of squares
//   B  every CTA rebuilds the global norm from the per-CTA partials in a fixed order (deterministic), applies clip + Adam to its
//      own entries; lr / step are read from slot [0], the new values go to slot [1]
//   C  slot [1] -> [0]; all CTAs re-pack the parameters into the tensor-core operand tiles
namespace cgx = cooperative_groups;
__global__ void __launch_bounds__(256) finish_tc_kernel(const float* __restrict__ scal, int g1, const float* __restrict__ mat, int g2, int D,
                                                       float* __restrict__ grads, PpoLossParams lp, float* __restrict__ prm,
                                                       float* __restrict__ m, float* __restrict__ v, float* __restrict__ lr,
                                                       int* __restrict__ step, PpoAdamParams ap, float* __restrict__ pk,
                                                       float* __restrict__ part_ss) {
  cgx::grid_group grid = cgx::this_grid();
  __shared__ float sm[4][64];
  __shared__ float sc[16];
  __shared__ float s_red[8];
  __shared__ float s_coef, s_norm;
  const Layout L(D);
  const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + col;
  // ---- A: reduce --------------------------------------------------------------------------------------------------------
  float acc = 0.f;
  if (e < L.P)
    for (int c = grp; c < g2; c += 4) acc += mat[(size_t)c * L.P + e];
  sm[grp][col] = acc;
  __syncthreads();
  const bool scalar_entry = e < L.w1 || e >= L.bmu || e == L.bv;
  float gval = 0.f;
  if (grp == 0 && e < L.P && !scalar_entry) {
    gval = (sm[0][col] + sm[1][col]) + (sm[2][col] + sm[3][col]);
    grads[e] = gval;
  }
  if (blockIdx.x == 0 && threadIdx.x < 10) {
    float x = 0.f;
    for (int c = 0; c < g1; ++c) x += scal[c * 16 + threadIdx.x];
    sc[threadIdx.x] = x;
  }
  __syncthreads();
  float ss = gval * gval;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float* q = sc;
    grads[L.P + PPO_STAT_A_LOSS] = q[0]; grads[L.P + PPO_STAT_C_LOSS] = q[1]; grads[L.P + PPO_STAT_ENTROPY] = q[2];
    grads[L.P + PPO_STAT_B_LOSS] = q[3]; grads[L.P + PPO_STAT_KL] = q[4];
    grads[L.P + PPO_STAT_LOSS] = q[0] + 0.5f * q[1] * lp.critic_coef - q[2] * lp.entropy_coef + q[3] * lp.bounds_loss_coef;
    grads[L.sigma] = q[5]; grads[L.sigma + 1] = q[6]; grads[L.bmu] = q[7]; grads[L.bmu + 1] = q[8]; grads[L.bv] = q[9];
    ss += (q[5] * q[5] + q[6] * q[6]) + (q[7] * q[7] + q[8] * q[8]) + q[9] * q[9];
  }
  if (grp == 0) {   // 64 threads = 2 warps hold the squares
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) part_ss[blockIdx.x] = s_red[0] + s_red[1];
  __threadfence();
  grid.sync();
  // ---- B: norm, clip, Adam ------------------------------------------------------------------------------------------------
  if (threadIdx.x < 32) {
    float x = 0.f;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) x += part_ss[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) {
      const float norm = sqrtf(x) * ap.inv_world;
      s_norm = norm;
      s_coef = (ap.grad_norm > 0.f) ? fminf(ap.grad_norm / (norm + 1e-6f), 1.0f) : 1.0f;
    }
  }
  __syncthreads();
  const float coef = s_coef * ap.inv_world;
  const float lr0 = lr[0];
  const int stp = step[0] + 1;
  if (grp == 0 && e < L.P) {
    const double bc1 = 1.0 - pow((double)ap.beta1, (double)stp), bc2 = 1.0 - pow((double)ap.beta2, (double)stp);
    const float step_size = (float)((double)lr0 / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float gr = (scalar_entry ? grads[e] : gval) * coef;
    const float mm = m[e] + (gr - m[e]) * (1.0f - ap.beta1);
    const float vv = v[e] * ap.beta2 + (1.0f - ap.beta2) * gr * gr;
    m[e] = mm;
    v[e] = vv;
    prm[e] = prm[e] - step_size * (mm / (sqrtf(vv) / bc2_sqrt + ap.eps));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float kl = grads[L.P + PPO_STAT_KL] * ap.inv_world;
    for (int q = 0; q < PPO_STAT_COUNT; ++q) grads[L.P + q] *= ap.inv_world;
    grads[L.P + PPO_STAT_GRAD_NORM] = s_norm;
    grads[L.P + PPO_STAT_LR] = lr0;
    float nl = lr0;
    if (ap.adaptive_lr) {
      if (kl > 2.0f * ap.kl_threshold) nl = fmaxf(lr0 / 1.5f, ap.min_lr);
      if (kl < 0.5f * ap.kl_threshold) nl = fminf(lr0 * 1.5f, ap.max_lr);
    }
    lr[1] = nl;
    step[1] = stp;
  }
  __threadfence();
  grid.sync();
  // ---- C: roll lr / step, re-pack -------------------------------------------------------------------------------------------
  if (blockIdx.x == 0 && threadIdx.x == 0) { lr[0] = lr[1]; step[0] = step[1]; }
  for (int q = blockIdx.x * 256 + threadIdx.x; q < PK_TOTAL; q += gridDim.x * 256) {
    int o, i;
    float x = 0.f;
    if (q < PK_W2T) { tile_inv(q - PK_W2, H / 4, o, i); x = prm[L.w2 + o * H + i]; }
    else if (q < PK_W1) { tile_inv(q - PK_W2T, H / 4, o, i); x = prm[L.w2 + i * H + o]; }
    else if (q < PK_W3) { tile_inv(q - PK_W1, DP / 4, o, i); x = i < D ? prm[L.w1 + o * D + i] : (i == D ? prm[L.b1 + o] : 0.f); }
    else if (q < PK_B2) { tile_inv(q - PK_W3, H / 4, o, i); x = o < 2 ? prm[L.wmu + o * H + i] : (o == 2 ? prm[L.wv + i] : 0.f); }
    else if (q < PK_B3) { x = prm[L.b2 + (q - PK_B2)]; }
    else { const int j = q - PK_B3; x = j < 2 ? prm[L.bmu + j] : (j == 2 ? prm[L.bv] : 0.f); }
    pk[q] = x;
  }
}

// ---- T2: weight gradients, K = samples ------------------------------------------------------------------------------
constexpr int WK = 64;   // samples per CTA chunk
// feature-major global slab [rows][ld] (columns c0..c0+WK) -> K-major operand tile (outer = feature row, inner = sample)
__device__ __forceinline__ void stage_slab(float* tile, const float* __restrict__ g, int rows, int64_t ld, int64_t c0) {
  // walk the tile in shared-memory order (consecutive lanes -> consecutive 16 B chunks: conflict-free); the matching global
  // reads are 8 rows x 64 B per warp instruction, i.e. whole 32 B sectors
  for (int q = threadIdx.x; q < rows * (WK / 4); q += NT) {
    const int og = q / (8 * (WK / 4)), rem = q - og * (8 * (WK / 4));
    const int ic = rem >> 3, ol = rem & 7;
    cp_async16(tile + q * 4, g + (int64_t)(og * 8 + ol) * ld + c0 + ic * 4);   // asynchronous: no register round trip
  }
}
__host__ __device__ inline size_t wgrad_smem_bytes() { return (size_t)(4 * H * WK + 2 * 16 * WK) * sizeof(float) + 64; }

__global__ void __launch_bounds__(NT, 1) wgrad_tc_kernel(TrainWs ws, int D, float* __restrict__ partial, int part0, int64_t nchunks) {
  extern __shared__ __align__(1024) float smem[];
  const Layout L(D);
  float* dz2 = smem;
  float* h1 = dz2 + H * WK;
  float* dz1 = h1 + H * WK;
  float* h2 = dz1 + H * WK;
  float* xt = h2 + H * WK;
  float* dz3 = xt + 16 * WK;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(dz3 + 16 * WK);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int t = threadIdx.x, row = t & 127, half = t >> 7;
  const uint32_t lane_base = (uint32_t)((t >> 5) & 3) * 32u;
  if (t == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (t < 32) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t acc_w2 = tmem, acc_w1 = tmem + 128, acc_w3 = tmem + 160, acc_b2 = tmem + 192;
  uint32_t phase = 0, first = 0;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, first = 1) {
    const int64_t c0 = ch * WK;
    stage_slab(dz2, ws.dz2t, H, ws.ld, c0);
    stage_slab(h1, ws.h1t, H, ws.ld, c0);
    stage_slab(dz1, ws.dz1t, H, ws.ld, c0);
    stage_slab(h2, ws.h2t, H, ws.ld, c0);
    stage_slab(xt, ws.xt, 16, ws.ld, c0);
    stage_slab(dz3, ws.dz3t, 16, ws.ld, c0);
    cp_async_wait_all();
    fence_async_smem(); tc_fence_before(); __syncthreads();
    if (t == 0) {
      tc_fence_after();
      const uint32_t i128 = make_idesc(128, 128, 0, 0), i16 = make_idesc(128, 16, 0, 0);
      for (int k = 0; k < WK; k += 8) {
        const uint32_t acc = (k > 0) ? 1u : first;
        const uint32_t ko = (k >> 2) * 128;
        const uint64_t d_dz2 = make_desc(smem_u32(dz2) + ko, 128, (WK / 4) * 128), d_h1 = make_desc(smem_u32(h1) + ko, 128, (WK / 4) * 128);
        const uint64_t d_dz1 = make_desc(smem_u32(dz1) + ko, 128, (WK / 4) * 128), d_h2 = make_desc(smem_u32(h2) + ko, 128, (WK / 4) * 128);
        const uint64_t d_x = make_desc(smem_u32(xt) + ko, 128, (WK / 4) * 128), d_dz3 = make_desc(smem_u32(dz3) + ko, 128, (WK / 4) * 128);
        umma_tf32(acc_w2, d_dz2, d_h1, i128, acc);    // dW2[o][i]   += sum_r dz2[r][o] H1[r][i]
        umma_tf32(acc_w1, d_dz1, d_x, i16, acc);      // dW1[o][d]   += sum_r dz1[r][o] X[r][d]   (column D: db1)
        umma_tf32(acc_w3, d_h2, d_dz3, i16, acc);     // dW3^T[k][j] += sum_r H2[r][k] dz3[r][j]
        umma_tf32(acc_b2, d_dz2, d_x, i16, acc);      // column D: db2[o] = sum_r dz2[r][o]
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1; tc_fence_after();
    tc_fence_before();
    __syncthreads();
  }
  // ---- accumulators -> this CTA's partial-gradient slot ----
  float* out = partial + (size_t)part0 + (size_t)blockIdx.x * L.P;   // matrix gradients only; scalars come from the T1 slots
  tc_fence_after();
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 32) {
    const int col = half * 64 + c0;
    float v[32];
    tmem_ld32(acc_w2 + (lane_base << 16) + (uint32_t)col, v);
#pragma unroll
    for (int q = 0; q < 32; ++q) out[L.w2 + row * H + col + q] = v[q];
  }
  if (t < 128) {
    float v[16];
    tmem_ld16(acc_w1 + (lane_base << 16), v);
    for (int d = 0; d < D; ++d) out[L.w1 + t * D + d] = v[d];
    out[L.b1 + t] = v[D];
    tmem_ld16(acc_w3 + (lane_base << 16), v);
    out[L.wmu + t] = v[0];
    out[L.wmu + H + t] = v[1];
    out[L.wv + t] = v[2];
    tmem_ld16(acc_b2 + (lane_base << 16), v);
    out[L.b2 + t] = v[D];
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

extern "C" int64_t ppo_packed_weight_floats(void) { return PK_TOTAL; }

// params -> packed operand tiles; call after every parameter update (the policy host class does)
extern "C" int ppo_pack_weights_tc(const float* params, int32_t obs_dim, float* packed, void* stream) {
  if (!params || !packed) return USV_E_NULL;
  if (obs_dim < 1 || obs_dim >= DP) return USV_E_SIZE;
  if ((uintptr_t)packed & 15) return USV_E_ALIGN;
  pack_weights_kernel<<<(PK_TOTAL + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, obs_dim, packed);
  return usv::finish_launch();
}

extern "C" int ppo_policy_forward_tc(const float* params, const float* packed, const float* obs, int32_t obs_dim, const float* obs_mean,
                                     const float* obs_var, const float* value_mean, const float* value_var, uint64_t seed,
                                     uint64_t counter, const uint64_t* counter_offset, int64_t row_offset, float* actions, float* neglogp, float* values,
                                     float* mus, float* sigmas, int64_t M, void* stream) {
  if (M < 0 || obs_dim < 1 || obs_dim >= DP) return USV_E_SIZE;   // one padded column is reserved for the bias
  if (M == 0) return USV_OK;
  if (!params || !packed || !obs || !obs_mean || !obs_var) return USV_E_NULL;
  if ((value_mean == nullptr) != (value_var == nullptr)) return USV_E_NULL;
  if ((uintptr_t)packed & 15) return USV_E_ALIGN;
  const size_t smem = fwd_smem_bytes();
  cudaFuncSetAttribute(forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t ntiles = (M + TM - 1) / TM;
  const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
  forward_tc_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(params, packed, obs, obs_dim, obs_mean, obs_var, value_mean, value_var, seed,
                                                              counter, counter_offset, row_offset, actions, neglogp, values, mus, sigmas, M);
  return usv::finish_launch();
}

namespace ppo { void launch_reduce(const float* scratch, int nparts, int n, float* grads, const PpoLossParams& lp, int P, cudaStream_t s); }

extern "C" int64_t ppo_train_tc_workspace_floats(int64_t M) {
  const int64_t ld = (M + 127) / 128 * 128;
  return (4 * (int64_t)H + 2 * 16) * ld;
}

// T1 + T2 + the fused cooperative tail: one whole PPO minibatch step (gradient, clip, Adam, adaptive lr, re-pack) in 3 launches
extern "C" int ppo_minibatch_step_tc(float* params, float* packed, const float* obs, int32_t obs_dim, const float* obs_mean,
                                     const float* obs_var, const float* actions, const float* old_neglogp, const float* advantages,
                                     const float* old_values, const float* returns, float* old_mu, float* old_sigma,
                                     const PpoLossParams* lp, float* grads, float* scratch, float* workspace, float* exp_avg,
                                     float* exp_avg_sq, float* lr, int32_t* step, const PpoAdamParams* ap, int64_t M, void* stream) {
  if (M <= 0 || obs_dim < 1 || obs_dim >= DP) return USV_E_SIZE;
  if (!params || !packed || !obs || !obs_mean || !obs_var || !actions || !old_neglogp || !advantages || !old_values || !returns ||
      !old_mu || !old_sigma || !lp || !grads || !scratch || !workspace || !exp_avg || !exp_avg_sq || !lr || !step || !ap)
    return USV_E_NULL;
  if (((uintptr_t)workspace | (uintptr_t)packed) & 15) return USV_E_ALIGN;
  const Layout L(obs_dim);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ld = (M + 127) / 128 * 128;
  TrainWs ws;
  ws.ld = ld;
  ws.h1t = workspace; ws.h2t = ws.h1t + H * ld; ws.dz2t = ws.h2t + H * ld; ws.dz1t = ws.dz2t + H * ld;
  ws.xt = ws.dz1t + H * ld; ws.dz3t = ws.xt + 16 * ld;
  cudaMemsetAsync(ws.dz3t + 3 * ld, 0, sizeof(float) * 13 * ld, st);  // rows 3..15 of dz3^T are structurally zero
  const int64_t ntiles = ld / TM, nchunks = ld / WK;
  int g1 = (int)(ntiles < num_sms() ? ntiles : num_sms()), g2 = (int)(nchunks < num_sms() ? nchunks : num_sms());
  if (g2 > 64) g2 = 64;
  cudaFuncSetAttribute(train_fwd_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_smem_bytes());
  cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_bytes());
  LossInTc in{actions, old_neglogp, advantages, old_values, returns, old_mu, old_sigma};
  train_fwd_bwd_tc_kernel<<<g1, NT, train_smem_bytes(), st>>>(params, packed, obs, obs_dim, obs_mean, obs_var, in, *lp, ws, scratch, M);
  const int mat0 = 16 * 160;
  wgrad_tc_kernel<<<g2, NT, wgrad_smem_bytes(), st>>>(ws, obs_dim, scratch, mat0, nchunks);
  // the per-CTA sum-of-squares partials live behind the matrix slots of `scratch` (ppo_train_scratch_floats covers 160 slots)
  const int fgrid = (L.P + 63) / 64;
  float* part_ss = scratch + mat0 + (size_t)64 * L.P;
  const float* scal = scratch;
  const float* mat = scratch + mat0;
  int D = obs_dim;
  PpoLossParams lpv = *lp;
  PpoAdamParams apv = *ap;
  void* args[] = {(void*)&scal, (void*)&g1, (void*)&mat, (void*)&g2, (void*)&D, (void*)&grads, (void*)&lpv, (void*)&params,
                  (void*)&exp_avg, (void*)&exp_avg_sq, (void*)&lr, (void*)&step, (void*)&apv, (void*)&packed, (void*)&part_ss};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)finish_tc_kernel, dim3(fgrid), dim3(256), args, 0, st);
  if (e != cudaSuccess) return (int)e;
  return usv::finish_launch(3);
}

extern "C" int ppo_minibatch_grad_tc(const float* params, const float* packed, const float* obs, int32_t obs_dim, const float* obs_mean,
                                     const float* obs_var, const float* actions, const float* old_neglogp, const float* advantages,
                                     const float* old_values, const float* returns, float* old_mu, float* old_sigma,
                                     const PpoLossParams* lp, float* grads, float* scratch, float* workspace, int64_t M, void* stream) {
  if (M <= 0 || obs_dim < 1 || obs_dim >= DP) return USV_E_SIZE;
  if (!params || !packed || !obs || !obs_mean || !obs_var || !actions || !old_neglogp || !advantages || !old_values || !returns ||
      !old_mu || !old_sigma || !lp || !grads || !scratch || !workspace)
    return USV_E_NULL;
  if (((uintptr_t)workspace | (uintptr_t)packed) & 15) return USV_E_ALIGN;
  const Layout L(obs_dim);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ld = (M + 127) / 128 * 128;
  TrainWs ws;
  ws.ld = ld;
  ws.h1t = workspace; ws.h2t = ws.h1t + H * ld; ws.dz2t = ws.h2t + H * ld; ws.dz1t = ws.dz2t + H * ld;
  ws.xt = ws.dz1t + H * ld; ws.dz3t = ws.xt + 16 * ld;
  cudaMemsetAsync(ws.dz3t, 0, sizeof(float) * 16 * ld, st);           // rows 3..15 of dz3^T are structurally zero
  const int64_t ntiles = ld / TM, nchunks = ld / WK;
  int g1 = (int)(ntiles < num_sms() ? ntiles : num_sms()), g2 = (int)(nchunks < num_sms() ? nchunks : num_sms());
  if (g2 > 64) g2 = 64;                                               // fewer, longer wgrad CTAs: the second stage reads g2 x P floats
  cudaFuncSetAttribute(train_fwd_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_smem_bytes());
  cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_bytes());
  LossInTc in{actions, old_neglogp, advantages, old_values, returns, old_mu, old_sigma};
  train_fwd_bwd_tc_kernel<<<g1, NT, train_smem_bytes(), st>>>(params, packed, obs, obs_dim, obs_mean, obs_var, in, *lp, ws, scratch, M);
  const int mat0 = 16 * 160;                                          // scalar slots first, then the matrix slots
  wgrad_tc_kernel<<<g2, NT, wgrad_smem_bytes(), st>>>(ws, obs_dim, scratch, mat0, nchunks);
  reduce_tc_kernel<<<(L.P + 63) / 64, 256, 0, st>>>(scratch, g1, scratch + mat0, g2, obs_dim, grads, *lp);
  return usv::finish_launch(3);
}
