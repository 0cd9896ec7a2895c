// Probe of tcgen05 MN-major (transposed) tf32 operands with un-swizzled tiles.  Build: nvcc -arch=sm_100a ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define PPO_HIDDEN 128
__device__ __forceinline__ int tile_off(int outer, int inner, int kc) { return (outer >> 3) * (kc * 32) + (inner >> 2) * 32 + (outer & 7) * 4 + (inner & 3); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; return d; }
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory"); }
__device__ __forceinline__ int sw_off(int outer, int inner, int outer_extent) {   // floats; rows of 32 floats, 8-row atoms, 128B swizzle
  const int colstride = (outer_extent / 8) * 256;
  return (inner / 32) * colstride + (outer / 8) * 256 + (outer % 8) * 32 + ((((inner % 32) / 4) ^ (outer % 8)) * 4) + (inner % 4);
}
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo) { return make_desc(saddr, lbo, sbo) | ((uint64_t)2 << 61); }
__global__ void probe(const float* A, const float* B, float* C, int N, int variant, int tmem_col) {
  extern __shared__ __align__(1024) float smem[];
  float* at = smem; float* bt = smem + 128 * 128; uint64_t* mbar = (uint64_t*)(bt + 128 * 128); uint32_t* slot = (uint32_t*)(mbar + 1);
  const int t = threadIdx.x; const int bkc = N / 4;
  // A tile: outer = k (128), inner = m (128);  B tile: outer = k (128), inner = n (N)   [A[k][m], B[k][n] row-major in global]
  // variants 0,1: MN-major NONE. 2: K-major NONE (tiles transposed: outer = m/n, inner = k). 3: MN-major SW128. 4: K-major SW128.
  for (int e = t; e < 128 * 128; e += blockDim.x) { int k = e / 128, m = e % 128;
    if (variant <= 1 || variant == 5) at[tile_off(k, m, 32)] = A[e]; else if (variant == 2 || variant == 6) at[tile_off(m, k, 32)] = A[e];
    else if (variant == 3) at[sw_off(k, m, 128)] = A[e]; else at[sw_off(m, k, 128)] = A[e]; }
  for (int e = t; e < 128 * N; e += blockDim.x) { int k = e / N, n = e % N;
    if (variant <= 1 || variant == 6) bt[tile_off(k, n, bkc)] = B[e]; else if (variant == 2 || variant == 5) bt[tile_off(n, k, 32)] = B[e];
    else if (variant == 3) bt[sw_off(k, n, 128)] = B[e]; else bt[sw_off(n, k, N)] = B[e]; }
  if (t == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (t < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot + tmem_col;
  if (t == 0) {
    const uint32_t idesc = (variant == 2 || variant == 4) ? make_idesc(128, N, 0, 0) : (variant == 5 ? make_idesc(128, N, 1, 0) : (variant == 6 ? make_idesc(128, N, 0, 1) : make_idesc(128, N, 1, 1)));
    for (int k = 0; k < 128; k += 8) {
      uint64_t ad, bd;
      if (variant == 2) {
        ad = make_desc(smem_u32(at) + (k >> 2) * 128, 128, 32 * 128);
        bd = make_desc(smem_u32(bt) + (k >> 2) * 128, 128, 32 * 128);
      } else if (variant == 5) {
        ad = make_desc(smem_u32(at) + (k >> 3) * (32 * 128), 32 * 128, 128);
        bd = make_desc(smem_u32(bt) + (k >> 2) * 128, 128, 32 * 128);
      } else if (variant == 6) {
        ad = make_desc(smem_u32(at) + (k >> 2) * 128, 128, 32 * 128);
        bd = make_desc(smem_u32(bt) + (k >> 3) * (bkc * 128), bkc * 128, 128);
      } else if (variant == 3) {   // MN-major SW128: K rows of 128 B; MN atoms (32 elems) stride LBO = (128/8)*1024
        ad = make_desc_sw(smem_u32(at) + (k >> 3) * 1024, 16 * 1024, 1024);
        bd = make_desc_sw(smem_u32(bt) + (k >> 3) * 1024, 16 * 1024, 1024);
      } else if (variant == 4) {   // K-major SW128: MN rows of 128 B (32 k's); k atoms stride = (rows/8)*1024
        ad = make_desc_sw(smem_u32(at) + (k / 32) * (16 * 1024) + (k % 32) * 4, 16, 1024);
        bd = make_desc_sw(smem_u32(bt) + (k / 32) * ((N / 8) * 1024) + (k % 32) * 4, 16, 1024);
      } else if (variant == 0) {        // LBO = K-group stride, SBO = MN-group stride (CUTLASS INTERLEAVE convention)
        ad = make_desc(smem_u32(at) + (k >> 3) * (32 * 128), 32 * 128, 128);
        bd = make_desc(smem_u32(bt) + (k >> 3) * (bkc * 128), bkc * 128, 128);
      } else {                   // swapped
        ad = make_desc(smem_u32(at) + (k >> 3) * (32 * 128), 128, 32 * 128);
        bd = make_desc(smem_u32(bt) + (k >> 3) * (bkc * 128), 128, bkc * 128);
      }
      umma(tmem, ad, bd, idesc, k > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  uint32_t done = 0; while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (t < 128) {
    uint32_t r[16];
    for (int c0 = 0; c0 < N; c0 += 16) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem + ((uint32_t)(t / 32 * 32) << 16) + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 16; ++q) C[t * N + c0 + q] = __uint_as_float(r[q]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "r"(512));
}
int main() {
  for (int N : {16, 128}) for (int col : {0}) for (int variant : {5, 6}) {
    std::vector<float> A(128 * 128), B(128 * N), C(128 * N), R(128 * N, 0.f);
    for (int k = 0; k < 128; ++k) for (int m = 0; m < 128; ++m) A[k * 128 + m] = (float)((k * 3 + m * 7) % 5 - 2);
    for (int k = 0; k < 128; ++k) for (int n = 0; n < N; ++n) B[k * N + n] = (float)((k * 5 + n * 11) % 7 - 3);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 128; ++k) s += A[k * 128 + m] * B[k * N + n]; R[m * N + n] = s; }
    float *dA, *dB, *dC; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice); cudaMemset(dC, 0xff, C.size() * 4);
    size_t smem = 2 * 128 * 128 * 4 + 64; cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 256, smem>>>(dA, dB, dC, N, variant, col);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, zeros = 0; for (size_t i = 0; i < C.size(); ++i) { if (C[i] != R[i]) ++bad; if (C[i] == 0.f) ++zeros; }
    printf("N=%3d col=%3d variant=%d err=%s mismatches=%d/%zu zeros=%d  C[0..3]=%g %g %g %g  ref=%g %g %g %g\n", N, col, variant, cudaGetErrorString(e), bad, C.size(), zeros, C[0], C[1], C[2], C[3], R[0], R[1], R[2], R[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
  }
  return 0;
}
