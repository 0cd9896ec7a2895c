// Shared host/device helpers for libusv_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/usv_b200.h"

namespace usv {

extern std::atomic<int64_t> g_launch_count;

inline int finish_launch(int n_launches = 1) {
  g_launch_count.fetch_add(n_launches, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? USV_OK : (int)cudaGetLastError();
}

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// rotation-matrix rows needed for R^T v with R = quaternion_to_matrix(q), q = (r,i,j,k) real-first,
// two_s = 2/|q|^2  (public PyTorch3D formula; see oracle/ref_shim.py for the restatement we pin)
struct Rot3 {
  float m00, m01, m02, m10, m11, m12, m20, m21, m22;
};

__device__ __forceinline__ Rot3 quat_to_matrix(float r, float i, float j, float k) {
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  Rot3 R;
  R.m00 = 1.0f - two_s * (j * j + k * k);
  R.m01 = two_s * (i * j - k * r);
  R.m02 = two_s * (i * k + j * r);
  R.m10 = two_s * (i * j + k * r);
  R.m11 = 1.0f - two_s * (i * i + k * k);
  R.m12 = two_s * (j * k - i * r);
  R.m20 = two_s * (i * k - j * r);
  R.m21 = two_s * (j * k + i * r);
  R.m22 = 1.0f - two_s * (i * i + j * j);
  return R;
}

// (R^T v)
__device__ __forceinline__ void rot_t_apply(const Rot3& R, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = R.m00 * x + R.m10 * y + R.m20 * z;
  oy = R.m01 * x + R.m11 * y + R.m21 * z;
  oz = R.m02 * x + R.m12 * y + R.m22 * z;
}

// thruster LUT index: clamp(round_half_even(((cmd+1)/2)*(n-1)), 0, n-1) with torch's op order
// (no FMA contraction: the index must be bit-exact)  [ref: OIGE/envs/USV/ThrusterDynamics.py:187-191]
__device__ __forceinline__ int lut_index(float cmd, int n_lut) {
  float t = __fmul_rn(__fmul_rn(__fadd_rn(cmd, 1.0f), 0.5f), (float)(n_lut - 1));
  float rr = rintf(t);  // round-half-to-even == torch.round
  int idx = (int)rr;
  // NaN -> (long)NaN is UB in the reference; we pin it to 0
  if (!(rr >= 0.0f)) idx = 0;
  if (idx > n_lut - 1) idx = n_lut - 1;
  return idx;
}

}  // namespace usv
