// Fused ASV env step for sm_100a: reset-if-flagged -> action -> thruster LUT -> n_substeps x
// {first-order thruster lag, hydrodynamic damping, disturbances, planar rigid-body semi-implicit
// Euler} -> CaptureXY observation (13) -> reward + penalties -> kills / done.  One launch replaces
// VecEnvRLGames.step  [ref: OIGE/envs/vec_env_rlgames.py:120-217] and everything below it.
//
// Layout: env state / per-episode constants are structure-of-arrays in HBM (field-major), so a
// warp's 32 envs read each field as one 128 B line; the (n,13) row-major observation rl_games wants
// is staged through shared memory (stride 13 is odd -> conflict-free) and written as float4 lines.
// The thruster LUTs (2 x n_lut floats, gathered at random) are staged once per CTA in shared memory.
// RNG is Philox4x32-10 keyed (seed; env, step, stream): no RNG state in HBM.
//
// Roofline: HBM-bound by design -- algorithmic bytes per env-step are tabulated in DESIGN.md.
#include <math_constants.h>
#include "philox.cuh"
#include "usv_common.cuh"

namespace usv {

#ifndef USV_BLOCK
#define USV_BLOCK 256
#endif
#ifndef USV_MINB
#define USV_MINB 3
#endif
constexpr int kBlock = USV_BLOCK;
constexpr int kObs = 13;
// python: math.pi / 2*math.pi are doubles that meet fp32 tensors -> rounded to fp32
#define USV_PI_F 3.14159274101257324f
#define USV_2PI_F 6.28318548202514648f

struct EnvState {
  float x, y, psi, vx, vy, r, thrL, thrR, prev_d, prev_w, prev_asum;
  int goal_cnt, progress;
};
struct EnvConst {
  float tx, ty, mass, linu, linv, linr, quadu, quadv, quadr, kdrag, mL, mR, kiz;
  float fcx, fcy, fxf, fyf, fxs, fys, famp, tc, tf, ts, tamp;
};
struct StepOut {
  float obs[kObs];
  float rew;
  int done;
  bool finite;
  // diagnostics for stats
  float dist_rew, align_rew, speed_rew, d, speed, bpen, bdist;
  float pen_lin, pen_ang, pen_angvar, pen_energy, pen_actvar, absw, asum;
};

// AoSoA: envs are grouped in tiles of 32 (one warp); inside a tile the fields are consecutive 128 B lines:
//   field f of env i lives at base[((i >> 5) * COUNT + f) * 32 + (i & 31)].
// A warp reads/writes one full line per field, and every field of an env is an IMMEDIATE offset from one
// per-thread base pointer (plain field-major SoA with a runtime stride cost two 64-bit adds per access,
// ~100 of the 1678 instructions per env-step in the r01 profile).
constexpr int kTile = 32;
__device__ __forceinline__ int64_t tile_base(int64_t i, int count) { return (i >> 5) * (int64_t)(count * kTile) + (i & 31); }
#define stride kTile

__device__ __forceinline__ void load_state(const float* __restrict__ s0, int64_t /*cap*/, int64_t i0, EnvState& e) {
  const float* __restrict__ s = s0 + tile_base(i0, USV_S_COUNT);
  constexpr int i = 0;
  e.x = s[USV_S_X * stride + i];
  e.y = s[USV_S_Y * stride + i];
  e.psi = s[USV_S_PSI * stride + i];
  e.vx = s[USV_S_VX * stride + i];
  e.vy = s[USV_S_VY * stride + i];
  e.r = s[USV_S_R * stride + i];
  e.thrL = s[USV_S_THR_L * stride + i];
  e.thrR = s[USV_S_THR_R * stride + i];
  e.prev_d = s[USV_S_PREV_D * stride + i];
  e.prev_w = s[USV_S_PREV_W * stride + i];
  e.prev_asum = s[USV_S_PREV_ASUM * stride + i];
  e.goal_cnt = __float_as_int(s[USV_S_GOAL_CNT * stride + i]);
  e.progress = __float_as_int(s[USV_S_PROGRESS * stride + i]);
}

__device__ __forceinline__ void store_state(float* __restrict__ s0, int64_t /*cap*/, int64_t i0, const EnvState& e) {
  float* __restrict__ s = s0 + tile_base(i0, USV_S_COUNT);
  constexpr int i = 0;
  s[USV_S_X * stride + i] = e.x;
  s[USV_S_Y * stride + i] = e.y;
  s[USV_S_PSI * stride + i] = e.psi;
  s[USV_S_VX * stride + i] = e.vx;
  s[USV_S_VY * stride + i] = e.vy;
  s[USV_S_R * stride + i] = e.r;
  s[USV_S_THR_L * stride + i] = e.thrL;
  s[USV_S_THR_R * stride + i] = e.thrR;
  s[USV_S_PREV_D * stride + i] = e.prev_d;
  s[USV_S_PREV_W * stride + i] = e.prev_w;
  s[USV_S_PREV_ASUM * stride + i] = e.prev_asum;
  s[USV_S_GOAL_CNT * stride + i] = __int_as_float(e.goal_cnt);
  s[USV_S_PROGRESS * stride + i] = __int_as_float(e.progress);
}

template <bool kDisturb>
__device__ __forceinline__ void load_consts(const float* __restrict__ c0, int64_t /*cap*/, int64_t i0, EnvConst& k) {
  const float* __restrict__ c = c0 + tile_base(i0, USV_C_COUNT);
  constexpr int i = 0;
  k.tx = c[USV_C_TX * stride + i];
  k.ty = c[USV_C_TY * stride + i];
  k.mass = c[USV_C_MASS * stride + i];
  k.linu = c[USV_C_LIN_U * stride + i];
  k.linv = c[USV_C_LIN_V * stride + i];
  k.linr = c[USV_C_LIN_R * stride + i];
  k.quadu = c[USV_C_QUAD_U * stride + i];
  k.quadv = c[USV_C_QUAD_V * stride + i];
  k.quadr = c[USV_C_QUAD_R * stride + i];
  k.kdrag = c[USV_C_KDRAG * stride + i];
  k.mL = c[USV_C_THR_ML * stride + i];
  k.mR = c[USV_C_THR_MR * stride + i];
  k.kiz = c[USV_C_KIZ * stride + i];
  if (kDisturb) {
    k.fcx = c[USV_C_FCX * stride + i];
    k.fcy = c[USV_C_FCY * stride + i];
    k.fxf = c[USV_C_FXF * stride + i];
    k.fyf = c[USV_C_FYF * stride + i];
    k.fxs = c[USV_C_FXS * stride + i];
    k.fys = c[USV_C_FYS * stride + i];
    k.famp = c[USV_C_FAMP * stride + i];
    k.tc = c[USV_C_TC * stride + i];
    k.tf = c[USV_C_TF * stride + i];
    k.ts = c[USV_C_TS * stride + i];
    k.tamp = c[USV_C_TAMP * stride + i];
  } else {
    k.fcx = k.fcy = k.fxf = k.fyf = k.fxs = k.fys = k.famp = k.tc = k.tf = k.ts = k.tamp = 0.0f;
  }
}

template <bool kDisturb>
__device__ __forceinline__ void store_consts(float* __restrict__ c0, int64_t /*cap*/, int64_t i0, const EnvConst& k) {
  float* __restrict__ c = c0 + tile_base(i0, USV_C_COUNT);
  constexpr int i = 0;
  c[USV_C_TX * stride + i] = k.tx;
  c[USV_C_TY * stride + i] = k.ty;
  c[USV_C_MASS * stride + i] = k.mass;
  c[USV_C_LIN_U * stride + i] = k.linu;
  c[USV_C_LIN_V * stride + i] = k.linv;
  c[USV_C_LIN_R * stride + i] = k.linr;
  c[USV_C_QUAD_U * stride + i] = k.quadu;
  c[USV_C_QUAD_V * stride + i] = k.quadv;
  c[USV_C_QUAD_R * stride + i] = k.quadr;
  c[USV_C_KDRAG * stride + i] = k.kdrag;
  c[USV_C_THR_ML * stride + i] = k.mL;
  c[USV_C_THR_MR * stride + i] = k.mR;
  c[USV_C_KIZ * stride + i] = k.kiz;
  if (kDisturb) {
    c[USV_C_FCX * stride + i] = k.fcx;
    c[USV_C_FCY * stride + i] = k.fcy;
    c[USV_C_FXF * stride + i] = k.fxf;
    c[USV_C_FYF * stride + i] = k.fyf;
    c[USV_C_FXS * stride + i] = k.fxs;
    c[USV_C_FYS * stride + i] = k.fys;
    c[USV_C_FAMP * stride + i] = k.famp;
    c[USV_C_TC * stride + i] = k.tc;
    c[USV_C_TF * stride + i] = k.tf;
    c[USV_C_TS * stride + i] = k.ts;
    c[USV_C_TAMP * stride + i] = k.tamp;
  }
}

__device__ __forceinline__ float urange(float u, float lo, float hi) { return u * (hi - lo) + lo; }

// Branch-free sin/cos for |x| < ~1e5 (every angle in this kernel is bounded: headings, phases x*f+shift):
// 3-term Cody-Waite reduction by pi/2 + the classic single-precision minimax polynomials on [-pi/4, pi/4]
// (~1 ulp).  libdevice's sinf/cosf carry a Payne-Hanek slow path behind a branch + convergence barrier per
// call; 22 calls per env-step made that ~10% of the issued instructions (profiles/r01_step_kernel.md).
__device__ __forceinline__ void trig_reduce(float x, float& r, int& q) {
  const float j = rintf(x * 0.636619772f);
  q = (int)j;
  r = fmaf(j, -1.57079601e+00f, x);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
}
__device__ __forceinline__ float poly_sin(float r, float r2) {
  float p = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  p = fmaf(p, r2, -1.6666654611e-1f);
  return fmaf(p * r2, r, r);
}
__device__ __forceinline__ float poly_cos(float r2) {
  float p = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  p = fmaf(p, r2, 4.166664568298827e-2f);
  p = fmaf(p, r2, -0.5f);
  return fmaf(p, r2, 1.0f);
}
__device__ __forceinline__ void fsincos(float x, float* sp, float* cp) {
  float r; int q;
  trig_reduce(x, r, q);
  const float r2 = r * r;
  const float s = poly_sin(r, r2), c = poly_cos(r2);
  const float ss = (q & 1) ? c : s;
  const float cc = (q & 1) ? s : c;
  // sign flips as sign-bit XORs: sin negated in quadrants 2,3; cos in quadrants 1,2
  *sp = __int_as_float(__float_as_int(ss) ^ ((q & 2) << 30));
  *cp = __int_as_float(__float_as_int(cc) ^ (((q + 1) & 2) << 30));
}
// sin() for the sinusoidal force/torque disturbances, evaluated 3x per physics sub-step: explicit 2-term
// Cody-Waite reduction to [-pi, pi], then the SFU (MUFU.SIN, abs error <= 2^-21.4 on that interval).  With
// amplitudes <= 1.77 N / 1 Nm the force error is < 1e-6 N against drag/thrust forces of 1..100 N, i.e. far inside
// the 1e-5 parity bar, and it replaces ~17 issue slots by 6 (the loop was 48% of the kernel, r01 profile).
__device__ __forceinline__ float fsin_sfu(float x) {
  const float k = rintf(x * 0.159154943f);
  float r = fmaf(k, -6.28318548e+00f, x);
  r = fmaf(k, 1.74845553e-07f, r);
  return __sinf(r);
}
__device__ __forceinline__ float fsin(float x) {
  float r; int q;
  trig_reduce(x, r, q);
  const float r2 = r * r;
  const float v = (q & 1) ? poly_cos(r2) : poly_sin(r, r2);
  return __int_as_float(__float_as_int(v) ^ ((q & 2) << 30));
}

// wrap an angle into (-pi, pi]  (the branch torch.atan2 returns for the yaw read-back)
__device__ __forceinline__ float wrap_pi(float a) {
  // in-range values pass through bit-exactly (rintf gives 0): same as the oracle's masked wrap
  a = a - USV_2PI_F * rintf(a * (1.0f / USV_2PI_F));
  a = (a > USV_PI_F) ? a - USV_2PI_F : a;
  a = (a <= -USV_PI_F) ? a + USV_2PI_F : a;
  return a;
}

__device__ __forceinline__ float penalty_scalar(const UsvPenaltyTerm& t, float x) {
  switch (t.form) {
    case USV_PEN_NEG_ABS: return -fabsf(x) * t.c1 + t.c2;
    case USV_PEN_NEG_DEADZONE: return -fmaxf(fabsf(x) - t.k, 0.0f) * t.c1;
    case USV_PEN_EXP_NEG_ABS: return (__expf(-t.k * fabsf(x)) - 1.0f) * t.c1;
    default: return 0.0f;
  }
}

// reset_idx for one env  [ref: SNAP/USV_Virtual.py:750-817 ; OIGE/tasks/USV_Virtual.py:1502-1618]
template <bool kDisturb>
__device__ __forceinline__ void reset_env(EnvState& e, EnvConst& k, const UsvStepParams& p, uint64_t gid, uint64_t step) {
  const Uniform4 r0 = philox_uniform4(p.seed, gid, step, RS_RESET_0);
  const Uniform4 r1 = philox_uniform4(p.seed, gid, step, RS_RESET_1);
  // task.reset / get_spawns: goal counter cleared  [ref SNAP/USV_capture_xy.py:308-310,342]
  e.goal_cnt = 0;
  if (kDisturb) {
    // UF.generate_force / TD.generate_torque  [ref OIGE/tasks/USV/USV_disturbances.py:327-384,469-508]
    if (p.use_force_disturbance) {
      const Uniform4 r5 = philox_uniform4(p.seed, gid, step, RS_RESET_5);
      const Uniform4 r6 = philox_uniform4(p.seed, gid, step, RS_RESET_6);
      const Uniform4 r7 = philox_uniform4(p.seed, gid, step, RS_RESET_7);
      if (p.use_sin_force) {
        k.fxf = urange(r5.c, p.force_min_freq, p.force_max_freq);
        k.fyf = urange(r5.d, p.force_min_freq, p.force_max_freq);
        k.fxs = urange(r6.a, p.force_min_shift, p.force_max_shift);
        k.fys = urange(r6.b, p.force_min_shift, p.force_max_shift);
        k.famp = urange(r6.c, p.force_sin_min, p.force_sin_max);
      }
      if (p.use_const_force) {
        const float rr = urange(r6.d, p.force_const_min, p.force_const_max);
        const float th = r7.a * USV_PI_F * 2.0f;
        float sth_, cth_;
        fsincos(th, &sth_, &cth_);
        k.fcx = cth_ * rr;
        k.fcy = sth_ * rr;
      }
    }
    if (p.use_torque_disturbance) {
      const Uniform4 r7 = philox_uniform4(p.seed, gid, step, RS_RESET_7);
      const Uniform4 r8 = philox_uniform4(p.seed, gid, step, RS_RESET_8);
      if (p.use_sin_torque) {
        k.tf = urange(r7.b, p.torque_min_freq, p.torque_max_freq);
        k.ts = urange(r7.c, p.torque_min_shift, p.torque_max_shift);
        k.tamp = urange(r7.d, p.torque_sin_min, p.torque_sin_max);
      }
      if (p.use_const_torque) {
        float rr = urange(r8.a, p.torque_const_min, p.torque_const_max);
        if (r8.b > 0.5f) rr *= -1.0f;
        k.tc = rr;
      }
    }
  }
  // MDD.randomize_masses  [ref USV_disturbances.py:127-151]
  k.mass = p.mass_rand ? urange(r1.d, p.mass_min, p.mass_max) : p.mass_base;
  // hydrodynamics.reset_coefficients  [ref OIGE/envs/USV/Hydrodynamics.py:136-174]
  if (p.drag_rand) {
    const Uniform4 r3 = philox_uniform4(p.seed, gid, step, RS_RESET_3);
    const Uniform4 r4 = philox_uniform4(p.seed, gid, step, RS_RESET_4);
    k.linu = p.lin_base[0] + (r3.a * 2.0f - 1.0f) * p.lin_rand[0];
    k.linv = p.lin_base[1] + (r3.b * 2.0f - 1.0f) * p.lin_rand[1];
    k.linr = p.lin_base[2] + (r3.c * 2.0f - 1.0f) * p.lin_rand[2];
    k.quadu = p.quad_base[0] + (r4.a * 2.0f - 1.0f) * p.quad_rand[0];
    k.quadv = p.quad_base[1] + (r4.b * 2.0f - 1.0f) * p.quad_rand[1];
    k.quadr = p.quad_base[2] + (r4.c * 2.0f - 1.0f) * p.quad_rand[2];
  }
  if (p.kdrag_rand) {  // _sample_k_drag [ref Hydrodynamics.py:119-134]
    const Uniform4 r2 = philox_uniform4(p.seed, gid, step, RS_RESET_2);
    if (p.kdrag_log) {
      const float l0 = logf(p.kdrag_min), l1 = logf(p.kdrag_max);
      k.kdrag = expf(l0 + r2.d * (l1 - l0));
    } else {
      k.kdrag = p.kdrag_min + r2.d * (p.kdrag_max - p.kdrag_min);
    }
  }
  // thrusters.reset_thruster_randomization  [ref OIGE/envs/USV/ThrusterDynamics.py:112-127]
  if (p.thr_rand) {
    if (p.thr_separate) {
      const Uniform4 r4 = philox_uniform4(p.seed, gid, step, RS_RESET_4);
      const Uniform4 r5 = philox_uniform4(p.seed, gid, step, RS_RESET_5);
      k.mL = r4.d * 2.0f * p.thr_left_frac + (1.0f - p.thr_left_frac);
      k.mR = r5.a * 2.0f * p.thr_right_frac + (1.0f - p.thr_right_frac);
    } else {
      const Uniform4 r3 = philox_uniform4(p.seed, gid, step, RS_RESET_3);
      k.mL = k.mR = r3.d * 2.0f * p.thr_rand_frac + (1.0f - p.thr_rand_frac);
    }
  }
  // _apply_mass_driven_coupling  [ref OIGE/tasks/USV_Virtual.py:988-1040]
  if (p.mass_coupling) {
    const float denom = fmaxf(p.couple_mass_max - p.mass_base, 1e-6f);
    const float rr = fminf(fmaxf((k.mass - p.mass_base) / denom, 0.0f), 1.0f);
    k.kdrag = p.kdrag_min + rr * (p.kdrag_max - p.kdrag_min);
    const float s = fminf(fmaxf(1.0f - rr * p.couple_thr_a, 1.0f - p.couple_thr_a), 1.0f);
    k.mL = k.mR = s;
    k.kiz = p.couple_kiz_min + rr * (p.couple_kiz_max - p.couple_kiz_min);
  }
  // goals (live calls set_targets at the end of reset_idx; classic only at post_reset)
  if (p.retarget_on_reset) {  // [ref SNAP/USV_capture_xy.py:312-326]
    k.tx = r0.a * p.goal_random_position * 2.0f - p.goal_random_position;
    k.ty = r0.b * p.goal_random_position * 2.0f - p.goal_random_position;
  }
  // get_spawns  [ref SNAP/USV_capture_xy.py:330-394]: annulus around the target, yaw on a half circle
  const float sr = r0.c * (p.spawn_max_dist - p.spawn_min_dist) + p.spawn_min_dist;
  const float sth = r0.d * 2.0f * USV_PI_F;
  float ssp, csp;
  fsincos(sth, &ssp, &csp);
  e.x = sr * csp + k.tx;
  e.y = sr * ssp + k.ty;
  // quaternion (cos(a/2),0,0,sin(a/2)) with a ~ U[0,pi)  ->  yaw = a
  e.psi = r1.a * USV_PI_F;
  // root velocities: zero, then vx,vy ~ U(-1.5,1.5) in the world frame  [ref SNAP/USV_Virtual.py:786-794]
  e.vx = r1.b * (2.0f * p.spawn_vel_range) - p.spawn_vel_range;
  e.vy = r1.c * (2.0f * p.spawn_vel_range) - p.spawn_vel_range;
  e.r = 0.0f;
  e.progress = 0;
  // NOTE: thruster lag state (current_forces), prev_d, prev_w and prev_asum are NOT reset
  // (reference quirks 4 and 7).
}

// planar force model for one physics sub-step; returns body wrench and world acceleration
template <bool kDisturb>
__device__ __forceinline__ void planar_wrench(const EnvState& e, const EnvConst& k, const UsvStepParams& p, float ox,
                                              float oy, float inv_m, float inv_iz, float s, float c, float& du, float& dv,
                                              float& dr, float& Fx, float& Fy, float& Tz, float& ax, float& ay,
                                              float& rdot) {
  // R^T v (world -> body)  [ref Hydrodynamics.py:213-222, planar quaternion]
  const float u = c * e.vx + s * e.vy;
  const float v = -s * e.vx + c * e.vy;
  const float w = e.r;
  // ComputeDampingMatrix  [ref Hydrodynamics.py:176-205]
  const float fwd_u = p.lin_fwd[0] + p.offset_lin_forward_damping_speed;
  const float fwd_v = p.lin_fwd[1] + p.offset_lin_forward_damping_speed;
  const float fwd_r = p.lin_fwd[2] + p.offset_lin_forward_damping_speed;
  float Du = (((k.linu + p.offset_linear_damping) - fwd_u) + (k.quadu + p.offset_nonlin_damping) * fabsf(u)) * p.scaling_damping;
  float Dv = (((k.linv + p.offset_linear_damping) - fwd_v) + (k.quadv + p.offset_nonlin_damping) * fabsf(v)) * p.scaling_damping;
  float Dr = (((k.linr + p.offset_linear_damping) - fwd_r) + (k.quadr + p.offset_nonlin_damping) * fabsf(w)) * p.scaling_damping;
  if (p.use_drag_scale) { Du *= k.kdrag; Dv *= k.kdrag; Dr *= k.kdrag; }
  du = -1.0f * Du * u;
  dv = -1.0f * Dv * v;
  dr = -1.0f * Dr * w;
  // disturbances: functions of the WORLD position, applied in the BODY frame (is_global=False)
  // [ref USV_disturbances.py:386-410,510-530 ; SNAP/USV_Virtual.py:621-650]
  float fdx = 0.0f, fdy = 0.0f, td = 0.0f;
  if (kDisturb) {
    if (p.use_const_force) { fdx = k.fcx; fdy = k.fcy; }
    if (p.use_sin_force) {
      fdx = k.fcx + fsin_sfu((e.x + ox) * k.fxf + k.fxs) * k.famp;
      fdy = k.fcy + fsin_sfu((e.y + oy) * k.fyf + k.fys) * k.famp;
    }
    if (p.use_const_torque) td = k.tc;
    if (p.use_sin_torque) td = k.tc + fsin_sfu(((e.x + ox) + (e.y + oy)) * k.tf + k.ts) * k.tamp;
  }
  // net wrench at the base link; thrusters push along body x at (thr_x, thr_y_*)  (heron.urdf:167,242)
  Fx = fdx + du + e.thrL + e.thrR;
  Fy = fdy + dv;
  Tz = td + dr - p.thr_y_left * e.thrL - p.thr_y_right * e.thrR;
  ax = (c * Fx - s * Fy) * inv_m;
  ay = (s * Fx + c * Fy) * inv_m;
  rdot = Tz * inv_iz;
}

// one full control step for one env, state in registers
template <bool kDisturb, bool kLutGlobal>
__device__ __forceinline__ void control_step(EnvState& e, EnvConst& k, const UsvStepParams& p, bool do_reset,
                                             float2 act, uint64_t gid, int64_t lid, uint64_t step, bool first_call,
                                             const float* __restrict__ s_lutL, const float* __restrict__ s_lutR,
                                             StepOut& o) {
  // ---- pre_physics_step ------------------------------------------------------------------
  if (do_reset) reset_env<kDisturb>(e, k, p, gid, step);
  const Uniform8 nz = philox_uniform8x16(p.seed, gid, step, RS_STEP_A);
  const float u_a0 = nz.v[0], u_a1 = nz.v[1], u_vx = nz.v[2], u_vy = nz.v[3], u_w = nz.v[4], u_h = nz.v[5],
              u_px = nz.v[6], u_py = nz.v[7];
  // VecEnvRLGames.step clamps to +-clipActions  [ref vec_env_rlgames.py:136-140]
  float a0 = fminf(fmaxf(act.x, -p.clip_actions), p.clip_actions);
  float a1 = fminf(fmaxf(act.y, -p.clip_actions), p.clip_actions);
  float pa0, pa1;  // what Penalties sees as `actions`
  float c0, c1;    // command that indexes the LUT
  if (!p.action_affine) {
    // classic [ref SNAP/USV_Virtual.py:589-615]: AN.add_noise_on_act works IN PLACE on the tensor that
    // self.actions aliases -> penalties see the noisy, unclamped action; resets zero only the thrust.
    if (p.action_noise) {
      a0 += urange(u_a0, p.action_noise_min, p.action_noise_max);
      a1 += urange(u_a1, p.action_noise_min, p.action_noise_max);
    }
    pa0 = a0; pa1 = a1;
    c0 = fminf(fmaxf(a0, -1.0f), 1.0f);
    c1 = fminf(fmaxf(a1, -1.0f), 1.0f);
  } else {
    // live [ref OIGE/tasks/USV_Virtual.py:1064-1097]
    float t0 = a0 + p.action_bias, t1 = a1 + p.action_bias;
    if (p.action_noise) {
      t0 += urange(u_a0, p.action_noise_min, p.action_noise_max);
      t1 += urange(u_a1, p.action_noise_min, p.action_noise_max);
    }
    t0 = fminf(fmaxf(t0, -1.0f), 1.0f);
    t1 = fminf(fmaxf(t1, -1.0f), 1.0f);
    c0 = fminf(fmaxf(0.5f * (t0 + 1.0f), 0.0f), 1.0f);
    c1 = fminf(fmaxf(0.5f * (t1 + 1.0f), 0.0f), 1.0f);
    pa0 = p.penalties_use_u ? c0 : a0;
    pa1 = p.penalties_use_u ? c1 : a1;
  }
  if (do_reset) { c0 = 0.0f; c1 = 0.0f; }
  // set_target_force -> get_cmd_interpolated  [ref ThrusterDynamics.py:179-219]
  const int iL = lut_index(c0, p.n_lut), iR = lut_index(c1, p.n_lut);
  const float tgtL = (kLutGlobal ? __ldg(s_lutL + iL) : s_lutL[iL]) * k.mL;
  const float tgtR = (kLutGlobal ? __ldg(s_lutR + iR) : s_lutR[iR]) * k.mR;

  // ---- physics sub-steps -----------------------------------------------------------------
  float ox = 0.0f, oy = 0.0f;
  if (kDisturb && p.envs_per_row > 0) {
    const int row = (int)(lid / p.envs_per_row), col = (int)(lid % p.envs_per_row);
    ox = p.grid_row_offset - (float)row * p.env_spacing;
    oy = (float)col * p.env_spacing - p.grid_col_offset;
  }
  const float oma = 1.0f - p.lag_alpha;
  const float inv_m = 1.0f / k.mass, inv_iz = 1.0f / (p.izz * k.kiz);
  // heading (cos psi, sin psi): one full evaluation per control step, then advanced by the small per-sub-step yaw
  // increment with a rotation by (cos d, sin d) from short Taylor polynomials (|d| = dt*|r| <= 0.5: error < 5e-9)
  float hsn, hcs;
  fsincos(e.psi, &hsn, &hcs);
  for (int ss = 0; ss < p.n_substeps; ++ss) {
    // apply_forces(): update_forces() advances the lag BEFORE the wrench is applied
    // [ref ThrusterDynamics.py:129-141; SNAP/USV_Virtual.py:640]
    e.thrL = __fadd_rn(__fmul_rn(e.thrL, p.lag_alpha), __fmul_rn(oma, tgtL));
    e.thrR = __fadd_rn(__fmul_rn(e.thrR, p.lag_alpha), __fmul_rn(oma, tgtR));
    float du, dv, dr, Fx, Fy, Tz, ax, ay, rdot;
    planar_wrench<kDisturb>(e, k, p, ox, oy, inv_m, inv_iz, hsn, hcs, du, dv, dr, Fx, Fy, Tz, ax, ay, rdot);
    // world.step(): semi-implicit Euler (velocities first, then positions)
    e.vx += p.dt * ax;
    e.vy += p.dt * ay;
    e.r += p.dt * rdot;
    e.x += p.dt * e.vx;
    e.y += p.dt * e.vy;
    const float dpsi = p.dt * e.r;
    e.psi += dpsi;
    if (fabsf(dpsi) <= 0.5f) {
      const float d2 = dpsi * dpsi;
      const float sd = dpsi * fmaf(d2, fmaf(d2, fmaf(d2, -1.0f / 5040.0f, 1.0f / 120.0f), -1.0f / 6.0f), 1.0f);
      const float cd = fmaf(d2, fmaf(d2, fmaf(d2, fmaf(d2, 1.0f / 40320.0f, -1.0f / 720.0f), 1.0f / 24.0f), -0.5f), 1.0f);
      const float nc = hcs * cd - hsn * sd;
      hsn = hsn * cd + hcs * sd;
      hcs = nc;
    } else {
      fsincos(e.psi, &hsn, &hcs);
    }
  }
  e.psi = wrap_pi(e.psi);

  // ---- post_physics_step ------------------------------------------------------------------
  e.progress += 1;  // [ref rl_task.py:294]
  // update_state: observation noise  [ref SNAP/USV_Virtual.py:476-530 ; USV_disturbances.py:552-601]
  float pxn = e.x, pyn = e.y, vxn = e.vx, vyn = e.vy, wn = e.r, yawn = e.psi;
  if (p.noise_pos) { pxn += urange(u_px, p.pos_noise_min, p.pos_noise_max); pyn += urange(u_py, p.pos_noise_min, p.pos_noise_max); }
  if (p.noise_vel) {
    vxn += urange(u_vx, p.vel_noise_min, p.vel_noise_max);
    vyn += urange(u_vy, p.vel_noise_min, p.vel_noise_max);
    wn += urange(u_w, p.vel_noise_min, p.vel_noise_max);
  }
  if (p.noise_heading) yawn += urange(u_h, p.heading_noise_min, p.heading_noise_max);
  float hs, hc;
  fsincos(yawn, &hs, &hc);
  // get_state_observations  [ref SNAP/USV_capture_xy.py:80-97]
  const float ex = k.tx - pxn, ey = k.ty - pyn;
  const float theta = wrap_pi(yawn);  // == atan2(sin, cos) up to rounding, same (-pi,pi] branch
  const float beta = atan2f(ey, ex);
  // torch.fmod(x, 2pi) has C semantics (sign of the dividend).  beta, theta in [-pi, pi] -> x in [-pi, 3pi]:
  // fmod only acts for x >= 2pi, where x - 2pi is exact; negative x stays unwrapped (reference quirk).
  const float xa = beta - theta + USV_PI_F;
  const float alpha = ((xa >= USV_2PI_F) ? xa - USV_2PI_F : xa) - USV_PI_F;
  const float herr = fabsf(alpha);
  float sa, ca;
  fsincos(alpha, &sa, &ca);
  const float d = sqrtf(ex * ex + ey * ey);
  // Core.update_observation_tensor, "local" frame  [ref SNAP/USV_core.py:41-54]
  o.obs[0] = hc * vxn + hs * vyn;
  o.obs[1] = -hs * vxn + hc * vyn;
  o.obs[2] = wn;
  o.obs[3] = ca;
  o.obs[4] = sa;
  o.obs[5] = d;
  o.obs[6] = vxn;
  o.obs[7] = vyn;
  o.obs[8] = 0.0f;
  o.obs[9] = vxn;
  o.obs[10] = vyn;
  o.obs[11] = 0.0f;
  o.obs[12] = 0.0f;
  // compute_reward  [ref SNAP/USV_capture_xy.py:101-227 ; SNAP/USV_task_rewards.py:40-76]
  const float speed = sqrtf(vxn * vxn + vyn * vyn);
  const int goal = (d < p.position_tolerance) && (speed < p.goal_speed_gate);
  e.goal_cnt = e.goal_cnt * goal + goal;
  float dist_rew;
  if (p.reward_mode == USV_REWARD_LINEAR) {
    dist_rew = p.position_scale * (e.prev_d - d);
  } else if (p.reward_mode == USV_REWARD_SQUARE) {
    dist_rew = p.position_scale * (e.prev_d * e.prev_d - d * d);
  } else {
    dist_rew = p.position_scale * (expf(-d / p.exponential_reward_coeff) - expf(-e.prev_d / p.exponential_reward_coeff));
  }
  const float h2 = herr * herr;
  const float align = p.align_la1 * (__expf(p.align_la2 * (h2 * h2)) + __expf(p.align_la3 * h2));
  if (do_reset) dist_rew = 0.0f;  // distance_reward[just_had_been_reset] = 0
  float speed_rew;
  const float sclamp = 1.0f - fminf(fmaxf(speed / 1.0f, 0.0f), 1.0f);
  if (d > 3.5f) {
    const bool in_range = (speed >= 0.8f) && (speed <= 1.5f);
    const float ds = speed - 1.15f;
    speed_rew = in_range ? 0.1f : __expf(-(ds * ds) * 5.0f) * 0.1f;
  } else if (d > 2.5f) {
    speed_rew = sclamp * 0.15f;
  } else if (d > 1.5f) {
    speed_rew = sclamp * 0.25f;
  } else {
    speed_rew = sclamp * 0.35f;
  }
  if (!(d == d)) speed_rew = 0.0f;  // NaN distance matches no mask in the reference -> zeros_like
  const float goal_rew = (float)e.goal_cnt * p.goal_reward;
  const float task_rew = dist_rew + align + speed_rew + goal_rew + p.time_reward;
  e.prev_d = d;
  // Penalties.compute_penalty  [ref SNAP/USV_task_rewards.py:422-506]
  const float asum = pa0 + pa1;
  const float dw = first_call ? 0.0f : (wn - e.prev_w);
  const float dasum = first_call ? 0.0f : (asum - e.prev_asum);
  float pen_lin = 0.0f, pen_ang = 0.0f, pen_angvar = 0.0f, pen_energy = 0.0f, pen_actvar = 0.0f;
  if (p.pen_linear_vel.form != USV_PEN_OFF) pen_lin = penalty_scalar(p.pen_linear_vel, speed);
  if (p.pen_angular_vel.form != USV_PEN_OFF) pen_ang = penalty_scalar(p.pen_angular_vel, wn);
  if (p.pen_angular_vel_variation.form != USV_PEN_OFF) pen_angvar = penalty_scalar(p.pen_angular_vel_variation, dw);
  if (p.pen_energy.form == USV_PEN_NEG_SUM) pen_energy = -(pa0 + pa1) * p.pen_energy.c1 + p.pen_energy.c2;
  else if (p.pen_energy.form == USV_PEN_EXP_NEG_SUMSQ) pen_energy = (__expf(-(pa0 * pa0 + pa1 * pa1)) - 1.0f) * p.pen_energy.c1;
  if (p.pen_action_variation.form != USV_PEN_OFF) pen_actvar = penalty_scalar(p.pen_action_variation, dasum);
  const float penalties = pen_lin + pen_ang + pen_angvar + pen_energy + pen_actvar;
  e.prev_w = wn;
  e.prev_asum = asum;
  o.rew = task_rew + penalties;  // [ref SNAP/USV_Virtual.py:844]
  // update_kills + is_done  [ref SNAP/USV_capture_xy.py:231-275 ; SNAP/USV_Virtual.py:855-866]
  int die = (d > p.kill_dist) ? 1 : 0;
  if ((e.goal_cnt >= p.kill_after_n_steps_in_tolerance) && (speed < p.goal_speed_gate)) die = 1;
  o.done = (e.progress >= p.max_episode_length - 1) ? 1 : die;
  // diagnostics
  o.dist_rew = dist_rew; o.align_rew = align; o.speed_rew = speed_rew; o.d = d; o.speed = speed;
  o.bdist = d - p.kill_dist;
  o.bpen = -expf(-o.bdist / 0.25f) * p.boundary_cost;
  o.pen_lin = pen_lin; o.pen_ang = pen_ang; o.pen_angvar = pen_angvar; o.pen_energy = pen_energy; o.pen_actvar = pen_actvar;
  o.absw = fabsf(wn); o.asum = asum;
  // NaN probe on the un-clamped obs and the reward [ref vec_env_rlgames.py:187-192]: x*0 is 0 for finite x, NaN otherwise
  float chk = o.rew * 0.0f;
#pragma unroll
  for (int j = 0; j < kObs; ++j) chk = fmaf(o.obs[j], 0.0f, chk);
  o.finite = (chk == 0.0f);
  // _process_data: clamp obs to +-clipObservations  [ref vec_env_rlgames.py:82-95]
#pragma unroll
  for (int j = 0; j < kObs; ++j) o.obs[j] = fminf(fmaxf(o.obs[j], -p.clip_obs), p.clip_obs);
}

__device__ __forceinline__ void accumulate_stats(float* __restrict__ st0, int64_t /*cap*/, int64_t i0, bool was_reset,
                                                 const StepOut& o, const UsvStepParams& p) {
  float* __restrict__ st = st0 + tile_base(i0, USV_ST_COUNT);
  constexpr int i = 0;
  // episode_sums[k][i] += term  [ref SNAP/USV_capture_xy.py:278-306 ; SNAP/USV_task_rewards.py:526-540 ;
  // SNAP/USV_Virtual.py:819-830]; sums are cleared by the reset that precedes this step.
  const float v[USV_ST_COUNT] = {o.dist_rew, o.align_rew, o.speed_rew, o.d, o.speed, o.bpen, o.bdist,
                                 o.pen_lin, o.pen_ang, o.pen_angvar, o.pen_energy, o.pen_actvar,
                                 o.speed, o.absw, o.asum};
#pragma unroll
  for (int f = 0; f < USV_ST_COUNT; ++f) {
    const float prev = was_reset ? 0.0f : st[f * stride + i];
    st[f * stride + i] = prev + v[f];
  }
}

#undef stride

// Each WARP stages its own 32 x 13 observation tile in smem (stride 13 is odd -> conflict-free) and writes it
// out as one contiguous 1664 B run of float4 lines: only a __syncwarp, no CTA barrier on the store path.
__device__ __forceinline__ void write_obs_tile(float* s_obs, const StepOut& o, bool active, float* __restrict__ obs,
                                               int64_t block_start, int64_t n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sw = s_obs + warp * (32 * kObs);
  if (active) {
#pragma unroll
    for (int j = 0; j < kObs; ++j) sw[lane * kObs + j] = o.obs[j];
  }
  __syncwarp();
  const int64_t warp_start = block_start + (int64_t)warp * 32;
  if (warp_start >= n) return;
  const int rows = (int)min((int64_t)32, n - warp_start);
  const int total = rows * kObs;
  float* g = obs + warp_start * kObs;
  if (rows == 32 && (((uintptr_t)g & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(sw);
    float4* g4 = reinterpret_cast<float4*>(g);
#pragma unroll
    for (int q = lane; q < (32 * kObs) / 4; q += 32) g4[q] = s4[q];
  } else {
    for (int q = lane; q < total; q += 32) g[q] = sw[q];
  }
  __syncwarp();
}

template <bool kDisturb, bool kStats>
__global__ void __launch_bounds__(kBlock, USV_MINB) step_fused_kernel(UsvEnvBuffers b, const float2* __restrict__ actions,
                                                            float* __restrict__ obs, float* __restrict__ rew,
                                                            int64_t n, const __grid_constant__ UsvStepParams p) {
  extern __shared__ __align__(16) float smem[];
  float* s_obs = smem;                       // [kBlock*13]
  // The two 4 KB thruster LUTs are gathered at 2 random indices per env-step.  Staging them per CTA cost 16 LDG+STS
  // per thread plus a CTA barrier in this one-tile-per-CTA kernel (r01 profile); the read-only path keeps the 8 KB
  // resident in L1 instead.  (The multi-step rollout kernel, whose CTAs live for T steps, does stage them in smem.)
  const float* __restrict__ s_lutL = b.lut_left;
  const float* __restrict__ s_lutR = b.lut_right;
  const int64_t block_start = (int64_t)blockIdx.x * kBlock;
  const int64_t i = block_start + threadIdx.x;
  const bool active = i < n;
  EnvState e;
  EnvConst k;
  bool do_reset = false;
  float2 act = make_float2(0.f, 0.f);
  if (active) {
    load_state(b.state, b.state_stride, i, e);
    load_consts<kDisturb>(b.consts, b.consts_stride, i, k);
    do_reset = b.reset_buf[i] != 0;
    act = actions[i];
  }
  StepOut o;
  if (active) {
    control_step<kDisturb, true>(e, k, p, do_reset, act, (uint64_t)(p.env_id_offset + i), i, p.step_counter,
                                 p.first_call != 0, s_lutL, s_lutR, o);
    store_state(b.state, b.state_stride, i, e);
    if (do_reset) store_consts<kDisturb>(b.consts, b.consts_stride, i, k);
    if (kStats) accumulate_stats(b.stats, b.stats_stride, i, do_reset, o, p);
    rew[i] = o.rew;
    b.reset_buf[i] = (int64_t)o.done;
    if (b.nonfinite_flag) {
      if (!o.finite) atomicOr(b.nonfinite_flag, 1u);
      // torch.clamp propagates NaN actions and the reference fails fast on them [ref: vec_env_rlgames.py:143];
      // fminf/fmaxf would silently scrub them, so flag them here
      if (!isfinite(act.x) || !isfinite(act.y)) atomicOr(b.nonfinite_flag, 2u);
    }
  }
  write_obs_tile(s_obs, o, active, obs, block_start, n);
}

// T control steps per launch, state in registers between steps.
template <bool kDisturb, bool kStats>
__global__ void __launch_bounds__(kBlock) rollout_fused_kernel(UsvEnvBuffers b, const float2* __restrict__ actions,
                                                               float* __restrict__ obs, float* __restrict__ rew,
                                                               int64_t* __restrict__ done, int T, int64_t n,
                                                               const __grid_constant__ UsvStepParams p) {
  extern __shared__ __align__(16) float smem[];
  float* s_obs = smem;
  float* s_lutL = smem + kBlock * kObs;
  float* s_lutR = s_lutL + p.n_lut;
  for (int t = threadIdx.x; t < p.n_lut; t += kBlock) {
    s_lutL[t] = b.lut_left[t];
    s_lutR[t] = b.lut_right[t];
  }
  const int64_t block_start = (int64_t)blockIdx.x * kBlock;
  const int64_t i = block_start + threadIdx.x;
  const bool active = i < n;
  EnvState e;
  EnvConst k;
  bool do_reset = false;
  bool consts_dirty = false;
  if (active) {
    load_state(b.state, b.state_stride, i, e);
    load_consts<kDisturb>(b.consts, b.consts_stride, i, k);
    do_reset = b.reset_buf[i] != 0;
  }
  __syncthreads();
  bool first_call = p.first_call != 0;
  bool bad = false;
  for (int t = 0; t < T; ++t) {
    StepOut o;
    if (active) {
      const float2 act = actions[(int64_t)t * n + i];
      control_step<kDisturb, false>(e, k, p, do_reset, act, (uint64_t)(p.env_id_offset + i), i, p.step_counter + (uint64_t)t,
                             first_call, s_lutL, s_lutR, o);
      consts_dirty |= do_reset;
      if (kStats) accumulate_stats(b.stats, b.stats_stride, i, do_reset, o, p);
      if (rew) rew[(int64_t)t * n + i] = o.rew;
      if (done) done[(int64_t)t * n + i] = (int64_t)o.done;
      bad |= !o.finite || !isfinite(act.x) || !isfinite(act.y);
      do_reset = o.done != 0;
    }
    first_call = false;
    if (obs) {
      write_obs_tile(s_obs, o, active, obs + (int64_t)t * n * kObs, block_start, n);
    }
  }
  if (active) {
    store_state(b.state, b.state_stride, i, e);
    if (consts_dirty) store_consts<kDisturb>(b.consts, b.consts_stride, i, k);
    b.reset_buf[i] = do_reset ? 1 : 0;
    if (b.nonfinite_flag && bad) atomicOr(b.nonfinite_flag, 1u);
  }
}

template <bool kDisturb>
__global__ void __launch_bounds__(kBlock) planar_forces_kernel(UsvEnvBuffers b, float* __restrict__ out, int64_t n,
                                                               const __grid_constant__ UsvStepParams p) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  EnvState e;
  EnvConst k;
  load_state(b.state, b.state_stride, i, e);
  load_consts<kDisturb>(b.consts, b.consts_stride, i, k);
  float ox = 0.0f, oy = 0.0f;
  if (kDisturb && p.envs_per_row > 0) {
    ox = p.grid_row_offset - (float)(int)(i / p.envs_per_row) * p.env_spacing;
    oy = (float)(int)(i % p.envs_per_row) * p.env_spacing - p.grid_col_offset;
  }
  float du, dv, dr, Fx, Fy, Tz, ax, ay, rdot, hsn, hcs;
  fsincos(e.psi, &hsn, &hcs);
  planar_wrench<kDisturb>(e, k, p, ox, oy, 1.0f / k.mass, 1.0f / (p.izz * k.kiz), hsn, hcs, du, dv, dr, Fx, Fy, Tz, ax, ay, rdot);
  float* o = out + i * 8;
  o[0] = du; o[1] = dv; o[2] = dr; o[3] = Fx; o[4] = Fy; o[5] = Tz; o[6] = ax; o[7] = ay;
}

static int check_common(const UsvEnvBuffers* b, int64_t n, const UsvStepParams* p) {
  if (!b || !p) return USV_E_NULL;
  if (n < 0) return USV_E_SIZE;
  if (!b->state || !b->consts || !b->reset_buf || !b->lut_left || !b->lut_right) return USV_E_NULL;
  if (b->state_stride < n || b->consts_stride < n || (b->state_stride & 31) || (b->consts_stride & 31)) return USV_E_SIZE;
  if (b->stats && (b->stats_stride < n || (b->stats_stride & 31))) return USV_E_SIZE;
  if (p->n_lut < 2 || p->n_lut > 8192) return USV_E_PARAM;
  if (p->n_substeps < 0 || p->n_substeps > 1024) return USV_E_PARAM;
  if (!(p->izz > 0.0f)) return USV_E_PARAM;
  if (p->reward_mode < USV_REWARD_LINEAR || p->reward_mode > USV_REWARD_EXPONENTIAL) return USV_E_PARAM;
  return USV_OK;
}

static bool wants_disturb(const UsvStepParams* p) {
  return p->use_force_disturbance || p->use_torque_disturbance || p->use_const_force || p->use_sin_force ||
         p->use_const_torque || p->use_sin_torque;
}

static size_t step_smem(const UsvStepParams* p) { return (size_t)(kBlock * kObs + 2 * p->n_lut) * sizeof(float); }
static size_t step_smem_nolut() { return (size_t)(kBlock * kObs) * sizeof(float); }

template <typename K>
static void ensure_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace usv

using namespace usv;

extern "C" {

int usv_step_fused_f32(const UsvEnvBuffers* b, const float* actions, float* obs, float* rew, int64_t n,
                       const UsvStepParams* p, void* stream) {
  int rc = check_common(b, n, p);
  if (rc) return rc;
  if (n == 0) return USV_OK;
  if (!actions || !obs || !rew) return USV_E_NULL;
  if ((uintptr_t)actions & 7) return USV_E_ALIGN;
  const size_t smem = step_smem_nolut();
  const int grid = grid_for(n, kBlock);
  const bool dis = wants_disturb(p), st = b->stats != nullptr;
  cudaStream_t s = (cudaStream_t)stream;
#define USV_LAUNCH_STEP(D, S)                                                                              \
  do {                                                                                                     \
    ensure_smem(step_fused_kernel<D, S>, smem);                                                            \
    step_fused_kernel<D, S><<<grid, kBlock, smem, s>>>(*b, (const float2*)actions, obs, rew, n, *p);       \
  } while (0)
  if (dis && st) USV_LAUNCH_STEP(true, true);
  else if (dis) USV_LAUNCH_STEP(true, false);
  else if (st) USV_LAUNCH_STEP(false, true);
  else USV_LAUNCH_STEP(false, false);
#undef USV_LAUNCH_STEP
  return finish_launch();
}

int usv_rollout_fused_f32(const UsvEnvBuffers* b, const float* actions, float* obs, float* rew, int64_t* done,
                          int32_t T, int64_t n, const UsvStepParams* p, void* stream) {
  int rc = check_common(b, n, p);
  if (rc) return rc;
  if (T < 0) return USV_E_SIZE;
  if (n == 0 || T == 0) return USV_OK;
  if (!actions) return USV_E_NULL;
  if ((uintptr_t)actions & 7) return USV_E_ALIGN;
  const size_t smem = step_smem(p);
  const int grid = grid_for(n, kBlock);
  const bool dis = wants_disturb(p), st = b->stats != nullptr;
  cudaStream_t s = (cudaStream_t)stream;
#define USV_LAUNCH_ROLL(D, S)                                                                              \
  do {                                                                                                     \
    ensure_smem(rollout_fused_kernel<D, S>, smem);                                                         \
    rollout_fused_kernel<D, S><<<grid, kBlock, smem, s>>>(*b, (const float2*)actions, obs, rew, done, T, n, *p); \
  } while (0)
  if (dis && st) USV_LAUNCH_ROLL(true, true);
  else if (dis) USV_LAUNCH_ROLL(true, false);
  else if (st) USV_LAUNCH_ROLL(false, true);
  else USV_LAUNCH_ROLL(false, false);
#undef USV_LAUNCH_ROLL
  return finish_launch();
}

int usv_planar_forces_f32(const UsvEnvBuffers* b, float* out, int64_t n, const UsvStepParams* p, void* stream) {
  int rc = check_common(b, n, p);
  if (rc) return rc;
  if (n == 0) return USV_OK;
  if (!out) return USV_E_NULL;
  const int grid = grid_for(n, kBlock);
  if (wants_disturb(p))
    planar_forces_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(*b, out, n, *p);
  else
    planar_forces_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>(*b, out, n, *p);
  return finish_launch();
}

}  // extern "C"
