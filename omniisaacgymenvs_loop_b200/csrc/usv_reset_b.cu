// Scene rebuild of the LIVE CaptureXY task for the envs that reset: obstacle placement by rejection
// sampling  [ref: OIGE/tasks/USV/USV_capture_xy_static_obs.py:936-1060]  and the per-env potential field of BatchedMapGPU
// [ref: OIGE/tasks/USV/d_multi_gemini.py:66-104 occupancy + SDF, :135-192 cost-to-go wavefront, :194-271 potential field].
//
// The reference runs 225 Jacobi sweeps of an 8-neighbour min-plus relaxation as ~30 torch ops per sweep over the whole
// (B,150,150) batch in HBM.  Here ONE CTA owns one env: both 150x150 cost buffers live in shared memory (2 x 92 KB with an
// +inf halo), a sweep is one pass over smem, and the CTA stops as soon as a sweep changes nothing (Jacobi iterates are then
// fixed, so the result is bit-identical to running all 225 sweeps).  The reference's BATCH-GLOBAL maxima (max finite cost,
// max repulsion: quirk 9 of SURVEY appendix C) need two grid-wide reductions, hence three kernels:
//   scene_cost_kernel  : obstacles (warp 0) -> free mask -> wavefront -> raw cost into field[env], atomicMax(max cost)
//   scene_jmax_kernel  : repulsion J per cell with the global max cost -> atomicMax(max J), any-inside flag
//   scene_field_kernel : per-env min/max of G and J -> field[env] = norm(G) + 0.5 norm(J)
// All three are persistent grids striding over the compacted list of resetting envs (or over a dense batch for the
// standalone builder entry), so their cost is ~0 when nothing resets.
#include <math_constants.h>
#include "philox.cuh"
#include "usv_step_core.cuh"

namespace usv {

constexpr int kG = USV_B_GRID;       // 150
constexpr int kGP = kG + 2;          // padded row (halo of +inf)
constexpr int kCells = kG * kG;
constexpr int kSceneThreads = 1024;  // 32 warps: warp w owns rows w, w+32, ...; lane l owns columns l, l+32, ...
constexpr int kRowIters = (kG + 31) / 32;  // 5
constexpr int kColIters = (kG + 31) / 32;  // 5
constexpr int kMaxSweeps = (int)(kG * 1.5);  // 225  (d_multi_gemini.py:160)
constexpr float kObstR = 0.5f;
constexpr int kChunks = 5, kBandRows = 8, kBands = (kG + kBandRows - 1) / kBandRows, kTiles = kChunks * kBands;  // wavefront tiles: 32 columns x 8 rows (95)
constexpr int kActStride = 96;       // tile-active flags per sweep parity
constexpr float kReach = 1.7f + 1e-3f;  // an obstacle matters to a cell only within 0.5 (radius) + 0.5 + 0.7 (influence) of its centre

// counters (uint32[8]) in the workspace
enum { CW_COUNT = 0, CW_MAXCOST = 1, CW_MAXJ = 2, CW_INSIDE = 3, CW_HAVE = 4, CW_WORDS = 8 };

struct SceneIO {
  // list mode (list != nullptr): env id = list[j]; obstacles in bconsts (AoSoA), target in consts, field[env]
  const int32_t* list;
  const uint32_t* count;  // number of list entries (device)
  float* bconsts;
  const float* consts;
  // dense mode: j-th scene of a batch of m
  const float* obstacles;  // [m,16,2]
  const float* targets;    // [m,2]
  int64_t m;
  float* field;            // [*,150,150]
  float* cost_out;         // dense mode only, or NULL
  const float* lin;        // [150] cell-centre coordinates (torch.linspace(-14.9, 14.9, 150))
};

__device__ __forceinline__ int64_t scene_count(const SceneIO& io) { return io.list ? (int64_t)*io.count : io.m; }

__device__ __forceinline__ float block_reduce_max(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  v = s_red[lane];  // kSceneThreads / 32 == 32 partials
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float block_reduce_min(float v, float* s_red) { return -block_reduce_max(-v, s_red); }

// min over the 16 obstacles of the centre distance, minus the radius  (d_multi_gemini.py:84-92).  torch.norm over a
// 2-vector evaluates sqrt(fma(y, y, x*x)) (ATen's sum-of-squares accumulation, CPU and CUDA): spelled out, because the
// occupancy is the sign of this and the repulsion term amplifies 1 ulp of it by ~1e3 next to an obstacle
__device__ __forceinline__ float cell_sdf(float cx, float cy, const float* s_ox, const float* s_oy) {
  float best = CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < USV_B_OBSTACLES; ++j) {
    const float dx = __fsub_rn(cx, s_ox[j]), dy = __fsub_rn(cy, s_oy[j]);
    best = fminf(best, sqrtf(__fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
  }
  return __fsub_rn(best, kObstR);
}

// bit j set: obstacle j is within kReach of grid row cy (|dy| <= dist), i.e. it can make a cell of that row occupied / repelled
__device__ __forceinline__ uint32_t row_obstacle_mask(float cy, const float* s_oy) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < USV_B_OBSTACLES; ++j) m |= (fabsf(cy - s_oy[j]) <= kReach) ? (1u << j) : 0u;
  return m;
}
// cell_sdf over the obstacles of `mask` only: exact wherever the result is below kReach - 0.5, +inf-ish (>= that) elsewhere --
// every consumer only distinguishes values below 1.2 (occupancy: <= 0; repulsion: sdf - 0.5 < 0.7)
__device__ __forceinline__ float cell_sdf_masked(float cx, float cy, const float* s_ox, const float* s_oy, uint32_t mask) {
  float best = CUDART_INF_F;
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1;
    const float dx = __fsub_rn(cx, s_ox[j]), dy = __fsub_rn(cy, s_oy[j]);
    best = fminf(best, sqrtf(__fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
  }
  return __fsub_rn(best, kObstR);
}

// stage the 16 obstacle centres + target of scene j in smem (s_sc[0..15]=x, [16..31]=y, [32]=tx, [33]=ty); returns env id
__device__ __forceinline__ int64_t load_scene(const SceneIO& io, int64_t j, float* s_sc) {
  const int64_t env = io.list ? (int64_t)io.list[j] : j;
  if (threadIdx.x < USV_B_OBSTACLES) {
    const int t = threadIdx.x;
    if (io.list) {
      const float* bc = io.bconsts + tile_base(env, USV_BC_COUNT);
      s_sc[t] = bc[(USV_BC_OBST + 2 * t) * kTile];
      s_sc[16 + t] = bc[(USV_BC_OBST + 2 * t + 1) * kTile];
    } else {
      s_sc[t] = io.obstacles[(j * USV_B_OBSTACLES + t) * 2];
      s_sc[16 + t] = io.obstacles[(j * USV_B_OBSTACLES + t) * 2 + 1];
    }
  } else if (threadIdx.x == 32) {
    if (io.list) {
      const float* c = io.consts + tile_base(env, USV_C_COUNT);
      s_sc[32] = c[USV_C_TX * kTile];
      s_sc[33] = c[USV_C_TY * kTile];
    } else {
      s_sc[32] = io.targets[j * 2];
      s_sc[33] = io.targets[j * 2 + 1];
    }
  }
  return env;
}

// get_spawns obstacle placement for one env by warp 0 (lanes 0..15 = obstacles)  [ref :970-1047]
__device__ __forceinline__ void place_obstacles(const UsvStepParams& p, uint64_t step, uint64_t gid, float tx, float ty, float& ox, float& oy) {
  const int j = threadIdx.x & 31;
  const Uniform4 r0 = philox_uniform4(p.seed, gid, step, RS_RESET_0);
  float sx, sy;
  spawn_xy(p, r0, tx, ty, sx, sy);  // the same draw reset_env() of the step kernel will use for the pose
  const float mnx = tx - 12.0f, mny = ty - 12.0f;
  const float spx = (tx + 12.0f) - mnx, spy = (ty + 12.0f) - mny;
  auto draw = [&](int round) {
    const Uniform4 u = philox_uniform4(p.seed, gid, step, RS_OBST + (uint32_t)round * 8u + (uint32_t)((j & 15) >> 1));
    ox = __fadd_rn(__fmul_rn((j & 1) ? u.c : u.a, spx), mnx);  // rand * (max - min) + min: two roundings, as torch
    oy = __fadd_rn(__fmul_rn((j & 1) ? u.d : u.b, spy), mny);
  };
  auto invalid = [&]() {
    const float dsx = ox - sx, dsy = oy - sy, dtx = ox - tx, dty = oy - ty;
    bool inv = (sqrtf(__fmaf_rn(dsy, dsy, __fmul_rn(dsx, dsx))) < 3.0f) || (sqrtf(__fmaf_rn(dty, dty, __fmul_rn(dtx, dtx))) < 3.0f);
    const bool vj = ox < 900.0f;
#pragma unroll
    for (int i = 0; i < USV_B_OBSTACLES; ++i) {
      const float xi = __shfl_sync(0xffffffffu, ox, i), yi = __shfl_sync(0xffffffffu, oy, i);
      const float ddx = ox - xi, ddy = oy - yi;
      // S1: only the higher index of a conflicting pair is re-drawn
      if (i < j && vj && xi < 900.0f && __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)) < 2.5f * 2.5f) inv = true;
    }
    return inv && j < USV_B_OBSTACLES;
  };
  draw(0);
  for (int it = 0; it < 20; ++it) {
    const bool inv = invalid();
    if (!(__ballot_sync(0xffffffffu, inv) & 0xffffu)) break;
    if (inv) draw(it + 1);
  }
  if (invalid()) { ox = 999.0f; oy = 999.0f; }  // leftovers go to limbo
}

__global__ void __launch_bounds__(kSceneThreads, 1) scene_cost_kernel(SceneIO io, uint32_t* __restrict__ counters, int place,
                                                                      const __grid_constant__ UsvStepParams p,
                                                                      const uint64_t* __restrict__ step_offset) {
  const uint64_t step = p.step_counter + (step_offset ? *step_offset : 0ull);   // device-side addend: CUDA-graph replays
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;
  float* bufB = smem + kGP * kGP;
  float* s_sc = bufB + kGP * kGP;     // 34 floats
  float* s_red = s_sc + 64;           // 32 floats
  int* s_act = reinterpret_cast<int*>(s_red + 32);        // [2][96] tile-active flags of the current / next sweep
  uint32_t* s_rowmask = reinterpret_cast<uint32_t*>(s_act + 2 * kActStride);  // [150] obstacles that can matter on a grid row
  unsigned char* s_free = reinterpret_cast<unsigned char*>(s_rowmask + 160);  // [95][32] free bits of a lane's 8 cells of a tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = scene_count(io);
  for (int64_t j = blockIdx.x; j < total; j += gridDim.x) {
    int64_t env = io.list ? (int64_t)io.list[j] : j;
    if (place) {
      // B5: re-draw the obstacles of this env (list mode only), publish them to bconsts and smem
      if (warp == 0) {
        const float* c = io.consts + tile_base(env, USV_C_COUNT);
        const float tx = c[USV_C_TX * kTile], ty = c[USV_C_TY * kTile];
        float ox = 0.0f, oy = 0.0f;
        place_obstacles(p, step, (uint64_t)(p.env_id_offset + env), tx, ty, ox, oy);
        if (lane < USV_B_OBSTACLES) {
          float* bc = io.bconsts + tile_base(env, USV_BC_COUNT);
          bc[(USV_BC_OBST + 2 * lane) * kTile] = ox;
          bc[(USV_BC_OBST + 2 * lane + 1) * kTile] = oy;
          s_sc[lane] = ox;
          s_sc[16 + lane] = oy;
        }
        if (lane == 0) { s_sc[32] = tx; s_sc[33] = ty; }
      }
    } else {
      env = load_scene(io, j, s_sc);
    }
    // both buffers +inf (halo included)
    for (int q = threadIdx.x; q < kGP * kGP; q += kSceneThreads) { bufA[q] = CUDART_INF_F; bufB[q] = CUDART_INF_F; }
    if (threadIdx.x < 2 * kActStride) s_act[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < kG) s_rowmask[threadIdx.x] = row_obstacle_mask(io.lin[threadIdx.x], s_sc + 16);
    if (threadIdx.x == 0) {
      // target cell: ((pos + map/2) / cell).long().clamp(0, 149)   (d_multi_gemini.py:148-155).  It is written into BOTH buffers:
      // the tile-skipping below relies on "a tile that is not swept holds the same values in both buffers"
      const int txi = min(max((int)__fdiv_rn(s_sc[32] + 15.0f, 0.2f), 0), kG - 1);
      const int tyi = min(max((int)__fdiv_rn(s_sc[33] + 15.0f, 0.2f), 0), kG - 1);
      bufA[(tyi + 1) * kGP + txi + 1] = 0.0f;
      bufB[(tyi + 1) * kGP + txi + 1] = 0.0f;
      const int tc = txi >> 5, tb = tyi / kBandRows;
      for (int db = -1; db <= 1; ++db)
        for (int dc = -1; dc <= 1; ++dc)
          if (tc + dc >= 0 && tc + dc < kChunks && tb + db >= 0 && tb + db < kBands) s_act[(tb + db) * kChunks + tc + dc] = 1;
    }
    __syncthreads();
    // Tiles of 32 columns x 8 rows (95 per grid), dealt round-robin to the 32 warps so that the ring of tiles the wavefront is
    // crossing is spread over all of them (one 32 x 25 tile per warp left most warps waiting at the sweep barrier: 7 barrier
    // stalls per issue in the r01 capture).  A lane walks down its column of the tile with a rolling 3x3 window in registers:
    // 3 shared loads + 1 store per cell instead of 9 + 1.  The per-cell "free" bits of a tile live in shared memory.
    for (int t = warp; t < kTiles; t += 32) {
      const int chunk = t % kChunks, band = t / kChunks;
      const int x = chunk * 32 + lane, y0 = band * kBandRows;
      uint32_t fm = 0;
      if (x < kG) {
#pragma unroll 1
        for (int i = 0; i < kBandRows && y0 + i < kG; ++i) {
          const int y = y0 + i;
          const bool border = (y == 0) || (y == kG - 1) || (x == 0) || (x == kG - 1);
          const float sdf = cell_sdf_masked(io.lin[x], io.lin[y], s_sc, s_sc + 16, s_rowmask[y]);
          if (!border && !(sdf <= 0.0f)) fm |= 1u << i;
        }
      }
      s_free[t * 32 + lane] = (unsigned char)fm;
    }
    __syncthreads();
    float* cur = bufA;
    float* nxt = bufB;
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
      int* act_cur = s_act + (sweep & 1) * kActStride;
      int* act_nxt = s_act + ((sweep + 1) & 1) * kActStride;
      int any_change = 0;
      for (int t = warp; t < kTiles; t += 32) {
        if (!act_cur[t]) continue;                 // warp-uniform
        __syncwarp();
        if (lane == 0) act_cur[t] = 0;             // consumed; writers of this sweep only touch act_nxt
        const int chunk = t % kChunks, band = t / kChunks;
        const int x = chunk * 32 + lane, y0 = band * kBandRows;
        const int rows = min(kBandRows, kG - y0);
        const bool xin = x < kG;
        const int xc = min(x, kG - 1);
        const uint32_t freemask = s_free[t * 32 + lane];
        const float* c0 = cur + y0 * kGP + xc + 1;  // (row y0 - 1, column x) in padded coordinates
        float* n0 = nxt + (y0 + 1) * kGP + xc + 1;
        float ul = c0[-1], uc = c0[0], ur = c0[1];
        float ml = c0[kGP - 1], mc = c0[kGP], mr = c0[kGP + 1];
        uint32_t chg = 0;
#pragma unroll
        for (int i = 0; i < kBandRows; ++i) {
          if (i < rows) {
            const float* d = c0 + (i + 2) * kGP;
            const float dl = d[-1], dc = d[0], dr = d[1];
            float best = CUDART_INF_F;
            if ((freemask >> i) & 1u) {
              const float a = fminf(fminf(ml, mr), fminf(uc, dc)) + 1.0f;
              const float b = fminf(fminf(ul, ur), fminf(dl, dr)) + 1.414f;
              best = fminf(mc, fminf(a, b));
            }
            if (xin) {
              n0[i * kGP] = best;
              chg |= (best != mc) ? (1u << i) : 0u;
            }
            ul = ml; uc = mc; ur = mr;
            ml = dl; mc = dc; mr = dr;
          }
        }
        // activate for the next sweep exactly the tiles whose inputs changed: this one, and a neighbour only when a cell on the
        // shared edge changed
        const uint32_t any_b = __ballot_sync(0xffffffffu, chg != 0u);
        if (any_b) {
          any_change = 1;
          const bool left = any_b & 1u, right = (any_b >> 31) & 1u;
          const bool up = __any_sync(0xffffffffu, chg & 1u), down = __any_sync(0xffffffffu, (chg >> (rows - 1)) & 1u);
          if (lane < 9) {
            const int dc = lane % 3 - 1, db = lane / 3 - 1;
            const bool need = (dc == 0 || (dc < 0 ? left : right)) && (db == 0 || (db < 0 ? up : down));
            const int c2 = chunk + dc, b2 = band + db;
            if (need && c2 >= 0 && c2 < kChunks && b2 >= 0 && b2 < kBands) act_nxt[b2 * kChunks + c2] = 1;
          }
        }
      }
      const int any = __syncthreads_or(any_change);
      float* tsw = cur; cur = nxt; nxt = tsw;
      if (!any) break;
    }
    // raw cost -> field[env] (and the optional dense cost output); max finite cost of the whole batch
    float mx = -1.0f;
    float* out = io.field + env * (int64_t)kCells;
#pragma unroll 1
    for (int ri = 0; ri < kRowIters; ++ri) {
      const int y = warp + 32 * ri;
      if (y < kG) {
#pragma unroll
        for (int ci = 0; ci < kColIters; ++ci) {
          const int x = lane + 32 * ci;
          if (x < kG) {
            const float c = cur[(y + 1) * kGP + x + 1];
            out[y * kG + x] = c;
            if (io.cost_out) io.cost_out[j * (int64_t)kCells + y * kG + x] = c;
            if (c < CUDART_INF_F) mx = fmaxf(mx, c);
          }
        }
      }
    }
    mx = block_reduce_max(mx, s_red);
    if (threadIdx.x == 0 && mx >= 0.0f) {  // non-negative floats order like their bit patterns
      atomicMax(counters + CW_MAXCOST, __float_as_uint(mx));
      atomicOr(counters + CW_HAVE, 1u);
    }
    __syncthreads();
  }
}

struct CellTerms { float vis, J, edge; };

__device__ __forceinline__ CellTerms cell_terms(float cost, float sdf, float vis_inf) {
  CellTerms t;
  t.vis = (cost == CUDART_INF_F) ? vis_inf : cost;  // torch.where(isinf(cost), max_val*1.5, cost)
  t.edge = __fsub_rn(sdf, kObstR);                  // dist_to_edge = sdf - obstacle_radius  (:220)
  const float rep = fminf(fmaxf(__fdiv_rn(__fmul_rn(t.vis, 0.2f), 3.0f), 0.0f), 1.0f);
  t.J = 0.0f;
  if (t.edge < 0.7f) {
    const float dcl = fmaxf(t.edge, 1e-3f);
    const float q = __fsub_rn(__fdiv_rn(1.0f, dcl), 1.42857146f);  // 1.0/d - fp32(1/0.7)
    t.J = __fmul_rn(__fmul_rn(20.0f, __fmul_rn(q, q)), rep);
  }
  return t;
}

__device__ __forceinline__ float global_vis_inf(const uint32_t* counters, int have) {
  // max_val = cost[finite].max() over the BATCH (fallback 100), times 1.5
  const float max_val = have ? __uint_as_float(counters[CW_MAXCOST]) : 100.0f;
  return __fmul_rn(max_val, 1.5f);
}

__global__ void __launch_bounds__(kSceneThreads, 1) scene_jmax_kernel(SceneIO io, uint32_t* __restrict__ counters) {
  __shared__ float s_sc[64];
  __shared__ float s_red[32];
  __shared__ uint32_t s_rowmask[160];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = scene_count(io);
  const float vis_inf = global_vis_inf(counters, counters[CW_HAVE] != 0u);
  for (int64_t j = blockIdx.x; j < total; j += gridDim.x) {
    const int64_t env = load_scene(io, j, s_sc);
    __syncthreads();
    if (threadIdx.x < kG) s_rowmask[threadIdx.x] = row_obstacle_mask(io.lin[threadIdx.x], s_sc + 16);
    __syncthreads();
    const float* f = io.field + env * (int64_t)kCells;
    float mj = 0.0f;
    int inside = 0;
#pragma unroll 1
    for (int ri = 0; ri < kRowIters; ++ri) {
      const int y = warp + 32 * ri;
      if (y < kG) {
#pragma unroll
        for (int ci = 0; ci < kColIters; ++ci) {
          const int x = lane + 32 * ci;
          if (x < kG) {
            const CellTerms t = cell_terms(f[y * kG + x], cell_sdf_masked(io.lin[x], io.lin[y], s_sc, s_sc + 16, s_rowmask[y]), vis_inf);
            mj = fmaxf(mj, t.J);
            inside |= (t.edge <= 0.0f) ? 1 : 0;
          }
        }
      }
    }
    mj = block_reduce_max(mj, s_red);
    inside = __syncthreads_or(inside);
    if (threadIdx.x == 0) {
      atomicMax(counters + CW_MAXJ, __float_as_uint(mj));
      if (inside) atomicOr(counters + CW_INSIDE, 1u);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSceneThreads, 1) scene_field_kernel(SceneIO io, const uint32_t* __restrict__ counters) {
  __shared__ float s_sc[64];
  __shared__ float s_red[32];
  __shared__ uint32_t s_rowmask[160];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t total = scene_count(io);
  const float vis_inf = global_vis_inf(counters, counters[CW_HAVE] != 0u);
  const bool any_inside = counters[CW_INSIDE] != 0u;
  const float cur_max = __uint_as_float(counters[CW_MAXJ]);
  const float high = (cur_max > 1e-6f) ? __fmul_rn(cur_max, 10.0f) : 100.0f;
  for (int64_t j = blockIdx.x; j < total; j += gridDim.x) {
    const int64_t env = load_scene(io, j, s_sc);
    __syncthreads();
    if (threadIdx.x < kG) s_rowmask[threadIdx.x] = row_obstacle_mask(io.lin[threadIdx.x], s_sc + 16);
    __syncthreads();
    float* f = io.field + env * (int64_t)kCells;
    float gmn = CUDART_INF_F, gmx = -CUDART_INF_F, jmn = CUDART_INF_F, jmx = -CUDART_INF_F;
#pragma unroll 1
    for (int ri = 0; ri < kRowIters; ++ri) {
      const int y = warp + 32 * ri;
      if (y < kG) {
#pragma unroll
        for (int ci = 0; ci < kColIters; ++ci) {
          const int x = lane + 32 * ci;
          if (x < kG) {
            CellTerms t = cell_terms(f[y * kG + x], cell_sdf_masked(io.lin[x], io.lin[y], s_sc, s_sc + 16, s_rowmask[y]), vis_inf);
            if (any_inside && t.edge <= 0.0f) t.J = high;
            gmn = fminf(gmn, t.vis); gmx = fmaxf(gmx, t.vis);
            jmn = fminf(jmn, t.J); jmx = fmaxf(jmx, t.J);
          }
        }
      }
    }
    gmn = block_reduce_min(gmn, s_red);
    gmx = block_reduce_max(gmx, s_red);
    jmn = block_reduce_min(jmn, s_red);
    jmx = block_reduce_max(jmx, s_red);
    const float gden = __fadd_rn(__fsub_rn(gmx, gmn), 1e-6f), jden = __fadd_rn(__fsub_rn(jmx, jmn), 1e-6f);
#pragma unroll 1
    for (int ri = 0; ri < kRowIters; ++ri) {
      const int y = warp + 32 * ri;
      if (y < kG) {
#pragma unroll
        for (int ci = 0; ci < kColIters; ++ci) {
          const int x = lane + 32 * ci;
          if (x < kG) {
            CellTerms t = cell_terms(f[y * kG + x], cell_sdf_masked(io.lin[x], io.lin[y], s_sc, s_sc + 16, s_rowmask[y]), vis_inf);
            if (any_inside && t.edge <= 0.0f) t.J = high;
            const float gn = __fdiv_rn(__fsub_rn(t.vis, gmn), gden);
            const float jn = __fdiv_rn(__fsub_rn(t.J, jmn), jden);
            f[y * kG + x] = __fadd_rn(gn, __fmul_rn(0.5f, jn));
          }
        }
      }
    }
    __syncthreads();
  }
}

// reset_buf -> compact list of env ids (order irrelevant: scenes are independent, the batch maxima commutative)
__global__ void compact_resets_kernel(const int64_t* __restrict__ reset_buf, int64_t n, int32_t* __restrict__ list,
                                      uint32_t* __restrict__ counters) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (reset_buf[i] != 0) list[atomicAdd(counters + CW_COUNT, 1u)] = (int32_t)i;
  }
}

static size_t cost_smem_bytes() { return (size_t)(2 * kGP * kGP + 64 + 32 + 2 * kActStride + 160) * sizeof(float) + (size_t)kTiles * 32; }

static int scene_grid() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

static int launch_scene(const SceneIO& io, uint32_t* counters, int place, const UsvStepParams* p, const uint64_t* step_offset,
                        cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(scene_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cost_smem_bytes());
    attr = true;
  }
  const int grid = scene_grid();
  scene_cost_kernel<<<grid, kSceneThreads, cost_smem_bytes(), s>>>(io, counters, place, *p, step_offset);
  scene_jmax_kernel<<<grid, kSceneThreads, 0, s>>>(io, counters);
  scene_field_kernel<<<grid, kSceneThreads, 0, s>>>(io, counters);
  return finish_launch(3);
}

}  // namespace usv

using namespace usv;

extern "C" {

int64_t usv_live_scene_workspace_bytes(int64_t n) { return (int64_t)CW_WORDS * 4 + (n < 0 ? 0 : n) * 4; }

int usv_live_reset_scene_f32(const UsvEnvBuffers* b, const UsvLiveBuffers* lb, const float* cell_centres, void* workspace,
                             int64_t n, const UsvStepParams* p, void* stream) {
  if (!b || !lb || !p || !cell_centres || !workspace) return USV_E_NULL;
  if (n < 0 || n > 0x7fffffffLL) return USV_E_SIZE;
  if (!b->consts || !b->reset_buf || !lb->bconsts || !lb->field) return USV_E_NULL;
  if (b->consts_stride < n || lb->bconsts_stride < n || (b->consts_stride & 31) || (lb->bconsts_stride & 31)) return USV_E_SIZE;
  if ((uintptr_t)workspace & 15) return USV_E_ALIGN;
  if (n == 0) return USV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* counters = (uint32_t*)workspace;
  int32_t* list = (int32_t*)(counters + CW_WORDS);
  cudaMemsetAsync(counters, 0, CW_WORDS * sizeof(uint32_t), s);
  const int cgrid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  compact_resets_kernel<<<cgrid, 256, 0, s>>>(b->reset_buf, n, list, counters);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  SceneIO io{};
  io.list = list;
  io.count = counters + CW_COUNT;
  io.bconsts = lb->bconsts;
  io.consts = b->consts;
  io.field = lb->field;
  io.lin = cell_centres;
  return launch_scene(io, counters, 1, p, b->step_offset, s);
}

int usv_live_build_fields_f32(const float* obstacles, const float* targets, const float* cell_centres, float* field,
                              float* cost_out, void* workspace, int64_t m, void* stream) {
  if (!obstacles || !targets || !cell_centres || !field || !workspace) return USV_E_NULL;
  if (m < 0) return USV_E_SIZE;
  if ((uintptr_t)workspace & 15) return USV_E_ALIGN;
  if (m == 0) return USV_OK;
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* counters = (uint32_t*)workspace;
  cudaMemsetAsync(counters, 0, CW_WORDS * sizeof(uint32_t), s);
  SceneIO io{};
  io.obstacles = obstacles;
  io.targets = targets;
  io.m = m;
  io.field = field;
  io.cost_out = cost_out;
  io.lin = cell_centres;
  UsvStepParams p{};
  return launch_scene(io, counters, 0, &p, nullptr, s);
}

}  // extern "C"
