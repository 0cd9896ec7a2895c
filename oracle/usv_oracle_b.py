"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the live CaptureXY task with static obstacles.

torch-on-CPU restatement, one function per reference function, pinned against tests/golden/capture_xy_live.npz (produced by
running the reference's own CaptureXYTask / BatchedMapGPU under oracle/ref_shim.py).
OIGE = omniisaacgymenvs/ ; file = OIGE/tasks/USV/USV_capture_xy_static_obs.py unless stated otherwise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import philox
from .usv_oracle import ClassicEnvOracle, EnvConfig, com_disc, penalties

GRID, MAP_SIZE, OBST_R = 150, 30.0, 0.5          # :30-32
CELL = MAP_SIZE / GRID
N_OBST, N_CLOSEST = 16, 5
COLLISION_TH = 1.2                               # :103
F32 = torch.float32


@dataclass
class LiveTaskConfig:
    """IROS2024/USV_Virtual_CaptureXY_SysID-TEST.yaml task/reward section + the live reward dataclass defaults
    [OIGE/tasks/USV/USV_task_rewards.py:27-32]."""
    position_tolerance: float = 1.0
    kill_after_n_steps_in_tolerance: int = 1
    kill_dist: float = 20.0
    boundary_cost: float = 25.0
    goal_reward: float = 20.0
    time_reward: float = -0.05
    position_scale: float = 1.5
    align_la1: float = 0.04
    align_la2: float = -10.0
    align_la3: float = -0.1
    spawn_min_dist: float = 9.0
    spawn_max_dist: float = 12.0
    goal_random_position: float = 0.0


# B1  get_state_observations :193-299 + Core.update_observation_tensor OIGE/tasks/USV/USV_core.py:55-125
def live_observation(state, target, obstacles, prev_action, priv):
    n = target.shape[0]
    err = target - state["position"]
    theta = torch.atan2(state["orientation"][:, 1], state["orientation"][:, 0])
    beta = torch.atan2(err[:, 1], err[:, 0])
    alpha = torch.fmod(beta - theta + math.pi, 2 * math.pi) - math.pi
    td = torch.zeros((n, 5 + 3 * N_CLOSEST), dtype=F32)
    td[:, 0], td[:, 1], td[:, 2] = torch.cos(alpha), torch.sin(alpha), torch.norm(err, dim=1)
    rel = obstacles - state["position"].unsqueeze(1)                      # (n,16,2)
    dist = torch.norm(rel, dim=-1)
    cd, ci = torch.topk(dist, k=N_CLOSEST, dim=1, largest=False)          # ascending centre distance
    cv = torch.gather(rel, 1, ci.unsqueeze(-1).expand(-1, -1, 2))
    ct, st = torch.cos(theta).unsqueeze(1), torch.sin(theta).unsqueeze(1)
    xb = cv[:, :, 0] * ct + cv[:, :, 1] * st
    yb = -cv[:, :, 0] * st + cv[:, :, 1] * ct
    nf = torch.sqrt(xb ** 2 + yb ** 2 + 1e-6)
    for i in range(N_CLOSEST):
        td[:, 5 + 3 * i] = cd[:, i] - OBST_R
        td[:, 6 + 3 * i] = -xb[:, i] / nf[:, i]
        td[:, 7 + 3 * i] = -yb[:, i] / nf[:, i]
    obs = torch.zeros((n, 33), dtype=F32)
    c, s = state["orientation"][:, 0], state["orientation"][:, 1]
    v = state["linear_velocity"]
    obs[:, 0] = c * v[:, 0] + s * v[:, 1]
    obs[:, 1] = -s * v[:, 0] + c * v[:, 1]
    obs[:, 2] = state["angular_velocity"]
    obs[:, 3:23] = td
    obs[:, 23:25] = prev_action
    obs[:, 25:33] = priv
    return obs, {"err": err, "alpha": alpha, "herr": torch.abs(alpha), "d": torch.sqrt(torch.square(err).sum(-1))}


# B2  _get_potential_values :302-326: F.grid_sample(bilinear, align_corners=False, padding_mode='border'), restated
def sample_potential(field, pos):
    n, Hh, Ww = field.shape
    gx, gy = 2.0 * pos[:, 0] / MAP_SIZE, 2.0 * pos[:, 1] / MAP_SIZE
    ix = torch.clamp(((gx + 1) * Ww - 1) / 2, 0, Ww - 1)                 # un-normalise, then clip to the border
    iy = torch.clamp(((gy + 1) * Hh - 1) / 2, 0, Hh - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wx1, wy1 = ix - x0, iy - y0
    wx0, wy0 = 1 - wx1, 1 - wy1
    b = torch.arange(n)

    def tap(yy, xx):
        inb = (xx >= 0) & (xx <= Ww - 1) & (yy >= 0) & (yy <= Hh - 1)
        v = field[b, yy.clamp(0, Hh - 1).long(), xx.clamp(0, Ww - 1).long()]
        return torch.where(inb, v, torch.zeros_like(v))

    return tap(y0, x0) * wx0 * wy0 + tap(y0, x1) * wx1 * wy0 + tap(y1, x0) * wx0 * wy1 + tap(y1, x1) * wx1 * wy1


class LiveRewardState:
    """The cross-step buffers of the live task (Appendix A of SURVEY): lazily created `prev_*`, goal counter, outcomes."""

    def __init__(self, n):
        self.goal_reached = torch.zeros(n, dtype=torch.int32)
        self.done_success = torch.zeros(n, dtype=torch.int32)
        self.done_collision = torch.zeros(n, dtype=torch.int32)
        self.prev_d = None          # CaptureXYTask.prev_position_dist
        self.prev_err = None        # CaptureXYReward.prev_position_error (separate buffer!)
        self.prev_h = None
        self.prev_pot = None
        self.prev_danger = None
        self.just_reset = torch.arange(n)

    def reset(self, ids):            # reset() :767-783 -- note prev_potential = None for EVERY env
        self.goal_reached[ids] = 0
        self.done_success[ids] = 0
        self.done_collision[ids] = 0
        self.just_reset = ids.clone()
        self.prev_pot = None


# B3  compute_reward :335-657 (+ CaptureXYReward.compute_reward OIGE/tasks/USV/USV_task_rewards.py:44-80)
def live_reward(c: LiveTaskConfig, S: LiveRewardState, aux, state, obstacles, field):
    d, herr, alpha, err = aux["d"], aux["herr"], aux["alpha"], aux["err"]
    pos, vel, w = state["position"], state["linear_velocity"], state["angular_velocity"]
    goal = (d < c.position_tolerance).int()                                   # :354-358 (no speed gate in the live task)
    S.goal_reached *= goal
    S.goal_reached += goal
    if S.prev_d is None:
        S.prev_d = d
    if S.prev_err is None:
        S.prev_err = d
    dist_rew = c.position_scale * (S.prev_err - d)                            # linear mode
    align = c.align_la1 * (torch.exp(c.align_la2 * herr.pow(4)) + torch.exp(c.align_la3 * herr.pow(2)))
    S.prev_err = d
    rm = S.just_reset
    dist_rew = dist_rew.clone()
    dist_rew[rm] = 0                                                          # :371-372
    if len(rm) > 0:
        S.prev_d = S.prev_d.clone()
        S.prev_d[rm] = d[rm]                                                  # :373-380
    pot = sample_potential(field, pos)
    pn = torch.clamp(pot, 0.0, 1.0)
    x = torch.clamp((pn - 0.6) / (0.9 - 0.6 + 1e-6), 0.0, 1.0)
    danger = x * x * (3.0 - 2.0 * x)                                          # :390-394 smoothstep
    if S.prev_danger is None:
        S.prev_danger = danger.clone()
    align = align * torch.maximum(torch.tensor(0.3), 1.0 - danger)            # :416-423
    dist_rew = dist_rew * torch.maximum(torch.tensor(0.6), 1.0 - danger * 0.5)
    g = torch.clamp(torch.cos(herr), min=0.0, max=1.0)                        # :429-432
    dist_rew = torch.clamp(dist_rew, max=0.0) + g * torch.clamp(dist_rew, min=0.0)
    if S.prev_h is None:
        S.prev_h = herr.clone()
    if len(rm) > 0:
        S.prev_h[rm] = herr[rm]
    h_imp = torch.clamp(S.prev_h - herr, -0.4, 0.4)                           # :436-446
    h_imp_rew = h_imp * 0.05
    S.prev_h = herr.clone()
    if S.prev_pot is None:
        S.prev_pot = pot.clone()                                              # :448-452 -- after ANY reset: zero shaping for all
    if len(rm) > 0:
        S.prev_pot[rm] = pot[rm]
    praw = (S.prev_pot - pot) * 100.0
    praw = torch.where(praw.abs() < 0.01, torch.zeros_like(praw), praw)
    pa1 = 2.0 * torch.tanh(praw / (2.0 + 1e-6))
    gdir = err / (d.unsqueeze(-1) + 1e-6)
    v_toward = torch.sum(vel * gdir, dim=-1)
    vtp = torch.clamp(v_toward, min=0.0)
    dd_pos = torch.clamp(S.prev_d - d, min=0.0)
    g_v = torch.clamp((vtp - 0.02) / (0.15 - 0.02 + 1e-6), 0.0, 1.0)
    g_d = torch.clamp(dd_pos / (0.01 + 1e-6), 0.0, 1.0)
    g_gate = torch.maximum(g_v, g_d) * torch.pow(g, 1.0)
    ppos, pneg = torch.clamp(pa1, min=0.0), torch.clamp(pa1, max=0.0)
    gate_pos = torch.where(ppos < 0.5, torch.ones_like(g_gate), g_gate)       # :505-517
    shaping = gate_pos * ppos + pneg
    S.prev_danger = danger.clone()
    worsening = shaping < -0.05                                               # :532-547
    turning = w.abs() > 0.2
    v_fwd = torch.sum(vel * state["orientation"], dim=-1).abs()
    speed_factor = torch.clamp((v_fwd - 0.15) / (0.60 - 0.15 + 1e-6), 0.0, 1.0)
    hazard = (worsening & turning).float() * (-10.0) * (g * g) * speed_factor
    S.prev_pot = pot.clone()
    S.just_reset = torch.tensor([], dtype=torch.long)
    speed_rew = (1.0 - torch.exp(-vtp / (0.8 + 1e-6))) * 0.05                 # :566-572
    tgt_w = torch.where(herr.abs() > 1.0, torch.sign(alpha) * 1.0, torch.sign(alpha) * 0.2)
    ang_rew = torch.exp(-((w - tgt_w) ** 2) / 0.2) * 0.03                     # :578-583
    coll = torch.zeros_like(d)
    for i in range(N_OBST):                                                   # :610-617
        od = torch.norm(obstacles[:, i] - pos, dim=1)
        coll = coll + (od < COLLISION_TH).float() * (-10.0) * 10.0
    goal_rew = (S.goal_reached * c.goal_reward).float() * 5.0                 # :620
    S.prev_d = d
    total = (dist_rew * 0.5 + align * 0.5 + shaping * 2.0 + hazard + goal_rew + c.time_reward + coll + speed_rew + ang_rew + h_imp_rew)
    return {"reward": total, "distance_reward": dist_rew, "alignment_reward": align, "potential_shaping": shaping, "turn_hazard": hazard,
            "speed_reward": speed_rew, "angular_reward": ang_rew, "heading_improve": h_imp_rew, "collision_penalty": coll,
            "goal_reward": goal_rew, "danger": danger, "potential": pot, "g_gate": gate_pos, "danger_hi": (danger > 0.5).float(),
            # :349-351 (diagnostic only)
            "boundary_penalty": -torch.expm1(torch.clamp(torch.clamp(d - c.kill_dist, min=0.0) / 0.25, max=20.0)) * c.boundary_cost}


# B4  update_kills :661-706
def live_kills(c: LiveTaskConfig, S: LiveRewardState, d, pos, obstacles):
    od = torch.norm(obstacles - pos.unsqueeze(1), dim=-1)
    collision = od.min(dim=1).values < COLLISION_TH
    success = S.goal_reached >= c.kill_after_n_steps_in_tolerance
    term = (d > c.kill_dist) | collision | success
    S.done_collision[term] = collision[term].to(torch.int32)
    S.done_success[term] = (success & ~collision)[term].to(torch.int32)
    return term.long()


# B6  BatchedMapGPU  [OIGE/tasks/USV/d_multi_gemini.py:66-104,135-271]
def grid_coords():
    lin = torch.linspace(-MAP_SIZE / 2 + CELL / 2, MAP_SIZE / 2 - CELL / 2, GRID)
    yg, xg = torch.meshgrid(lin, lin, indexing="ij")                          # row = y, column = x
    return xg, yg


def occupancy_and_sdf(obstacles):
    xg, yg = grid_coords()
    dx = xg[None, :, :, None] - obstacles[:, None, None, :, 0]
    dy = yg[None, :, :, None] - obstacles[:, None, None, :, 1]
    sdf = torch.norm(torch.stack([dx, dy], dim=-1), dim=-1).min(dim=-1).values - OBST_R   # ATen: sqrt(fma(y, y, x*x))
    occ = (sdf <= 0).float()
    occ[:, 0, :] = 1; occ[:, -1, :] = 1; occ[:, :, 0] = 1; occ[:, :, -1] = 1   # border walls
    return occ, sdf


def cost_to_go(occ, target, sweeps=int(GRID * 1.5)):
    """225 Jacobi sweeps of the 8-neighbour min-plus relaxation (1 / 1.414), obstacles = +inf."""
    B = occ.shape[0]
    inf = float("inf")
    cost = torch.full((B, GRID, GRID), inf)
    ti = ((target + MAP_SIZE / 2) / CELL).long().clamp(0, GRID - 1)
    cost[torch.arange(B), ti[:, 1], ti[:, 0]] = 0.0
    free = occ < 0.5
    for _ in range(sweeps):
        p = torch.nn.functional.pad(cost, (1, 1, 1, 1), value=inf)
        best = cost
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == 0 and dy == 0:
                    continue
                step = 1.0 if (dx == 0 or dy == 0) else 1.414
                best = torch.minimum(best, p[:, 1 + dy:1 + dy + GRID, 1 + dx:1 + dx + GRID] + step)
        cost = torch.where(free, best, torch.full_like(best, inf))
    return cost


def potential_field(cost, sdf, influence=0.7, eta=20.0):
    """g_norm + 0.5*J_norm; NOTE the batch-global maxima (`max_val`, `J_obs.max()`) of the reference :204-210,257-260."""
    B = cost.shape[0]
    finite = torch.isfinite(cost)
    max_val = cost[finite].max() if finite.any() else torch.tensor(100.0)
    vis = torch.where(torch.isinf(cost), max_val * 1.5, cost)
    mn, mx = vis.view(B, -1).min(1)[0].view(-1, 1, 1), vis.view(B, -1).max(1)[0].view(-1, 1, 1)
    g_norm = (vis - mn) / (mx - mn + 1e-6)
    edge = sdf - OBST_R
    rep_mask = torch.clamp(vis * CELL / 3.0, 0.0, 1.0)
    J = torch.zeros_like(edge)
    infl = edge < influence
    dcl = edge.clamp(min=1e-3)
    J = torch.where(infl, eta * (1.0 / dcl - 1.0 / influence) ** 2 * rep_mask, J)
    inside = edge <= 0
    if inside.any():
        cur = J.max()
        J = torch.where(inside, (cur * 10.0 if cur > 1e-6 else torch.tensor(100.0)).expand_as(J), J)
    jn, jx = J.view(B, -1).min(1)[0].view(-1, 1, 1), J.view(B, -1).max(1)[0].view(-1, 1, 1)
    return g_norm + 0.5 * (J - jn) / (jx - jn + 1e-6)


def build_field(obstacles, target):
    occ, sdf = occupancy_and_sdf(obstacles)
    return potential_field(cost_to_go(occ, target), sdf), occ, sdf


# ------------------------------------------------------------------------------------------------------------------
# Full live env step: the shared dynamics of ClassicEnvOracle + the Variant-B task, resets included (B5 with the CUDA
# path's Philox streams; RNG parity with torch's generator is statistical only, as for the classic task).
@dataclass
class LivePrivConfig:
    """Privileged tail + CoM randomisation  [OIGE/tasks/USV_Virtual.py:97-151,837-984 ; USV_disturbances.py:88-124,153-194]."""
    priv_mode: int = 2                       # 0 raw, 1 centered, 2 minmax
    mass_obs_relative: bool = True
    com_obs_scaled: bool = True
    com_scale: tuple = (1.3, 1.0, 1.0)       # (box_length, box_width, max(heron_zero_height, 1))
    priv_a: tuple = (1.0, 0.5, 0.5, 1.0)     # minmax: min of [k_drag, thr_L, thr_R, k_Iz]
    priv_b: tuple = (0.5, 0.5, 0.5, 0.5)     # minmax: max - min
    priv_active: tuple = (True, True, True, True)
    com_rand: bool = True
    com_base: tuple = (0.0, 0.0, 0.0)
    com_disp: tuple = (0.15, 0.05, 0.02)
    collision_threshold: float = COLLISION_TH
    fixed_horizon_eval: bool = False
    masscom_obs_base: bool = False           # mass.masscom_obs_source == 'base'  [OIGE/tasks/USV_Virtual.py:840-880]


def priv_encode(pc: LivePrivConfig, j: int, x: torch.Tensor) -> torch.Tensor:
    if pc.priv_mode == 0:
        return x
    if pc.priv_mode == 1:
        return torch.clamp((x - pc.priv_a[j]) / pc.priv_b[j], -1.0, 1.0)
    if not pc.priv_active[j]:
        return torch.zeros_like(x)
    z = (x - pc.priv_a[j]) / pc.priv_b[j]
    return torch.clamp(2.0 * z - 1.0, -1.0, 1.0)


def place_obstacles(seed: int, env_ids: np.ndarray, step: int, start: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """get_spawns :970-1047: 16 centres ~ U(target +- 12), <= 20 rounds of re-drawing the invalid ones (within 3 m of start /
    target, or < 2.5 m from a lower-index obstacle), leftovers -> (999, 999)."""
    m = len(env_ids)
    mn = target - 12.0
    span = (target + 12.0) - mn

    def draw(rnd):
        out = torch.zeros((m, N_OBST, 2), dtype=F32)
        for pair in range(N_OBST // 2):
            u = torch.from_numpy(philox.uniform4(seed, env_ids, step, philox.RS_OBST + rnd * 8 + pair))
            out[:, 2 * pair] = u[:, 0:2] * span + mn
            out[:, 2 * pair + 1] = u[:, 2:4] * span + mn
        return out

    def invalid(obs):
        ds = obs - start.unsqueeze(1)
        dt = obs - target.unsqueeze(1)
        inv = (torch.norm(ds, dim=-1) < 3.0) | (torch.norm(dt, dim=-1) < 3.0)
        valid = obs[..., 0] < 900.0
        delta = obs.unsqueeze(2) - obs.unsqueeze(1)                  # [n, a, b] = obs[a] - obs[b]
        d2 = delta[..., 0] * delta[..., 0] + delta[..., 1] * delta[..., 1]
        triu = torch.triu(torch.ones((N_OBST, N_OBST), dtype=torch.bool), diagonal=1).unsqueeze(0)
        conf = (d2 < 2.5 * 2.5) & valid.unsqueeze(2) & valid.unsqueeze(1) & triu
        return inv | conf.any(dim=1)

    obs = draw(0)
    active = torch.ones(m, dtype=torch.bool)                         # per-env early exit == the batch-wide one: valid ones never move
    for it in range(20):
        inv = invalid(obs) & active.unsqueeze(1)
        active = inv.any(dim=1)
        if not active.any():
            break
        obs = torch.where(inv.unsqueeze(-1), draw(it + 1), obs)
    obs = torch.where(invalid(obs).unsqueeze(-1), torch.tensor([999.0, 999.0]), obs)
    return obs


class LiveEnvOracle(ClassicEnvOracle):
    """VecEnvRLGames.step over the live USVVirtual + CaptureXYTask (static obstacles)."""

    def __init__(self, cfg: EnvConfig, task: LiveTaskConfig, priv: LivePrivConfig, num_envs: int, env_id_offset: int = 0):
        super().__init__(cfg, num_envs, env_id_offset)
        self.task, self.priv = task, priv
        n = num_envs
        self.obstacles = torch.zeros((n, N_OBST, 2), dtype=F32)
        self.field = torch.zeros((n, GRID, GRID), dtype=F32)
        self.com = torch.tensor([priv.com_base] * n, dtype=F32)
        self.S = LiveRewardState(n)
        self.outcome_at_reset = {"success": torch.zeros(0), "collision": torch.zeros(0)}

    def reset_idx(self, ids: torch.Tensor, step: int):
        if ids.numel() == 0:
            return
        c = self.cfg
        gids = self.env_ids[ids.numpy()]
        self.outcome_at_reset = {"success": self.S.done_success[ids].float(), "collision": self.S.done_collision[ids].float()}
        self.S.reset(ids)
        if self.priv.com_rand:
            rc = torch.from_numpy(philox.uniform4(c.seed, gids, step, philox.RS_RESET_COM))
            if int(self.priv.com_rand) == 2:      # legacy XY disc, radius bound in com_disp[0]  [USV_disturbances.py:108-124]
                self.com[ids] = com_disc(self.priv.com_base, rc[:, 0], rc[:, 1], float(self.priv.com_disp[0]))
            else:
                self.com[ids] = torch.tensor(self.priv.com_base, dtype=F32) + (rc[:, 0:3] * 2 - 1) * torch.tensor(self.priv.com_disp, dtype=F32)
        # scene: spawn point of this reset (same draw as ClassicEnvOracle.reset_idx) and the CURRENT target
        r0 = torch.from_numpy(philox.uniform4(c.seed, gids, step, philox.RS_RESET[0]))
        rmin32, rmax32 = torch.tensor(c.spawn_min_dist, dtype=F32), torch.tensor(c.spawn_max_dist, dtype=F32)
        sr = r0[:, 2] * (rmax32 - rmin32) + rmin32                       # fp32 parameters, as the kernels (and ClassicEnvOracle) form it
        th = r0[:, 3] * 2 * math.pi
        start = torch.stack([sr * torch.cos(th), sr * torch.sin(th)], 1)
        if not c.spawn_about_origin:
            start = start + self.target[ids]
        tgt = self.target[ids].clone()
        self.obstacles[ids] = place_obstacles(c.seed, gids, step, start, tgt)
        self.field[ids] = build_field(self.obstacles[ids], tgt)[0]
        super().reset_idx(ids, step)

    def priv_tail(self):
        c, pc = self.cfg, self.priv
        if pc.masscom_obs_base:
            # ablation source: base mass / CoM encodings (USV_disturbances.py:196-250) and neutral dynamics parameters -- the mid-range
            # in minmax mode, 1.0 otherwise (USV_Virtual.py:859-880) -- for every env, then the same encoders
            n = self.mass.shape[0]
            mass = torch.zeros(n, dtype=F32) if pc.mass_obs_relative else torch.full((n,), float(c.mass_base), dtype=F32)
            com = torch.tensor(pc.com_base, dtype=F32).unsqueeze(0).repeat(n, 1)
            if pc.com_obs_scaled:
                com = com / (torch.tensor(pc.com_scale, dtype=F32) + 1e-6)
            neutral = [0.5 * (float(a) + (float(a) + float(b))) if pc.priv_mode == 2 else 1.0 for a, b in zip(pc.priv_a, pc.priv_b)]
            cols = [mass.unsqueeze(1), com] + [priv_encode(pc, j, torch.ones(n, dtype=F32) * x).unsqueeze(1) for j, x in enumerate(neutral)]
            return torch.cat(cols, dim=1)
        mass = (self.mass - c.mass_base) / max(abs(c.mass_base), 1e-6) if pc.mass_obs_relative else self.mass
        com = self.com / (torch.tensor(pc.com_scale, dtype=F32) + 1e-6) if pc.com_obs_scaled else self.com
        cols = [mass.unsqueeze(1), com] + [priv_encode(pc, j, x).unsqueeze(1) for j, x in
                                           enumerate((self.drag_scale[:, 0], self.thr_mult_left, self.thr_mult_right, self.k_iz))]
        return torch.cat(cols, dim=1)

    def step(self, actions: torch.Tensor):
        c, t = self.cfg, self.task
        state, dyn = self.dynamics(actions)
        reset_ids, w = dyn["reset_ids"], state["angular_velocity"]
        prev_action = dyn["raw_actions"].clone()
        prev_action[reset_ids] = 0.0
        obs, aux = live_observation(state, self.target, self.obstacles, prev_action, self.priv_tail())
        out = live_reward(t, self.S, aux, state, self.obstacles, self.field)
        pen = penalties(c, state, dyn["pen_actions"], self.prev_w, self.prev_asum, self.first_call)
        self.prev_w = w
        self.prev_asum = pen["asum"]
        self.first_call = False
        rew = out["reward"] + pen["total"]
        self.goal_reached = self.S.goal_reached
        self.prev_d = aux["d"]
        die = live_kills(t, self.S, aux["d"], state["position"], self.obstacles)
        if self.priv.fixed_horizon_eval:
            die = torch.zeros_like(die)
        ones = torch.ones_like(self.reset_buf)
        self.reset_buf = torch.where(self.progress_buf >= c.max_episode_length - 1, ones, die)
        obs = torch.clamp(obs, -c.clip_obs, c.clip_obs)
        self.step_counter += 1
        self.last = {**out, **pen, **aux, "reset_ids": reset_ids, "before_rect": dyn["before_rect"], "unit": dyn["unit"],
                     "raw_actions": dyn["raw_actions"], "state": state}
        return obs, rew, self.reset_buf.clone()
