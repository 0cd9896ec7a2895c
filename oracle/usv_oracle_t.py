"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the Tier-3 tasks: GoToPose, KeepXY, TrackXYVelocity behind the live
USVVirtual's 33-dim observation.  torch-on-CPU restatement, pinned against tests/golden/tier3_tasks.npz (the reference's own task
classes driven method by method under oracle/ref_shim.py; they cannot run end-to-end in the reference).
OIGE = omniisaacgymenvs/."""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import philox
from .usv_oracle import ClassicEnvOracle, EnvConfig, com_disc, penalties
from .usv_oracle_b import LivePrivConfig, priv_encode

F32 = torch.float32
GO_TO_POSE, KEEP_XY, TRACK_XY_VELOCITY = 1, 2, 3
RS_RESET_TASK = 12


@dataclass
class Tier3Config:
    """Task + reward dataclasses  [OIGE/tasks/USV/USV_task_parameters.py:95-177 ; USV_task_rewards.py:170-325]."""
    task: int = GO_TO_POSE
    position_tolerance: float = 0.01
    kill_after_n_steps_in_tolerance: int = 500
    kill_dist: float = 10.0
    reward_mode: int = 2                 # position / velocity reward: 0 linear, 1 square, 2 exponential
    exponential_reward_coeff: float = 0.25
    position_scale: float = 1.0
    heading_reward_mode: int = 2
    heading_exponential_reward_coeff: float = 0.25
    heading_scale: float = 5.0
    sig_gain: float = 3.0
    lin_vel_tolerance: float = 0.01
    goal_random_velocity: float = 0.75


def mode_reward(mode: int, err: torch.Tensor, coeff: float) -> torch.Tensor:
    if mode == 0:
        return 1.0 / (1.0 + err)
    if mode == 1:
        return 1.0 / (1.0 + err * err)
    return torch.exp(-err / coeff)


def task_observation(c: Tier3Config, state, target, target_heading, target_vel, prev_action, priv):
    """get_state_observations + Core.update_observation_tensor  [USV_go_to_pose.py:81-125 ; USV_keep_xy.py:80-116 ;
    USV_track_xy_velocity.py:64-88 ; USV_core.py:55-125]."""
    n = state["position"].shape[0]
    td = torch.zeros((n, 20), dtype=F32)
    aux = {}
    if c.task == TRACK_XY_VELOCITY:
        verr = target_vel - state["linear_velocity"]
        td[:, :2] = verr
        aux.update(verr=verr, perr=state["position"])
    else:
        err = target - state["position"]
        theta = torch.atan2(state["orientation"][:, 1], state["orientation"][:, 0])
        beta = torch.atan2(err[:, 1], err[:, 0])
        alpha = torch.fmod(beta - theta + math.pi, 2 * math.pi) - math.pi
        td[:, 0], td[:, 1], td[:, 2] = torch.cos(alpha), torch.sin(alpha), torch.norm(err, dim=1)
        aux.update(perr=err)
        if c.task == GO_TO_POSE:
            h = torch.fmod(target_heading - theta + math.pi, 2 * math.pi) - math.pi
            herr = torch.atan2(torch.sin(h), torch.cos(h))
            td[:, 3], td[:, 4] = torch.cos(herr), torch.sin(herr)
            aux.update(herr=herr)
    obs = torch.zeros((n, 33), dtype=F32)
    cth, sth = state["orientation"][:, 0], state["orientation"][:, 1]
    v = state["linear_velocity"]
    obs[:, 0] = cth * v[:, 0] + sth * v[:, 1]
    obs[:, 1] = -sth * v[:, 0] + cth * v[:, 1]
    obs[:, 2] = state["angular_velocity"]
    obs[:, 3:23] = td
    obs[:, 23:25] = prev_action
    obs[:, 25:33] = priv
    return obs, aux


def task_reward_and_kills(c: Tier3Config, aux, state, actions, goal_reached, prev_d):
    """compute_reward + update_kills; returns (reward, die, new prev_d); goal_reached is updated in place."""
    pd = torch.sqrt(torch.square(aux["perr"]).sum(-1))
    if c.task == TRACK_XY_VELOCITY:                                            # [USV_track_xy_velocity.py:90-128]
        vd = torch.sqrt(torch.square(aux["verr"]).sum(-1))
        goal = (vd < c.lin_vel_tolerance).int()
        goal_reached *= goal
        goal_reached += goal
        rew = mode_reward(c.reward_mode, vd, c.exponential_reward_coeff)
        die = ((pd > c.kill_dist) | (goal_reached > c.kill_after_n_steps_in_tolerance)).long()
        return rew, die, prev_d
    if c.task == GO_TO_POSE:                                                   # [USV_go_to_pose.py:129-209]
        hd = torch.abs(aux["herr"])
        progress = 2.0 * (prev_d - pd).clamp(min=-2, max=2)
        speed = torch.norm(state["linear_velocity"], dim=-1)
        goal = ((pd < c.position_tolerance) & (speed < 0.1)).int()
        goal_reached *= goal
        goal_reached += goal
        hw = 1.0 - 1 / (1 + torch.exp(-c.sig_gain * (pd - 2)))                 # GoToPoseReward.compute_reward :206-255
        pos_rew = c.position_scale * mode_reward(c.reward_mode, pd, c.exponential_reward_coeff)
        head_rew = hw * c.heading_scale * mode_reward(c.heading_reward_mode, hd, c.heading_exponential_reward_coeff)
        rew = pos_rew + head_rew + progress + 2.0 * goal.float() + (-0.05 * torch.abs(actions).sum(dim=-1))
        prev_d = pd.clone()
    else:                                                                      # KeepXY [USV_keep_xy.py:118-179]
        rew = mode_reward(c.reward_mode, pd, c.exponential_reward_coeff)
    die = ((pd > c.kill_dist) | (goal_reached >= c.kill_after_n_steps_in_tolerance)).long()
    return rew, die, prev_d


class Tier3EnvOracle(ClassicEnvOracle):
    """VecEnvRLGames.step over the live USVVirtual with a Tier-3 task plugged in."""

    def __init__(self, cfg: EnvConfig, task: Tier3Config, priv: LivePrivConfig, num_envs: int, env_id_offset: int = 0):
        super().__init__(cfg, num_envs, env_id_offset)
        self.task, self.priv = task, priv
        n = num_envs
        self.com = torch.tensor([priv.com_base] * n, dtype=F32)
        self.target_heading = torch.zeros(n, dtype=F32)
        self.target_vel = torch.zeros((n, 2), dtype=F32)

    def reset_idx(self, ids: torch.Tensor, step: int):
        if ids.numel() == 0:
            return
        c = self.cfg
        gids = self.env_ids[ids.numpy()]
        rc = torch.from_numpy(philox.uniform4(c.seed, gids, step, philox.RS_RESET_COM))
        if self.priv.com_rand:
            if int(self.priv.com_rand) == 2:      # legacy XY disc, radius bound in com_disp[0]  [USV_disturbances.py:108-124]
                self.com[ids] = com_disc(self.priv.com_base, rc[:, 0], rc[:, 1], float(self.priv.com_disp[0]))
            else:
                self.com[ids] = torch.tensor(self.priv.com_base, dtype=F32) + (rc[:, 0:3] * 2 - 1) * torch.tensor(self.priv.com_disp, dtype=F32)
        if not c.reset_pose_external:
            if self.task.task == GO_TO_POSE:
                self.target_heading[ids] = rc[:, 3] * math.pi * 2
            if self.task.task == TRACK_XY_VELOCITY:
                rt = torch.from_numpy(philox.uniform4(c.seed, gids, step, RS_RESET_TASK))
                g = self.task.goal_random_velocity
                self.target_vel[ids] = rt[:, 0:2] * g * 2 - g
        if self.task.task == GO_TO_POSE:
            self.prev_d[ids] = 0                                               # reset(): prev_position_dist[env_ids] = 0
        super().reset_idx(ids, step)

    def priv_tail(self):
        c, pc = self.cfg, self.priv
        mass = (self.mass - c.mass_base) / max(abs(c.mass_base), 1e-6) if pc.mass_obs_relative else self.mass
        com = self.com / (torch.tensor(pc.com_scale, dtype=F32) + 1e-6) if pc.com_obs_scaled else self.com
        cols = [mass.unsqueeze(1), com] + [priv_encode(pc, j, x).unsqueeze(1) for j, x in
                                           enumerate((self.drag_scale[:, 0], self.thr_mult_left, self.thr_mult_right, self.k_iz))]
        return torch.cat(cols, dim=1)

    def step(self, actions: torch.Tensor):
        c = self.cfg
        state, dyn = self.dynamics(actions)
        reset_ids, w = dyn["reset_ids"], state["angular_velocity"]
        prev_action = dyn["raw_actions"].clone()
        prev_action[reset_ids] = 0.0
        obs, aux = task_observation(self.task, state, self.target, self.target_heading, self.target_vel, prev_action, self.priv_tail())
        rew, die, self.prev_d = task_reward_and_kills(self.task, aux, state, dyn["raw_actions"], self.goal_reached, self.prev_d)
        pen = penalties(c, state, dyn["pen_actions"], self.prev_w, self.prev_asum, self.first_call)
        self.prev_w = w
        self.prev_asum = pen["asum"]
        self.first_call = False
        rew = rew + pen["total"]
        if self.priv.fixed_horizon_eval:
            die = torch.zeros_like(die)
        ones = torch.ones_like(self.reset_buf)
        self.reset_buf = torch.where(self.progress_buf >= c.max_episode_length - 1, ones, die)
        obs = torch.clamp(obs, -c.clip_obs, c.clip_obs)
        self.step_counter += 1
        return obs, rew, self.reset_buf.clone()
