"""Classic CaptureXY task class (13-dim observation) on the C-ABI kernel
[ref: SNAP/USV_capture_xy.py:30-399 ; the live factory's obstacle variant is tasks/USV_Virtual.py + csrc/usv_step_b.cu].

Same constructor, methods, argument meaning and cross-call state as the reference class; every method that computes launches
usv_capturexy_obs_reward_done_f32 -- the SAME device function the fused env step runs (csrc/usv_step.cu:post_classic) -- with the bit
of that call, so the reference's state tensors go in and its outputs come out one call at a time.  The fused env never goes through
this class (one launch per control step); this is the drop-in and the parity surface."""
from __future__ import annotations

import math

import torch

from ... import _lib
from ...config import UsvEnvConfig, reward_section_kwargs, task_section_kwargs
from .USV_core import parse_data_dict
from .USV_task_parameters import CaptureXYParameters
from .USV_task_rewards import CXY, CaptureXYReward, launch_capturexy, state_pointers


class CaptureXYTask:
    def __init__(self, task_param: dict, reward_param: dict, num_envs: int, device: str, priv_dim: int = 0) -> None:
        self._num_envs, self._device = int(num_envs), torch.device(device)
        if self._device.type != "cuda":
            raise _lib.UsvLibraryError("CaptureXYTask runs on CUDA only (no CPU fallback)")
        self._task_parameters = parse_data_dict(CaptureXYParameters(), dict(task_param))
        self._reward_parameters = parse_data_dict(CaptureXYReward(), dict(reward_param))
        cfg = UsvEnvConfig(**task_section_kwargs(self._task_parameters.as_section()),
                           **reward_section_kwargs(self._reward_parameters.as_section()), clip_obs=float("inf"))
        self._cfg, self._params = cfg, cfg.to_params()
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=self._device)
        self._goal_reached = z(self._num_envs, dt=torch.int32)
        self._target_positions = z(self._num_envs, 2)
        self._just_reset = torch.ones(self._num_envs, dtype=torch.uint8, device=self._device)    # every env starts "just reset"
        self._prev_position_dist = z(self._num_envs)
        self._first_reward = True
        self._num_observations = 13
        self._obs = z(self._num_envs, 13)
        self._terms = z(self._num_envs, 3)
        self.current_state = None

    # the reference exposes the reset set as an index tensor
    @property
    def just_had_been_reset(self) -> torch.Tensor:
        return self._just_reset.nonzero().flatten()

    def _io(self, state: dict):
        io = _lib.UsvCaptureXYIO()
        keep = state_pointers(io, state)
        io.target = self._target_positions.data_ptr()
        io.goal_reached = self._goal_reached.data_ptr()
        io.kill_dist = float(self._task_parameters.kill_dist)
        return io, keep

    def get_state_observations(self, current_state: dict, observation_frame: str = "local") -> torch.Tensor:
        if observation_frame != "local":
            raise NotImplementedError("only the 'local' observation frame exists on the USV path (every USV YAML sets it)")
        self.current_state = current_state
        io, keep = self._io(current_state)
        io.obs, io.what = self._obs.data_ptr(), CXY["OBS"]
        launch_capturexy(io, self._num_envs, self._params)
        return self._obs

    def compute_reward(self, current_state: dict, actions: torch.Tensor) -> torch.Tensor:
        io, keep = self._io(current_state)
        rew = torch.empty(self._num_envs, dtype=torch.float32, device=self._device)
        io.reward, io.reward_terms = rew.data_ptr(), self._terms.data_ptr()
        io.prev_position_dist, io.just_reset = self._prev_position_dist.data_ptr(), self._just_reset.data_ptr()
        io.what, io.first_reward = CXY["REWARD"], int(self._first_reward)
        launch_capturexy(io, self._num_envs, self._params)
        self._first_reward = False
        self._just_reset.zero_()
        self.distance_reward, self.alignment_reward, self.a = self._terms.unbind(1)
        self.position_dist = self._prev_position_dist          # compute_reward leaves the current distance there
        return rew

    def update_kills(self, step: float = 0) -> torch.Tensor:
        if self.current_state is None:
            raise RuntimeError("update_kills before get_state_observations")
        io, keep = self._io(self.current_state)
        die = torch.empty(self._num_envs, dtype=torch.int64, device=self._device)
        _, _, io.kill_dist = self._cfg.curriculum(step)
        io.die, io.what = die.data_ptr(), CXY["KILLS"]
        launch_capturexy(io, self._num_envs, self._params)
        return die

    def reset(self, env_ids: torch.Tensor) -> None:
        self._goal_reached[env_ids] = 0
        self._just_reset.zero_()
        self._just_reset[env_ids] = 1

    def get_goals(self, env_ids: torch.Tensor, targets_position: torch.Tensor, targets_orientation: torch.Tensor):
        """New targets ~ U(-goal_random_position, +goal_random_position)^2 for `env_ids`; added onto `targets_position[:, :2]`."""
        g = float(self._task_parameters.goal_random_position)
        self._target_positions[env_ids] = (torch.rand((len(env_ids), 2), device=self._device) * 2.0 - 1.0) * g
        targets_position[env_ids, :2] += self._target_positions[env_ids]
        return targets_position, targets_orientation

    def get_spawns(self, env_ids: torch.Tensor, initial_position: torch.Tensor, initial_orientation: torch.Tensor, step: int = 0):
        """Spawn on an annulus around the target (curriculum-resolved radii), heading ~ U(0, pi) as a (w, 0, 0, z) quaternion."""
        n = len(env_ids)
        self._goal_reached[env_ids] = 0
        rmin, rmax, _ = self._cfg.curriculum(step)
        r = torch.rand(n, device=self._device) * (rmax - rmin) + rmin
        th = torch.rand(n, device=self._device) * (2.0 * math.pi)
        initial_position[env_ids, 0] += r * torch.cos(th) + self._target_positions[env_ids, 0]
        initial_position[env_ids, 1] += r * torch.sin(th) + self._target_positions[env_ids, 1]
        half = torch.rand(n, device=self._device) * (0.5 * math.pi)
        initial_orientation[env_ids, 0] = torch.cos(half)
        initial_orientation[env_ids, 3] = torch.sin(half)
        return initial_position, initial_orientation
