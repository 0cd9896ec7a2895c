"""VecEnvRLGames: the rl_games-facing vec-env of the reference [ref: OIGE/envs/vec_env_rlgames.py:38-230] without Isaac Sim.
step(actions) -> ({"obs": {"state": (N,13)}, "states": (N,0)}, rew (N,), resets (N,) int64, extras)."""
from __future__ import annotations

import torch


class _World:
    """The reference drives `self._world.step(render=...)` between sub-steps; here the sub-steps run inside the fused
    kernel, so the world is a counter."""

    def __init__(self):
        self.steps = 0

    def is_playing(self):
        return True

    def step(self, render=False):
        self.steps += 1

    def reset(self):
        pass


class VecEnvRLGames:
    def __init__(self, headless: bool = True, sim_device: int = 0, enable_livestream: bool = False, enable_viewport: bool = False):
        self._world = _World()
        self._render = not headless
        self.sim_frame_count = 0
        self._task = None

    def set_task(self, task, backend="torch", sim_params=None, init_sim=True) -> None:
        self._task = task
        task._env = self
        self.num_envs = task.num_envs
        self.observation_space = task.observation_space
        self.action_space = task.action_space
        self.num_states = task.num_states
        self.state_space = task.state_space

    def _process_data(self):
        # the fused kernel already clamps obs to +-clipObservations; tensors are handed over as clones on rl_device
        # [ref :82-112]
        dev = self._task.rl_device
        self._obs = {k: v.to(dev).clone() for k, v in self._obs.items()}
        self._rew = self._rew.to(dev).clone()
        self._resets = self._resets.to(dev).clone()
        self._extras = dict(self._extras)

    def step(self, actions):
        """[ref :120-217]"""
        t = self._task
        actions = torch.clamp(actions, -t.clip_actions, t.clip_actions).to(t.device).clone()
        t.pre_physics_step(actions)
        # [ref :154-171] (controlFrequencyInv-1) x {apply_forces; world.step; update_state} + the final apply_forces/world.step
        # run inside the fused kernel launched by post_physics_step()
        self.sim_frame_count += t.control_frequency_inv
        self._obs, self._rew, self._resets, self._extras = t.post_physics_step()
        self._states = t.get_states()
        self._process_data()
        return {"obs": self._obs, "states": self._states}, self._rew, self._resets, self._extras

    def reset(self):
        """[ref :219-230] flags every env and takes one zero-action step to produce observations."""
        self._task.reset()
        actions = torch.zeros((self.num_envs, self._task.num_actions), device=self._task.rl_device)
        obs_dict, _, _, _ = self.step(actions)
        return obs_dict
