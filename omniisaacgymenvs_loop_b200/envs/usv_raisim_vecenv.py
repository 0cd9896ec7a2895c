"""Raisim-style adapter over VecEnvRLGames: `observe()` / `step(action) -> (reward, dones)` as the loopz training loop expects.

Mirrors `omniisaacgymenvs/envs/usv_raisim_vecenv.py:43-250` (`USVRaisimVecEnv`).  The reference adapter moves every
observation / reward / done vector to the host as numpy; this one does the same when it is handed numpy actions and keeps
everything on the device when it is handed CUDA tensors (`observe(as_numpy=False)`), which is the path the GPU learner uses."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch


class USVRaisimVecEnv:
    def __init__(self, base_env: Any, *, reward_info_size: int = 16, device=None) -> None:
        self._env = base_env
        self._task = getattr(base_env, "_task", None)
        if self._task is None:
            raise ValueError("base_env must be a VecEnvRLGames-like env with attribute `_task`.")
        self._device = torch.device(device) if device is not None else torch.device(getattr(self._task, "rl_device", self._task.device))
        self.num_envs = int(getattr(self._env, "num_envs", self._task.num_envs))
        self.num_obs = int(self._task.num_observations)
        self.num_acts = int(self._task.num_actions)
        self._reward_info_size = int(reward_info_size)
        self._last_obs_torch: Optional[torch.Tensor] = None
        self._last_reward_torch: Optional[torch.Tensor] = None
        self._last_dones_torch: Optional[torch.Tensor] = None
        self._last_extras: Dict[str, Any] = {}

    @staticmethod
    def _extract_obs_tensor(obs_dict) -> torch.Tensor:
        if isinstance(obs_dict, dict):
            obs = obs_dict.get("obs")
            if obs is None:
                raise KeyError("obs_dict does not contain key 'obs'.")
            if isinstance(obs, dict):
                if "state" in obs and torch.is_tensor(obs["state"]):
                    return obs["state"]
                vals = [v for v in obs.values() if torch.is_tensor(v)]
                if len(vals) == 1:
                    return vals[0]
                raise TypeError("obs_dict['obs'] is a dict but no single tensor could be inferred")
            return obs
        if torch.is_tensor(obs_dict):
            return obs_dict
        raise TypeError("Unsupported observation type returned from base_env.reset/step")

    def reset(self) -> None:
        self._last_obs_torch = self._extract_obs_tensor(self._env.reset())

    def observe(self, *_args: Any, as_numpy: bool = True, **_kwargs: Any):
        if self._last_obs_torch is None:
            self.reset()
        if not as_numpy:
            return self._last_obs_torch           # non-finite entries are zeroed by the kernels that read it
        return torch.nan_to_num(self._last_obs_torch, nan=0.0, posinf=0.0, neginf=0.0).cpu().numpy().astype(np.float32, copy=False)

    def step(self, action):
        as_numpy = isinstance(action, np.ndarray)
        a = torch.from_numpy(action) if as_numpy else action
        if not torch.is_tensor(a):
            raise TypeError("action must be a torch.Tensor or np.ndarray")
        obs_dict, rew, resets, extras = self._env.step(a.to(self._device, torch.float32))
        self._last_obs_torch = self._extract_obs_tensor(obs_dict)
        self._last_reward_torch, self._last_dones_torch = rew, resets
        self._last_extras = extras if isinstance(extras, dict) else {"extras": extras}
        if not as_numpy:
            return rew, resets
        rew_np = torch.nan_to_num(rew, nan=0.0, posinf=0.0, neginf=0.0).cpu().numpy().astype(np.float32, copy=False).reshape(-1)
        return rew_np, resets.cpu().numpy().reshape(-1).astype(np.bool_, copy=False)

    def get_reward_info(self) -> np.ndarray:
        info = np.zeros((self.num_envs, self._reward_info_size), dtype=np.float32)
        if self._last_reward_torch is not None:
            info[:, 0] = self._last_reward_torch.detach().cpu().numpy().reshape(-1)
        return info

    def get_extras(self) -> Dict[str, Any]:
        return self._last_extras

    def curriculum_callback(self) -> None:
        return None

    def close(self) -> None:
        fn = getattr(self._env, "close", None)
        if callable(fn):
            fn()
