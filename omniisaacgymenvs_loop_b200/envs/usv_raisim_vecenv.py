"""Raisim-style adapter over VecEnvRLGames: `observe()` / `step(action) -> (reward, dones)` as the loopz training loop expects.

Mirrors `omniisaacgymenvs/envs/usv_raisim_vecenv.py:43-250` (`USVRaisimVecEnv`).  The reference adapter moves every
observation / reward / done vector to the host as numpy; this one does the same when it is handed numpy actions and keeps
everything on the device when it is handed CUDA tensors (`observe(as_numpy=False)`), which is the path the GPU learner uses."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import torch


def _state_tensor(ret) -> torch.Tensor:
    """The observation tensor inside what VecEnvRLGames.reset / step hand back: {"obs": {"state": tensor}} on this path; a bare
    tensor or {"obs": tensor} from other rl_games-style envs is accepted too."""
    seen = ret
    for key in ("obs", "state"):
        if isinstance(seen, dict):
            if key not in seen:
                if key == "state" and len(seen) == 1:            # a single observation group under another name
                    seen = next(iter(seen.values()))
                    continue
                raise KeyError(f"observation dict has no '{key}' entry (keys: {sorted(seen)})")
            seen = seen[key]
    if not torch.is_tensor(seen):
        raise TypeError(f"no observation tensor in a {type(ret).__name__} returned by the wrapped env")
    return seen


class USVRaisimVecEnv:
    """observe() / step(action) -> (reward, dones) over a VecEnvRLGames.  The learner-facing sizes come from the task behind it."""

    def __init__(self, base_env: Any, *, reward_info_size: int = 16, device=None) -> None:
        task = getattr(base_env, "_task", None)
        if task is None:
            raise ValueError("USVRaisimVecEnv wraps a VecEnvRLGames whose task has been set (no `_task` on the given env)")
        self._env, self._task = base_env, task
        self._device = torch.device(device if device is not None else getattr(task, "rl_device", task.device))
        self.num_envs, self.num_obs, self.num_acts = (int(getattr(base_env, "num_envs", task.num_envs)), int(task.num_observations),
                                                      int(task.num_actions))
        self._reward_info_size = int(reward_info_size)
        # what the last reset / step left behind (device tensors; the numpy views are made on demand)
        self._last_obs_torch = self._last_reward_torch = self._last_dones_torch = None
        self._last_extras: Dict[str, Any] = {}

    _extract_obs_tensor = staticmethod(_state_tensor)

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


class USVSysIDVecEnv(USVRaisimVecEnv):
    """The SysID / DAgger wrapper [ref: omniisaacgymenvs/envs/usv_raisim_vecenv.py:387-617]: full observation unchanged, plus the
    non-privileged slice, a rolling history of it (N, history_len, obs_nonpriv_dim) and the privileged tail the teacher reads.

    The history lives on the device and every accessor takes `as_numpy` (default True = the reference's numpy return; False = the
    device tensor / view, no host trip).  Reference behaviours kept: history is zero-filled ("zeros") or filled with the first frame
    ("repeat") at `reset()`; every step rolls the window by one and writes the current frame last; for envs that finished on the step,
    "repeat" refills the window with the post-reset frame, while "zeros" leaves the old frames in place -- the reference zeroes an
    advanced-indexing COPY (`self._history_torch[done_mask].zero_()`, :613), so nothing is cleared."""

    def __init__(self, base_env: Any, *, history_len: int = 50, priv_dim: int = 4, fill_history_on_reset: str = "zeros",
                 reward_info_size: int = 16, device=None) -> None:
        super().__init__(base_env, reward_info_size=reward_info_size, device=device)
        self.history_len, self.priv_dim = int(history_len), int(priv_dim)
        if self.history_len <= 0:
            raise ValueError(f"history_len must be > 0, got {self.history_len}")
        if self.priv_dim <= 0:
            raise ValueError(f"priv_dim must be > 0, got {self.priv_dim}")
        self.obs_nonpriv_dim = int(self.num_obs - self.priv_dim)
        if self.obs_nonpriv_dim <= 0:
            raise ValueError(f"Invalid dims: num_obs={self.num_obs}, priv_dim={self.priv_dim} => obs_nonpriv_dim={self.obs_nonpriv_dim}")
        if fill_history_on_reset not in {"repeat", "zeros"}:
            raise ValueError("fill_history_on_reset must be 'repeat' or 'zeros'")
        self._fill_history_on_reset = fill_history_on_reset
        self._history_torch = torch.zeros((self.num_envs, self.history_len, self.obs_nonpriv_dim), device=self._device, dtype=torch.float32)

    def _current(self) -> torch.Tensor:
        if self._last_obs_torch is None:
            self.reset()
        return self._last_obs_torch

    @staticmethod
    def _out(t: torch.Tensor, as_numpy: bool):
        return t.detach().cpu().numpy().astype(np.float32, copy=False) if as_numpy else t

    def reset(self) -> None:
        super().reset()
        cur = self._current()[:, : self.obs_nonpriv_dim]
        if self._fill_history_on_reset == "repeat":
            self._history_torch[:] = cur.unsqueeze(1).expand(-1, self.history_len, -1)
        else:
            self._history_torch.zero_()
            self._history_torch[:, -1, :] = cur

    def step(self, action):
        out = super().step(action)
        cur = self._current()[:, : self.obs_nonpriv_dim]
        self._history_torch = torch.roll(self._history_torch, shifts=-1, dims=1)
        self._history_torch[:, -1, :] = cur
        if self._fill_history_on_reset == "repeat" and self._last_dones_torch is not None:
            done = self._last_dones_torch.view(-1).bool().to(self._device)
            # no host sync: a masked select instead of the reference's `if done_mask.any()` + boolean-index assignment
            self._history_torch = torch.where(done.view(-1, 1, 1), cur.unsqueeze(1).expand(-1, self.history_len, -1), self._history_torch)
        return out

    def get_priv_tail(self) -> torch.Tensor:
        """(N, priv_dim) encoded privileged tail of the current observation: what the teacher's mass encoder reads."""
        return self._current()[:, -self.priv_dim:]

    def get_priv_tail_from_obs(self, obs_full) -> torch.Tensor:
        obs_t = torch.from_numpy(obs_full) if isinstance(obs_full, np.ndarray) else obs_full
        if not torch.is_tensor(obs_t):
            raise TypeError("obs_full must be torch.Tensor or np.ndarray")
        obs_t = obs_t.to(self._device, dtype=torch.float32)
        if obs_t.ndim != 2 or obs_t.shape[1] < self.priv_dim:
            raise ValueError(f"obs_full must be [N, >=priv_dim], got {tuple(obs_t.shape)} priv_dim={self.priv_dim}")
        return obs_t[:, -self.priv_dim:]

    def get_masscom(self) -> torch.Tensor:
        """(N,4) teacher mass / CoM encodings.  The fused live step writes them into obs[:, -8:-4] (UsvLiveParams: mass_obs_relative,
        com_obs_scaled), i.e. the values `task.MDD.get_masses(...)` returns in the reference (:445-476)."""
        if self.priv_dim < 4:
            raise ValueError("get_masscom needs the mass + CoM columns (priv_dim >= 4)")
        return self._current()[:, -self.priv_dim: -self.priv_dim + 4] if self.priv_dim > 4 else self._current()[:, -4:]

    def observe_nonpriv(self, as_numpy: bool = True):
        return self._out(self._current()[:, : self.obs_nonpriv_dim], as_numpy)

    def observe_history(self, as_numpy: bool = True):
        return self._out(self._history_torch.reshape(self.num_envs, -1), as_numpy)

    def observe_sysid_obs(self, as_numpy: bool = True):
        """[history_flat | current non-privileged obs], (N, T*D + D): the student's input (dagger.py:50-60)."""
        return self._out(torch.cat([self._history_torch.reshape(self.num_envs, -1), self._current()[:, : self.obs_nonpriv_dim]], dim=1), as_numpy)
