"""Reward / penalty parameter classes of the classic CaptureXY task on the C-ABI kernel
[ref: SNAP/USV_task_rewards.py:16-76 (CaptureXYReward), :381-506 (Penalties)].

`CaptureXYReward` is parameters only: its arithmetic runs inside usv_capturexy_obs_reward_done_f32 together with the rest of
CaptureXYTask.compute_reward.  `Penalties.compute_penalty(state, actions, step)` keeps the reference's call and its cross-call state
(previous angular velocity and previous action sum, both device tensors) and launches the same entry point with the PENALTY bit."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Optional

import torch

from ... import _lib
from ...config import UsvEnvConfig, penalty_section_kwargs

CXY = {k[len("USV_CXY_"):]: v for k, v in _lib.ENUMS.items() if k.startswith("USV_CXY_")}


@dataclass
class CaptureXYReward:
    reward_mode: str = "exponential"
    position_scale: float = 1.0
    exponential_reward_coeff: float = 0.25
    align_la1: float = 0.02
    align_la2: float = -10.0
    align_la3: float = -0.1

    def __post_init__(self) -> None:
        if str(self.reward_mode).lower() not in ("linear", "square", "exponential"):
            raise ValueError(f"reward_mode {self.reward_mode!r}: linear, square or exponential")

    def as_section(self) -> dict:
        return {k: getattr(self, k) for k in self.__dataclass_fields__}


def launch_capturexy(io: "_lib.UsvCaptureXYIO", n: int, params) -> None:
    rc = _lib.lib().usv_capturexy_obs_reward_done_f32(ctypes.byref(io), ctypes.c_int64(n), ctypes.byref(params), _lib.stream())
    _lib.check(rc, "usv_capturexy_obs_reward_done_f32")


def state_pointers(io, state: dict) -> None:
    """Fills the four state pointers of a UsvCaptureXYIO from the reference's `current_state` dict (fp32 CUDA tensors)."""
    f32 = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
    keep = []
    for key, attr in (("position", "position"), ("orientation", "heading"), ("linear_velocity", "linear_velocity"),
                      ("angular_velocity", "angular_velocity")):
        if key in state and state[key] is not None:
            t = f32(state[key])
            keep.append(t)
            setattr(io, attr, t.data_ptr())
            if not t.is_cuda:
                raise _lib.UsvLibraryError("the USV task classes run on CUDA tensors only (no CPU fallback)")
    return keep


@dataclass
class Penalties:
    """The YAML keys of `penalties_parameters`; the `_fn` strings must come from the closed set config.parse_penalty_lambda knows."""
    penalize_linear_velocities: bool = False
    penalize_linear_velocities_fn: str = "lambda x,step : -torch.norm(x, dim=-1)*c1 + c2"
    penalize_linear_velocities_c1: float = 0.01
    penalize_linear_velocities_c2: float = 0.0
    penalize_angular_velocities: bool = False
    penalize_angular_velocities_fn: str = "lambda x,step : -torch.abs(x)*c1 + c2"
    penalize_angular_velocities_c1: float = 0.01
    penalize_angular_velocities_c2: float = 0.0
    penalize_angular_velocities_variation: bool = False
    penalize_angular_velocities_variation_fn: str = "lambda x,step: torch.exp(c1 * torch.abs(x)) - 1.0"
    penalize_angular_velocities_variation_c1: float = -0.033
    penalize_energy: bool = False
    penalize_energy_fn: str = "lambda x,step : -torch.sum(x**2)*c1 + c2"
    penalize_energy_c1: float = 0.01
    penalize_energy_c2: float = 0.0
    penalize_action_variation: bool = False
    penalize_action_variation_fn: str = "lambda x,step: torch.exp(c1 * torch.abs(x)) - 1.0"
    penalize_action_variation_c1: float = -0.033
    _prev_w: Optional[torch.Tensor] = field(default=None, repr=False)
    _prev_asum: Optional[torch.Tensor] = field(default=None, repr=False)
    _params: object = field(default=None, repr=False)

    def __post_init__(self) -> None:
        sect = {k: getattr(self, k) for k in self.__dataclass_fields__ if not k.startswith("_")}
        cfg = UsvEnvConfig(**penalty_section_kwargs(sect))       # raises on a lambda outside the closed set
        self._params = cfg.to_params()

    def compute_penalty(self, state: dict, actions: torch.Tensor, step: int) -> torch.Tensor:
        n = actions.shape[0]
        dev = actions.device
        first = self._prev_w is None
        if first:
            self._prev_w = torch.zeros(n, dtype=torch.float32, device=dev)
            self._prev_asum = torch.zeros(n, dtype=torch.float32, device=dev)
        io = _lib.UsvCaptureXYIO()
        keep = state_pointers(io, state)
        act = actions if (actions.dtype == torch.float32 and actions.is_contiguous()) else actions.float().contiguous()
        zeros2 = torch.zeros((n, 2), dtype=torch.float32, device=dev)
        if not io.position:
            io.position = zeros2.data_ptr()
        io.target = zeros2.data_ptr()
        out = torch.empty(n, dtype=torch.float32, device=dev)
        terms = torch.empty((n, 5), dtype=torch.float32, device=dev)
        io.actions, io.penalty, io.penalty_terms = _lib.ptr(act).value, out.data_ptr(), terms.data_ptr()
        io.prev_angular_velocity, io.prev_action_sum = self._prev_w.data_ptr(), self._prev_asum.data_ptr()
        io.what, io.first_penalty = CXY["PENALTY"], int(first)
        launch_capturexy(io, n, self._params)
        del keep
        (self.linear_vel_penalty, self.angular_vel_penalty, self.angular_vel_variation_penalty, self.energy_penalty,
         self.action_variation_penalty) = terms.unbind(1)
        return out

    def get_stats_name(self) -> list:
        pairs = (("penalize_linear_velocities", "linear_vel_penalty"), ("penalize_angular_velocities", "angular_vel_penalty"),
                 ("penalize_angular_velocities_variation", "angular_vel_variation_penalty"), ("penalize_energy", "energy_penalty"),
                 ("penalize_action_variation", "action_variation_penalty"))
        return [name for flag, name in pairs if getattr(self, flag)]
