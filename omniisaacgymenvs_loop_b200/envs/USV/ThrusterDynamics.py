"""Dynamics / DynamicsZeroOrder / DynamicsFirstOrder: same surface as the reference
[ref: OIGE/envs/USV/ThrusterDynamics.py:4-277]; LUT build, LUT lookup and the first-order lag are
sm_100a kernels (LUT staged in shared memory)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ... import _lib
from .Utils import f32c, randomize_rows, require_cuda


class Dynamics:
    def __init__(self, num_envs, device):
        self.num_envs = num_envs
        self.device = require_cuda(device)
        self.thrusters = torch.zeros((num_envs, 6), dtype=torch.float32, device=self.device)
        self.current_forces = torch.zeros((num_envs, 2), dtype=torch.float32, device=self.device)
        self.Reset()

    def update(self, cmd, dt):
        raise NotImplementedError()

    def Reset(self):
        self.current_forces[:, :] = 0.0


class DynamicsZeroOrder(Dynamics):
    def update(self, cmd):
        return cmd


class DynamicsFirstOrder(Dynamics):
    def __init__(self, task_cfg, num_envs, device, timeConstant, dt, numberOfPointsForInterpolation,
                 interpolationPointsFromRealDataLeft, interpolationPointsFromRealDataRight, coeff_neg_commands,
                 coeff_pos_commands, cmd_lower_range, cmd_upper_range):
        super().__init__(num_envs, device)
        self._lib = _lib.lib()
        self.tau = timeConstant
        self.idx_matrix = torch.zeros((num_envs, 2), dtype=torch.float32, device=self.device)
        self.dt = dt
        self._use_thruster_randomization = task_cfg["use_thruster_randomization"]
        self._thruster_rand = task_cfg["thruster_rand"]
        self._use_separate_randomization = task_cfg["use_separate_randomization"]
        self._left_rand = task_cfg["left_rand"]
        self._right_rand = task_cfg["right_rand"]
        ones = lambda: torch.ones((num_envs, 1), dtype=torch.float32, device=self.device)
        self.thruster_multiplier = ones()
        self.thruster_left_multiplier = ones()
        self.thruster_right_multiplier = ones()
        self.reset_thruster_randomization(torch.arange(num_envs, device=self.device), num_envs)
        self.commands = torch.linspace(cmd_lower_range, cmd_upper_range, steps=len(interpolationPointsFromRealDataLeft),
                                       device=self.device)
        self.numberOfPointsForInterpolation = numberOfPointsForInterpolation
        self.interpolationPointsFromRealDataLeft = torch.tensor(interpolationPointsFromRealDataLeft, dtype=torch.float32, device=self.device)
        self.interpolationPointsFromRealDataRight = torch.tensor(interpolationPointsFromRealDataRight, dtype=torch.float32, device=self.device)
        self.thruster_forces_before_dynamics = torch.zeros((num_envs, 2), dtype=torch.float32, device=self.device)
        self.thruster_forces_after_randomization = torch.zeros((num_envs, 2), dtype=torch.float32, device=self.device)
        self.coeff_neg_commands = torch.tensor(coeff_neg_commands, device=self.device)
        self.coeff_pos_commands = torch.tensor(coeff_pos_commands, device=self.device)
        self.interpolate_on_field_data()

    # ---- A6 ------------------------------------------------------------------------------
    def reset_thruster_randomization(self, env_ids: torch.Tensor, num_resets: int) -> None:
        """[ref :112-127] mult = U(0,1)*2r + (1-r)."""
        if self._use_thruster_randomization:
            if self._use_separate_randomization:
                randomize_rows(self.thruster_left_multiplier, env_ids, 1 - self._left_rand, 0.0, 2 * self._left_rand, stream_id=4)
                randomize_rows(self.thruster_right_multiplier, env_ids, 1 - self._right_rand, 0.0, 2 * self._right_rand, stream_id=5)
            else:
                randomize_rows(self.thruster_multiplier, env_ids, 1 - self._thruster_rand, 0.0, 2 * self._thruster_rand, stream_id=6)

    # ---- A5 ------------------------------------------------------------------------------
    def update(self, thruster_forces_before_dynamics, dt):
        """[ref :129-141] cur = cur*alpha + (1-alpha)*target with alpha = exp(fp32(-dt/tau))."""
        alpha = float(np.exp(np.float32(-dt / self.tau), dtype=np.float32))
        tgt = f32c(thruster_forces_before_dynamics, self.device)
        rc = self._lib.usv_thruster_lag_f32(_lib.ptr(self.current_forces), _lib.ptr(tgt), ctypes.c_float(alpha),
                                            _lib.ptr(self.thrusters), ctypes.c_int64(self.num_envs), _lib.stream())
        _lib.check(rc, "usv_thruster_lag_f32")
        return self.current_forces

    def compute_thrusters_constant_force(self):
        self.thrusters[:, 0] = 400
        self.thrusters[:, 3] = -400
        return self.thrusters

    def _build(self, pts):
        lut = torch.empty(self.numberOfPointsForInterpolation, dtype=torch.float32, device=self.device)
        rc = self._lib.usv_thruster_build_lut_f32(_lib.ptr(pts), ctypes.c_int32(pts.numel()), _lib.ptr(lut),
                                                  ctypes.c_int32(lut.numel()), _lib.stream())
        _lib.check(rc, "usv_thruster_build_lut_f32")
        return lut

    def interpolate_on_field_data(self):
        """[ref :152-177] 21 (or 11) field points -> numberOfPointsForInterpolation by linear interpolation."""
        self.x_linear_interp = torch.linspace(float(self.commands.min()), float(self.commands.max()), self.numberOfPointsForInterpolation)
        self.y_linear_interp_left = self._build(self.interpolationPointsFromRealDataLeft)
        self.y_linear_interp_right = self._build(self.interpolationPointsFromRealDataRight)
        self.n_left = self.numberOfPointsForInterpolation
        self.n_right = self.numberOfPointsForInterpolation

    # ---- A4 ------------------------------------------------------------------------------
    def get_cmd_interpolated(self, cmd_value):
        """[ref :179-213] LUT lookup with round-half-to-even, then the multipliers."""
        if self._use_thruster_randomization and not self._use_separate_randomization:
            mL = mR = self.thruster_multiplier
        else:
            mL, mR = self.thruster_left_multiplier, self.thruster_right_multiplier
        after = self.thruster_forces_after_randomization if self._use_thruster_randomization else None
        rc = self._lib.usv_thruster_target_f32(
            _lib.ptr(f32c(cmd_value, self.device)), _lib.ptr(self.y_linear_interp_left), _lib.ptr(self.y_linear_interp_right),
            ctypes.c_int32(self.n_left), _lib.ptr(mL), _lib.ptr(mR), _lib.ptr(self.thruster_forces_before_dynamics),
            _lib.ptr(after), ctypes.c_int64(self.num_envs), _lib.stream())
        _lib.check(rc, "usv_thruster_target_f32")

    def set_target_force(self, commands):
        self.get_cmd_interpolated(commands)

    def update_forces(self):
        """[ref :221-234] thrusters[:, [0,3]] = lag(target)."""
        if self._use_thruster_randomization:
            self.update(self.thruster_forces_after_randomization, self.dt)
        else:
            self.update(self.thruster_forces_before_dynamics, self.dt)
        return self.thrusters

    def command_to_thrusters_force_lsm(self, left_thruster_command, right_thruster_command):
        raise NotImplementedError("scalar least-squares thruster model is not on the vectorised hot path "
                                  "[ref: OIGE/envs/USV/ThrusterDynamics.py:238-277]")
