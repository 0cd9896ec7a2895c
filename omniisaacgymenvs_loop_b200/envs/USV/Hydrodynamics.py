"""HydrodynamicsObject: same class / ctor / attribute / method surface as the reference
[ref: OIGE/envs/USV/Hydrodynamics.py:6-245]; ComputeHydrodynamicsEffects is one sm_100a kernel."""
from __future__ import annotations

import ctypes

import torch

from ... import _lib
from .Utils import f32c, randomize_rows, require_cuda


class HydrodynamicsObject:
    def __init__(self, task_cfg, num_envs, device, water_density, gravity, linear_damping, quadratic_damping,
                 linear_damping_forward_speed, offset_linear_damping, offset_lin_forward_damping_speed, offset_nonlin_damping,
                 scaling_damping, offset_added_mass, scaling_added_mass, alpha, last_time):
        self.device = require_cuda(device)
        self._lib = _lib.lib()
        self._use_drag_randomization = task_cfg["use_drag_randomization"]
        self._use_drag_scale_randomization = bool(task_cfg.get("use_drag_scale_randomization", False))
        self._k_drag_min = float(task_cfg.get("k_drag_min", 1.0))
        self._k_drag_max = float(task_cfg.get("k_drag_max", 1.0))
        self._k_drag_sample_space = str(task_cfg.get("k_drag_sample_space", "linear"))
        if self._k_drag_sample_space not in ("linear", "log"):
            raise ValueError(f"k_drag_sample_space must be 'linear' or 'log', got {self._k_drag_sample_space}")
        t = lambda v: torch.tensor(v, dtype=torch.float32, device=self.device)
        names = ("u", "v", "w", "p", "q", "r")
        self._linear_rand = t([task_cfg[f"{k}_linear_rand"] * linear_damping[i] for i, k in enumerate(names)])
        self._quad_rand = t([task_cfg[f"{k}_quad_rand"] * quadratic_damping[i] for i, k in enumerate(names)])
        self._num_envs = num_envs
        self.drag = torch.zeros((num_envs, 6), dtype=torch.float32, device=self.device)
        self.drag_scale = torch.ones((num_envs, 1), dtype=torch.float32, device=self.device)
        self.linear_damping_base = linear_damping
        self.quadratic_damping_base = quadratic_damping
        self.linear_damping = t([linear_damping] * num_envs)
        self.quadratic_damping = t([quadratic_damping] * num_envs)
        self.linear_damping_forward_speed = t(linear_damping_forward_speed)
        self.offset_linear_damping = offset_linear_damping
        self.offset_lin_forward_damping_speed = offset_lin_forward_damping_speed
        self.offset_nonlin_damping = offset_nonlin_damping
        self.scaling_damping = scaling_damping
        all_ids = torch.arange(num_envs, device=self.device)
        if self._use_drag_scale_randomization:
            self.drag_scale[:, :] = self._sample_k_drag(num_envs)
        if self._use_drag_randomization:
            self._redraw_coefficients(all_ids)
        # added mass: allocated and stored, never used by any computation (as in the reference :105-115)
        self._Ca = torch.zeros((6, 6), device=self.device)
        self.added_mass = torch.zeros((num_envs, 6), device=self.device)
        self.offset_added_mass = offset_added_mass
        self.scaling_added_mass = scaling_added_mass
        self.alpha = alpha
        self._filtered_acc = torch.zeros(6, device=self.device)
        self._last_time = last_time
        self._last_vel_rel = torch.zeros(6, device=self.device)
        self.local_velocities = torch.zeros((num_envs, 6), dtype=torch.float32, device=self.device)

    # ---- randomisation (A3) --------------------------------------------------------------
    def _sample_k_drag(self, n: int) -> torch.Tensor:
        """[ref :119-134] (n,1) samples of k_drag (uniform or log-uniform)."""
        if n <= 0:
            return torch.ones((0, 1), dtype=torch.float32, device=self.device)
        kmin, kmax = float(self._k_drag_min), float(self._k_drag_max)
        if kmin <= 0.0 or kmax <= 0.0:
            raise ValueError(f"k_drag_min/max must be > 0, got {kmin}, {kmax}")
        if kmax < kmin:
            raise ValueError(f"k_drag_max must be >= k_drag_min, got {kmin}, {kmax}")
        out = torch.empty((n, 1), dtype=torch.float32, device=self.device)
        randomize_rows(out, torch.arange(n, device=self.device), 0.0, kmin, kmax, stream_id=1,
                       log_space=self._k_drag_sample_space == "log")
        return out

    def _redraw_coefficients(self, env_ids):
        base_l = torch.tensor(self.linear_damping_base, dtype=torch.float32, device=self.device)
        base_q = torch.tensor(self.quadratic_damping_base, dtype=torch.float32, device=self.device)
        randomize_rows(self.linear_damping, env_ids, base_l, -self._linear_rand, self._linear_rand, stream_id=2)
        randomize_rows(self.quadratic_damping, env_ids, base_q, -self._quad_rand, self._quad_rand, stream_id=3)

    def reset_coefficients(self, env_ids: torch.Tensor, num_resets: int) -> None:
        """[ref :136-174]"""
        if self._use_drag_randomization:
            self._redraw_coefficients(env_ids)
        if self._use_drag_scale_randomization:
            self.drag_scale[env_ids, :] = self._sample_k_drag(num_resets)

    # ---- forces (A2) -----------------------------------------------------------------------
    def _params(self, use_water_current=False, flow_vel=(0.0, 0.0, 0.0)):
        p = _lib.UsvHydrodynamicsParams()
        fwd = self.linear_damping_forward_speed.tolist() if torch.is_tensor(self.linear_damping_forward_speed) else list(self.linear_damping_forward_speed)
        for i in range(6):
            p.linear_damping_forward_speed[i] = float(fwd[i])
        p.offset_linear_damping = float(self.offset_linear_damping)
        p.offset_lin_forward_damping_speed = float(self.offset_lin_forward_damping_speed)
        p.offset_nonlin_damping = float(self.offset_nonlin_damping)
        p.scaling_damping = float(self.scaling_damping)
        p.use_drag_scale = int(bool(self._use_drag_scale_randomization))
        p.use_water_current = int(bool(use_water_current))
        fv = flow_vel.tolist() if torch.is_tensor(flow_vel) else list(flow_vel)
        for i in range(3):
            p.flow_vel[i] = float(fv[i])
        return p

    def _launch(self, quat, vel6, drag, local, damp, p):
        n = vel6.shape[0]
        rc = self._lib.usv_hydrodynamics_f32(_lib.ptr(quat), _lib.ptr(vel6), _lib.ptr(self.linear_damping), _lib.ptr(self.quadratic_damping),
                                             _lib.ptr(self.drag_scale), _lib.ptr(drag), _lib.ptr(local), _lib.ptr(damp),
                                             ctypes.c_int64(n), ctypes.byref(p), _lib.stream())
        _lib.check(rc, "usv_hydrodynamics_f32")

    def ComputeDampingMatrix(self, vel):
        """[ref :176-205] (N,6) diagonal damping for body velocities `vel`."""
        n = vel.shape[0]
        ident = torch.zeros((n, 4), dtype=torch.float32, device=self.device)
        ident[:, 0] = 1.0
        damp = torch.empty((n, 6), dtype=torch.float32, device=self.device)
        scratch = torch.empty((n, 6), dtype=torch.float32, device=self.device)
        self._launch(ident, f32c(vel, self.device), scratch, None, damp, self._params())
        return damp

    def ComputeHydrodynamicsEffects(self, time, quaternions, world_vel, use_water_current, flow_vel):
        """[ref :207-245] drag (N,6) = -D(v_body) * v_body, v_body = [R^T v, R^T w] (minus the current)."""
        n = world_vel.shape[0]
        drag = torch.empty((n, 6), dtype=torch.float32, device=self.device)
        local = torch.empty((n, 6), dtype=torch.float32, device=self.device)
        p = self._params(use_water_current, flow_vel if use_water_current else (0.0, 0.0, 0.0))
        self._launch(f32c(quaternions, self.device), f32c(world_vel, self.device), drag, local, None, p)
        self.local_velocities = local
        self.local_lin_velocities = local[:, :3]
        self.local_ang_velocities = local[:, 3:]
        self.drag = drag
        return self.drag
