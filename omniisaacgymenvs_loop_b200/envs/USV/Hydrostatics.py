"""HydrostaticsObject: same class / ctor / method surface as the reference
[ref: OIGE/envs/USV/Hydrostatics.py:12-133], one sm_100a kernel per call instead of ~14 ATen launches."""
from __future__ import annotations

import ctypes

import torch

from ... import _lib
from .Utils import f32c, require_cuda


class HydrostaticsObject:
    def __init__(self, num_envs, device, water_density, gravity, metacentric_width, metacentric_length,
                 average_hydrostatics_force_value, amplify_torque, offset_added_mass, scaling_added_mass, alpha, last_time):
        self._num_envs = num_envs
        self.device = require_cuda(device)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self.drag = z(num_envs, 6)
        self.water_density = water_density
        self.gravity = gravity
        self.metacentric_width = metacentric_width
        self.metacentric_length = metacentric_length
        self.archimedes_force_global = z(num_envs, 3)
        self.archimedes_torque_global = z(num_envs, 3)
        self.archimedes_force_local = z(num_envs, 3)
        self.archimedes_torque_local = z(num_envs, 3)
        self.average_hydrostatics_force_value = average_hydrostatics_force_value
        self.amplify_torque = amplify_torque
        # added-mass / acceleration members are stored but inert in the reference too (:52-57)
        self.offset_added_mass = offset_added_mass
        self.scaling_added_mass = scaling_added_mass
        self.alpha = alpha
        self._filtered_acc = z(6)
        self._last_time = last_time
        self._last_vel_rel = z(6)
        self._lib = _lib.lib()

    def _params(self):
        return _lib.UsvHydrostaticsParams(self.water_density, self.gravity, self.metacentric_width, self.metacentric_length,
                                          self.average_hydrostatics_force_value, self.amplify_torque)

    def _run(self, submerged_volume, rpy, quaternions):
        n = submerged_volume.shape[0]
        out = torch.empty((n, 6), dtype=torch.float32, device=self.device)
        p = self._params()
        rc = self._lib.usv_hydrostatics_f32(
            _lib.ptr(f32c(submerged_volume, self.device)), _lib.ptr(f32c(rpy, self.device)), _lib.ptr(f32c(quaternions, self.device)),
            _lib.ptr(out), _lib.ptr(self.archimedes_force_global), _lib.ptr(self.archimedes_torque_global),
            ctypes.c_int64(n), ctypes.byref(p), _lib.stream())
        _lib.check(rc, "usv_hydrostatics_f32")
        return out

    def compute_archimedes_metacentric_global(self, submerged_volume, rpy):
        """[ref :63-98] returns (force_global, torque_global), both (N,3)."""
        n = submerged_volume.shape[0]
        ident = torch.zeros((n, 4), dtype=torch.float32, device=self.device)
        ident[:, 0] = 1.0
        self._run(submerged_volume, rpy, ident)
        return self.archimedes_force_global, self.archimedes_torque_global

    def compute_archimedes_metacentric_local(self, submerged_volume, rpy, quaternions):
        """[ref :100-133] (N,6) = [R^T F_global, torque_global * amplify_torque]."""
        out = self._run(submerged_volume, rpy, quaternions)
        self.archimedes_force_local = out[:, :3]
        self.archimedes_torque_local = self.archimedes_torque_global      # [ref :123] not rotated
        return out
