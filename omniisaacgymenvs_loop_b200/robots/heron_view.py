"""PlanarHeronView: the simulator surface the reference's USVVirtual drives, on the planar rigid-body integrator
[ref: omniisaacgymenvs/robots/articulations/views/heron_view.py ; calls in OIGE/tasks/USV_Virtual.py:772-774 (get_world_poses /
get_velocities), :1119-1133 (base / thruster_left / thruster_right .apply_forces_and_torques_at_pos), :1177-1182, :1567-1572 (set_*),
world.step() in OIGE/envs/vec_env_rlgames.py:154-171].

The fused engines never go through this class (one launch per control step).  It exists so that code written against the reference's
view -- an unmodified `apply_forces()` that evaluates the force modules and pushes wrenches per sub-step -- can be pointed at the
same integrator the fused kernel uses: state (x, y, psi, vx, vy, r) per env, heave / roll / pitch frozen,
only the planar components of the pushed 6-DOF wrenches enter.  Two kernels: usv_planar_wrench_accumulate_f32 (every apply_* call) and
usv_planar_rigid_step_f32 (world.step())."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from .. import _lib

# thruster mounts in the body frame (heron.urdf:167,242): left = +y
THRUSTER_X, THRUSTER_Y = -0.53, 0.377654


class _Body:
    """One rigid prim of the articulation as the task sees it: apply_forces_and_torques_at_pos on the shared planar wrench."""

    def __init__(self, view: "PlanarHeronView", offset_x: float = 0.0, offset_y: float = 0.0):
        self._view, self._ox, self._oy = view, float(offset_x), float(offset_y)

    def apply_forces_and_torques_at_pos(self, forces: Optional[torch.Tensor] = None, torques: Optional[torch.Tensor] = None,
                                        positions: Optional[torch.Tensor] = None, indices=None, is_global: bool = True) -> None:
        if positions is not None or indices is not None:
            raise NotImplementedError("the USV task applies wrenches at the prim origins of all envs (USV_Virtual.py:1119-1133)")
        v = self._view
        f = None if forces is None else forces.to(v.device, torch.float32).contiguous()
        t = None if torques is None else torques.to(v.device, torch.float32).contiguous()
        rc = _lib.lib().usv_planar_wrench_accumulate_f32(_lib.ptr(v.wrench), _lib.ptr(f), _lib.ptr(t), _lib.ptr(v.pose), ctypes.c_float(self._ox),
                                                         ctypes.c_float(self._oy), ctypes.c_int32(int(bool(is_global))), ctypes.c_int64(v.count),
                                                         _lib.stream())
        _lib.check(rc, "usv_planar_wrench_accumulate_f32")

    # mass / inertia access used by MassDistributionDisturbances.set_masses and the k_Iz randomisation (USV_Virtual.py:180-289,1526)
    def get_masses(self, indices=None, clone: bool = True) -> torch.Tensor:
        m = self._view.mass if indices is None else self._view.mass[indices]
        return m.clone() if clone else m

    def set_masses(self, masses: torch.Tensor, indices=None) -> None:
        v = self._view
        if indices is None:
            v.mass.copy_(masses.reshape(-1))
        else:
            v.mass[indices] = masses.reshape(-1).to(v.device, torch.float32)

    def get_inertias(self, indices=None, clone: bool = True) -> torch.Tensor:
        """(n, 9) row-major inertia tensors; only Izz (entry 8) is dynamic here."""
        v = self._view
        izz = v.izz if indices is None else v.izz[indices]
        out = torch.zeros((izz.shape[0], 9), dtype=torch.float32, device=v.device)
        out[:, 0] = out[:, 4] = 1.0
        out[:, 8] = izz
        return out

    def set_inertias(self, inertias: torch.Tensor, indices=None) -> None:
        v = self._view
        izz = inertias.reshape(-1, 9)[:, 8].to(v.device, torch.float32)
        if indices is None:
            v.izz.copy_(izz)
        else:
            v.izz[indices] = izz


class PlanarHeronView:
    def __init__(self, num_envs: int, device="cuda:0", mass: float = 34.96, izz: float = 10.0, name: str = "heron_view"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.UsvLibraryError("PlanarHeronView runs on CUDA only (no CPU fallback)")
        self.count, self.name, self.num_dof = int(num_envs), name, 2
        f32 = dict(dtype=torch.float32, device=self.device)
        self.pose, self.vel, self.wrench = torch.zeros((self.count, 3), **f32), torch.zeros((self.count, 3), **f32), torch.zeros((self.count, 3), **f32)
        self.mass, self.izz = torch.full((self.count,), float(mass), **f32), torch.full((self.count,), float(izz), **f32)
        self.z = torch.zeros(self.count, **f32)                       # frozen heave: reported back, never integrated
        self.base = _Body(self)
        self.thruster_left, self.thruster_right = _Body(self, THRUSTER_X, THRUSTER_Y), _Body(self, THRUSTER_X, -THRUSTER_Y)
        self._joints = torch.zeros((self.count, self.num_dof), **f32)

    # ---- what USVVirtual.update_state / post_reset read ---------------------------------------------------------------------
    def get_world_poses(self, indices=None, clone: bool = True):
        """(positions [n,3], orientations [n,4] as (w, x, y, z)): planar pose, yaw-only quaternion."""
        p = self.pose if indices is None else self.pose[indices]
        z = self.z if indices is None else self.z[indices]
        half = 0.5 * p[:, 2]
        zero = torch.zeros_like(half)
        return torch.stack([p[:, 0], p[:, 1], z], 1), torch.stack([torch.cos(half), zero, zero, torch.sin(half)], 1)

    def get_velocities(self, indices=None, clone: bool = True) -> torch.Tensor:
        """[n,6]: linear (vx, vy, 0) and angular (0, 0, r) world velocities."""
        v = self.vel if indices is None else self.vel[indices]
        out = torch.zeros((v.shape[0], 6), dtype=torch.float32, device=self.device)
        out[:, 0], out[:, 1], out[:, 5] = v[:, 0], v[:, 1], v[:, 2]
        return out

    def set_world_poses(self, positions=None, orientations=None, indices=None) -> None:
        idx = slice(None) if indices is None else indices
        if positions is not None:
            self.pose[idx, 0], self.pose[idx, 1] = positions[:, 0], positions[:, 1]
            self.z[idx] = positions[:, 2]
        if orientations is not None:                                    # yaw of a (w, x, y, z) quaternion
            w, x, y, z = orientations.unbind(1)
            self.pose[idx, 2] = torch.atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))

    def set_velocities(self, velocities: torch.Tensor, indices=None) -> None:
        idx = slice(None) if indices is None else indices
        self.vel[idx, 0], self.vel[idx, 1], self.vel[idx, 2] = velocities[:, 0], velocities[:, 1], velocities[:, 5]

    def get_joint_positions(self, indices=None, clone: bool = True) -> torch.Tensor:
        return (self._joints if indices is None else self._joints[indices]).clone()

    get_joint_velocities = get_joint_positions

    def set_joint_positions(self, positions, indices=None) -> None:
        return None                                                     # the thruster joints are fixed in the planar model

    set_joint_velocities = set_joint_positions

    # ---- world.step() ----------------------------------------------------------------------------------------------------
    def step(self, dt: float) -> None:
        rc = _lib.lib().usv_planar_rigid_step_f32(_lib.ptr(self.pose), _lib.ptr(self.vel), _lib.ptr(self.wrench), _lib.ptr(self.mass),
                                                  _lib.ptr(self.izz), ctypes.c_float(dt), ctypes.c_int64(self.count), _lib.stream())
        _lib.check(rc, "usv_planar_rigid_step_f32")


class PlanarWorld:
    """`world.step(render=False)` / `is_playing()` / `get_physics_dt()` of the Isaac Sim World for the views registered with it."""

    def __init__(self, physics_dt: float):
        self._dt, self._views = float(physics_dt), []
        self.current_time_step_index = 0

    def add(self, view: PlanarHeronView) -> PlanarHeronView:
        self._views.append(view)
        return view

    def is_playing(self) -> bool:
        return True

    def get_physics_dt(self) -> float:
        return self._dt

    def step(self, render: bool = False) -> None:
        for v in self._views:
            v.step(self._dt)
        self.current_time_step_index += 1
