"""NPZ scene replay for the live task: deterministic scenes (goal, obstacles, start pose / velocity) instead of the random spawn
[ref: OIGE/tasks/USV_Virtual.py:1329-1457 (_scene_replay_load_npz / _take_scene_indices / _apply) ;
 OIGE/tasks/USV/USV_capture_xy_static_obs.py:785-905 (CaptureXYTask.apply_scene)].

The file format is the reference's: arrays `obstacles_xy (S,k,2+)`, `obstacles_count (S,)`, `start_pos (S,2+)`, `start_yaw (S,)`,
`start_vel (S,2+)`, `goal_pos (S,2+)`, all in the env-local frame.  As in the reference, the env ids that reset are read on the host
(one sync per control step: this is the evaluation path, not the training hot path); their scenes are written into the engine's
buffers, the potential fields of exactly that reset batch are rebuilt by the scene kernels (batch-global maxima, like
BatchedMapGPU on the subset), and the fused live step then runs with `reset_pose_external` so that it keeps the replayed pose."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from .engine import FusedUsvLiveEnv

E = _lib.ENUMS
REQUIRED = ("obstacles_xy", "obstacles_count", "start_pos", "start_yaw", "start_vel", "goal_pos")


class SceneReplay:
    def __init__(self, env: FusedUsvLiveEnv, npz_path: str, cycle: bool = True):
        if not env.cfg.reset_pose_external:
            raise ValueError("scene replay needs UsvEnvConfig.reset_pose_external=True (the kernel must keep the replayed pose)")
        if env.task != 0:
            raise ValueError("scene replay drives the obstacle task (UsvLiveConfig.task == 0)")
        path = os.path.abspath(npz_path)
        if not os.path.exists(path):
            raise FileNotFoundError(f"scene_replay.npz_path not found: {path}")
        with np.load(path, allow_pickle=True) as npz:
            missing = [k for k in REQUIRED if k not in npz.files]
            if missing:
                raise KeyError(f"scene_replay npz missing keys={missing}; found={list(npz.files)}")
            self.data = {k: np.array(npz[k]) for k in REQUIRED}
        self.num_scenes = int(self.data["start_pos"].shape[0])
        if self.num_scenes <= 0:
            raise ValueError(f"scene_replay npz has no scenes: start_pos.shape={self.data['start_pos'].shape}")
        self.env, self.cycle, self.path = env, bool(cycle), path
        self.next_scene_idx = torch.zeros(env.num_envs, dtype=torch.long)
        self.last_scene_idx = torch.full((env.num_envs,), -1, dtype=torch.long)

    def take_scene_indices(self, env_ids: torch.Tensor) -> torch.Tensor:
        """One scene index per env in env_ids; every env walks the file with its own counter  [ref :1372-1393]."""
        env_cpu = env_ids.detach().cpu().long()
        idx = self.next_scene_idx[env_cpu].clone()
        self.next_scene_idx[env_cpu] = idx + 1
        if self.cycle:
            idx = idx % self.num_scenes
        elif bool((idx < 0).any()) or bool((idx >= self.num_scenes).any()):
            raise IndexError(f"scene_replay index out of range: idx={idx.tolist()} num_scenes={self.num_scenes}")
        self.last_scene_idx[env_cpu] = idx
        return idx

    def scenes(self, idx: torch.Tensor):
        """(start_pos, start_yaw, start_vel, goal, obstacles (n,16,2) with the unused slots in limbo) of the scene indices."""
        d, i = self.data, idx.numpy().astype(np.int64)
        f = lambda k, cols: torch.from_numpy(np.asarray(d[k], dtype=np.float32)[i][..., :cols] if cols else np.asarray(d[k], dtype=np.float32)[i])
        obst = torch.from_numpy(np.asarray(d["obstacles_xy"], dtype=np.float32)[i][..., :2])
        count = torch.from_numpy(np.asarray(d["obstacles_count"], dtype=np.int64)[i].reshape(-1))
        big, n, k = E["USV_B_OBSTACLES"], obst.shape[0], obst.shape[1]
        limbo = torch.tensor([999.0, 999.0])
        obst = torch.cat([obst, limbo.view(1, 1, 2).repeat(n, big - k, 1)], dim=1) if k < big else obst[:, :big]
        keep = torch.arange(big).view(1, -1) < count.view(-1, 1).clamp(min=0, max=big)        # apply_scene :838-841
        obst = torch.where(keep.unsqueeze(-1), obst, limbo.view(1, 1, 2))
        return f("start_pos", 2), f("start_yaw", 0).reshape(-1), f("start_vel", 2), f("goal_pos", 2), obst

    def apply(self, env_ids: torch.Tensor) -> Optional[torch.Tensor]:
        """Writes the next scene of every env in env_ids into the engine (goal, obstacles, pose, velocity) and rebuilds the
        potential fields of this batch.  Returns the scene indices used."""
        if env_ids.numel() == 0:
            return None
        env, dev = self.env, self.env.device
        idx = self.take_scene_indices(env_ids)
        pos, yaw, vel, goal, obst = self.scenes(idx)
        ids = env_ids.to(dev, torch.long)
        for name, v in (("USV_C_TX", goal[:, 0]), ("USV_C_TY", goal[:, 1]), ("USV_S_X", pos[:, 0]), ("USV_S_Y", pos[:, 1]),
                        ("USV_S_PSI", yaw), ("USV_S_VX", vel[:, 0]), ("USV_S_VY", vel[:, 1]), ("USV_S_R", torch.zeros_like(yaw))):
            env.set_field(name, v.contiguous().to(dev), ids)
        flat = obst.reshape(obst.shape[0], -1).to(dev)
        for j in range(flat.shape[1]):
            env.bconsts[ids >> 5, E["USV_BC_OBST"] + j, ids & 31] = flat[:, j]
        env.potential[ids] = env.build_fields(obst.to(dev), goal.to(dev))
        return idx

    def step(self, actions: torch.Tensor):
        """One control step of the live env with replayed scenes for the envs that reset (== the reference's reset_idx with
        scene_replay_enabled followed by the step)."""
        env = self.env
        ids = env.reset_buf.nonzero(as_tuple=False).squeeze(-1)       # host sync, as `reset_buf.nonzero()` in the reference
        self.apply(ids)
        return env.step(actions, rebuild_scene=False)
