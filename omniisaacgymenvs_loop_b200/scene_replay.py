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


# ---- writing scene files: the reference's scene builder  [ref: OIGE/scripts/build_usv_scenes.py:560-740] -------------------------------
def snapshot_scenes(engine, seed: int = 0, max_obstacles: Optional[int] = None) -> dict:
    """The scene of EVERY env of a live engine right after `env.reset()` (flag + one zero-action step), in the arrays of the reference's
    scene file.  The reference builder resets a vec-env N times and snapshots env 0 each time (:580-640); here one reset of an N-env
    engine yields N scenes.  Obstacles keep their slots, limbo ones included (the reference stores `xunlian_pos` as it is and sets
    `obstacles_count` to the number of slots, :598-611)."""
    n = int(engine.num_envs)
    f = lambda name: engine.field(name).detach().float().cpu().numpy().astype(np.float32)
    obst = engine.obstacles.detach().float().cpu().numpy().astype(np.float32)
    k = int(obst.shape[1]) if max_obstacles is None else min(int(obst.shape[1]), int(max_obstacles))
    return {
        "num_episodes": n,
        "episode_idx": np.arange(n, dtype=np.int32),
        "seed": (int(seed) + np.arange(n)).astype(np.int64),
        "max_obstacles": k,
        "obstacles_xy": np.ascontiguousarray(obst[:, :k, :2]),
        "obstacles_count": np.full((n,), k, dtype=np.int32),
        "start_pos": np.stack([f("USV_S_X"), f("USV_S_Y")], axis=1),
        "start_yaw": f("USV_S_PSI"),
        "start_vel": np.stack([f("USV_S_VX"), f("USV_S_VY")], axis=1),
        "goal_pos": np.stack([f("USV_C_TX"), f("USV_C_TY")], axis=1),
    }


def save_scenes(out_dir: str, scenes: dict, task_name: str = "USV", generator_cfg: Optional[dict] = None) -> str:
    """Writes `<task>__scenes__N<N>__seed<seed>.npz` (compressed, via a temporary file) and its `.sha1` side-car, with the keys of the
    reference builder (:704-737); returns the path of the npz."""
    import datetime
    import hashlib
    import json
    import re

    missing = [k for k in REQUIRED if k not in scenes]
    if missing:
        raise KeyError(f"scene dict misses keys={missing}")
    n = int(np.asarray(scenes["start_pos"]).shape[0])
    seed0 = int(np.asarray(scenes.get("seed", [0])).reshape(-1)[0])
    safe = re.sub(r"[^A-Za-z0-9_.-]+", "_", str(task_name)).strip("_") or "task"
    os.makedirs(out_dir, exist_ok=True)
    final_path = os.path.join(out_dir, f"{safe}__scenes__N{n}__seed{seed0}.npz")
    tmp_path = final_path + ".tmp.npz"
    cfg = {"task_name": str(task_name), "num_episodes": n, "seed": seed0, "max_obstacles": int(scenes.get("max_obstacles", np.asarray(scenes["obstacles_xy"]).shape[1]))}
    cfg.update(generator_cfg or {})
    data = {k: scenes[k] for k in ("episode_idx", "seed", *REQUIRED) if k in scenes}
    data.update(num_episodes=n, max_obstacles=cfg["max_obstacles"], generator_cfg=json.dumps(cfg), created_at=datetime.datetime.now().isoformat())
    np.savez_compressed(tmp_path, **data)
    sha1 = hashlib.sha1()
    with open(tmp_path, "rb") as fp:
        for chunk in iter(lambda: fp.read(8192), b""):
            sha1.update(chunk)
    os.replace(tmp_path, final_path)
    with open(final_path + ".sha1", "w") as fp:
        fp.write(sha1.hexdigest())
    return final_path


def load_scenes(npz_path: str, verify_sha1: bool = False) -> dict:
    """The arrays `SceneReplay` consumes (same key check as `_scene_replay_load_npz`, USV_Virtual.py:1329-1370); `verify_sha1` compares
    the file with its `.sha1` side-car when one exists."""
    import hashlib

    path = os.path.abspath(npz_path)
    if not os.path.exists(path):
        raise FileNotFoundError(f"scene_replay.npz_path not found: {path}")
    if verify_sha1 and os.path.exists(path + ".sha1"):
        want = open(path + ".sha1").read().strip()
        got = hashlib.sha1(open(path, "rb").read()).hexdigest()
        if want != got:
            raise ValueError(f"scene file {path} does not match its .sha1 side-car ({got} != {want})")
    with np.load(path, allow_pickle=True) as npz:
        missing = [k for k in REQUIRED if k not in npz.files]
        if missing:
            raise KeyError(f"scene_replay npz missing keys={missing}; found={list(npz.files)}")
        return {k: np.array(npz[k]) for k in npz.files}
