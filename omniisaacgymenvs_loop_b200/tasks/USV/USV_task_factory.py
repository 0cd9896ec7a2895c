"""task_factory.get(task_dict, reward_dict, num_envs, device, priv_dim)  [ref: OIGE/tasks/USV/USV_task_factory.py:28-64].

The step-by-step task-class surface exists for the classic CaptureXY task (tasks/USV/USV_capture_xy.py).  The live factory's tasks
(CaptureXY with static obstacles, GoToPose, KeepXY, TrackXYVelocity) run fused inside tasks/USV_Virtual.py -> engine.FusedUsvLiveEnv
(`UsvLiveParams.task`); asking the factory for one of them names that entry instead of returning a half-working object."""
from __future__ import annotations

from .USV_capture_xy import CaptureXYTask

FUSED_ONLY = ("GoToPose", "KeepXY", "TrackXYVelocity")


class TaskFactory:
    def __init__(self):
        self.creators = {}

    def register(self, name: str, task) -> None:
        self.creators[name] = task

    def get(self, task_dict: dict, reward_dict: dict, num_envs: int, device: str, priv_dim: int = 4):
        if task_dict["name"] != reward_dict["name"]:
            raise ValueError("task_parameters.name and reward_parameters.name must match")
        mode = task_dict["name"]
        if mode in FUSED_ONLY:
            raise NotImplementedError(f"{mode} runs fused in tasks.USV_Virtual.USVVirtual (engine.FusedUsvLiveEnv); it has no per-call task class")
        if mode not in self.creators:
            raise KeyError(f"unknown task mode {mode!r}")
        return self.creators[mode](task_dict, reward_dict, num_envs, device, priv_dim=priv_dim)


task_factory = TaskFactory()
task_factory.register("CaptureXY", CaptureXYTask)
