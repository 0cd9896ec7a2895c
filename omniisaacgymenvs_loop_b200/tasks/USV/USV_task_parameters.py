"""Task parameter sections of the USV task YAMLs  [ref: OIGE/tasks/USV/USV_task_parameters.py:17-52].  Field names are the YAML keys."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class CaptureXYParameters:
    position_tolerance: float = 0.1
    kill_after_n_steps_in_tolerance: int = 1
    goal_random_position: float = 0.0
    max_spawn_dist: float = 11.0
    min_spawn_dist: float = 0.5
    kill_dist: float = 20.0
    boundary_cost: float = 25.0
    goal_reward: float = 100.0
    time_reward: float = -0.1
    spawn_curriculum: bool = False
    spawn_curriculum_min_dist: float = 0.2
    spawn_curriculum_max_dist: float = 3.0
    spawn_curriculum_kill_dist: float = 30.0
    spawn_curriculum_mode: str = "linear"
    spawn_curriculum_warmup: int = 250
    spawn_curriculum_end: int = 1000

    def __post_init__(self) -> None:
        if str(self.spawn_curriculum_mode).lower() != "linear":
            raise ValueError("spawn_curriculum_mode: only 'linear' exists")

    def as_section(self) -> dict:
        """The YAML section back (what config.task_section_kwargs reads)."""
        return {k: getattr(self, k) for k in self.__dataclass_fields__ if k != "spawn_curriculum_mode"}
