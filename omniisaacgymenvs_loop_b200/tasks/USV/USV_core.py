"""Helpers shared by the USV task classes  [ref: OIGE/tasks/USV/USV_core.py:17-60, parse_data_dict].

The reference's `Core` owns the per-env observation buffers and writes them with torch ops; here the observation tensor is produced
by the kernels (csrc/usv_step.cu:post_classic, csrc/usv_step_b.cu:post_live), so only the small pieces a caller touches remain."""
from __future__ import annotations

import dataclasses
from typing import Any


def parse_data_dict(target: Any, data: dict, ask_for_validation: bool = False) -> Any:
    """Copies the keys of a YAML section onto a parameter dataclass instance; unknown keys are an error (the reference ignores them
    with a warning and optionally asks on stdin -- a silent typo in a reward parameter is not something a training run should survive).
    The `name` key of task / reward sections is the factory selector, not a parameter."""
    names = {f.name for f in dataclasses.fields(target)} if dataclasses.is_dataclass(target) else set(vars(target))
    unknown = [k for k in data if k not in names and k != "name"]
    if unknown:
        raise KeyError(f"{type(target).__name__}: unknown parameter(s) {unknown}")
    for k, v in data.items():
        if k != "name":
            setattr(target, k, v)
    post = getattr(target, "__post_init__", None)
    if post is not None:
        post()
    return target
