"""Shared helpers for the GPU parity tests (oracle <-> engine state transfer)."""
import dataclasses

import numpy as np
import torch

from oracle import usv_oracle as O


def oracle_cfg(cfg) -> "O.EnvConfig":
    """product UsvEnvConfig -> oracle EnvConfig (same field names; the oracle stays independent of the product)."""
    d = dataclasses.asdict(cfg)
    fields = {f.name for f in dataclasses.fields(O.EnvConfig)}
    kw = {}
    for k, v in d.items():
        if k not in fields:
            continue
        if isinstance(v, dict) and set(v) == {"form", "c1", "c2", "k"}:
            v = O.PenaltyTerm(**v)
        elif isinstance(v, list):
            v = tuple(v)
        kw[k] = v
    return O.EnvConfig(**kw)


STATE_MAP = [("USV_S_X", lambda o: o.pos[:, 0]), ("USV_S_Y", lambda o: o.pos[:, 1]), ("USV_S_PSI", lambda o: o.psi),
             ("USV_S_VX", lambda o: o.vel[:, 0]), ("USV_S_VY", lambda o: o.vel[:, 1]), ("USV_S_R", lambda o: o.r),
             ("USV_S_THR_L", lambda o: o.current_forces[:, 0]), ("USV_S_THR_R", lambda o: o.current_forces[:, 1]),
             ("USV_S_PREV_D", lambda o: o.prev_d), ("USV_S_PREV_W", lambda o: o.prev_w), ("USV_S_PREV_ASUM", lambda o: o.prev_asum)]
CONST_MAP = [("USV_C_TX", lambda o: o.target[:, 0]), ("USV_C_TY", lambda o: o.target[:, 1]), ("USV_C_MASS", lambda o: o.mass),
             ("USV_C_LIN_U", lambda o: o.linear_damping[:, 0]), ("USV_C_LIN_V", lambda o: o.linear_damping[:, 1]),
             ("USV_C_LIN_R", lambda o: o.linear_damping[:, 5]), ("USV_C_QUAD_U", lambda o: o.quadratic_damping[:, 0]),
             ("USV_C_QUAD_V", lambda o: o.quadratic_damping[:, 1]), ("USV_C_QUAD_R", lambda o: o.quadratic_damping[:, 5]),
             ("USV_C_KDRAG", lambda o: o.drag_scale[:, 0]), ("USV_C_THR_ML", lambda o: o.thr_mult_left),
             ("USV_C_THR_MR", lambda o: o.thr_mult_right), ("USV_C_KIZ", lambda o: o.k_iz),
             ("USV_C_FCX", lambda o: o.f_const[:, 0]), ("USV_C_FCY", lambda o: o.f_const[:, 1]),
             ("USV_C_FXF", lambda o: o.f_freq[:, 0]), ("USV_C_FYF", lambda o: o.f_freq[:, 1]),
             ("USV_C_FXS", lambda o: o.f_shift[:, 0]), ("USV_C_FYS", lambda o: o.f_shift[:, 1]), ("USV_C_FAMP", lambda o: o.f_amp),
             ("USV_C_TC", lambda o: o.t_const), ("USV_C_TF", lambda o: o.t_freq), ("USV_C_TS", lambda o: o.t_shift),
             ("USV_C_TAMP", lambda o: o.t_amp)]


def push_oracle_state(orc, env):
    """Overwrites the engine's SoA state with the oracle's AoS state (so a step starts from identical inputs)."""
    dev = env.device
    for name, get in STATE_MAP + CONST_MAP:
        env.set_field(name, get(orc).to(dev))
    env.set_field("USV_S_GOAL_CNT", orc.goal_reached.to(dev))
    env.set_field("USV_S_PROGRESS", orc.progress_buf.to(torch.int32).to(dev))
    env.reset_buf.copy_(orc.reset_buf.to(dev))
    env.step_counter = orc.step_counter
    env.curriculum_step = getattr(orc, "curriculum_step", 0.0)
    env.first_call = orc.first_call


def engine_state(env):
    """dict of cpu tensors keyed like STATE_MAP/CONST_MAP names."""
    out = {name: env.field(name).detach().cpu().clone() for name, _ in STATE_MAP + CONST_MAP}
    out["goal"] = env.goal_reached.cpu().clone()
    out["progress"] = env.progress_buf.cpu().clone()
    out["reset"] = env.reset_buf.cpu().clone()
    return out


def oracle_state(orc):
    out = {name: get(orc).clone() for name, get in STATE_MAP + CONST_MAP}
    out["goal"] = orc.goal_reached.clone()
    out["progress"] = orc.progress_buf.clone()
    out["reset"] = orc.reset_buf.clone()
    return out


def assert_close(got, want, rtol=1e-5, atol=1e-6, what=""):
    got = torch.as_tensor(got).detach().cpu().double()
    want = torch.as_tensor(want).detach().cpu().double()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} outside rtol={rtol} atol={atol}; worst flat idx {i}: "
                             f"got {got.flatten()[i].item():.9g} want {want.flatten()[i].item():.9g}")
