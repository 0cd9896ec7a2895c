"""Host-side configuration of the fused USV env: the reference's task YAML -> UsvStepParams.

`UsvEnvConfig.from_task_cfg` reads the same YAML tree the reference's USVVirtual.__init__ reads
[ref: SNAP/USV_Virtual.py:62-215 ; OIGE/tasks/USV_Virtual.py:295-649] (PyYAML dict; hydra
interpolations such as ${resolve_default:512,...} are resolved by `load_task_yaml`).
"""
from __future__ import annotations

import dataclasses
import math
import re
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import _lib

PEN_OFF, PEN_NEG_SUM, PEN_EXP_NEG_SUMSQ, PEN_NEG_ABS, PEN_NEG_DEADZONE, PEN_EXP_NEG_ABS = range(6)
REWARD_MODES = {"linear": 0, "square": 1, "exponential": 2}


@dataclass
class PenaltyTerm:
    form: int = PEN_OFF
    c1: float = 0.0
    c2: float = 0.0
    k: float = 0.0


_NUM = r"([-+]?\d*\.?\d+(?:[eE][-+]?\d+)?)"


def parse_penalty_lambda(src: str, c1: float = 0.0, c2: float = 0.0) -> PenaltyTerm:
    """Maps a YAML penalty lambda string onto the closed set the kernel implements.

    The reference eval()s these strings [ref: OIGE/tasks/USV/USV_task_rewards.py:429-438]; a CUDA
    kernel cannot, so only the forms that appear under cfg/task/USV/** are accepted and anything
    else raises (no silent torch fallback).  `c1`/`c2` are the dataclass constants the default
    lambdas refer to."""
    s = re.sub(r"\s+", "", src)
    s = re.sub(r"^lambdax,step:", "", s)
    s = s.replace("c1", repr(float(c1))).replace("c2", repr(float(c2)))
    pats = [
        (rf"^-torch\.sum\(x,dim=-1\)\*{_NUM}(?:\+{_NUM})?$", lambda m: PenaltyTerm(PEN_NEG_SUM, float(m[1]), float(m[2] or 0))),
        (rf"^\(torch\.exp\(-torch\.sum\(x\*\*2,dim=-1\)\)-1\.0\)\*{_NUM}$", lambda m: PenaltyTerm(PEN_EXP_NEG_SUMSQ, float(m[1]))),
        (rf"^-torch\.norm\(x,dim=-1\)\*{_NUM}(?:\+{_NUM})?$", lambda m: PenaltyTerm(PEN_NEG_ABS, float(m[1]), float(m[2] or 0))),
        (rf"^-torch\.abs\(x\)\*{_NUM}(?:\+{_NUM})?$", lambda m: PenaltyTerm(PEN_NEG_ABS, float(m[1]), float(m[2] or 0))),
        (rf"^-torch\.clamp\(torch\.abs\(x\)-{_NUM},min=0(?:\.0)?\)\*{_NUM}$", lambda m: PenaltyTerm(PEN_NEG_DEADZONE, float(m[2]), 0.0, float(m[1]))),
        (rf"^\(torch\.exp\(-{_NUM}\*torch\.abs\(x\)\)-1\.0\)\*{_NUM}$", lambda m: PenaltyTerm(PEN_EXP_NEG_ABS, float(m[2]), 0.0, float(m[1]))),
        (rf"^torch\.exp\({_NUM}\*torch\.abs\(x\)\)-1\.0$", lambda m: PenaltyTerm(PEN_EXP_NEG_ABS, 1.0, 0.0, -float(m[1]))),
    ]
    for pat, mk in pats:
        m = re.match(pat, s)
        if m:
            return mk(m)
    raise NotImplementedError(f"penalty lambda not in the supported closed set: {src!r}")


def task_section_kwargs(tp: dict) -> dict:
    """UsvEnvConfig fields of a `task_parameters` YAML section (CaptureXYParameters [ref: OIGE/tasks/USV/USV_task_parameters.py:17-52])."""
    return dict(
        position_tolerance=tp.get("position_tolerance", 0.1),
        kill_after_n_steps_in_tolerance=int(tp.get("kill_after_n_steps_in_tolerance", 1)),
        kill_dist=tp.get("kill_dist", 20.0), boundary_cost=tp.get("boundary_cost", 25.0),
        goal_reward=tp.get("goal_reward", 100.0), time_reward=tp.get("time_reward", -0.1),
        goal_random_position=tp.get("goal_random_position", 0.0),
        spawn_min_dist=tp.get("min_spawn_dist", 0.5), spawn_max_dist=tp.get("max_spawn_dist", 11.0),
        spawn_curriculum=bool(tp.get("spawn_curriculum", False)), spawn_curriculum_min_dist=tp.get("spawn_curriculum_min_dist", 0.2),
        spawn_curriculum_max_dist=tp.get("spawn_curriculum_max_dist", 3.0), spawn_curriculum_kill_dist=tp.get("spawn_curriculum_kill_dist", 30.0),
        spawn_curriculum_warmup=int(tp.get("spawn_curriculum_warmup", 250)), spawn_curriculum_end=int(tp.get("spawn_curriculum_end", 1000)))


def reward_section_kwargs(rp: dict) -> dict:
    """UsvEnvConfig fields of a `reward_parameters` section (CaptureXYReward [ref: SNAP/USV_task_rewards.py:16-38])."""
    return dict(
        reward_mode=REWARD_MODES[str(rp.get("reward_mode", "exponential")).lower()],
        position_scale=rp.get("position_scale", 1.0), exponential_reward_coeff=rp.get("exponential_reward_coeff", 0.25),
        align_la1=rp.get("align_la1", 0.02), align_la2=rp.get("align_la2", -10.0), align_la3=rp.get("align_la3", -0.1))


def penalty_section_kwargs(pp: dict) -> dict:
    """UsvEnvConfig fields of a `penalties_parameters` section (Penalties [ref: SNAP/USV_task_rewards.py:381-420]); the lambda strings are
    mapped onto the kernel's closed set (parse_penalty_lambda)."""
    def pen(name, default_src, c1d, c2d=0.0):
        if not pp.get(f"penalize_{name}", False):
            return PenaltyTerm()
        return parse_penalty_lambda(pp.get(f"penalize_{name}_fn", default_src), pp.get(f"penalize_{name}_c1", c1d),
                                    pp.get(f"penalize_{name}_c2", c2d))

    return dict(
        pen_linear_vel=pen("linear_velocities", "lambda x,step : -torch.norm(x, dim=-1)*c1 + c2", 0.01),
        pen_angular_vel=pen("angular_velocities", "lambda x,step : -torch.abs(x)*c1 + c2", 0.01),
        pen_angular_vel_variation=pen("angular_velocities_variation", "lambda x,step: torch.exp(c1 * torch.abs(x)) - 1.0", -0.033),
        pen_energy=pen("energy", "lambda x,step : -torch.sum(x**2)*c1 + c2", 0.01),
        pen_action_variation=pen("action_variation", "lambda x,step: torch.exp(c1 * torch.abs(x)) - 1.0", -0.033))


@dataclass
class UsvEnvConfig:
    """Mirror of UsvStepParams with the classic snapshot's defaults
    [ref: SNAP/USV_Virtual_CaptureXY_SysID-TEST.yaml]."""
    seed: int = 1234
    num_envs: int = 512
    dt: float = 0.02
    n_substeps: int = 5
    max_episode_length: int = 3000
    horizon_length: int = 16
    clip_actions: float = 1.0
    clip_obs: float = 12.0
    izz: float = 10.0                # yaw inertia: not recoverable from the tree (heron.urdf:69 placeholder)
    thr_y_left: float = 0.377654     # heron.urdf:242
    thr_y_right: float = -0.377654   # heron.urdf:167
    time_constant: float = 0.05
    env_spacing: float = 15.0
    envs_per_row: int = 0
    grid_row_offset: float = 0.0
    grid_col_offset: float = 0.0
    lin_fwd: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    offset_linear_damping: float = 0.0
    offset_lin_forward_damping_speed: float = 0.0
    offset_nonlin_damping: float = 0.0
    scaling_damping: float = 1.0
    use_drag_scale: bool = False
    n_lut: int = 1000
    lut_points_left: Tuple[float, ...] = (-3.8, -3.8, -3.6, -3.6, -1.6, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                                          0.0, 4.0, 10.0, 15.0, 21.0, 23.0, 22.0)
    lut_points_right: Tuple[float, ...] = (-5.0, -5.0, -5.0, -4.6, -2.2, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                                           0.0, 4.6, 10.0, 17.0, 24.0, 24.0, 23.0)
    action_affine: bool = False
    action_noise: bool = True
    action_noise_min: float = -0.05
    action_noise_max: float = 0.05
    action_bias: float = 0.0
    action_bias_steps: int = 0       # live: the bias is applied to the first N control steps only (host-side counter)
    penalties_use_u: bool = False
    noise_pos: bool = False
    pos_noise_min: float = -0.03
    pos_noise_max: float = 0.03
    noise_vel: bool = True
    vel_noise_min: float = -0.03
    vel_noise_max: float = 0.03
    noise_heading: bool = True
    heading_noise_min: float = -0.025
    heading_noise_max: float = 0.025
    use_force_disturbance: bool = False
    use_const_force: bool = False
    use_sin_force: bool = False
    use_torque_disturbance: bool = False
    use_const_torque: bool = False
    use_sin_torque: bool = False
    position_tolerance: float = 0.1
    kill_after_n_steps_in_tolerance: int = 1
    kill_dist: float = 20.0
    boundary_cost: float = 25.0
    goal_reward: float = 30.0
    time_reward: float = -0.2
    goal_speed_gate: float = 0.05
    reward_mode: int = 0
    position_scale: float = 1.0
    exponential_reward_coeff: float = 0.25
    align_la1: float = 0.02
    align_la2: float = -10.0
    align_la3: float = -0.1
    pen_linear_vel: PenaltyTerm = field(default_factory=PenaltyTerm)
    pen_angular_vel: PenaltyTerm = field(default_factory=PenaltyTerm)
    pen_angular_vel_variation: PenaltyTerm = field(default_factory=lambda: PenaltyTerm(PEN_EXP_NEG_ABS, 1.0, 0.0, 0.033))
    pen_energy: PenaltyTerm = field(default_factory=lambda: PenaltyTerm(PEN_EXP_NEG_SUMSQ, 0.01))
    pen_action_variation: PenaltyTerm = field(default_factory=PenaltyTerm)
    goal_random_position: float = 0.0
    retarget_on_reset: bool = False
    spawn_min_dist: float = 0.3
    spawn_max_dist: float = 12.0
    # optional linear curriculum on the task's `step` (= control steps / horizon_length)  [ref: SNAP/USV_capture_xy.py:231-275,346-380]
    spawn_curriculum: bool = False
    spawn_curriculum_min_dist: float = 0.2
    spawn_curriculum_max_dist: float = 3.0
    spawn_curriculum_kill_dist: float = 30.0
    spawn_curriculum_warmup: int = 250
    spawn_curriculum_end: int = 1000
    spawn_about_origin: bool = False   # live task (Variant B): annulus around the env origin, not the target
    retarget_after_spawn: bool = False # live reset order: spawn around the old target, re-draw the target last
    reset_pose_external: bool = False  # scene replay: the host writes pose / velocity / target of resetting envs
    spawn_vel_range: float = 1.5
    mass_rand: bool = False
    mass_min: float = 34.96
    mass_max: float = 36.96
    mass_base: float = 35.96
    drag_rand: bool = False
    lin_base: Tuple[float, float, float] = (0.0, 99.99, 0.82985084)
    quad_base: Tuple[float, float, float] = (17.257603, 99.99, 17.33600724)
    lin_rand_frac: Tuple[float, float, float] = (0.1, 0.1, 0.1)
    quad_rand_frac: Tuple[float, float, float] = (0.1, 0.1, 0.1)
    kdrag_rand: bool = False
    kdrag_min: float = 1.0
    kdrag_max: float = 1.0
    kdrag_log: bool = False
    thr_rand: bool = False
    thr_separate: bool = False
    thr_rand_frac: float = 0.5
    thr_left_frac: float = 0.5
    thr_right_frac: float = 0.5
    mass_coupling: bool = False
    couple_mass_max: float = 54.96
    couple_thr_a: float = 0.5
    couple_kiz_min: float = 1.0              # inertia.k_Iz_min / k_Iz_max: the range of the coupled AND of the independent k_Iz
    couple_kiz_max: float = 1.5
    couple_targets: int = 7                  # coupling.mass_driven.targets as bits: 1 drag_scale, 2 thruster, 4 yaw_inertia
    kiz_rand: bool = False                   # inertia.use_yaw_inertia_randomization (independent draw; a yaw_inertia coupling target wins)
    kiz_log: bool = False                    # inertia.k_Iz_sample_space == "log"
    use_water_current: bool = False          # env.water_current  [ref: OIGE/tasks/USV_Virtual.py:444-445]
    flow_vel_xy: Tuple[float, float] = (0.0, 0.0)
    force_const_min: float = 0.0
    force_const_max: float = 2.5
    force_sin_min: float = 0.0
    force_sin_max: float = 2.5
    force_min_freq: float = 0.25
    force_max_freq: float = 3.0
    force_min_shift: float = 0.0
    force_max_shift: float = 100.0
    torque_const_min: float = 0.0
    torque_const_max: float = 1.0
    torque_sin_min: float = 0.0
    torque_sin_max: float = 1.0
    torque_min_freq: float = 0.25
    torque_max_freq: float = 3.0
    torque_min_shift: float = 0.0
    torque_max_shift: float = 100.0

    # ------------------------------------------------------------------------------------
    def full_dr(self) -> "UsvEnvConfig":
        """DR50 flag set: every per-env randomisation switched on ('A, full DR' of SURVEY 8(d))."""
        return dataclasses.replace(
            self, use_force_disturbance=True, use_const_force=True, use_sin_force=True, use_torque_disturbance=True,
            use_const_torque=True, use_sin_torque=True, noise_pos=True, mass_rand=True, drag_rand=True, thr_rand=True)

    def curriculum(self, step: float):
        """(min_spawn_dist, max_spawn_dist, kill_dist) in force at curriculum step `step`: the reference interpolates linearly between
        the curriculum values and the final ones from `warmup` to `end` (python doubles, as the reference)."""
        if not self.spawn_curriculum:
            return self.spawn_min_dist, self.spawn_max_dist, self.kill_dist
        w, e = self.spawn_curriculum_warmup, self.spawn_curriculum_end
        if step < w:
            return self.spawn_curriculum_min_dist, self.spawn_curriculum_max_dist, self.spawn_curriculum_kill_dist
        if step > e:
            return self.spawn_min_dist, self.spawn_max_dist, self.kill_dist
        r = (step - w) / (e - w)
        lerp = lambda a, b: r * (b - a) + a
        return (lerp(self.spawn_curriculum_min_dist, self.spawn_min_dist), lerp(self.spawn_curriculum_max_dist, self.spawn_max_dist),
                lerp(self.spawn_curriculum_kill_dist, self.kill_dist))

    @property
    def lag_alpha(self) -> float:
        # torch.exp(torch.tensor(-dt/tau)) evaluated in fp32 [ref: OIGE/envs/USV/ThrusterDynamics.py:132]
        import torch
        return float(torch.exp(torch.tensor(-self.dt / self.time_constant)))

    def to_params(self, step_counter: int = 0, env_id_offset: int = 0, first_call: bool = False):
        p = _lib.UsvStepParams()
        sq = lambda v: math.sqrt(v ** 2 / 2)   # ForceDisturbance.__init__ [ref: USV_disturbances.py:281-289]
        special = {
            "seed": self.seed, "step_counter": step_counter, "env_id_offset": env_id_offset,
            "lag_alpha": self.lag_alpha, "first_call": int(first_call),
            "lin_rand": tuple(f * b for f, b in zip(self.lin_rand_frac, self.lin_base)),
            "quad_rand": tuple(f * b for f, b in zip(self.quad_rand_frac, self.quad_base)),
            "force_const_min": sq(self.force_const_min), "force_const_max": sq(self.force_const_max),
            "force_sin_min": sq(self.force_sin_min), "force_sin_max": sq(self.force_sin_max),
        }
        for name, ctype in p._fields_:
            v = special[name] if name in special else getattr(self, name)
            if isinstance(v, PenaltyTerm):
                t = getattr(p, name)
                t.form, t.c1, t.c2, t.k = int(v.form), float(v.c1), float(v.c2), float(v.k)
            elif isinstance(v, (tuple, list)):
                arr = getattr(p, name)
                for i, x in enumerate(v):
                    arr[i] = float(x)
            elif isinstance(v, bool):
                setattr(p, name, int(v))
            else:
                setattr(p, name, v)
        return p

    # ------------------------------------------------------------------------------------
    def to_task_cfg(self) -> dict:
        """The reference's task-YAML tree (env / sim / dynamics) for this config: what USVVirtual.__init__ reads."""
        inv_mode = {v: k for k, v in REWARD_MODES.items()}
        lam = {PEN_NEG_SUM: lambda t: f"lambda x,step : -torch.sum(x, dim=-1)*{t.c1!r} + {t.c2!r}",
               PEN_EXP_NEG_SUMSQ: lambda t: f"lambda x,step : (torch.exp(-torch.sum(x**2, dim=-1)) - 1.0) * {t.c1!r}",
               PEN_NEG_ABS: lambda t: f"lambda x,step : -torch.abs(x)*{t.c1!r} + {t.c2!r}",
               PEN_NEG_DEADZONE: lambda t: f"lambda x,step: -torch.clamp(torch.abs(x)-{t.k!r}, min=0)*{t.c1!r}",
               PEN_EXP_NEG_ABS: lambda t: f"lambda x,step: (torch.exp(-{t.k!r} * torch.abs(x)) - 1.0) * {t.c1!r}"}
        pen = {}
        for name, t in (("linear_velocities", self.pen_linear_vel), ("angular_velocities", self.pen_angular_vel),
                        ("angular_velocities_variation", self.pen_angular_vel_variation), ("energy", self.pen_energy),
                        ("action_variation", self.pen_action_variation)):
            pen[f"penalize_{name}"] = t.form != PEN_OFF
            if t.form != PEN_OFF:
                src = lam[t.form](t)
                if name == "linear_velocities":
                    src = src.replace("torch.abs(x)", "torch.norm(x, dim=-1)")
                pen[f"penalize_{name}_fn"] = src
        L = [self.lin_base[0], self.lin_base[1], 99.99, 13.0, 13.0, self.lin_base[2]]
        Q = [self.quad_base[0], self.quad_base[1], 10.0, 5.0, 5.0, self.quad_base[2]]
        return {
            "name": "USVVirtual",
            "env": {"numEnvs": self.num_envs, "envSpacing": self.env_spacing, "maxEpisodeLength": self.max_episode_length,
                    "action_mode": "Continuous", "horizon_length": self.horizon_length, "observation_frame": "local",
                    "controlFrequencyInv": self.n_substeps, "clipObservations": {"state": self.clip_obs}, "clipActions": self.clip_actions,
                    "water_current": {"use_water_current": self.use_water_current,
                                      "flow_velocity": [self.flow_vel_xy[0], self.flow_vel_xy[1], 0.0]},
                    "disturbances": {
                        "forces": dict(use_force_disturbance=self.use_force_disturbance, use_constant_force=self.use_const_force,
                                       use_sinusoidal_force=self.use_sin_force, force_const_min=self.force_const_min,
                                       force_const_max=self.force_const_max, force_sin_min=self.force_sin_min, force_sin_max=self.force_sin_max,
                                       force_min_freq=self.force_min_freq, force_max_freq=self.force_max_freq,
                                       force_min_shift=self.force_min_shift, force_max_shift=self.force_max_shift),
                        "torques": dict(use_torque_disturbance=self.use_torque_disturbance, use_constant_torque=self.use_const_torque,
                                        use_sinusoidal_torque=self.use_sin_torque, torque_const_min=self.torque_const_min,
                                        torque_const_max=self.torque_const_max, torque_sin_min=self.torque_sin_min,
                                        torque_sin_max=self.torque_sin_max, torque_min_freq=self.torque_min_freq,
                                        torque_max_freq=self.torque_max_freq, torque_min_shift=self.torque_min_shift,
                                        torque_max_shift=self.torque_max_shift),
                        "observations": dict(add_noise_on_pos=self.noise_pos, position_noise_min=self.pos_noise_min,
                                             position_noise_max=self.pos_noise_max, add_noise_on_vel=self.noise_vel,
                                             velocity_noise_min=self.vel_noise_min, velocity_noise_max=self.vel_noise_max,
                                             add_noise_on_heading=self.noise_heading, heading_noise_min=self.heading_noise_min,
                                             heading_noise_max=self.heading_noise_max),
                        "actions": dict(add_noise_on_act=self.action_noise, min_action_noise=self.action_noise_min,
                                        max_action_noise=self.action_noise_max),
                        "mass": dict(add_mass_disturbances=self.mass_rand, min_mass=self.mass_min, max_mass=self.mass_max,
                                     CoM_max_displacement=0.0, base_mass=self.mass_base),
                        "drag": dict(use_drag_randomization=self.drag_rand, u_linear_rand=self.lin_rand_frac[0],
                                     v_linear_rand=self.lin_rand_frac[1], w_linear_rand=0.0, p_linear_rand=0.0, q_linear_rand=0.0,
                                     r_linear_rand=self.lin_rand_frac[2], u_quad_rand=self.quad_rand_frac[0],
                                     v_quad_rand=self.quad_rand_frac[1], w_quad_rand=0.0, p_quad_rand=0.0, q_quad_rand=0.0,
                                     r_quad_rand=self.quad_rand_frac[2], use_drag_scale_randomization=self.kdrag_rand,
                                     k_drag_min=self.kdrag_min, k_drag_max=self.kdrag_max,
                                     k_drag_sample_space="log" if self.kdrag_log else "linear"),
                        "thruster": dict(use_thruster_randomization=self.thr_rand, thruster_rand=self.thr_rand_frac,
                                         use_separate_randomization=self.thr_separate, left_rand=self.thr_left_frac,
                                         right_rand=self.thr_right_frac)},
                    "task_parameters": dict(name="CaptureXY", position_tolerance=self.position_tolerance,
                                            kill_after_n_steps_in_tolerance=self.kill_after_n_steps_in_tolerance,
                                            goal_random_position=self.goal_random_position, max_spawn_dist=self.spawn_max_dist,
                                            min_spawn_dist=self.spawn_min_dist, kill_dist=self.kill_dist, boundary_cost=self.boundary_cost,
                                            goal_reward=self.goal_reward, time_reward=self.time_reward,
                                            spawn_curriculum=self.spawn_curriculum, spawn_curriculum_min_dist=self.spawn_curriculum_min_dist,
                                            spawn_curriculum_max_dist=self.spawn_curriculum_max_dist,
                                            spawn_curriculum_kill_dist=self.spawn_curriculum_kill_dist,
                                            spawn_curriculum_warmup=self.spawn_curriculum_warmup,
                                            spawn_curriculum_end=self.spawn_curriculum_end),
                    "reward_parameters": dict(name="CaptureXY", reward_mode=inv_mode[self.reward_mode], position_scale=self.position_scale,
                                              exponential_reward_coeff=self.exponential_reward_coeff, align_la1=self.align_la1,
                                              align_la2=self.align_la2, align_la3=self.align_la3),
                    "penalties_parameters": pen},
            "sim": {"dt": self.dt, "gravity": [0.0, 0.0, -9.81]},
            "dynamics": {
                "thrusters": {"cmd_lower_range": -1.0, "cmd_upper_range": 1.0, "timeConstant": self.time_constant,
                              "interpolation": {"numberOfPointsForInterpolation": self.n_lut,
                                                "interpolationPointsFromRealDataLeft": list(self.lut_points_left),
                                                "interpolationPointsFromRealDataRight": list(self.lut_points_right)},
                              "leastSquareMethod": {"neg_cmd_coeff": [0.0] * 5, "pos_cmd_coeff": [0.0] * 5}},
                "hydrodynamics": {"linear_damping": L, "quadratic_damping": Q,
                                  "linear_damping_forward_speed": [self.lin_fwd[0], self.lin_fwd[1], 0.0, 0.0, 0.0, self.lin_fwd[2]],
                                  "offset_linear_damping": self.offset_linear_damping,
                                  "offset_lin_forward_damping_speed": self.offset_lin_forward_damping_speed,
                                  "offset_nonlin_damping": self.offset_nonlin_damping, "scaling_damping": self.scaling_damping,
                                  "offset_added_mass": 0.0, "scaling_added_mass": 1.0},
                "hydrostatics": {"average_hydrostatics_force_value": 275, "amplify_torque": 1.0, "material_density": 133,
                                 "water_density": 1000, "mass": self.mass_base, "box_width": 1.0, "box_length": 1.3,
                                 "waterplane_area": 0.233333, "heron_zero_height": 0.24},
                "acceleration": {"alpha": 0.3, "last_time": -10.0}},
        }

    @classmethod
    def from_task_cfg(cls, task_cfg: dict, **overrides) -> "UsvEnvConfig":
        """Builds the config from the reference's task YAML dict (env/sim/dynamics sections)."""
        env, sim, dyn = task_cfg["env"], task_cfg["sim"], task_cfg["dynamics"]
        dist = env["disturbances"]
        f, t, o, a, m, dr, th = (dist["forces"], dist["torques"], dist["observations"], dist["actions"], dist["mass"],
                                 dist["drag"], dist["thruster"])
        tp, rp, pp = env["task_parameters"], env["reward_parameters"], env["penalties_parameters"]
        hd, hs, thr = dyn["hydrodynamics"], dyn["hydrostatics"], dyn["thrusters"]
        L, Q = hd["linear_damping"], hd["quadratic_damping"]
        fw = hd["linear_damping_forward_speed"]
        wc = env.get("water_current", {}) or {}

        num_envs = env["numEnvs"] if isinstance(env["numEnvs"], int) else 512
        clip_obs = env.get("clipObservations", {"state": 12.0})
        cfg = cls(
            num_envs=num_envs, dt=float(sim["dt"]), n_substeps=int(env["controlFrequencyInv"]),
            max_episode_length=int(env["maxEpisodeLength"]), horizon_length=int(env.get("horizon_length", 16)),
            clip_actions=float(env.get("clipActions", 1.0)),
            clip_obs=float(clip_obs["state"] if isinstance(clip_obs, dict) else clip_obs),
            time_constant=float(thr["timeConstant"]), env_spacing=float(env.get("envSpacing", 15)),
            lin_fwd=(fw[0], fw[1], fw[5]), offset_linear_damping=hd["offset_linear_damping"],
            offset_lin_forward_damping_speed=hd["offset_lin_forward_damping_speed"],
            offset_nonlin_damping=hd["offset_nonlin_damping"], scaling_damping=hd["scaling_damping"],
            use_drag_scale=bool(dr.get("use_drag_scale_randomization", False)),
            n_lut=int(thr["interpolation"]["numberOfPointsForInterpolation"]),
            lut_points_left=tuple(thr["interpolation"]["interpolationPointsFromRealDataLeft"]),
            lut_points_right=tuple(thr["interpolation"]["interpolationPointsFromRealDataRight"]),
            action_noise=bool(a["add_noise_on_act"]), action_noise_min=a["min_action_noise"], action_noise_max=a["max_action_noise"],
            noise_pos=bool(o["add_noise_on_pos"]), pos_noise_min=o["position_noise_min"], pos_noise_max=o["position_noise_max"],
            noise_vel=bool(o["add_noise_on_vel"]), vel_noise_min=o["velocity_noise_min"], vel_noise_max=o["velocity_noise_max"],
            noise_heading=bool(o["add_noise_on_heading"]), heading_noise_min=o["heading_noise_min"], heading_noise_max=o["heading_noise_max"],
            use_force_disturbance=bool(f["use_force_disturbance"]),
            use_const_force=bool(f["use_constant_force"]), use_sin_force=bool(f["use_sinusoidal_force"]),
            use_torque_disturbance=bool(t["use_torque_disturbance"]),
            use_const_torque=bool(t["use_constant_torque"]), use_sin_torque=bool(t["use_sinusoidal_torque"]),
            **task_section_kwargs(tp), **reward_section_kwargs(rp), **penalty_section_kwargs(pp),
            mass_rand=bool(m.get("add_mass_disturbances", False)), mass_min=float(m.get("min_mass", 0.0)),
            mass_max=float(m.get("max_mass", 0.0)), mass_base=float(m.get("base_mass", hs["mass"])),
            drag_rand=bool(dr["use_drag_randomization"]), lin_base=(L[0], L[1], L[5]), quad_base=(Q[0], Q[1], Q[5]),
            lin_rand_frac=(dr["u_linear_rand"], dr["v_linear_rand"], dr["r_linear_rand"]),
            quad_rand_frac=(dr["u_quad_rand"], dr["v_quad_rand"], dr["r_quad_rand"]),
            kdrag_rand=bool(dr.get("use_drag_scale_randomization", False)), kdrag_min=float(dr.get("k_drag_min", 1.0)),
            kdrag_max=float(dr.get("k_drag_max", 1.0)), kdrag_log=str(dr.get("k_drag_sample_space", "linear")) == "log",
            thr_rand=bool(th["use_thruster_randomization"]), thr_separate=bool(th["use_separate_randomization"]),
            thr_rand_frac=th["thruster_rand"], thr_left_frac=th["left_rand"], thr_right_frac=th["right_rand"],
            force_const_min=f["force_const_min"], force_const_max=f["force_const_max"], force_sin_min=f["force_sin_min"],
            force_sin_max=f["force_sin_max"], force_min_freq=f["force_min_freq"], force_max_freq=f["force_max_freq"],
            force_min_shift=f["force_min_shift"], force_max_shift=f["force_max_shift"],
            torque_const_min=t["torque_const_min"], torque_const_max=t["torque_const_max"], torque_sin_min=t["torque_sin_min"],
            torque_sin_max=t["torque_sin_max"], torque_min_freq=t["torque_min_freq"], torque_max_freq=t["torque_max_freq"],
            torque_min_shift=t["torque_min_shift"], torque_max_shift=t["torque_max_shift"],
            use_water_current=bool(wc.get("use_water_current", False)),
            flow_vel_xy=tuple(float(x) for x in (list(wc.get("flow_velocity", [0.0, 0.0, 0.0])) + [0.0, 0.0])[:2]),
        )
        return dataclasses.replace(cfg, **overrides)


@dataclass
class UsvLiveConfig:
    """Mirror of UsvLiveParams: what the live task (Variant B, CaptureXY with static obstacles) adds on top of UsvEnvConfig
    [ref: OIGE/tasks/USV_Virtual.py:97-151,440-527,837-984 ; OIGE/tasks/USV/USV_disturbances.py:88-124,153-194 ;
    OIGE/tasks/USV/USV_capture_xy_static_obs.py:30-32,103].  Defaults = IROS2024/USV_Virtual_CaptureXY_SysID-TEST.yaml."""
    priv_mode: int = 2                       # 0 raw, 1 centered, 2 minmax
    mass_obs_relative: bool = True
    com_obs_scaled: bool = True
    com_scale: Tuple[float, float, float] = (1.3, 1.0, 1.0)   # box_length, box_width, max(heron_zero_height, 1)
    priv_a: Tuple[float, float, float, float] = (1.0, 0.5, 0.5, 1.0)
    priv_b: Tuple[float, float, float, float] = (0.5, 0.5, 0.5, 0.5)
    priv_active: Tuple[bool, bool, bool, bool] = (True, True, True, True)
    com_rand: int = 1                        # 0 off, 1 box (com_displacement_xyz), 2 legacy XY disc (CoM_max_displacement in com_disp[0])
    com_base: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    com_disp: Tuple[float, float, float] = (0.15, 0.05, 0.02)
    collision_threshold: float = 1.2
    map_size: float = 30.0
    fixed_horizon_eval: bool = False
    masscom_obs_base: bool = False           # mass.masscom_obs_source == "base": the tail shows base / neutral values (evaluation ablation)
    priv_dim: int = 8                        # env.priv_dim / env.mass_dim: 8 = [mass, CoM, k_drag, thr_L, thr_R, k_Iz], 4 = [mass, CoM] (obs 29 wide)
    # Tier-3 tasks behind the same live step: 0 CaptureXY+obstacles, 1 GoToPose, 2 KeepXY, 3 TrackXYVelocity
    # [ref: OIGE/tasks/USV/USV_task_rewards.py:170-325 ; USV_task_parameters.py:95-177]
    task: int = 0
    heading_reward_mode: int = 2             # GoToPoseReward.heading_reward_mode: exponential
    heading_exponential_reward_coeff: float = 0.25
    heading_scale: float = 5.0
    sig_gain: float = 3.0
    goal_random_velocity: float = 0.75       # TrackXYVelocityParameters
    lin_vel_tolerance: float = 0.01

    def to_params(self):
        lp = _lib.UsvLiveParams()
        lp.task, lp.heading_reward_mode = int(self.task), int(self.heading_reward_mode)
        lp.heading_exponential_reward_coeff, lp.heading_scale, lp.sig_gain = self.heading_exponential_reward_coeff, self.heading_scale, self.sig_gain
        lp.goal_random_velocity, lp.lin_vel_tolerance = self.goal_random_velocity, self.lin_vel_tolerance
        lp.priv_mode, lp.mass_obs_relative, lp.com_obs_scaled = int(self.priv_mode), int(self.mass_obs_relative), int(self.com_obs_scaled)
        for j in range(3):
            # `com / (scale_t + eps)` with scale_t an fp32 tensor  [ref: USV_disturbances.py:183-190]
            lp.com_scale_eps[j] = float(np.float32(self.com_scale[j]) + np.float32(1e-6))
            lp.com_base[j], lp.com_disp[j] = float(self.com_base[j]), float(self.com_disp[j])
        for j in range(4):
            lp.priv_a[j], lp.priv_b[j], lp.priv_active[j] = float(self.priv_a[j]), float(self.priv_b[j]), int(self.priv_active[j])
        lp.com_rand = int(self.com_rand)
        lp.masscom_obs_base = int(self.masscom_obs_base)
        if int(self.priv_dim) not in (4, 8):
            raise ValueError(f"Unsupported priv_dim/mass_dim={self.priv_dim}. Supported: 4 or 8.")      # [ref: USV_Virtual.py:485-488]
        lp.priv_dim = int(self.priv_dim)
        for j in range(4):
            # minmax: `0.5 * (min + max)` in Python floats, then an fp32 tensor; raw / centered: ones  [ref: USV_Virtual.py:859-880]
            lp.priv_neutral[j] = float(np.float32(0.5 * (float(self.priv_a[j]) + (float(self.priv_a[j]) + float(self.priv_b[j]))))) \
                if int(self.priv_mode) == 2 else 1.0
        lp.collision_threshold, lp.map_size, lp.fixed_horizon_eval = self.collision_threshold, self.map_size, int(self.fixed_horizon_eval)
        return lp

    @classmethod
    def from_task_cfg(cls, task_cfg: dict) -> "UsvLiveConfig":
        env = task_cfg["env"]
        dist = env["disturbances"]
        m, dr, th, inr = dist["mass"], dist.get("drag", {}) or {}, dist.get("thruster", {}) or {}, dist.get("inertia", {}) or {}
        hs = task_cfg["dynamics"]["hydrostatics"]
        cp = (dist.get("coupling", {}) or {}).get("mass_driven", {}) or {}
        targets = set(cp.get("targets", []) or []) if cp.get("enabled", False) else set()
        pp = env.get("privileged_params", {}) or {}
        mode = {"raw": 0, "centered": 1, "minmax": 2}[str(pp.get("mode", "raw"))]
        nominal = float(pp.get("nominal", 1.0))
        a = float(th.get("thruster_rand", 0.0))
        rng = [(float(dr.get("k_drag_min", 1.0)), float(dr.get("k_drag_max", 1.0))),
               (1.0 - a, 1.0 if "thruster" in targets else 1.0 + a), (1.0 - a, 1.0 if "thruster" in targets else 1.0 + a),
               (float(inr.get("k_Iz_min", 1.0)), float(inr.get("k_Iz_max", 1.0)))]
        active = [bool(dr.get("use_drag_scale_randomization", False)) or "drag_scale" in targets,
                  bool(th.get("use_thruster_randomization", False)) or "thruster" in targets,
                  bool(th.get("use_thruster_randomization", False)) or "thruster" in targets,
                  bool(inr.get("use_yaw_inertia_randomization", False)) or "yaw_inertia" in targets]
        if mode == 1:
            pa = [nominal] * 4
            pb = [max(abs(lo - nominal), abs(hi - nominal), 1e-6) for lo, hi in rng]
        else:
            pa = [lo for lo, _ in rng]
            pb = [hi - lo for lo, hi in rng]
            active = [act and (hi - lo) > 1e-6 for act, (lo, hi) in zip(active, rng)]   # degenerate range -> neutral 0
            pb = [b if b > 1e-6 else 1.0 for b in pb]
        source = str(m.get("masscom_obs_source", "sim"))
        if source not in ("sim", "base"):
            raise ValueError(f"mass.masscom_obs_source must be 'sim' or 'base', got {source}")      # [ref: USV_Virtual.py:469-472]
        disp = m.get("com_displacement_xyz", None)
        legacy = float(m.get("CoM_max_displacement", 0.0) or 0.0)       # legacy XY disc, used when no box is given  [ref: USV_disturbances.py:108-124]
        com_rand = 0
        if bool(m.get("add_mass_disturbances", False)):
            com_rand = 1 if disp is not None else (2 if legacy > 0.0 else 0)
        if com_rand == 2:
            disp = [legacy, 0.0, 0.0]
        scale = m.get("com_obs_scale", None) or [hs["box_length"], hs["box_width"], max(hs["heron_zero_height"], 1.0)]
        return cls(priv_mode=mode, mass_obs_relative=str(m.get("mass_obs_mode", "raw")) == "relative",
                   com_obs_scaled=str(m.get("com_obs_mode", "raw")) == "scaled", com_scale=tuple(float(x) for x in scale),
                   priv_a=tuple(pa), priv_b=tuple(pb), priv_active=tuple(active),
                   com_rand=com_rand,
                   com_base=tuple(float(x) for x in m.get("base_com", [0.0, 0.0, 0.0])),
                   com_disp=tuple(float(x) for x in (disp or [0.0, 0.0, 0.0])),
                   fixed_horizon_eval=bool(env.get("fixed_horizon_eval", False)), masscom_obs_base=source == "base",
                   priv_dim=int(env.get("priv_dim", env.get("mass_dim", 4))))           # [ref: USV_Virtual.py:484]


COUPLE_BITS = {"drag_scale": 1, "thruster": 2, "yaw_inertia": 4}


def live_env_config(task_cfg: dict, **overrides) -> "UsvEnvConfig":
    """UsvEnvConfig for the LIVE USVVirtual (OIGE/tasks/USV_Virtual.py): from_task_cfg + the keys only the live file reads
    (action_processing :590-615, mass-driven coupling :415-438, live reward dataclass defaults USV_task_rewards.py:27-32)."""
    env = task_cfg["env"]
    dist = env["disturbances"]
    ap = env.get("action_processing", {}) or {}
    cp = (dist.get("coupling", {}) or {}).get("mass_driven", {}) or {}
    targets = set(cp.get("targets", []) or []) if cp.get("enabled", False) else set()
    unknown = targets - set(COUPLE_BITS)
    if unknown:
        raise ValueError(f"coupling.mass_driven.targets: unknown target(s) {sorted(unknown)}")
    rp, m, dr, th, inr = env["reward_parameters"], dist["mass"], dist["drag"], dist["thruster"], dist.get("inertia", {}) or {}
    space = str(inr.get("k_Iz_sample_space", "linear"))
    if space not in ("linear", "log"):
        raise ValueError(f"k_Iz_sample_space must be 'linear' or 'log', got {space}")           # [ref: USV_Virtual.py:426-429]
    live = dict(
        action_affine=bool(ap.get("use_affine_thrust_mapping", True)), penalties_use_u=bool(ap.get("penalties_use_thrust_u", False)),
        action_bias=float(ap.get("initial_action_bias", 0.0)), action_bias_steps=int(ap.get("initial_action_bias_steps", 0)),
        spawn_about_origin=True, retarget_on_reset=True, retarget_after_spawn=True, goal_speed_gate=float("inf"),
        position_scale=rp.get("position_scale", 1.5), align_la1=rp.get("align_la1", 0.04),
        mass_coupling=bool(targets), couple_mass_max=float(m.get("max_mass", 0.0)), couple_thr_a=float(th.get("thruster_rand", 0.0)),
        couple_kiz_min=float(inr.get("k_Iz_min", 1.0)), couple_kiz_max=float(inr.get("k_Iz_max", 1.0)),
        couple_targets=sum(COUPLE_BITS[t] for t in targets) if targets else 7,
        kiz_rand=bool(inr.get("use_yaw_inertia_randomization", False)), kiz_log=space == "log",
        use_drag_scale=bool(dr.get("use_drag_scale_randomization", False)) or "drag_scale" in targets,
    )
    live.update(overrides)
    return UsvEnvConfig.from_task_cfg(task_cfg, **live)


_LIVE_LUT = (0.0,) * 11 + (8.0, 16.0, 24.0, 32.0, 40.0, 48.0, 56.0, 64.0, 72.0, 80.0)


def live_default_config(**overrides) -> "UsvEnvConfig":
    """UsvEnvConfig of the live task as shipped  [ref: OIGE/cfg/task/USV/IROS2024/USV_Virtual_CaptureXY_SysID-TEST.yaml]."""
    base = dict(
        n_substeps=10, max_episode_length=200, env_spacing=20.0, use_drag_scale=True, lut_points_left=_LIVE_LUT, lut_points_right=_LIVE_LUT,
        action_affine=True, action_noise=False, action_bias=-0.6, action_bias_steps=20000, penalties_use_u=True,
        position_tolerance=1.0, goal_reward=20.0, time_reward=-0.05, goal_speed_gate=float("inf"), position_scale=1.5,
        exponential_reward_coeff=0.15, align_la1=0.04, pen_angular_vel=PenaltyTerm(PEN_NEG_DEADZONE, 0.02, 0.0, 0.4),
        pen_angular_vel_variation=PenaltyTerm(PEN_NEG_DEADZONE, 0.005, 0.0, 0.1), pen_energy=PenaltyTerm(PEN_NEG_SUM, 0.005, 0.0, 0.0),
        retarget_on_reset=True, retarget_after_spawn=True, spawn_min_dist=9.0, spawn_about_origin=True, mass_rand=True, mass_max=54.96, mass_base=34.96,
        kdrag_max=1.5, mass_coupling=True)
    base.update(overrides)
    return UsvEnvConfig(**base)


def live_task_cfg(cfg: Optional["UsvEnvConfig"] = None, live: Optional["UsvLiveConfig"] = None) -> dict:
    """The live task-YAML tree for (cfg, live): to_task_cfg() plus the keys only the live USVVirtual reads."""
    cfg = cfg if cfg is not None else live_default_config()
    live = live if live is not None else UsvLiveConfig()
    t = cfg.to_task_cfg()
    env, dist = t["env"], t["env"]["disturbances"]
    env["mass_dim"] = int(live.priv_dim)
    env["privileged_params"] = {"mode": ("raw", "centered", "minmax")[live.priv_mode], "nominal": 1.0}
    env["fixed_horizon_eval"] = live.fixed_horizon_eval
    env["action_processing"] = {"use_affine_thrust_mapping": cfg.action_affine, "initial_action_bias": cfg.action_bias,
                                "initial_action_bias_steps": cfg.action_bias_steps, "penalties_use_thrust_u": cfg.penalties_use_u}
    dist["coupling"] = {"mass_driven": {"enabled": cfg.mass_coupling,
                                        "targets": [t for t, bit in COUPLE_BITS.items() if int(cfg.couple_targets) & bit]}}
    dist["mass"].update(com_displacement_xyz=list(live.com_disp) if int(live.com_rand) == 1 else None,
                        CoM_max_displacement=float(live.com_disp[0]) if int(live.com_rand) == 2 else 0.0, base_com=list(live.com_base),
                        apply_com_to_sim=True, mass_obs_mode="relative" if live.mass_obs_relative else "raw",
                        com_obs_mode="scaled" if live.com_obs_scaled else "raw",
                        masscom_obs_source="base" if live.masscom_obs_base else "sim")
    dist["drag"]["use_drag_scale_randomization"] = cfg.kdrag_rand
    dist["thruster"]["thruster_rand"] = cfg.couple_thr_a if cfg.mass_coupling else cfg.thr_rand_frac
    dist["inertia"] = {"use_yaw_inertia_randomization": bool(cfg.kiz_rand), "k_Iz_min": cfg.couple_kiz_min, "k_Iz_max": cfg.couple_kiz_max,
                       "k_Iz_sample_space": "log" if cfg.kiz_log else "linear"}
    t["dynamics"]["hydrostatics"].update(box_length=live.com_scale[0], box_width=live.com_scale[1])
    return t


def load_task_yaml(path: str, num_envs: Optional[int] = None) -> dict:
    """PyYAML loader for the reference's task files; resolves the hydra interpolations the env reads
    (${resolve_default:D,${...x}} -> D or the override; everything else is left as a string)."""
    import yaml

    with open(path) as fh:
        txt = fh.read()

    def rd(m):
        return str(num_envs) if (num_envs is not None and "num_envs" in m.group(2)) else m.group(1)

    txt = re.sub(r"\$\{resolve_default:([^,]+),\$\{([^}]*)\}\}", rd, txt)
    return yaml.safe_load(txt)
