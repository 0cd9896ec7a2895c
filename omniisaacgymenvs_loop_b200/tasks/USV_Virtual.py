"""USVVirtual: the RLTask surface of the reference [ref: SNAP/USV_Virtual.py:56-866 (classic), OIGE/tasks/USV_Virtual.py,
OIGE/tasks/base/rl_task.py:53-303] over the fused sm_100a env step.

Where the reference interleaves ~250 eager torch launches, 5-10 PhysX steps and a dozen host syncs per control step,
this class launches ONE kernel in post_physics_step(); pre_physics_step() only records the actions, apply_forces() and
update_state() between sub-steps are no-ops because the sub-step loop runs inside the kernel.  The method names, buffers
(obs_buf / rew_buf / reset_buf / progress_buf / extras) and attributes the callers read stay as they are.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib
from ..config import UsvEnvConfig, UsvLiveConfig, live_env_config
from ..engine import FusedUsvEnv, FusedUsvLiveEnv, LIVE_OBS_DIM, OBS_DIM
from ..envs.USV.Hydrodynamics import HydrodynamicsObject
from ..envs.USV.Hydrostatics import HydrostaticsObject
from ..envs.USV.ThrusterDynamics import DynamicsFirstOrder
from ..utils import spaces

E = _lib.ENUMS
_STAT_KEYS = ["distance_reward", "alignment_reward", "velocity_reward", "position_error", "velocity_norm", "boundary_penalty",
              "boundary_dist", "linear_vel_penalty", "angular_vel_penalty", "angular_vel_variation_penalty", "energy_penalty",
              "action_variation_penalty", "normed_linear_vel", "normed_angular_vel", "actions_sum"]
_PENALTY_FLAGS = {"linear_vel_penalty": "pen_linear_vel", "angular_vel_penalty": "pen_angular_vel",
                  "angular_vel_variation_penalty": "pen_angular_vel_variation", "energy_penalty": "pen_energy",
                  "action_variation_penalty": "pen_action_variation"}


# live task (Variant B): episode_sums keys in the order of the kernel's USV_BST_* rows  [ref: OIGE/tasks/USV/
# USV_capture_xy_static_obs.py:130-187 ; OIGE/tasks/USV/USV_task_rewards.py Penalties.create_stats ; OIGE/tasks/USV_Virtual.py:586-601]
_LIVE_STAT_KEYS = [k[len("USV_BST_"):].lower() for k, _ in sorted(((k, v) for k, v in E.items() if k.startswith("USV_BST_") and k != "USV_BST_COUNT"),
                                                                 key=lambda kv: kv[1])]


def is_live_task_cfg(task_cfg: dict) -> bool:
    """The live USVVirtual reads env.action_processing / env.mass_dim / env.privileged_params; the classic snapshot has none."""
    env = task_cfg["env"]
    return any(k in env for k in ("action_processing", "mass_dim", "privileged_params"))


class SimConfig:
    """Minimal stand-in for omniisaacgymenvs.utils.config_utils.sim_config.SimConfig: `.config` and `.task_config`."""

    def __init__(self, config: dict):
        self.config = config
        self.task_config = config["task"]


class USVVirtual:
    def __init__(self, name: str, sim_config, env, offset=None, collect_stats: bool = True, variant: Optional[str] = None) -> None:
        self._sim_config = sim_config
        self._cfg = sim_config.config
        self._task_cfg = sim_config.task_config
        self._name = name
        self._env = env
        self._device = self._cfg.get("sim_device", "cuda:0")
        self.device = self._device
        self.rl_device = self._cfg.get("rl_device", self._device)
        # which USVVirtual this YAML belongs to: the classic snapshot (13-dim CaptureXY) or the live file (33-dim, obstacles)
        self._live = (variant == "live") if variant is not None else is_live_task_cfg(self._task_cfg)
        seed = int(self._cfg.get("seed", 1234))
        if self._live:
            self.cfg = live_env_config(self._task_cfg, seed=seed)
            self.live_cfg = UsvLiveConfig.from_task_cfg(self._task_cfg)
        else:
            self.cfg = UsvEnvConfig.from_task_cfg(self._task_cfg, seed=seed)
        envc = self._task_cfg["env"]
        self._num_envs = self.num_envs = int(self.cfg.num_envs)
        self._max_episode_length = self.cfg.max_episode_length
        self._observation_frame = envc.get("observation_frame", "local")
        if self._observation_frame != "local":
            raise NotImplementedError("only observation_frame='local' is valid in the reference (SNAP/USV_core.py:33-40 writes 14 "
                                      "columns into a 13-wide buffer in 'world' mode)")
        self._discrete_actions = envc.get("action_mode", "Continuous")
        if self._discrete_actions != "Continuous":
            raise NotImplementedError("only Continuous action_mode is on the USV PPO path")
        self.control_frequency_inv = self.cfg.n_substeps
        self.clip_obs = envc.get("clipObservations", {"state": 12.0})
        self.clip_actions = self.cfg.clip_actions
        self.randomize_actions = False
        self.randomize_observations = False
        # live: 3 + 20 + 2 + priv_dim (33, or 29 with env.mass_dim / priv_dim = 4)  [ref: OIGE/tasks/USV/USV_core.py:23-52]
        self._num_observations = self.num_observations = (LIVE_OBS_DIM - 8 + int(self.live_cfg.priv_dim)) if self._live else OBS_DIM
        self._num_actions = self.num_actions = 2
        self._max_actions = 2
        self.num_states = 0
        self.dt = self.cfg.dt
        self.step = 0
        self._nan_probe = os.getenv("USV_NAN_PROBE", "1") != "0"
        self._nan_probe_interval = int(self._cfg.get("nan_probe_interval", self.cfg.horizon_length))
        self._calls = 0
        off = int(self._cfg.get("env_id_offset", 0))
        if self._live:
            self.engine = FusedUsvLiveEnv(self.cfg, self.live_cfg, self._num_envs, self._device, env_id_offset=off, collect_stats=collect_stats)
        else:
            self.engine = FusedUsvEnv(self.cfg, self._num_envs, self._device, env_id_offset=off, collect_stats=collect_stats)
        self.set_action_and_observation_spaces()
        self.cleanup()
        self.actions = torch.zeros((self._num_envs, 2), device=self._device, dtype=torch.float32)
        self.episode_sums = {k: None for k in self._stat_names()}
        self.get_USV_dynamics()

    # ---- spaces / buffers  [ref: SNAP/USV_Virtual.py:272-348] ---------------------------------------
    def set_action_and_observation_spaces(self) -> None:
        self.observation_space = spaces.Dict({"state": spaces.Box(np.ones(self._num_observations) * -np.inf,
                                                                  np.ones(self._num_observations) * np.inf)})
        self.action_space = spaces.Box(low=np.array([-1.0, -1.0]), high=np.array([1.0, 1.0]), dtype=np.float32)
        self.state_space = spaces.Box(np.ones(self.num_states) * -np.inf, np.ones(self.num_states) * np.inf)

    def cleanup(self) -> None:
        self.obs_buf = {"state": self.engine.obs}
        self.states_buf = torch.zeros((self._num_envs, self.num_states), device=self._device, dtype=torch.float)
        self.rew_buf = self.engine.rew
        self.reset_buf = self.engine.reset_buf          # int64, starts at ones
        self.extras = {}

    @property
    def progress_buf(self) -> torch.Tensor:
        return self.engine.progress_buf

    # the evaluation scripts switch the observation source of the privileged tail at run time
    # [ref: OIGE/tasks/USV_Virtual.py:468-472,840-880 ; OIGE/scripts/rlgames_play_loopz.py:1100-1110 (_set_obs_source)]
    @property
    def _masscom_obs_source(self) -> str:
        return "base" if (self._live and self.live_cfg.masscom_obs_base) else "sim"

    @_masscom_obs_source.setter
    def _masscom_obs_source(self, mode: str) -> None:
        if mode not in ("sim", "base"):
            raise ValueError(f"mass.masscom_obs_source must be 'sim' or 'base', got {mode}")
        if not self._live:
            if mode == "base":
                raise NotImplementedError("the classic 13-dim observation has no privileged tail")
            return
        import dataclasses
        self.live_cfg = dataclasses.replace(self.live_cfg, masscom_obs_base=(mode == "base"))
        self.engine.live = self.live_cfg
        self.engine._live_params = self.live_cfg.to_params()

    def _stat_names(self):
        on = lambda k: k not in _PENALTY_FLAGS or getattr(self.cfg, _PENALTY_FLAGS[k]).form != 0
        return [k for k in (_LIVE_STAT_KEYS if self._live else _STAT_KEYS) if on(k)]

    # ---- force-layer objects kept as attributes  [ref: SNAP/USV_Virtual.py:419-468] -------------------
    def get_USV_dynamics(self):
        dyn, dist = self._task_cfg["dynamics"], self._task_cfg["env"]["disturbances"]
        hs, hd, th = dyn["hydrostatics"], dyn["hydrodynamics"], dyn["thrusters"]
        acc = dyn.get("acceleration", {"alpha": 0.3, "last_time": -10.0})
        n, dev = self._num_envs, self._device
        self.hydrostatics = HydrostaticsObject(n, dev, hs["water_density"], self._task_cfg["sim"]["gravity"][2], hs["box_width"] / 2,
                                               hs["box_length"] / 2, hs["average_hydrostatics_force_value"], hs["amplify_torque"],
                                               hd["offset_added_mass"], hd["scaling_added_mass"], acc["alpha"], acc["last_time"])
        self.hydrodynamics = HydrodynamicsObject(dist["drag"], n, dev, hs["water_density"], self._task_cfg["sim"]["gravity"][2],
                                                 hd["linear_damping"], hd["quadratic_damping"], hd["linear_damping_forward_speed"],
                                                 hd["offset_linear_damping"], hd["offset_lin_forward_damping_speed"],
                                                 hd["offset_nonlin_damping"], hd["scaling_damping"], hd["offset_added_mass"],
                                                 hd["scaling_added_mass"], acc["alpha"], acc["last_time"])
        it = th["interpolation"]
        self.thrusters_dynamics = DynamicsFirstOrder(dist["thruster"], n, dev, th["timeConstant"], self.dt,
                                                     it["numberOfPointsForInterpolation"], it["interpolationPointsFromRealDataLeft"],
                                                     it["interpolationPointsFromRealDataRight"], th["leastSquareMethod"]["neg_cmd_coeff"],
                                                     th["leastSquareMethod"]["pos_cmd_coeff"], th["cmd_lower_range"], th["cmd_upper_range"])

    # ---- RLTask surface -------------------------------------------------------------------------------
    def reset(self):
        """RLTask.reset: flag every env for reset  [ref: OIGE/tasks/base/rl_task.py:268-270]."""
        self.reset_buf.fill_(1)
        if self._live:
            self.engine.mark_host_reset()

    def pre_physics_step(self, actions: torch.Tensor) -> None:
        """Records the (already clamped) actions; resets of flagged envs, action noise, LUT lookup run inside the fused kernel
        in exactly this position  [ref: SNAP/USV_Virtual.py:571-617]."""
        self._fill_episode_extras()
        self.actions = actions.to(self._device, torch.float32).contiguous()

    def apply_forces(self) -> None:      # [ref: SNAP/USV_Virtual.py:619-650] -- inside the kernel's sub-step loop
        return

    def update_state(self) -> None:
        """current_state dict (un-noised read-back of the SoA state)  [ref: SNAP/USV_Virtual.py:470-530]."""
        e = self.engine
        psi = e.field("USV_S_PSI")
        self.heading = torch.stack([torch.cos(psi), torch.sin(psi)], 1)
        self.current_state = {"position": torch.stack([e.field("USV_S_X"), e.field("USV_S_Y")], 1), "orientation": self.heading,
                              "linear_velocity": torch.stack([e.field("USV_S_VX"), e.field("USV_S_VY")], 1),
                              "angular_velocity": e.field("USV_S_R")}

    def post_physics_step(self):
        """ONE fused launch: reset-if-flagged, action path, sub-steps, observation, reward + penalties, kills, progress
        [ref: OIGE/tasks/base/rl_task.py:283-303]."""
        self.engine.curriculum_step = self.step
        self.engine.step(self.actions)
        self.step += 1 / self.cfg.horizon_length                    # [ref: SNAP/USV_Virtual.py:838]
        self._calls += 1
        if self._nan_probe and self._calls % self._nan_probe_interval == 0:
            self.engine.check_finite()                              # sticky device flag: read every N steps, never lost
        return self.obs_buf, self.rew_buf, self.reset_buf, self.extras

    def get_observations(self) -> Dict[str, torch.Tensor]:
        return self.obs_buf

    def calculate_metrics(self) -> None:
        return

    def is_done(self) -> None:
        return

    def get_states(self):
        return self.states_buf

    def get_extras(self):
        return self.extras

    def reset_idx(self, env_ids: torch.Tensor) -> None:
        """Flags `env_ids`; the kernel performs reset_idx for them at the start of the next step, in the reference's position
        (pre_physics_step)  [ref: SNAP/USV_Virtual.py:750-817]."""
        self.reset_buf[env_ids] = 1
        if self._live:
            self.engine.mark_host_reset()

    # ---- extras["episode"]: mean over the envs being reset of episode_sums / maxEpisodeLength -----------
    def _fill_episode_extras(self) -> None:
        st = self.engine.bstats if self._live else self.engine.stats
        if st is None:
            return
        prefix = "USV_BST_" if self._live else "USV_ST_"
        n = self._num_envs
        mask = torch.zeros(self.engine.stride, dtype=torch.float32, device=self._device)
        mask[:n] = (self.reset_buf != 0).float()
        cnt = mask.sum()
        tot = (st * mask.view(-1, 1, 32)).sum(dim=(0, 2))                        # (USV_ST_COUNT,)
        new = tot / cnt.clamp(min=1.0) / self._max_episode_length
        # values live in persistent 0-dim tensors (updated in place): a step without resets keeps the previous value, and a
        # captured rollout graph keeps reading / writing the same storage across replays
        ep = self.extras.get("episode")
        if ep is None:
            ep = self.extras["episode"] = {}
        has = cnt > 0

        def put(key, value):
            if key not in ep:
                ep[key] = torch.zeros((), dtype=torch.float32, device=self._device)
            ep[key].copy_(torch.where(has, value, ep[key]))

        for k in self._stat_names():
            put(k, new[E[prefix + k.upper()]])                                    # [ref: SNAP/USV_Virtual.py:810-817]
        if self._live:
            # episode outcome events: plain means over the envs being reset, read BEFORE the reset clears the latches
            # [ref: OIGE/tasks/USV_Virtual.py:1508-1516,1581-1596]; "g_safe_mean" is declared but never fed (:497 commented out)
            oc = self.engine.bstate[:, E["USV_BS_OUTCOME"], :].reshape(-1).view(torch.int32)
            for k, bit in (("success", 0), ("collision", 1)):
                put(k, (((oc >> bit) & 1).float() * mask).sum() / cnt.clamp(min=1.0))
            if "g_safe_mean" not in ep:
                ep["g_safe_mean"] = torch.zeros((), dtype=torch.float32, device=self._device)
