"""Per-episode evaluation records in the reference's CSV format, for ALL envs of a fused env at once.

The reference's play script follows env 0 of a vec-env through one episode at a time, reading the task's tensors back to the host on
every control step, and writes one CSV row per episode [ref: OIGE/scripts/rlgames_play_loopz.py:970-1031 (columns), :1123-1405 (episode
loop), :174-302 (_compute_episode_conditions), :305-433 (_compute_step_metrics), :493-501 (_infer_done_reason), :149-171
(_hash_obstacles_xy), :504-533 (_bootstrap_mean_ci), :1046-1098 (_summarize_eval)].  Here the same per-episode quantities accumulate
on the device for every env (a handful of elementwise torch ops per control step: this is the evaluation path, not the hot path) and
only the rows of the episodes that finished on a step are read back.  Column names, definitions and the done-reason priority
(collision > out_of_bounds > goal_tolerance > other, all from the TERMINAL state) are the reference's.

Episode boundaries of the fused env: the kernel resets a flagged env at the START of the next control step and runs that step with a
zero action (USV_Virtual.py:1064-1101) -- the reference's `env.reset()` (flag + one zero-action step) -- so the state after that step
is the episode's start snapshot and its reward is not part of the episode, exactly as in the reference's loop.
"""
from __future__ import annotations

import csv
import hashlib
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

FIELDNAMES = [
    "run_id", "ckpt", "seed", "obs_source", "episode_idx",
    "start_x", "start_y", "start_yaw", "start_vx", "start_vy", "start_wz", "goal_x", "goal_y", "min_obs_dist_start",
    "sim_mass_raw", "sim_mass_rel", "sim_com_raw_x", "sim_com_raw_y", "sim_com_raw_z", "sim_com_scaled_x", "sim_com_scaled_y",
    "sim_com_scaled_z", "mass_obs_mode", "com_obs_mode", "com_scale_x", "com_scale_y", "com_scale_z",
    "thruster_mul", "thruster_left_mul", "thruster_right_mul", "k_drag", "k_Iz",
    "obstacles_count", "obstacles_limbo_count", "obstacles_hash_quant_m", "obstacles_hash",
    "success", "done_reason", "episode_len_steps", "time_to_goal_sec", "return_raw", "return_scaled", "path_length",
    "straight_line_dist", "path_efficiency", "action_smoothness_mean", "action_smoothness_sum", "action_saturation_rate",
    "collision", "out_of_bounds", "control_dt", "reward_scale",
]
SUMMARY_METRICS = ["success", "time_to_goal_sec", "path_efficiency", "action_smoothness_mean", "action_saturation_rate", "return_raw"]
DONE_REASONS = ("collision", "out_of_bounds", "goal_tolerance", "other")


def quantize_xy(xy: np.ndarray, quant_m: float) -> np.ndarray:
    """int32 grid coordinates in units of `quant_m` (round half to even, like np.rint)  [ref: rlgames_play_loopz.py:135-146]."""
    q = float(quant_m)
    if not np.isfinite(q) or q <= 0.0:
        q = 0.01
    return np.rint(np.asarray(xy, dtype=np.float64) / q).astype(np.int32, copy=False)


def hash_obstacles_xy(xy: np.ndarray, quant_m: float = 0.01) -> str:
    """SHA-1 of the quantised obstacle centres sorted lexicographically by (x, y): the same digest as the reference's for the same
    layout, whatever the obstacle order  [ref: rlgames_play_loopz.py:149-171]."""
    xy_q = quantize_xy(xy, quant_m)
    if xy_q.ndim != 2 or xy_q.shape[1] != 2:
        return ""
    order = np.lexsort((xy_q[:, 1], xy_q[:, 0]))
    h = hashlib.sha1()
    h.update(str(float(quant_m)).encode("utf-8"))
    h.update(b"|")
    h.update(np.ascontiguousarray(xy_q[order], dtype=np.int32).tobytes(order="C"))
    return h.hexdigest()


def bootstrap_mean_ci(values, num_boot: int = 2000, alpha: float = 0.05, rng: Optional[np.random.Generator] = None) -> Tuple[float, float, float]:
    """(mean, lo, hi) percentile-bootstrap interval of the mean over the finite values  [ref: rlgames_play_loopz.py:504-533]."""
    rng = rng if rng is not None else np.random.default_rng(0)
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    v = v[np.isfinite(v)]
    if v.size == 0:
        return float("nan"), float("nan"), float("nan")
    mean = float(np.mean(v))
    if v.size == 1:
        return mean, mean, mean
    idx = rng.integers(0, v.size, size=(int(num_boot), v.size))
    boot = np.mean(v[idx], axis=1)
    return mean, float(np.quantile(boot, alpha / 2.0)), float(np.quantile(boot, 1.0 - alpha / 2.0))


def infer_done_reason(collision: bool, out_of_bounds: bool, in_goal_tolerance: bool) -> str:
    """[ref: rlgames_play_loopz.py:493-501]"""
    if collision:
        return "collision"
    if out_of_bounds:
        return "out_of_bounds"
    if in_goal_tolerance:
        return "goal_tolerance"
    return "other"


def apply_mass_mode_to_obs(obs: torch.Tensor, mode: str, rng: np.random.Generator) -> torch.Tensor:
    """The play script's control-side intervention on the LAST 4 observation columns (`LOOPZ_PLAY_MASS_MODE` with
    `LOOPZ_PLAY_MASS_APPLY=control`): normal -> unchanged, zero -> 0, shuffle -> a fresh permutation across envs every step (drawn from
    the same numpy generator call as the reference, so the same seed gives the same permutation), swap -> env pairs (0<->1, 2<->3, ...)
    exchanged; shuffle / swap with one env fall back to zero, unknown modes to normal  [ref: rlgames_play_loopz.py:560-621].
    Works on the device tensor the policy reads (no host round trip); returns a new tensor unless the mode is normal."""
    mode_l = (mode or "normal").strip().lower()
    if obs.dim() != 2 or obs.shape[0] < 1 or mode_l in ("normal", "none", "off", ""):
        return obs
    n = int(obs.shape[0])
    if mode_l in ("shuffle", "swap") and n < 2:
        mode_l = "zero"
    if mode_l not in ("zero", "shuffle", "swap"):
        return obs
    out = obs.clone()
    if mode_l == "zero":
        out[:, -4:] = 0.0
    elif mode_l == "shuffle":
        perm = torch.as_tensor(rng.permutation(n), dtype=torch.long, device=obs.device)
        out[:, -4:] = obs[:, -4:][perm]
    else:
        idx = torch.arange(n, device=obs.device)
        pair = idx ^ 1
        pair = torch.where(pair < n, pair, idx)          # an odd env count leaves the last env alone
        out[:, -4:] = obs[:, -4:][pair]
    return out


class EpisodeRecorder:
    """Call `record(actions, rewards, dones)` after every `engine.step(actions)`; finished episodes append to `rows`.

    `engine`: FusedUsvLiveEnv (obstacle columns filled) or FusedUsvEnv (obstacle columns empty).  `action_scale`: the actor's
    `distribution.action_scale` (saturation = |a| > 0.95 scale).  `reward_scale`: the learner's reward scale (`return_scaled`)."""

    def __init__(self, engine, reward_scale: float = 1.0, action_scale: float = 1.0, run_id: str = "", ckpt: str = "", seed: int = 0,
                 obs_source: str = "sim", quant_m: float = 0.01):
        self.eng = engine
        self.n, self.dev = int(engine.num_envs), engine.device
        cfg = engine.cfg
        self.control_dt = float(cfg.dt) * int(cfg.n_substeps)
        self.reward_scale, self.action_scale = float(reward_scale), float(action_scale)
        self.meta = {"run_id": run_id, "ckpt": ckpt, "seed": int(seed), "obs_source": obs_source}
        self.quant_m = float(quant_m)
        self.live = getattr(engine, "live", None)
        self.has_obstacles = self.live is not None and int(getattr(engine, "task", 0)) == 0
        self.kill_dist, self.pos_tol = float(cfg.kill_dist), float(cfg.position_tolerance)
        self.collision_threshold = float(self.live.collision_threshold) if self.live is not None else float("nan")
        n, f64 = self.n, dict(dtype=torch.float64, device=self.dev)
        self.starting = torch.ones(n, dtype=torch.bool, device=self.dev)     # every env starts flagged (reset_buf = 1)
        self.start = torch.zeros((n, 7), **f64)                              # x, y, yaw, vx, vy, wz, min_obs_dist
        self.acc = torch.zeros((n, 7), **f64)                                # return, len, path, dsum, dcount, sat, sat_total
        self.prev_pos = torch.zeros((n, 2), **f64)
        self.prev_act = torch.zeros((n, 2), **f64)
        self.has_prev_act = torch.zeros(n, dtype=torch.bool, device=self.dev)
        self.rows: List[Dict] = []
        self.episode_idx = 0

    # ---- state read-back (device) ---------------------------------------------------------------------
    def _f(self, name: str) -> torch.Tensor:
        return self.eng.field(name).double()

    def _min_obs_dist(self, pos: torch.Tensor) -> torch.Tensor:
        if not self.has_obstacles:
            return torch.full((self.n,), float("nan"), dtype=torch.float64, device=self.dev)
        d = self.eng.obstacles.double() - pos[:, None, :]
        return torch.sqrt((d * d).sum(-1)).min(dim=1).values

    def record(self, actions: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor) -> int:
        """Accumulates one control step; returns the number of episodes that finished on it."""
        pos = torch.stack([self._f("USV_S_X"), self._f("USV_S_Y")], dim=1)
        st, go = self.starting, ~self.starting
        mind = self._min_obs_dist(pos)
        # envs whose reset ran inside this step: snapshot the start conditions, clear the accumulators
        snap = torch.stack([pos[:, 0], pos[:, 1], self._f("USV_S_PSI"), self._f("USV_S_VX"), self._f("USV_S_VY"), self._f("USV_S_R"), mind], 1)
        self.start = torch.where(st[:, None], snap, self.start)
        a = actions.to(self.dev).double().reshape(self.n, -1)
        step_len = torch.sqrt(((pos - self.prev_pos) ** 2).sum(-1))
        da = torch.sqrt(((a - self.prev_act) ** 2).sum(-1))
        hp = self.has_prev_act.double()
        inc = torch.stack([rewards.to(self.dev).double().reshape(-1), torch.ones_like(step_len), step_len, da * hp, hp,
                           (a.abs() > 0.95 * self.action_scale).double().sum(-1), torch.full_like(step_len, float(a.shape[1]))], 1)
        self.acc = torch.where(go[:, None], self.acc + inc, torch.zeros_like(self.acc))
        self.prev_pos = pos
        self.prev_act = a
        self.has_prev_act = go
        done = dones.to(self.dev).reshape(-1) != 0
        fin = (done & go).nonzero().reshape(-1)           # the one host read per step (the reference reads ~20 scalars per step)
        if fin.numel():
            self._emit(fin, pos, mind)
        self.starting = done
        return int(fin.numel())

    def _emit(self, idx: torch.Tensor, pos: torch.Tensor, mind: torch.Tensor) -> None:
        E = self._f
        goal = torch.stack([E("USV_C_TX"), E("USV_C_TY")], 1)
        cols = [self.start[idx], self.acc[idx], pos[idx], goal[idx], mind[idx, None],
                torch.stack([E("USV_C_MASS"), E("USV_C_THR_ML"), E("USV_C_THR_MR"), E("USV_C_KDRAG"), E("USV_C_KIZ")], 1)[idx]]
        if self.live is not None:
            cols.append(torch.stack([E("USV_BC_COM_X"), E("USV_BC_COM_Y"), E("USV_BC_COM_Z")], 1)[idx])
        M = torch.cat(cols, dim=1).cpu().numpy()
        obst = self.eng.obstacles[idx].cpu().numpy() if self.has_obstacles else None
        cfg, live = self.eng.cfg, self.live
        for k in range(M.shape[0]):
            r = M[k]
            sx, sy, syaw, svx, svy, swz, smin = r[0:7]
            ret, length, path, dsum, dcount, sat, sat_total = r[7:14]
            px, py, gx, gy, min_end = r[14:19]
            mass, thr_l, thr_r, kdrag, kiz = r[19:24]
            dist = math.sqrt((gx - px) ** 2 + (gy - py) ** 2)
            collision = bool(self.has_obstacles and min_end < self.collision_threshold)
            oob = dist > self.kill_dist
            reason = infer_done_reason(collision, oob, dist < self.pos_tol)
            success = int(reason == "goal_tolerance")
            straight = math.sqrt((gx - sx) ** 2 + (gy - sy) ** 2)
            self.episode_idx += 1
            row = dict(self.meta)
            row.update({
                "episode_idx": self.episode_idx, "start_x": sx, "start_y": sy, "start_yaw": syaw, "start_vx": svx, "start_vy": svy,
                "start_wz": swz, "goal_x": gx, "goal_y": gy, "min_obs_dist_start": smin, "sim_mass_raw": mass,
                "thruster_mul": thr_l, "thruster_left_mul": thr_l, "thruster_right_mul": thr_r, "k_drag": kdrag, "k_Iz": kiz,
                "success": success, "done_reason": reason, "episode_len_steps": int(length),
                "time_to_goal_sec": length * self.control_dt if success else float("nan"),
                "return_raw": ret, "return_scaled": ret * self.reward_scale, "path_length": path, "straight_line_dist": straight,
                "path_efficiency": straight / max(path, 1e-6),
                "action_smoothness_mean": dsum / dcount if dcount > 0 else float("nan"), "action_smoothness_sum": dsum,
                "action_saturation_rate": sat / sat_total if sat_total > 0 else float("nan"),
                "collision": int(collision), "out_of_bounds": int(oob), "control_dt": self.control_dt, "reward_scale": self.reward_scale})
            if live is not None:                      # MassDistributionDisturbances.get_masses  [ref: USV_disturbances.py:153-194]
                com, scale = r[24:27], np.asarray(live.com_scale, dtype=np.float64)
                rel = bool(live.mass_obs_relative)
                enc = com / (scale + 1e-6) if live.com_obs_scaled else com
                row.update({"sim_mass_rel": (mass - cfg.mass_base) / cfg.mass_base if rel else mass,
                            "sim_com_raw_x": com[0], "sim_com_raw_y": com[1], "sim_com_raw_z": com[2],
                            "sim_com_scaled_x": enc[0], "sim_com_scaled_y": enc[1], "sim_com_scaled_z": enc[2],
                            "mass_obs_mode": "relative" if rel else "raw", "com_obs_mode": "scaled" if live.com_obs_scaled else "raw",
                            "com_scale_x": scale[0], "com_scale_y": scale[1], "com_scale_z": scale[2]})
            if obst is not None:
                xy = obst[k]
                row.update({"obstacles_count": int(xy.shape[0]),
                            "obstacles_limbo_count": int(np.sum(np.all(np.isclose(xy, np.array([[999.0, 999.0]], dtype=np.float32)), axis=1))),
                            "obstacles_hash_quant_m": self.quant_m, "obstacles_hash": hash_obstacles_xy(xy, self.quant_m)})
            self.rows.append(row)

    # ---- output ------------------------------------------------------------------------------------------
    def write_csv(self, path: str) -> None:
        with open(path, "w", newline="") as fp:
            w = csv.DictWriter(fp, fieldnames=FIELDNAMES)
            w.writeheader()
            for row in self.rows:
                w.writerow(row)

    def summarize(self, seed: int = 0, log=print) -> Dict[str, Tuple[float, float, float]]:
        """mean / std / bootstrap 95 % interval of the reference's summary metrics  [ref: rlgames_play_loopz.py:1046-1073]."""
        return summarize_rows(self.rows, seed=seed, log=log)


def summarize_rows(rows: List[Dict], seed: int = 0, log=print) -> Dict[str, Tuple[float, float, float]]:
    out = {}
    if not rows:
        return out
    rng = np.random.default_rng(int(seed))
    if log:
        log(f"[loopz-play][EVAL] summary: episodes={len(rows)}")
    for k in SUMMARY_METRICS:
        arr = np.asarray([float(r.get(k, float("nan"))) for r in rows], dtype=np.float64)
        mean, lo, hi = bootstrap_mean_ci(arr, rng=rng)
        finite = arr[np.isfinite(arr)]
        sd = float(np.std(finite)) if finite.size > 0 else float("nan")
        out[k] = (mean, lo, hi)
        if log:
            log(f"  {k}: mean={mean:.6g} std={sd:.6g} 95%CI=[{lo:.6g}, {hi:.6g}]")
    return out
