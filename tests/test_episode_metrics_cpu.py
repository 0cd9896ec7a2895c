"""Evaluation records (utils/episode_metrics.py) against the reference play script's helpers (goldens from
oracle/make_golden_play.py, which executes the reference's own function definitions) and, for the batched recorder, against a
per-env scalar loop written the way the reference script follows env 0 [ref: OIGE/scripts/rlgames_play_loopz.py:1123-1405]."""
import csv
import json
import math
import os
from types import SimpleNamespace

import numpy as np
import torch

from omniisaacgymenvs_loop_b200.utils import episode_metrics as EM

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "play_metrics.json")))


def test_obstacle_hash_matches_reference():
    for lay in G["layouts"]:
        xy = np.asarray(lay["xy"], dtype=np.float32)
        assert EM.quantize_xy(xy, 0.01).tolist() == lay["quant"]
        assert EM.hash_obstacles_xy(xy, 0.01) == lay["hash"] == lay["hash_shuffled"]
        assert EM.hash_obstacles_xy(xy[::-1], 0.01) == lay["hash"]                # order-independent
        assert EM.hash_obstacles_xy(xy, 0.05) == lay["hash_q05"] != lay["hash"]


def test_bootstrap_ci_matches_reference():
    for b in G["bootstrap"]:
        v = np.asarray([np.nan if x is None else x for x in b["values"]])
        assert list(EM.bootstrap_mean_ci(v, rng=np.random.default_rng(3))) == b["ci"]      # same generator, same draws: bit-equal
    assert all(math.isnan(x) for x in EM.bootstrap_mean_ci(np.array([np.nan, np.nan])))


def test_done_reason_priority_matches_reference():
    for r in G["reasons"]:
        assert EM.infer_done_reason(r["collision"] > 0.5, r["out_of_bounds"] > 0.5, r["in_goal_tolerance"] > 0.5) == r["reason"]


def test_mass_mode_intervention_matches_reference():
    for c in G["mass_modes"]:
        obs = torch.tensor(c["obs"], dtype=torch.float32)
        keep = obs.clone()
        got = EM.apply_mass_mode_to_obs(obs, c["mode"], np.random.default_rng(c["seed"]))
        assert torch.equal(got, torch.tensor(c["out"], dtype=torch.float32)), (c["mode"], obs.shape)
        assert torch.equal(obs, keep)                                            # the input is never modified


class FakeEngine:
    """Just the attributes EpisodeRecorder reads from FusedUsvLiveEnv, with scripted trajectories."""

    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self.num_envs, self.device, self.task = n, torch.device("cpu"), 0
        self.cfg = SimpleNamespace(dt=0.02, n_substeps=10, kill_dist=6.0, position_tolerance=0.5, mass_base=34.96)
        self.live = SimpleNamespace(collision_threshold=1.2, com_scale=(1.3, 1.0, 1.0), mass_obs_relative=True, com_obs_scaled=True)
        self.g = g
        self.f = {k: torch.zeros(n) for k in ("USV_S_X", "USV_S_Y", "USV_S_PSI", "USV_S_VX", "USV_S_VY", "USV_S_R", "USV_C_TX", "USV_C_TY",
                                              "USV_C_MASS", "USV_C_THR_ML", "USV_C_THR_MR", "USV_C_KDRAG", "USV_C_KIZ", "USV_BC_COM_X",
                                              "USV_BC_COM_Y", "USV_BC_COM_Z")}
        self.obstacles = torch.zeros((n, 16, 2))
        self.reset_buf = torch.ones(n, dtype=torch.int64)

    def field(self, name):
        return self.f[name].clone()

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.g)

    def step(self, actions):
        n, f = self.num_envs, self.f
        r = self.reset_buf.bool()
        for k in ("USV_S_X", "USV_S_Y", "USV_C_TX", "USV_C_TY"):                    # reset at the start of the step
            f[k] = torch.where(r, self._rand(n) * 8 - 4, f[k])
        for k, lo, hi in (("USV_C_MASS", 35, 55), ("USV_C_THR_ML", .5, 1), ("USV_C_THR_MR", .5, 1), ("USV_C_KDRAG", 1, 1.5), ("USV_C_KIZ", 1, 1.5),
                          ("USV_BC_COM_X", -.15, .15), ("USV_BC_COM_Y", -.05, .05), ("USV_BC_COM_Z", -.02, .02), ("USV_S_PSI", -3, 3)):
            f[k] = torch.where(r, self._rand(n) * (hi - lo) + lo, f[k])
        ob = self._rand(n, 16, 2) * 24 - 12
        ob[:, 13:] = 999.0
        self.obstacles = torch.where(r[:, None, None], ob, self.obstacles)
        a = torch.where(r[:, None], torch.zeros_like(actions), actions)              # zero action on the reset step
        f["USV_S_VX"], f["USV_S_VY"] = a[:, 0] * 1.5, a[:, 1] * 1.5
        f["USV_S_X"] = f["USV_S_X"] + 0.2 * f["USV_S_VX"]
        f["USV_S_Y"] = f["USV_S_Y"] + 0.2 * f["USV_S_VY"]
        rew = self._rand(n) - 0.5
        pos = torch.stack([f["USV_S_X"], f["USV_S_Y"]], 1)
        d = (pos - torch.stack([f["USV_C_TX"], f["USV_C_TY"]], 1)).norm(dim=1)
        mind = (self.obstacles - pos[:, None]).norm(dim=2).min(1).values
        self.reset_buf = ((d > 6.0) | (d < 0.5) | (mind < 1.2) | (self._rand(n) < 0.02)).long()
        return rew, self.reset_buf.clone()


def test_recorder_equals_per_env_scalar_loop(tmp_path):
    n, steps = 24, 400
    eng = FakeEngine(n, 5)
    rec = EM.EpisodeRecorder(eng, reward_scale=0.01, action_scale=1.0, run_id="r", ckpt="c.pt", seed=3)
    g = torch.Generator().manual_seed(9)
    # the scalar loop of the reference script, one copy per env
    ref_rows = [[] for _ in range(n)]
    S = [None] * n
    for t in range(steps):
        act = torch.tanh(torch.randn((n, 2), generator=g) * 1.5)
        starting = eng.reset_buf.bool().clone()
        rew, done = eng.step(act)
        rec.record(act, rew, done)
        for i in range(n):
            px, py = float(eng.f["USV_S_X"][i]), float(eng.f["USV_S_Y"][i])
            if starting[i]:
                S[i] = dict(start=(px, py), prev=(px, py), ret=0.0, steps=0, path=0.0, pa=None, dsum=0.0, dcnt=0, sat=0, tot=0)
                continue
            s = S[i]
            a0 = act[i].double().numpy()
            if s["pa"] is not None:
                s["dsum"] += float(np.linalg.norm(a0 - s["pa"])); s["dcnt"] += 1
            s["pa"] = a0.copy()
            s["tot"] += a0.size; s["sat"] += int(np.sum(np.abs(a0) > 0.95))
            s["ret"] += float(rew[i]); s["steps"] += 1
            s["path"] += math.sqrt((px - s["prev"][0]) ** 2 + (py - s["prev"][1]) ** 2); s["prev"] = (px, py)
            if done[i]:
                gx, gy = float(eng.f["USV_C_TX"][i]), float(eng.f["USV_C_TY"][i])
                dist = math.hypot(gx - px, gy - py)
                mind = float((eng.obstacles[i].double() - torch.tensor([px, py], dtype=torch.float64)).norm(dim=1).min())
                reason = EM.infer_done_reason(mind < 1.2, dist > 6.0, dist < 0.5)
                straight = math.hypot(gx - s["start"][0], gy - s["start"][1])
                ref_rows[i].append(dict(env=i, reason=reason, steps=s["steps"], ret=s["ret"], path=s["path"], straight=straight,
                                        smooth=s["dsum"] / s["dcnt"] if s["dcnt"] else float("nan"), sat=s["sat"] / s["tot"],
                                        sx=s["start"][0], hash=EM.hash_obstacles_xy(eng.obstacles[i].numpy(), 0.01), mass=float(eng.f["USV_C_MASS"][i])))
    flat = sorted((r for rows in ref_rows for r in rows), key=lambda r: (r["sx"], r["steps"]))
    got = sorted(rec.rows, key=lambda r: (r["start_x"], r["episode_len_steps"]))
    assert len(flat) == len(got) > 100
    assert {r["done_reason"] for r in got} >= {"collision", "out_of_bounds", "goal_tolerance", "other"}
    for a, b in zip(flat, got):
        assert b["done_reason"] == a["reason"] and b["episode_len_steps"] == a["steps"] and b["obstacles_hash"] == a["hash"]
        assert b["success"] == int(a["reason"] == "goal_tolerance") and b["collision"] == int(a["reason"] == "collision")
        for x, y in ((b["return_raw"], a["ret"]), (b["path_length"], a["path"]), (b["straight_line_dist"], a["straight"]),
                     (b["action_saturation_rate"], a["sat"]), (b["sim_mass_raw"], a["mass"]), (b["return_scaled"], a["ret"] * 0.01)):
            assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (x, y)
        assert (math.isnan(b["action_smoothness_mean"]) and math.isnan(a["smooth"])) or abs(b["action_smoothness_mean"] - a["smooth"]) < 1e-9
        assert b["obstacles_count"] == 16 and b["obstacles_limbo_count"] == 3
        assert abs(b["sim_mass_rel"] - (a["mass"] - 34.96) / 34.96) < 1e-9 and b["control_dt"] == 0.2
        assert (b["time_to_goal_sec"] == a["steps"] * 0.2) if b["success"] else math.isnan(b["time_to_goal_sec"])
    path = tmp_path / "ep.csv"
    rec.write_csv(str(path))
    rows = list(csv.DictReader(open(path)))
    assert list(rows[0].keys()) == EM.FIELDNAMES and len(rows) == len(got)
    lines = []
    summ = rec.summarize(log=lines.append)
    assert set(summ) == set(EM.SUMMARY_METRICS) and 0.0 <= summ["success"][0] <= 1.0 and len(lines) == 1 + len(EM.SUMMARY_METRICS)


def test_scene_file_writer_round_trip(tmp_path):
    """The scene builder's file format [ref: OIGE/scripts/build_usv_scenes.py:704-737]: keys, dtypes, name, .sha1 side-car; what
    SceneReplay reads back is what the engine held."""
    import json

    from omniisaacgymenvs_loop_b200 import scene_replay as SR
    eng = FakeEngine(10, 2)
    eng.step(torch.zeros((10, 2)))                                   # env.reset(): flag + one zero-action step
    sc = SR.snapshot_scenes(eng, seed=40)
    assert sc["obstacles_xy"].shape == (10, 16, 2) and sc["obstacles_xy"].dtype == np.float32 and sc["obstacles_count"].tolist() == [16] * 10
    assert np.array_equal(sc["start_pos"][:, 0], eng.f["USV_S_X"].numpy()) and np.array_equal(sc["goal_pos"][:, 1], eng.f["USV_C_TY"].numpy())
    assert np.array_equal(sc["obstacles_xy"], eng.obstacles.numpy()) and sc["seed"].tolist() == list(range(40, 50))
    path = SR.save_scenes(str(tmp_path), sc, task_name="USV Virtual/CaptureXY", generator_cfg={"goal_random_position": 4.0})
    assert os.path.basename(path) == "USV_Virtual_CaptureXY__scenes__N10__seed40.npz" and os.path.exists(path + ".sha1")
    back = SR.load_scenes(path, verify_sha1=True)
    assert set(SR.REQUIRED) | {"num_episodes", "episode_idx", "seed", "max_obstacles", "generator_cfg", "created_at"} <= set(back)
    for k in SR.REQUIRED:
        assert np.array_equal(back[k], sc[k]) and back[k].dtype == sc[k].dtype, k
    assert json.loads(str(back["generator_cfg"]))["goal_random_position"] == 4.0 and int(back["num_episodes"]) == 10
    with open(path, "ab") as fp:                                    # a corrupted file is caught by the side-car
        fp.write(b"x")
    import pytest
    with pytest.raises(ValueError):
        SR.load_scenes(path, verify_sha1=True)
    with pytest.raises(KeyError):
        SR.save_scenes(str(tmp_path), {"start_pos": sc["start_pos"]})


def test_scene_file_built_on_b200_feeds_scene_replay_host_logic():
    """tests/golden/scenes_b200_N64_seed3.npz was written by scripts/build_usv_scenes.py on a B200 (64 envs, one reset).  The file obeys
    the spawn rules of the live task [ref: USV_capture_xy_static_obs.py:936-1047] and SceneReplay's host side (per-env scene counters,
    16-slot normalisation) consumes it; the device side of the replay is tests/test_gpu_live.py::test_scene_replay_npz_vs_oracle."""
    from omniisaacgymenvs_loop_b200 import scene_replay as SR
    path = os.path.join(os.path.dirname(__file__), "golden", "scenes_b200_N64_seed3.npz")
    d = SR.load_scenes(path, verify_sha1=True)
    ob, sp = d["obstacles_xy"], d["start_pos"]
    assert ob.shape == (64, 16, 2) and d["obstacles_count"].tolist() == [16] * 64
    real = ob[..., 0] < 900
    dist = np.linalg.norm(ob - sp[:, None, :], axis=2)
    assert dist[real].min() > 3.0 - 0.35                                  # 3 m from the start, minus one zero-action step of drift
    assert np.linalg.norm(ob - d["goal_pos"][:, None, :], axis=2)[real].min() >= 3.0 - 1e-4          # 3 m from the target
    for i in range(64):                                                   # 2.5 m between any two placed obstacles
        p = ob[i][real[i]]
        dd = np.linalg.norm(p[:, None] - p[None], axis=2) + np.eye(len(p)) * 99
        assert dd.min() >= 2.5 - 1e-4
    r = np.linalg.norm(sp, axis=1)
    assert 9.0 - 0.35 <= r.min() and r.max() <= 12.0 + 0.35 and 0.0 <= d["start_yaw"].min() and d["start_yaw"].max() < np.pi + 0.1
    rp = object.__new__(SR.SceneReplay)                                   # host logic only: no engine behind it
    rp.data, rp.num_scenes, rp.cycle = {k: d[k] for k in SR.REQUIRED}, 64, True
    rp.next_scene_idx, rp.last_scene_idx = torch.zeros(8, dtype=torch.long), torch.full((8,), -1, dtype=torch.long)
    rp.next_scene_idx[3] = 63
    idx = rp.take_scene_indices(torch.tensor([1, 3]))
    assert idx.tolist() == [0, 63] and rp.take_scene_indices(torch.tensor([3])).tolist() == [0]      # cycles past the end
    pos, yaw, vel, goal, obst = rp.scenes(idx)
    assert obst.shape == (2, 16, 2) and torch.equal(obst, torch.from_numpy(ob[[0, 63]])) and torch.equal(pos, torch.from_numpy(sp[[0, 63]]))
    assert yaw.shape == (2,) and vel.shape == (2, 2) and goal.shape == (2, 2)
