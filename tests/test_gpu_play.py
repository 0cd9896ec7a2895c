"""GPU test of the evaluation path: scripts/play_loopz.py + utils/episode_metrics.EpisodeRecorder over the fused live env, checked
against a per-env scalar loop written the way the reference script follows env 0 [ref: OIGE/scripts/rlgames_play_loopz.py:1123-1405]."""
import csv
import dataclasses
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg  # noqa: E402
from omniisaacgymenvs_loop_b200.utils import episode_metrics as EM  # noqa: E402

DEV = "cuda:0"


def test_play_loopz_rows_match_scalar_loop(tmp_path):
    from scripts.play_loopz import play
    from scripts.train_loopz import build_learner, make_env
    n, watch = 256, 12
    cfg = dataclasses.replace(live_default_config(num_envs=n, seed=21), max_episode_length=80)
    env = make_env(live_task_cfg(cfg), DEV, 21)
    ppo = build_learner(env, DEV, 16, 3, use_cuda_graph=False)
    eng = env._task.engine
    # wrap the recorder entry point: the scalar trackers see exactly the (action, reward, done) triples the recorder sees
    S, ref_rows = [None] * watch, []
    orig = EM.EpisodeRecorder.record

    def spy(self, actions, rewards, dones):
        starting = self.starting[:watch].cpu().clone()
        out = orig(self, actions, rewards, dones)
        X, Y = eng.field("USV_S_X")[:watch].double().cpu(), eng.field("USV_S_Y")[:watch].double().cpu()
        TX, TY = eng.field("USV_C_TX")[:watch].double().cpu(), eng.field("USV_C_TY")[:watch].double().cpu()
        A, R, D = actions[:watch].double().cpu().numpy(), rewards[:watch].double().cpu(), dones[:watch].cpu()
        OB = eng.obstacles[:watch].double().cpu()
        for i in range(watch):
            px, py = float(X[i]), float(Y[i])
            if starting[i]:
                S[i] = dict(start=(px, py), prev=(px, py), ret=0.0, steps=0, path=0.0, pa=None, dsum=0.0, dcnt=0, sat=0, tot=0)
                continue
            s = S[i]
            if s["pa"] is not None:
                s["dsum"] += float(np.linalg.norm(A[i] - s["pa"])); s["dcnt"] += 1
            s["pa"] = A[i].copy()
            s["tot"] += A[i].size; s["sat"] += int(np.sum(np.abs(A[i]) > 0.95))
            s["ret"] += float(R[i]); s["steps"] += 1
            s["path"] += math.hypot(px - s["prev"][0], py - s["prev"][1]); s["prev"] = (px, py)
            if D[i]:
                gx, gy = float(TX[i]), float(TY[i])
                dist = math.hypot(gx - px, gy - py)
                mind = float((OB[i] - torch.tensor([px, py], dtype=torch.float64)).norm(dim=1).min())
                ref_rows.append(dict(reason=EM.infer_done_reason(mind < 1.2, dist > cfg.kill_dist, dist < cfg.position_tolerance),
                                     steps=s["steps"], ret=s["ret"], path=s["path"], sx=s["start"][0],
                                     straight=math.hypot(gx - s["start"][0], gy - s["start"][1]),
                                     smooth=s["dsum"] / s["dcnt"] if s["dcnt"] else float("nan"), sat=s["sat"] / s["tot"],
                                     hash=EM.hash_obstacles_xy(OB[i].numpy(), 0.01)))
        return out

    EM.EpisodeRecorder.record = spy
    try:
        rec = play(env, ppo.actor, episodes=600, reward_scale=0.01, run_id="t", ckpt="none", seed=21)
    finally:
        EM.EpisodeRecorder.record = orig
    rows = rec.rows
    assert len(rows) >= 600 and len(ref_rows) >= 10
    key = {(round(r["start_x"], 9), r["episode_len_steps"]): r for r in rows}
    for a in ref_rows:
        b = key[(round(a["sx"], 9), a["steps"])]
        assert b["done_reason"] == a["reason"] and b["obstacles_hash"] == a["hash"]
        for x, y in ((b["return_raw"], a["ret"]), (b["path_length"], a["path"]), (b["straight_line_dist"], a["straight"]),
                     (b["action_saturation_rate"], a["sat"])):
            assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (x, y)
        assert (math.isnan(b["action_smoothness_mean"]) and math.isnan(a["smooth"])) or abs(b["action_smoothness_mean"] - a["smooth"]) < 1e-9
    for r in rows:
        assert 1 <= r["episode_len_steps"] <= 80 and r["done_reason"] in EM.DONE_REASONS
        assert r["obstacles_count"] == 16 and 0 <= r["obstacles_limbo_count"] <= 16 and len(r["obstacles_hash"]) == 40
        assert 34.9 <= r["sim_mass_raw"] <= 55.0 and 0.49 <= r["thruster_mul"] <= 1.001 and 0.99 <= r["k_drag"] <= 1.51
        assert math.isfinite(r["return_raw"]) and r["path_length"] >= 0 and r["straight_line_dist"] >= 0
        assert r["min_obs_dist_start"] > 1.2 and r["control_dt"] == pytest.approx(0.2)
        assert r["collision"] == int(r["done_reason"] == "collision")
    # the play script's obs-source switch (_set_obs_source): in "base" mode every env sees the same (base / neutral) privileged tail
    task = env._task
    assert task._masscom_obs_source == "sim"
    tail = env.observe(as_numpy=False)[:, 25:33]
    assert float(tail.std(dim=0).max()) > 0.01
    task._masscom_obs_source = "base"
    env.step(torch.zeros((n, 2), device=DEV))
    tail = env.observe(as_numpy=False)[:, 25:33]
    assert torch.equal(tail, torch.zeros_like(tail)) and task._masscom_obs_source == "base"      # minmax + relative / scaled: all neutral = 0
    task._masscom_obs_source = "sim"
    with pytest.raises(ValueError):
        task._masscom_obs_source = "other"
    path = tmp_path / "play.csv"
    rec.write_csv(str(path))
    back = list(csv.DictReader(open(path)))
    assert list(back[0].keys()) == EM.FIELDNAMES and len(back) == len(rows)
    assert set(rec.summarize(log=None)) == set(EM.SUMMARY_METRICS)
