#!/usr/bin/env python
"""Batched evaluation of a loopz actor with the reference's per-episode CSV [ref: OIGE/scripts/rlgames_play_loopz.py:860-1439]:
deterministic actions `tanh(mu) * action_scale` (:1219-1222), every env of the fused live task evaluated in parallel, one CSV row per
finished episode (columns and definitions: `utils/episode_metrics.py`), bootstrap summary at the end.
  python scripts/play_loopz.py --checkpoint full_100.pt --num-envs 1024 --episodes 2000 --csv runs/play.csv"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg, load_task_yaml
import numpy as np

from omniisaacgymenvs_loop_b200.utils.episode_metrics import EpisodeRecorder, apply_mass_mode_to_obs
from scripts.train_loopz import build_learner, make_env


def play(env, actor, episodes: int, reward_scale: float = 0.01, max_steps: int = 1_000_000, mass_mode: str = "normal", **meta) -> EpisodeRecorder:
    """Runs until `episodes` episodes have finished (all envs in parallel; the last step may overshoot); returns the recorder.
    `mass_mode`: the control-side intervention on the mass / CoM columns the policy sees (LOOPZ_PLAY_MASS_MODE, :1212-1218)."""
    rng = np.random.default_rng(int(meta.get("seed", 0)))
    rec = EpisodeRecorder(env._task.engine, reward_scale=reward_scale, action_scale=float(actor.distribution.action_scale.reshape(-1)[0]), **meta)
    # reset() flags every env and runs the zero-action reset step (VecEnvRLGames.reset): feed it to the recorder as the start snapshot
    env.reset()
    n = env.num_envs
    zero = torch.zeros((n, env.num_acts), device=actor.device)
    rec.record(zero, torch.zeros(n, device=actor.device), env._task.reset_buf)
    scale = actor.distribution.action_scale.to(actor.device)
    for _ in range(max_steps):
        if len(rec.rows) >= episodes:
            break
        obs = apply_mass_mode_to_obs(env.observe(as_numpy=False), mass_mode, rng)
        action = torch.tanh(actor.noiseless_action(obs)) * scale
        reward, dones = env.step(action)
        rec.record(action, reward, dones)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task-yaml", default=None)
    ap.add_argument("--num-envs", type=int, default=1024)
    ap.add_argument("--episodes", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--csv", default=None)
    ap.add_argument("--mass-mode", default=os.getenv("LOOPZ_PLAY_MASS_MODE", "normal"), help="normal | zero | shuffle | swap (policy-side only)")
    ap.add_argument("--obs-source", default="sim", choices=["sim", "base", "both"],
                    help="privileged-tail source (mass.masscom_obs_source); both = the reference's two-mode comparison (eval.modes)")
    args = ap.parse_args()
    device = "cuda:0"
    task_cfg = load_task_yaml(args.task_yaml, num_envs=args.num_envs) if args.task_yaml else live_task_cfg(live_default_config(num_envs=args.num_envs))
    env = make_env(task_cfg, device, args.seed)
    ppo = build_learner(env, device, 16, args.seed, use_cuda_graph=False)
    if args.checkpoint:
        ppo.load_state_dict(torch.load(args.checkpoint, map_location=device, weights_only=False))
    rows = []
    for mode in (["sim", "base"] if args.obs_source == "both" else [args.obs_source]):
        env._task._masscom_obs_source = mode                    # rlgames_play_loopz.py:1100-1110 (_set_obs_source)
        rec = play(env, ppo.actor, args.episodes, run_id="play", ckpt=str(args.checkpoint), seed=args.seed, obs_source=mode, mass_mode=args.mass_mode)
        print(f"[loopz-play][EVAL] mode={mode}")
        rec.summarize(seed=args.seed)
        rows += rec.rows
    if args.csv:
        os.makedirs(os.path.dirname(os.path.abspath(args.csv)), exist_ok=True)
        rec.rows = rows
        rec.write_csv(args.csv)
        print(f"[loopz-play][EVAL] wrote {len(rows)} rows to {args.csv}")

if __name__ == "__main__":
    main()
