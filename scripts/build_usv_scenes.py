#!/usr/bin/env python
"""Builds a scene file for NPZ scene replay [ref: OIGE/scripts/build_usv_scenes.py:520-740]: the reference resets a vec-env N times and
snapshots env 0 after each reset; here ONE reset of an N-env fused live env yields the N scenes (obstacles, start pose / velocity, goal),
written in the same npz schema with the same file name pattern and `.sha1` side-car.
  python scripts/build_usv_scenes.py --num-episodes 1000 --seed 0 --out-dir runs/scenes [--task-yaml cfg.yaml]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg, load_task_yaml
from omniisaacgymenvs_loop_b200.scene_replay import save_scenes, snapshot_scenes
from scripts.train_loopz import make_env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task-yaml", default=None)
    ap.add_argument("--num-episodes", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out-dir", default="runs/scenes")
    ap.add_argument("--task-name", default="USV_Virtual_CaptureXY")
    args = ap.parse_args()
    n = args.num_episodes
    task_cfg = load_task_yaml(args.task_yaml, num_envs=n) if args.task_yaml else live_task_cfg(live_default_config(num_envs=n))
    env = make_env(task_cfg, "cuda:0", args.seed)
    env.reset()                                                     # flag every env + one zero-action step (VecEnvRLGames.reset)
    eng = env._task.engine
    scenes = snapshot_scenes(eng, seed=args.seed)
    path = save_scenes(args.out_dir, scenes, task_name=args.task_name,
                       generator_cfg={"goal_random_position": float(eng.cfg.goal_random_position), "envs": n})
    print(f"[build_scenes] saved scenes to: {path}")
    print(f"[build_scenes] checksum written to: {path}.sha1")


if __name__ == "__main__":
    main()
