#!/usr/bin/env python
"""Thin launcher with the semantics of the reference's rl_games entry point [ref: OIGE/scripts/rlgames_train111.py:113-172]:
builds VecEnvRLGames + USVVirtual from a task YAML (or the built-in classic CaptureXY config), wraps it as RLGPUEnv and runs
the PPO loop.  Single GPU:  python scripts/train_usv.py --num-envs 4096 --epochs 200
Multi GPU:   torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_usv.py --num-envs 16384"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig, live_default_config, live_task_cfg, load_task_yaml
from omniisaacgymenvs_loop_b200.envs.vec_env_rlgames import VecEnvRLGames
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
from omniisaacgymenvs_loop_b200.tasks.USV_Virtual import SimConfig, USVVirtual
from omniisaacgymenvs_loop_b200.utils.rlgames.rlgames_utils import RLGPUEnv


def make_env(task_cfg: dict, device: str, seed: int, env_id_offset: int = 0, collect_stats: bool = True):
    env = VecEnvRLGames(headless=True)
    sim = SimConfig({"sim_device": device, "rl_device": device, "seed": seed, "env_id_offset": env_id_offset, "task": task_cfg})
    task = USVVirtual("USVVirtual", sim, env, collect_stats=collect_stats)
    env.set_task(task, backend="torch")
    return RLGPUEnv(env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task-yaml", default=None, help="a reference task YAML (cfg/task/USV/...); default: built-in classic CaptureXY")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--full-dr", action="store_true")
    ap.add_argument("--live", action="store_true", help="built-in LIVE CaptureXY (16 static obstacles, 33-dim obs) instead of the classic task")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--save", default=None)
    ap.add_argument("--play", action="store_true", help="the reference's test=True path: evaluate --checkpoint with PpoPlayerContinuous")
    ap.add_argument("--games", type=int, default=2000, help="--play: episodes to finish (player.games_num)")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    device = f"cuda:{local}"
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    if args.task_yaml:
        task_cfg = load_task_yaml(args.task_yaml, num_envs=args.num_envs)
    elif args.live:
        task_cfg = live_task_cfg(live_default_config(num_envs=args.num_envs))
    else:
        cfg = UsvEnvConfig(num_envs=args.num_envs)
        task_cfg = (cfg.full_dr() if args.full_dr else cfg).to_task_cfg()
    env = make_env(task_cfg, device, args.seed, env_id_offset=rank * args.num_envs)
    if args.play:                                   # [ref: OIGE/scripts/rlgames_train111.py test=True -> runner.run({'play': True})]
        from omniisaacgymenvs_loop_b200.rl.players import PpoPlayerContinuous
        player = PpoPlayerContinuous(env, {"games_num": args.games, "deterministic": True}, device, seed=args.seed)
        if args.checkpoint:
            player.restore(args.checkpoint)
        player.run()
        return
    agent = A2CAgent(env, PPOConfig(seed=args.seed), device, rank, world)
    if args.checkpoint:
        agent.restore(args.checkpoint)
    agent.train(args.epochs)
    if args.save and rank == 0:
        agent.save(args.save)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
