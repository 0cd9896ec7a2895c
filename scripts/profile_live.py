#!/usr/bin/env python
"""Short Variant-B run for ncu: a few control steps of the live task (scene rebuild + fused live step) at --envs envs."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import UsvLiveConfig, live_default_config
from omniisaacgymenvs_loop_b200.engine import FusedUsvLiveEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--episode", type=int, default=200)
args = ap.parse_args()
dev = "cuda:0"
env = FusedUsvLiveEnv(live_default_config(num_envs=args.envs, max_episode_length=args.episode), UsvLiveConfig(), args.envs, dev)
g = torch.Generator(device=dev).manual_seed(0)
for k in range(args.steps):
    env.step(torch.rand((args.envs, 2), device=dev, generator=g) * 2 - 1)
torch.cuda.synchronize()
env.check_finite()
print("ok", float(env.rew.mean()))
