#!/usr/bin/env python
"""Device-timed live step kernel alone (no scene rebuild: the fields of resetting envs go stale, which does not change the kernel's work) at
several env counts; USV_B200_LIB=<other .so> times another build on the same box (A/B)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import UsvLiveConfig, live_default_config
from omniisaacgymenvs_loop_b200.engine import FusedUsvLiveEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=str, default="16384,65536,131072,262144")
ap.add_argument("--steps", type=int, default=300)
args = ap.parse_args()
dev = "cuda:0"
out = {"lib": os.environ.get("USV_B200_LIB", "default")}
for n in [int(x) for x in args.envs.split(",")]:
    env = FusedUsvLiveEnv(live_default_config(num_envs=n), UsvLiveConfig(), n, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    acts = [torch.rand((n, 2), device=dev, generator=g) * 2 - 1 for _ in range(4)]
    env.step(acts[0])                              # builds every scene once
    for w in range(20):
        env.step(acts[w % 4], rebuild_scene=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        env.step(acts[k % 4], rebuild_scene=False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / args.steps * 1e3
    out[str(n)] = {"us": round(us, 2), "gbps": round(610 * n / us / 1e3, 1), "obs_sum": float(env.obs.double().sum())}
    del env
    torch.cuda.empty_cache()
print(json.dumps(out))
