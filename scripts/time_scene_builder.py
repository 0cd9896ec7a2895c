#!/usr/bin/env python
"""Device-timed latency / throughput of the live task's scene builder (usv_live_build_fields_f32: cost-to-go relaxation, J, potential field)
on dense batches of scenes placed by the task itself.  Small batches (<= 148) read as the latency of one scene, large ones as us per scene of throughput.
USV_B200_LIB=<other .so> times another build of the library on the same box (A/B)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import UsvLiveConfig, live_default_config
from omniisaacgymenvs_loop_b200.engine import FusedUsvLiveEnv

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=str, default="1,37,148,153,296,592,1184,4736")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--steady", type=int, default=0, help="also time N steady-state control steps (scene rebuild + live step) at --envs envs")
ap.add_argument("--envs", type=int, default=16384)
args = ap.parse_args()
dev = "cuda:0"
env = FusedUsvLiveEnv(live_default_config(num_envs=8192), UsvLiveConfig(), 8192, dev)
g = torch.Generator(device=dev).manual_seed(0)
out = {"lib": os.environ.get("USV_B200_LIB", "default")}
# scenes as the task places them: one control step resets every env, which draws obstacles around each env's spawn / target pair
env.step(torch.zeros((8192, 2), device=dev))
all_obst = env.obstacles
all_tgt = torch.stack([env.field("USV_C_TX"), env.field("USV_C_TY")], 1)
for m in [int(b) for b in args.batches.split(",")]:
    obst, tgt = all_obst[:m].contiguous(), all_tgt[:m].contiguous()
    for _ in range(3):
        f = env.build_fields(obst, tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        f = env.build_fields(obst, tgt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    out[str(m)] = {"us": round(ms * 1e3, 1), "us_per_scene": round(ms * 1e3 / m, 3), "checksum": float(f.double().sum())}
if args.steady:
    n = args.envs
    env2 = FusedUsvLiveEnv(live_default_config(num_envs=n), UsvLiveConfig(), n, dev)
    acts = [torch.rand((n, 2), device=dev, generator=g) * 2 - 1 for _ in range(8)]
    for w in range(230):                      # past the first episodes: resets spread over the control steps
        env2.step(acts[w % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steady):
        env2.step(acts[k % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steady
    out["steady"] = {"envs": n, "us_per_step": round(ms * 1e3, 1), "env_steps_per_s": round(n / (ms * 1e-3), 0),
                     "reset_fraction": float(env2.reset_buf.float().mean())}
print(json.dumps(out))
