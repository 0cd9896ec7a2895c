#!/usr/bin/env python
"""Device-timed latency / throughput of the live task's scene builder (usv_live_build_fields_f32: cost-to-go relaxation, J, potential field)
on dense batches of random scenes.  Small batches (<= 148) read as the latency of one scene, large ones as us per scene of throughput.
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
args = ap.parse_args()
dev = "cuda:0"
env = FusedUsvLiveEnv(live_default_config(num_envs=8192), UsvLiveConfig(), 8192, dev)
g = torch.Generator(device=dev).manual_seed(0)
out = {"lib": os.environ.get("USV_B200_LIB", "default")}
for m in [int(b) for b in args.batches.split(",")]:
    # obstacles / targets drawn like the placement rules' ranges (annulus around the origin, targets inside the map)
    r = 2.0 + 10.0 * torch.rand((m, 16), device=dev, generator=g)
    th = 6.2831853 * torch.rand((m, 16), device=dev, generator=g)
    obst = torch.stack([r * torch.cos(th), r * torch.sin(th)], -1)
    tgt = (torch.rand((m, 2), device=dev, generator=g) * 2 - 1) * 10.0
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
print(json.dumps(out))
