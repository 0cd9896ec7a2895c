#!/usr/bin/env python
"""BASELINE config C5 on one GPU: env-count sweep of the fused step (Variant A classic full-DR, Variant B live) with device
timing; prints one JSON line per point and a markdown table at the end.  python scripts/sweep_envs.py > gpurun_out/sweep.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig, UsvLiveConfig, live_default_config
from omniisaacgymenvs_loop_b200.engine import FusedUsvEnv, FusedUsvLiveEnv

dev = torch.device("cuda:0")
PEAK = 6544.7
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAK = float(json.load(f).get("hbm_gbs", PEAK))
except Exception:
    pass


def timed(fn, steps, warm):
    for w in range(warm):
        fn(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


rows = []
ONLY = sys.argv[1] if len(sys.argv) > 1 else "AB"
for lg in (range(10, 21) if "A" in ONLY else []):
    n = 1 << lg
    env = FusedUsvEnv(UsvEnvConfig().full_dr(), n, dev)
    g = torch.Generator(device=dev).manual_seed(lg)
    acts = [torch.rand((n, 2), device=dev, generator=g) * 2 - 1 for _ in range(4)]
    steps = 2000 if lg < 18 else 500
    ms = timed(lambda k: env.step(acts[k & 3]), steps, 50)
    r = {"variant": "A full-DR", "envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3), "gbps": 268 * n / (ms * 1e-3) / 1e9,
         "frac_of_hbm_peak": 268 * n / (ms * 1e-3) / 1e9 / PEAK}
    rows.append(r)
    print(json.dumps(r), flush=True)
    del env
for lg in (range(10, 19) if "B" in ONLY else []):
    n = 1 << lg
    env = FusedUsvLiveEnv(live_default_config(num_envs=n), UsvLiveConfig(), n, dev)
    g = torch.Generator(device=dev).manual_seed(lg)
    acts = [torch.rand((n, 2), device=dev, generator=g) * 2 - 1 for _ in range(4)]
    full = timed(lambda k: env.step(acts[k & 3]), 300, 210)
    ms = timed(lambda k: env.step(acts[k & 3], rebuild_scene=False), 500, 20)
    r = {"variant": "B live", "envs": n, "us_per_step": ms * 1e3, "env_steps_per_s": n / (ms * 1e-3), "gbps": 610 * n / (ms * 1e-3) / 1e9,
         "frac_of_hbm_peak": 610 * n / (ms * 1e-3) / 1e9 / PEAK, "steady_state_us_per_step": full * 1e3,
         "steady_state_env_steps_per_s": n / (full * 1e-3)}
    rows.append(r)
    print(json.dumps(r), flush=True)
    del env
    torch.cuda.empty_cache()
print("\n| variant | envs | us/step | env-steps/s | algorithmic GB/s | frac of HBM peak | steady-state env-steps/s (B: with scene rebuilds) |", file=sys.stderr)
print("|---|---|---|---|---|---|---|", file=sys.stderr)
for r in rows:
    print(f"| {r['variant']} | {r['envs']} | {r['us_per_step']:.1f} | {r['env_steps_per_s']:.3g} | {r['gbps']:.0f} | {r['frac_of_hbm_peak']:.3f} | "
          f"{r.get('steady_state_env_steps_per_s', float('nan')):.3g} |", file=sys.stderr)
