#!/usr/bin/env python
"""BASELINE config C5: env-count sweep 1 k - 1 M envs/GPU of the fused step (Variant A classic full-DR, Variant B live) at 1 / 2 / 4 / 8
GPUs (torchrun: one rank per GPU, envs sharded, no collective on the data path; device-timed, max over ranks), with the CPU oracle
port of the reference's torch path at the same env counts beside it.  One JSON line per point, a markdown table at the end (stderr).
    python scripts/sweep_envs.py AB cpu > gpurun_out/sweep_1gpu.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/sweep_envs.py A"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig, UsvLiveConfig, live_default_config
from omniisaacgymenvs_loop_b200.engine import FusedUsvEnv, FusedUsvLiveEnv

import time

import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
PEAK = 6544.7
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAK = float(json.load(f).get("hbm_gbs", PEAK))
except Exception:
    pass
ONLY = sys.argv[1] if len(sys.argv) > 1 else "AB"
WITH_CPU = "cpu" in sys.argv[2:]          # the CPU oracle port at every env count (rank 0; the reference arm of config C5)


def timed(fn, steps, warm):
    """Device time per step: CUDA events on the launching stream, barrier on both sides, max over ranks."""
    for w in range(warm):
        fn(w)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def cpu_rate(n):
    """env-steps/s of the oracle port (the reference's torch algorithm) on the host cores at n envs; bounded: ~1 s per point."""
    from oracle import usv_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    env = O.ClassicEnvOracle(O.EnvConfig().full_dr(), n)
    act = torch.zeros((n, 2))
    env.step(act)
    t0, k = time.perf_counter(), 0
    while k < 3 or (time.perf_counter() - t0 < 1.0 and k < 200):
        env.step(act)
        k += 1
    return n * k / (time.perf_counter() - t0)


rows = []
for lg in (range(10, 21) if "A" in ONLY else []):
    n = 1 << lg
    env = FusedUsvEnv(UsvEnvConfig().full_dr(), n, dev, env_id_offset=rank * n)
    g = torch.Generator(device=dev).manual_seed(lg + 100 * rank)
    acts = [torch.rand((n, 2), device=dev, generator=g) * 2 - 1 for _ in range(4)]
    steps = 2000 if lg < 18 else 500
    ms = timed(lambda k: env.step(acts[k & 3]), steps, 50)
    r = {"variant": "A full-DR", "n_gpus": world, "envs_per_gpu": n, "us_per_step": ms * 1e3, "env_steps_per_s": world * n / (ms * 1e-3),
         "gbps_per_gpu": 268 * n / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 268 * n / (ms * 1e-3) / 1e9 / PEAK}
    if WITH_CPU and rank == 0:
        r["cpu_port_env_steps_per_s"] = cpu_rate(n)
        r["cpu_threads"] = os.cpu_count()
    if world > 1:
        dist.barrier()
    rows.append(r)
    if rank == 0:
        sys.stdout.write(json.dumps(r) + "\n"); sys.stdout.flush()
    del env
for lg in (range(10, 19) if "B" in ONLY else []):
    n = 1 << lg
    env = FusedUsvLiveEnv(live_default_config(num_envs=n), UsvLiveConfig(), n, dev, env_id_offset=rank * n)
    g = torch.Generator(device=dev).manual_seed(lg + 100 * rank)
    acts = [torch.rand((n, 2), device=dev, generator=g) * 2 - 1 for _ in range(4)]
    full = timed(lambda k: env.step(acts[k & 3]), 300, 210)
    ms = timed(lambda k: env.step(acts[k & 3], rebuild_scene=False), 500, 20)
    r = {"variant": "B live", "n_gpus": world, "envs_per_gpu": n, "us_per_step": ms * 1e3, "env_steps_per_s": world * n / (ms * 1e-3),
         "gbps_per_gpu": 610 * n / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 610 * n / (ms * 1e-3) / 1e9 / PEAK,
         "steady_state_us_per_step": full * 1e3, "steady_state_env_steps_per_s": world * n / (full * 1e-3)}
    rows.append(r)
    if rank == 0:
        sys.stdout.write(json.dumps(r) + "\n"); sys.stdout.flush()
    del env
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank != 0:
    sys.exit(0)
print("\n| variant | GPUs | envs/GPU | us/step | env-steps/s (all GPUs) | algorithmic GB/s per GPU | frac of HBM peak | steady-state env-steps/s (B: with scene rebuilds) | CPU port env-steps/s |", file=sys.stderr)
print("|---|---|---|---|---|---|---|---|---|", file=sys.stderr)
for r in rows:
    print(f"| {r['variant']} | {r['n_gpus']} | {r['envs_per_gpu']} | {r['us_per_step']:.1f} | {r['env_steps_per_s']:.3g} | {r['gbps_per_gpu']:.0f} | "
          f"{r['frac_of_hbm_peak']:.3f} | {r.get('steady_state_env_steps_per_s', float('nan')):.3g} | {r.get('cpu_port_env_steps_per_s', float('nan')):.3g} |",
          file=sys.stderr)
