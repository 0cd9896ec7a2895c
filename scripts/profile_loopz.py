#!/usr/bin/env python
"""Short loopz-learner run for ncu / timing: one rollout on the live task and a few minibatch steps at --envs envs."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg
from scripts.train_loopz import build_learner, make_env, train

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--updates", type=int, default=1)
ap.add_argument("--time", action="store_true", help="device-time the kernels with CUDA events (no profiler)")
args = ap.parse_args()
dev = "cuda:0"
torch.manual_seed(0)
env = make_env(live_task_cfg(live_default_config(num_envs=args.envs)), dev, seed=0)
ppo = build_learner(env, dev, 16, seed=0, use_cuda_graph=False)
train(env, ppo, args.updates, 16, log_every=0, quiet=True)
torch.cuda.synchronize()
if args.time:
    obs = env.observe(as_numpy=False)
    ppo.storage.step = 16
    M = args.envs * 16 // 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, n=10):
        fn()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    print(f"act (actor sample, {args.envs} rows)      {timeit(lambda: ppo.actor.sample(obs)):9.1f} us")
    print(f"act (critic predict)                 {timeit(lambda: ppo.critic.predict(obs)):9.1f} us")
    print(f"returns + advantage standardisation  {timeit(lambda: ppo.storage.compute_returns(ppo.critic.predict(obs), 0.997, 0.95)):9.1f} us")
    print(f"minibatch step ({M} rows)          {timeit(lambda: ppo._minibatch(0, M)):9.1f} us")
print("ok", ppo.last_stats)
