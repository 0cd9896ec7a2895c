#!/usr/bin/env python
"""PPO epoch time on the LIVE task (33-dim obs, obstacles): rollout vs update, at --envs envs."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
from scripts.train_usv import make_env
ap = argparse.ArgumentParser(); ap.add_argument("--envs", type=int, default=16384); ap.add_argument("--epochs", type=int, default=10)
a = ap.parse_args()
dev = "cuda:0"
env = make_env(live_task_cfg(live_default_config(num_envs=a.envs)), dev, seed=5, collect_stats=False)
env.env._task._nan_probe = False
agent = A2CAgent(env, PPOConfig(seed=5), dev)
for _ in range(14):            # past the first 200-step episodes: steady-state reset rate
    agent.train_epoch()
torch.cuda.synchronize()
play = upd = 0.0
t0 = time.perf_counter()
for _ in range(a.epochs):
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with torch.no_grad():
        agent._play()                          # rollout graph replay when captured (live kernels take a device-side step offset)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    agent._graph.replay() if agent._graph is not None else agent.update()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    play += t2 - t1; upd += t3 - t2
tot = time.perf_counter() - t0
print(f"live PPO envs={a.envs} obs_dim={agent.obs_dim} tensor_cores={agent.policy.tensor_cores} rollout_graph={agent._graph_play is not None}: {tot/a.epochs*1e3:.2f} ms/epoch "
      f"(rollout {play/a.epochs*1e3:.2f}, update {upd/a.epochs*1e3:.2f}) = {a.envs*16/(tot/a.epochs):.3e} frames/s; reward {agent.episode_stats()}")
