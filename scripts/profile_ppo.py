#!/usr/bin/env python
"""Two PPO epochs at BASELINE config[2] size (16384 envs) for ncu launch lists / kernel captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
from scripts.train_usv import make_env

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
env = make_env(UsvEnvConfig(num_envs=n).to_task_cfg(), "cuda:0", seed=1, collect_stats=False)
env.env._task._nan_probe = False
agent = A2CAgent(env, PPOConfig(seed=1), "cuda:0")
for _ in range(epochs):
    agent.train_epoch()
torch.cuda.synchronize()
print("ok", agent.policy.stats())
