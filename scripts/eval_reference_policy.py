#!/usr/bin/env python
"""Closed-loop check of the env against the reference's own TRAINED classic-CaptureXY policy
(tests/golden/classic_policy.npz = the `model` dict of 811*/last_USV_ep_5450_rew_38.54975.pth, 5450 epochs in Isaac Sim/PhysX).
If the fused env reproduces the task the policy was trained on, the deterministic policy (a = mu) must drive most boats into
the 0.1 m / 0.05 m/s capture condition well before the 3000-step time-out."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
from omniisaacgymenvs_loop_b200.engine import FusedUsvEnv
from omniisaacgymenvs_loop_b200.rl.policy import PolicyMLP

def run(n=4096, steps=1500, izz=10.0, tensor_cores=False, device="cuda:0"):
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "classic_policy.npz")))
    cfg = UsvEnvConfig(num_envs=n, izz=izz)
    env = FusedUsvEnv(cfg, n, device)
    pol = PolicyMLP(13, device, tensor_cores=tensor_cores)
    pol.load_state_dict({k: torch.as_tensor(v) for k, v in G.items() if k not in ("epoch", "frame", "last_mean_rewards")})
    obs, _, _ = env.step(torch.zeros((n, 2), device=device))
    ep_ret = torch.zeros(n, device=device); ep_len = torch.zeros(n, device=device)
    fin_ret, fin_len, captured, killed, timeouts = [], [], 0, 0, 0
    for t in range(steps):
        out = pol.act(obs)
        obs, rew, done = env.step(torch.clamp(out["mus"], -1, 1))
        ep_ret += rew; ep_len += 1
        d = done.bool()
        if d.any():
            dist = obs[d, 5]
            captured += int((dist < 0.1).sum()); killed += int((dist > 19.9).sum()); timeouts += int((ep_len[d] >= cfg.max_episode_length - 1).sum())
            fin_ret.append(ep_ret[d].clone()); fin_len.append(ep_len[d].clone())
            ep_ret[d] = 0; ep_len[d] = 0
    fr = torch.cat(fin_ret) if fin_ret else torch.zeros(1); fl = torch.cat(fin_len) if fin_len else torch.zeros(1)
    return dict(episodes=int(fr.numel()), captured=captured, killed=killed, timeouts=timeouts, mean_return=float(fr.mean()), mean_len=float(fl.mean()))

if __name__ == "__main__":
    for izz in (2.0, 5.0, 10.0, 20.0):
        print("izz", izz, run(izz=izz))
    print("tensor cores, izz 10:", run(tensor_cores=True))
