#!/usr/bin/env python
"""GAE, policy forward and one PPO minibatch step at bench sizes, for ncu captures of the secondary kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omniisaacgymenvs_loop_b200.rl.a2c import gae
from omniisaacgymenvs_loop_b200.rl.policy import PolicyMLP
dev = "cuda:0"
T, n = 16, 1 << 22
g = torch.Generator(device=dev).manual_seed(0)
rew, val = torch.randn((T, n), device=dev, generator=g), torch.randn((T, n), device=dev, generator=g)
dones = (torch.rand((T, n), device=dev, generator=g) < 0.1).to(torch.uint8)
lv, ld = torch.randn(n, device=dev, generator=g), torch.zeros(n, dtype=torch.uint8, device=dev)
adv, ret = torch.empty_like(rew), torch.empty_like(rew)
for _ in range(3):
    gae(rew, val, dones, lv, ld, 0.99, 0.95, adv, ret)
M = 1 << 20
pol = PolicyMLP(13, dev)
obs = torch.randn((M, 13), device=dev, generator=g)
o = pol.act(obs)
for _ in range(2):
    pol.act(obs, o)
mb = 8192
s = slice(0, mb)
act, nlp = o["actions"][s].contiguous(), o["neglogpacs"][s].contiguous()
advm, ov, rt = torch.randn(mb, device=dev), torch.randn(mb, device=dev) * 0.3, torch.randn(mb, device=dev) * 0.5
mu, sg = o["mus"][s].clone(), o["sigmas"][s].clone()
for _ in range(3):
    pol.minibatch_grad(obs[s], act, nlp, advm, ov, rt, mu, sg)
    pol.optimizer_step()
torch.cuda.synchronize()
print("ok")
