#!/usr/bin/env python
"""Tensor-core loopz minibatch gradient: a few steps at 16384 envs x 16 (65 536-row minibatches) for the ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.test_gpu_loopz import _fill_random, build

_, _, ppo = build(n=16384, horizon=16, tensor_cores=True)
_fill_random(ppo, torch.Generator().manual_seed(0), 16384, 16, 33)
for _ in range(3):
    ppo._minibatch(0, 65536)
torch.cuda.synchronize()
print("ok", ppo.minibatch_statistics())
