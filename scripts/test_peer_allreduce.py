#!/usr/bin/env python
"""Multi-GPU check of the NVLink peer-memory all-reduce (run under torchrun, one rank per GPU, wrapped in `timeout`):
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/test_peer_allreduce.py
(1) values vs NCCL, bit-identity across ranks, eager and inside a replayed CUDA graph; (2) PPO epochs with the update phase captured
in one graph per rank: parameters stay identical on all ranks; prints ms/epoch for collective = peer vs nccl."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
from omniisaacgymenvs_loop_b200.rl.peer import PeerAllReduce
from scripts.train_usv import make_env

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
log = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)

# ---- (1) the collective itself ------------------------------------------------------------------------------------------
n = 18701
ar = PeerAllReduce(n, dev, rank, world)
g = torch.Generator(device=dev).manual_seed(100 + rank)
for it in range(20):
    x = torch.randn(n, device=dev, generator=g)
    want = x.clone()
    if world > 1:
        dist.all_reduce(want)
    got = ar(x.clone())
    torch.cuda.synchronize()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (it, float((got - want).abs().max()))
    if world > 1:
        all_got = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(all_got, got)
        assert all(torch.equal(all_got[0], t) for t in all_got), "ranks disagree bitwise"
ar.check()
x = torch.randn(n, device=dev, generator=g)
y = torch.empty_like(x)
for _ in range(3):
    ar(x, y)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(10):
        ar(x, y)
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
ar.check()
want = x.clone()
if world > 1:
    dist.all_reduce(want)
assert torch.allclose(y, want, rtol=1e-5, atol=1e-5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    graph.replay()
e1.record()
torch.cuda.synchronize()
t_peer = e0.elapsed_time(e1) / 200 * 1e3
e0.record()
for _ in range(200):
    if world > 1:
        dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
t_nccl = e0.elapsed_time(e1) / 200 * 1e3
log(f"[peer] world={world} all-reduce of {n} floats: peer kernel {t_peer:.1f} us (graph replay), nccl {t_nccl:.1f} us (eager)")
ar.close()

# ---- (2) PPO with the update phase in one graph per rank ----------------------------------------------------------------
envs = int(os.environ.get("PEER_TEST_ENVS", 16384))
for coll in ("peer", "nccl"):
    env = make_env(UsvEnvConfig(num_envs=envs).to_task_cfg(), str(dev), seed=5, env_id_offset=rank * envs, collect_stats=False)
    agent = A2CAgent(env, PPOConfig(seed=5), str(dev), rank, world, collective=coll)
    for _ in range(4):
        agent.train_epoch()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        agent.train_epoch()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = (time.perf_counter() - t0) / 8 * 1e3
    if agent.peer is not None:
        agent.peer.check()
    p = agent.policy.params.clone()
    if world > 1:
        ps = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(ps, p)
        assert all(torch.equal(ps[0], t) for t in ps), f"{coll}: parameters diverged across ranks"
    assert torch.isfinite(p).all()
    log(f"[ppo] collective={coll} world={world} envs/gpu={envs}: {ms:.2f} ms/epoch = {world * envs * 16 / ms * 1e3:.3e} frames/s "
        f"(graph={'yes' if agent._graph is not None else 'no'})")
    del agent, env
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
log("peer all-reduce OK")
