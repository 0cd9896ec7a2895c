#!/usr/bin/env python
"""Multi-GPU check of the NVLink peer-memory all-reduce (run under torchrun, one rank per GPU, wrapped in `timeout`):
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_peer_allreduce.py
(1) values vs NCCL, bit-identity across ranks, eager and inside a replayed CUDA graph; (1b) the minibatch step with the exchange
fused into its cooperative tail kernel vs minibatch_grad + NCCL all-reduce + Adam; (2) PPO epochs with the update phase captured
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

# ---- (1b) the fused minibatch step: exchange inside the cooperative tail kernel ----------------------------------------------------
from omniisaacgymenvs_loop_b200.rl.peer import PeerStepExchange
from omniisaacgymenvs_loop_b200.rl.policy import PolicyMLP

for Dobs in (13, 33):
    M = 8192
    ex = PeerStepExchange(Dobs, dev, rank, world)
    a = PolicyMLP(Dobs, dev, seed=5, tensor_cores=True, world_size=world)
    b = PolicyMLP(Dobs, dev, seed=5, tensor_cores=True, world_size=world)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    rn = lambda *shape: torch.randn(*shape, device=dev, generator=gen)
    obs = rn(M, Dobs) * 2
    inf = a.act(obs)
    act, nlp = (inf["actions"] + 0.2 * rn(M, 2)).contiguous(), (inf["neglogpacs"] + 0.1 * rn(M)).contiguous()
    adv, old_v, ret = rn(M), rn(M) * 0.3, rn(M) * 0.5
    mu_a, sg_a, mu_b, sg_b = inf["mus"].clone(), inf["sigmas"].clone(), inf["mus"].clone(), inf["sigmas"].clone()
    for it in range(6):
        a.minibatch_grad(obs, act, nlp, adv, old_v, ret, mu_a, sg_a)
        if world > 1:
            dist.all_reduce(a.grads)
        a.optimizer_step()
        b.minibatch_step(obs, act, nlp, adv, old_v, ret, mu_b, sg_b, peer=ex)
        torch.cuda.synchronize()
        ex.check()
        err = float(((a.params - b.params).abs() / (a.params.abs() + 1e-3)).max())
        # first step: same parameters on both paths, only summation orders differ (Adam's m / sqrt(v) turns 1e-7 of a tiny gradient into ~1e-5 of a tiny parameter); later steps: last-bit parameter differences
        # straddle TF32 rounding boundaries of the operands (2^-11 each), and Adam's m / sqrt(v) is sign-like for small gradients
        assert err < (5e-5 if it == 0 else 2e-3), f"fused step differs from grad + NCCL + Adam: {err} (D={Dobs}, it={it})"
        assert abs(float(a.lr) - float(b.lr)) < 1e-12 and int(a.step) == int(b.step)
        if world > 1:
            ps = [torch.empty_like(b.params) for _ in range(world)]
            dist.all_gather(ps, b.params)
            assert all(torch.equal(ps[0], t) for t in ps), "fused step: ranks disagree bitwise"
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(8):
            b.minibatch_step(obs, act, nlp, adv, old_v, ret, mu_b, sg_b, peer=ex)
    graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ex.check()
    t_fused = e0.elapsed_time(e1) / 80 * 1e3
    single = PolicyMLP(Dobs, dev, seed=5, tensor_cores=True)
    graph1 = torch.cuda.CUDAGraph()
    single.minibatch_step(obs, act, nlp, adv, old_v, ret, mu_a, sg_a)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph1):
        for _ in range(8):
            single.minibatch_step(obs, act, nlp, adv, old_v, ret, mu_a, sg_a)
    graph1.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        graph1.replay()
    e1.record()
    torch.cuda.synchronize()
    t_single = e0.elapsed_time(e1) / 80 * 1e3
    if world > 1:
        ps = [torch.empty_like(b.params) for _ in range(world)]
        dist.all_gather(ps, b.params)
        assert all(torch.equal(ps[0], t) for t in ps), "fused step (graph): ranks disagree bitwise"
    log(f"[fused] world={world} D={Dobs} M={M}: minibatch step {t_fused:.1f} us with the exchange fused, {t_single:.1f} us on one rank "
        f"(exchange cost {t_fused - t_single:+.1f} us)")
    ex.close()

# ---- (2) PPO with the update phase in one graph per rank ----------------------------------------------------------------
envs = int(os.environ.get("PEER_TEST_ENVS", 16384))
for coll in ("peer", "peer-unfused", "nccl"):
    env = make_env(UsvEnvConfig(num_envs=envs).to_task_cfg(), str(dev), seed=5, env_id_offset=rank * envs, collect_stats=False)
    agent = A2CAgent(env, PPOConfig(seed=5), str(dev), rank, world, collective=coll.split("-")[0])
    agent.fused_step = coll != "peer-unfused"
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
    agent.check_peers()
    assert agent.ranks_identical()
    p = agent.policy.params.clone()
    if world > 1:
        ps = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(ps, p)
        assert all(torch.equal(ps[0], t) for t in ps), f"{coll}: parameters diverged across ranks"
    assert torch.isfinite(p).all()
    log(f"[ppo] collective={coll} world={world} envs/gpu={envs}: {ms:.2f} ms/epoch = {world * envs * 16 / ms * 1e3:.3e} frames/s "
        f"(graph={'yes' if agent._graph is not None else 'no'}, launches/update={agent.graph_launches['update']})")
    del agent, env
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
log("peer all-reduce OK")
