"""CPU suite, world_size 2 over gloo: the host-side logic of the N>1 path -- env sharding by rank (global env ids keep the
Philox streams, so a sharded job reproduces the single-rank job), the one-span gradient+statistics all-reduce followed by the
1/world scaling, the start-up parameter broadcast, and the episode-statistics reduction.  (The CUDA kernels themselves are
exercised by the -m gpu suite; this checks what surrounds them.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_oracle as P
from oracle import usv_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        n = 48                                               # envs per rank
        cfg = O.EnvConfig(max_episode_length=6).full_dr()
        # --- env sharding: rank r owns global env ids [r*n, (r+1)*n)
        env = O.ClassicEnvOracle(cfg, n, env_id_offset=rank * n)
        g = torch.Generator().manual_seed(5)
        outs = []
        for _ in range(8):
            act = torch.rand((world * n, 2), generator=g) * 2 - 1   # same global action tensor on every rank
            obs, rew, done = env.step(act[rank * n:(rank + 1) * n])
            outs.append(torch.cat([obs, rew[:, None], done[:, None].float()], 1))
        mine = torch.stack(outs)                                   # (8, n, 15)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        # --- gradient span: [grads | stats] summed over ranks, then scaled by 1/world (a2c_common.py:311-323)
        span = torch.arange(10, dtype=torch.float32) * (rank + 1)
        dist.all_reduce(span, op=dist.ReduceOp.SUM)
        span = span / world
        # --- parameter broadcast at start-up (a2c_common.py:1350-1355)
        params = torch.full((5,), float(rank + 7))
        dist.broadcast(params, 0)
        # --- episode statistics: (sum of returns, sum of lengths, count) reduced once per epoch
        acc = torch.tensor([10.0 * (rank + 1), 100.0 * (rank + 1), 2.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        # --- PRODUCT host logic of the N > 1 path (rl/a2c.py), on CPU tensors over gloo: the cross-rank identity check that bench.py /
        #     train() run after the fused gradient exchange, and the episode-statistics reduction
        from omniisaacgymenvs_loop_b200.rl.a2c import ranks_hold_identical, reduce_episode_stats, state_signature
        same = [torch.arange(100, dtype=torch.float32) * 0.37, torch.full((7,), 3, dtype=torch.int32)]
        ident = ranks_hold_identical(same, world)
        off_by_one_ulp = [same[0].clone(), same[1]]
        if rank == 1:
            off_by_one_ulp[0][41] = torch.nextafter(off_by_one_ulp[0][41], torch.tensor(1e9))
        differ = ranks_hold_identical(off_by_one_ulp, world)
        swapped = [same[0].clone(), same[1]]
        if rank == 1:
            swapped[0][[3, 4]] = swapped[0][[4, 3]]                 # same multiset of words, another order: the weighted sum catches it
        differ2 = ranks_hold_identical(swapped, world)
        stats = reduce_episode_stats(torch.tensor([10.0 * (rank + 1), 100.0 * (rank + 1), 2.0 * (rank + 1)], dtype=torch.float64), world)
        empty = reduce_episode_stats(torch.zeros(3, dtype=torch.float64), world)
        assert state_signature(same).dtype == torch.int64
        if rank == 0:
            q.put((torch.cat(gathered, 1).numpy(), span.numpy(), params.numpy(), acc.numpy(), (ident, differ, differ2, stats, empty)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    world, n = 2, 48
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    sharded, span, params, acc, product = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank run of the same 96 envs
    cfg = O.EnvConfig(max_episode_length=6).full_dr()
    env = O.ClassicEnvOracle(cfg, world * n)
    g = torch.Generator().manual_seed(5)
    outs = []
    for _ in range(8):
        act = torch.rand((world * n, 2), generator=g) * 2 - 1
        obs, rew, done = env.step(act)
        outs.append(torch.cat([obs, rew[:, None], done[:, None].float()], 1))
    single = torch.stack(outs).numpy()
    assert np.array_equal(sharded, single)                       # bit-identical: sharding does not change any env's stream
    assert np.allclose(span, np.arange(10) * 1.5)                # (1x + 2x)/2
    assert np.all(params == 7.0)
    assert np.allclose(acc, [30.0, 300.0, 6.0])
    ident, differ, differ2, stats, empty = product
    assert ident is True and differ is False and differ2 is False            # one ulp on one rank, or two swapped words, is a divergence
    assert stats == (5.0, 50.0, 6) and empty[2] == 0 and np.isnan(empty[0])


def test_averaged_gradient_step_equals_big_batch_step():
    """Two ranks each holding half of a minibatch: mean of the per-rank mean-gradients == the full-batch gradient, so the
    all-reduce(SUM)/world + Adam step is the single-GPU step (what the adam kernel's inv_world implements)."""
    torch.manual_seed(0)
    D, M = 13, 64
    L = P.param_layout(D)
    params = torch.randn(L["P"]) * 0.05
    rms = P.RunningMeanStd((D,))
    obs = torch.randn((M, D))
    inf = P.policy_inference(params, obs, D, rms, P.RunningMeanStd((1,)), eps=torch.randn((M, 2)))
    batch = dict(obs=obs, actions=inf["actions"], old_logp_actions=inf["neglogpacs"] + 0.1 * torch.randn(M), advantages=torch.randn(M),
                 old_values=torch.randn((M, 1)) * 0.3, returns=torch.randn((M, 1)) * 0.5, mu=inf["mus"], sigma=inf["sigmas"])

    def grad(sl):
        p = params.clone().requires_grad_(True)
        loss, _ = P.minibatch_loss(p, {k: v[sl] for k, v in batch.items()}, D, rms)
        loss.backward()
        return p.grad

    full = grad(slice(0, M))
    halves = (grad(slice(0, M // 2)) + grad(slice(M // 2, M))) / 2
    assert torch.allclose(full, halves, rtol=1e-4, atol=1e-7)
