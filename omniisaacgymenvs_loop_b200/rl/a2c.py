"""A2CAgent: the rl_games continuous PPO loop [ref: RLG/common/a2c_common.py:65-1486, RLG/algos_torch/a2c_continuous.py]
re-hosted on the C-ABI kernels.  One process per GPU; envs are sharded across ranks; NCCL carries one all-reduce per
minibatch (gradient + KL + loss statistics in ONE span) and the episode statistics once per epoch.

Reference behaviours kept on purpose: env-major contiguous minibatches with no shuffling,
mu/sigma written back into the dataset after every minibatch, obs normaliser updated during mini-epoch 0 only, value
normaliser fed values then returns, rewards scaled by reward_shaper.scale_value at rollout time, critic weight
0.5*critic_coef, per-minibatch adaptive-KL learning rate ('legacy' schedule) -- here evaluated on device, so there is no
kl.item() host sync anywhere in the epoch.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from .. import _lib
from .policy import PolicyMLP

A = 2


@dataclass
class PPOConfig:
    """OIGE/cfg/train/USV/USV_PPOcontinuous_MLP.yaml `config:` section."""
    gamma: float = 0.99
    tau: float = 0.95
    learning_rate: float = 1e-4
    lr_schedule: str = "adaptive"
    kl_threshold: float = 0.016
    grad_norm: float = 1.0
    entropy_coef: float = 0.0
    truncate_grads: bool = True
    e_clip: float = 0.2
    horizon_length: int = 16
    minibatch_size: int = 8192
    mini_epochs: int = 8
    critic_coef: float = 0.5
    clip_value: bool = True
    bounds_loss_coef: float = 1e-4
    normalize_input: bool = True
    normalize_value: bool = True
    normalize_advantage: bool = True
    reward_scale: float = 0.01
    max_epochs: int = 3000
    games_to_track: int = 100
    seed: int = 42

    @classmethod
    def from_train_cfg(cls, cfg: dict) -> "PPOConfig":
        c = cfg["params"]["config"] if "params" in cfg else cfg
        kw = {k: c[k] for k in cls.__dataclass_fields__ if k in c and not isinstance(c[k], str) or k == "lr_schedule" and k in c}
        if "reward_shaper" in c:
            kw["reward_scale"] = c["reward_shaper"].get("scale_value", 1.0)
        return cls(**kw)


def gae(rewards, values, dones, last_values, last_dones, gamma, tau, adv, ret):
    """A2CBase.discount_values + returns  [ref: RLG/common/a2c_common.py:525-540,763] -- one kernel."""
    T, n = rewards.shape
    rc = _lib.lib().ppo_gae_f32(_lib.ptr(rewards), _lib.ptr(values), _lib.ptr(dones, torch.uint8), _lib.ptr(last_values),
                                _lib.ptr(last_dones, torch.uint8), ctypes.c_float(gamma), ctypes.c_float(tau), _lib.ptr(adv), _lib.ptr(ret),
                                ctypes.c_int32(T), ctypes.c_int64(n), _lib.stream())
    _lib.check(rc, "ppo_gae_f32")


class AverageMeter:
    """rl_games' windowed running mean [ref: RLG/algos_torch/torch_ext.py:281-307] with `current_size` kept on the device, so that
    update() needs no host sync (the reference slices the finished envs with `dones.nonzero()` first).  update(sum, count) takes
    the sum and the number of the values finishing at this step; the arithmetic is the reference's:
        size = clip(count, 0, max); old = min(max - size, current); mean = (mean*old + new_mean*size) / (old + size)."""

    def __init__(self, max_size: int, device):
        self.max_size = float(max_size)
        self.mean = torch.zeros((), dtype=torch.float32, device=device)
        self.current_size = torch.zeros((), dtype=torch.float32, device=device)

    def update(self, value_sum: torch.Tensor, count: torch.Tensor) -> None:
        has = count > 0
        new_mean = value_sum / count.clamp(min=1.0)
        size = count.clamp(0.0, self.max_size)
        old = torch.minimum(self.max_size - size, self.current_size)
        tot = old + size
        self.mean.copy_(torch.where(has, (self.mean * old + new_mean * size) / tot.clamp(min=1.0), self.mean))
        self.current_size.copy_(torch.where(has, tot, self.current_size))

    def clear(self) -> None:
        self.mean.zero_()
        self.current_size.zero_()

    def get_mean(self) -> float:
        return float(self.mean)

    def __len__(self) -> int:
        return int(self.current_size)


def state_signature(tensors) -> torch.Tensor:
    """Exact (bit-pattern) checksum of a list of fp32 / int32 tensors: [sum of words, position-weighted sum of words, word count] as int64.
    Two ranks hold bit-identical state iff their signatures are equal (up to the collision odds of two 64-bit sums)."""
    words = torch.cat([t.detach().reshape(-1).view(torch.int32) if t.dtype in (torch.float32, torch.int32) else t.detach().reshape(-1).to(torch.int32)
                       for t in tensors]).to(torch.int64)
    idx = torch.arange(1, words.numel() + 1, device=words.device, dtype=torch.int64)
    return torch.stack([words.sum(), (words * idx).sum(), torch.tensor(words.numel(), device=words.device, dtype=torch.int64)])


def ranks_hold_identical(tensors, world: int, group=None) -> bool:
    """All-gather of state_signature(): True when every rank of the group holds bit-identical copies (any backend: nccl on the GPUs, gloo in
    the CPU tests)."""
    if world <= 1:
        return True
    sig = state_signature(tensors)
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig, group=group)
    return all(torch.equal(s, sigs[0]) for s in sigs)


def reduce_episode_stats(acc: torch.Tensor, world: int, group=None):
    """(sum of returns, sum of lengths, count) of the finished episodes summed over ranks -> (mean return, mean length, count)
    [ref: the per-print reward / length statistics of RLG/common/a2c_common.py:1399-1418 at N ranks]."""
    acc = acc.clone()
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    s, l, c = acc.tolist()
    return (s / c, l / c, int(c)) if c > 0 else (float("nan"), float("nan"), 0)


class ScalarLog:
    """add_scalar(tag, value, step) into a JSON-lines file: the scalar stream the reference sends to tensorboardX's SummaryWriter
    [ref: RLG/common/a2c_common.py:343-362], for an image without tensorboard.  Any object with that method can be passed instead."""

    def __init__(self, path: str):
        import json
        import os
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        self._f, self._json = open(path, "a"), json

    def add_scalar(self, tag: str, value, step) -> None:
        self._f.write(self._json.dumps({"tag": tag, "value": float(value), "step": float(step)}) + "\n")

    def flush(self) -> None:
        self._f.flush()

    def close(self) -> None:
        self._f.close()


class A2CAgent:
    def __init__(self, vec_env, cfg: PPOConfig, device="cuda:0", rank: int = 0, world_size: int = 1, use_cuda_graph: bool = True,
                 collective: str = "peer"):
        """`vec_env` follows the rl_games IVecEnv contract (step/reset/get_env_info), e.g. RLGPUEnv(VecEnvRLGames)."""
        self.vec_env, self.cfg = vec_env, cfg
        self.device = torch.device(device)
        self.rank, self.world = rank, world_size
        self.multi_gpu = world_size > 1
        # gradient collective: "peer" = one-shot all-reduce over NVLink peer memory (rl/peer.py; stays inside the captured graph),
        # "nccl" = dist.all_reduce launched eagerly (NCCL inside a captured graph deadlocked on the 2-GPU box, r01 notes in DESIGN.md)
        self.collective = collective if world_size > 1 else "none"
        self.use_cuda_graph, self._graph = (use_cuda_graph and (world_size == 1 or collective == "peer")), None
        self._graph_play, self._probe_after_replay, self._graph_play_key = None, False, None
        self.graph_launches = {"play": 0, "update": 0}
        self.fused_step = True      # single rank + tensor cores: use PolicyMLP.minibatch_step (fused reduce / Adam / re-pack tail)
        info = vec_env.get_env_info()
        self.obs_dim = int(info["observation_space"]["state"].shape[0])
        self.num_actors = int(vec_env.env.num_envs)
        self.T = cfg.horizon_length
        self.batch_size = self.T * self.num_actors
        self.minibatch_size = min(cfg.minibatch_size, self.batch_size)
        assert self.batch_size % self.minibatch_size == 0, "batch_size must be a multiple of minibatch_size (as in rl_games)"
        self.num_minibatches = self.batch_size // self.minibatch_size
        self.policy = PolicyMLP(self.obs_dim, self.device, seed=cfg.seed, lr=cfg.learning_rate, e_clip=cfg.e_clip,
                                critic_coef=cfg.critic_coef, entropy_coef=cfg.entropy_coef, bounds_loss_coef=cfg.bounds_loss_coef,
                                clip_value=cfg.clip_value, grad_norm=cfg.grad_norm if cfg.truncate_grads else 0.0,
                                kl_threshold=cfg.kl_threshold, adaptive_lr=cfg.lr_schedule == "adaptive", world_size=world_size)
        self.policy.seed = cfg.seed + rank                                  # [ref: RLG/torch_runner.py:74-75]
        self.peer, self.peer_step = None, None
        if self.multi_gpu:                                                  # [ref: a2c_common.py:1350-1355]
            dist.broadcast(self.policy.params, 0)
            if self.collective == "peer":
                from .peer import PeerAllReduce, PeerStepExchange
                self.peer = PeerAllReduce(self.policy.grads.numel(), self.device, rank, world_size)
                if self.policy.tensor_cores:      # the all-reduce fused into the cooperative minibatch tail (3 launches per minibatch)
                    self.peer_step = PeerStepExchange(self.obs_dim, self.device, rank, world_size)
        N, T, D = self.num_actors, self.T, self.obs_dim
        f32 = dict(dtype=torch.float32, device=self.device)
        self.buf = dict(obses=torch.zeros((T, N, D), **f32), rewards=torch.zeros((T, N), **f32), values=torch.zeros((T, N, 1), **f32),
                        neglogpacs=torch.zeros((T, N), **f32), dones=torch.zeros((T, N), dtype=torch.uint8, device=self.device),
                        actions=torch.zeros((T, N, A), **f32), mus=torch.zeros((T, N, A), **f32), sigmas=torch.zeros((T, N, A), **f32))
        self.advs = torch.zeros((T, N), **f32)
        self.returns = torch.zeros((T, N), **f32)
        self.last_values = torch.zeros((N, 1), **f32)
        self.dones = torch.ones(N, dtype=torch.uint8, device=self.device)   # [ref: a2c_common.py:454]
        self._dones_next = torch.zeros((2, N), dtype=torch.uint8, device=self.device)   # double buffer: step n's flags are read at step n+1
        self.current_rewards = torch.zeros(N, **f32)
        self.current_lengths = torch.zeros(N, **f32)
        # finished-episode accumulators: [sum of returns, sum of lengths, count] (all-reduced once per epoch)
        self.episode_acc = torch.zeros(3, dtype=torch.float64, device=self.device)
        # the reference's meters: mean over the last `games_to_track` finished episodes  [ref: a2c_common.py:214-216,739-747]
        self.game_rewards = AverageMeter(cfg.games_to_track, self.device)
        self.game_lengths = AverageMeter(cfg.games_to_track, self.device)
        # RLGPUAlgoObserver: extras['episode'] of every step, averaged per print  [ref: OIGE/utils/rlgames/rlgames_utils.py:51-99]
        self.ep_info_sum, self.ep_info_n = {}, torch.zeros((), dtype=torch.float32, device=self.device)
        self.obs = None
        self.epoch_num = 0
        self.frame = 0
        self.mean_reward = float("nan")

    # ---- rollout  [ref: a2c_common.py:670-774] -------------------------------------------------------
    def play_steps(self):
        """T control steps of every env + policy inference, with no host sync and only static buffers across calls, so the whole
        rollout is capturable in a CUDA graph (train_epoch)."""
        b, pol, T = self.buf, self.policy, self.T
        if self.obs is None:
            self.obs = self.vec_env.reset()["obs"]["state"].clone()       # static across epochs (graph replays read it in place)
        obs, dones_u8 = self.obs, self.dones
        for n in range(T):
            b["obses"][n].copy_(obs)
            b["dones"][n].copy_(dones_u8)
            out = dict(actions=b["actions"][n], neglogpacs=b["neglogpacs"][n], values=b["values"][n], mus=b["mus"][n], sigmas=b["sigmas"][n])
            pol.act(b["obses"][n], out, row_offset=self.rank * self.num_actors)
            # preprocess_actions: clamp to [-1, 1]  [ref: a2c_common.py:1134-1144] (the stored action stays unclamped)
            obs_dict, rew, dones, infos = self.vec_env.step(torch.clamp(b["actions"][n], -1.0, 1.0))
            obs = obs_dict["obs"]["state"]
            # reward shaping (DefaultRewardsShaper [ref: tr_helpers.py:33-42]), uint8 dones and the episode bookkeeping [ref:
            # a2c_common.py:720-747] in ONE launch, no host sync: ~40 tiny elementwise / reduction launches per control step otherwise
            dones_u8 = self._dones_next[n & 1]
            rew_c, dones_c = rew.contiguous(), dones.contiguous()
            _lib.check(_lib.lib().ppo_rollout_bookkeep_f32(
                _lib.ptr(rew_c, torch.float32), _lib.ptr(dones_c, torch.int64), ctypes.c_float(self.cfg.reward_scale), _lib.ptr(b["rewards"][n]),
                _lib.ptr(dones_u8), _lib.ptr(self.current_rewards), _lib.ptr(self.current_lengths), _lib.ptr(self.episode_acc),
                ctypes.c_void_p(self.game_rewards.mean.data_ptr()), ctypes.c_void_p(self.game_rewards.current_size.data_ptr()),
                ctypes.c_void_p(self.game_lengths.mean.data_ptr()), ctypes.c_void_p(self.game_lengths.current_size.data_ptr()),
                ctypes.c_float(self.game_rewards.max_size), ctypes.c_int64(self.num_actors), _lib.stream()), "ppo_rollout_bookkeep_f32")
            ep = infos.get("episode") if isinstance(infos, dict) else None
            if ep:                                                       # observer.process_infos
                for k, v in ep.items():
                    if k not in self.ep_info_sum:
                        self.ep_info_sum[k] = torch.zeros((), dtype=torch.float32, device=self.device)
                    self.ep_info_sum[k] += v
                self.ep_info_n += 1
        self.obs.copy_(obs)
        self.dones.copy_(dones_u8)
        pol.values(self.obs, self.last_values)
        gae(b["rewards"], b["values"].view(T, -1), b["dones"], self.last_values.view(-1), self.dones, self.cfg.gamma, self.cfg.tau,
            self.advs, self.returns)

    # ---- dataset  [ref: a2c_common.py:1257-1320 ; swap_and_flatten01 :30-37] ---------------------------
    def prepare_dataset(self):
        b, pol = self.buf, self.policy
        flat = lambda x: x.transpose(0, 1).reshape(self.batch_size, *x.shape[2:]).contiguous()      # (T,N,..) -> env-major (N*T,..)
        values, returns = flat(b["values"]).view(-1), flat(self.returns)
        advantages = returns - values
        if self.cfg.normalize_value:                       # train(): update with values, normalise; then with returns
            pol.val_rms.update(values.view(-1, 1))
            values = pol.val_rms.normalize(values.view(-1, 1)).view(-1)
            pol.val_rms.update(returns.view(-1, 1))
            returns = pol.val_rms.normalize(returns.view(-1, 1)).view(-1)
        if self.cfg.normalize_advantage:
            advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
        self.ds = dict(obs=flat(b["obses"]), actions=flat(b["actions"]), old_logp_actions=flat(b["neglogpacs"]),
                       advantages=advantages.contiguous(), old_values=values.contiguous(), returns=returns.contiguous(),
                       mu=flat(b["mus"]), sigma=flat(b["sigmas"]))

    # ---- update  [ref: a2c_common.py:1197-1245 ; a2c_continuous.py:78-196] -----------------------------
    def update(self):
        """prepare_dataset + mini_epochs x num_minibatches PPO steps.  Touches only static buffers and device-side
        scalars (lr, Adam step, normaliser moments), so the whole phase is captured once into a CUDA graph
        (~800 launches, one NCCL all-reduce per minibatch) and replayed every epoch."""
        self.prepare_dataset()
        ds, pol, mb = self.ds, self.policy, self.minibatch_size
        for mini_ep in range(self.cfg.mini_epochs):
            for i in range(self.num_minibatches):
                s = slice(i * mb, (i + 1) * mb)
                if self.cfg.normalize_input and mini_ep == 0:
                    pol.obs_rms.update(ds["obs"][s])                    # train-mode forward updates the normaliser first
                args = (ds["obs"][s], ds["actions"][s], ds["old_logp_actions"][s], ds["advantages"][s], ds["old_values"][s],
                        ds["returns"][s], ds["mu"][s], ds["sigma"][s])
                if pol.tensor_cores and self.fused_step and (not self.multi_gpu or self.peer_step is not None):
                    # gradient (+ all-reduce over NVLink inside the tail kernel) + clip + Adam + lr + operand-tile refresh: 3 launches
                    pol.minibatch_step(*args, peer=self.peer_step)
                    continue
                pol.minibatch_grad(*args)
                if self.peer is not None:
                    self.peer(pol.grads)                                 # gradient + KL + loss stats in one span, over NVLink peer memory
                elif self.multi_gpu:
                    dist.all_reduce(pol.grads, op=dist.ReduceOp.SUM)
                pol.optimizer_step()

    def _rollout_graph_ok(self) -> bool:
        """The rollout is captured when the env is a fused engine (classic or live: their kernels take a device-side step offset) and
        the NaN probe can be deferred to one flag read per epoch."""
        task = getattr(getattr(self.vec_env, "env", None), "_task", None)
        eng = getattr(task, "engine", None)
        return (self.use_cuda_graph and eng is not None and getattr(eng, "_buffers", None) is not None
                and bool(eng._buffers.step_offset) and not eng.cfg.spawn_curriculum)

    def _play(self):
        if not self._rollout_graph_ok() or self.epoch_num < 2:
            self.play_steps()                                            # eager (first epochs: reset, allocator warm-up)
            return
        task = self.vec_env.env._task
        eng, pol, T = task.engine, self.policy, self.T
        key = eng.graph_key(T)               # host-side step parameters the capture bakes in (live task: the initial action bias)
        if key is None:
            self.play_steps()                                            # the window straddles a parameter change
            return
        if key != self._graph_play_key:
            self._graph_play, self._graph_play_key = None, key
        if self._graph_play is None:
            probe, task._nan_probe = task._nan_probe, False              # no host sync inside the capture: checked after replay
            saved = (eng.step_counter, eng.first_call, pol.sample_counter, task.step, task._calls)
            pol._packed_dirty = True                                     # the captured rollout re-packs the weights on every replay
            torch.cuda.synchronize(self.device)
            l0 = _lib.launch_count()
            self._graph_play = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_play):
                self.play_steps()
                eng.advance_step_offset(T)
                pol.advance_counter_offset(T)
            self.graph_launches["play"] = _lib.launch_count() - l0        # kernels of this library inside one replay
            # the capture ran the host code once without executing anything: rewind the host-side counters
            eng.step_counter, eng.first_call, pol.sample_counter, task.step, task._calls = saved
            self._probe_after_replay = probe
        self._graph_play.replay()
        eng.note_graph_replay(T)
        pol.note_graph_replay(T)
        task.step += T / task.cfg.horizon_length
        task._calls += T
        if self._probe_after_replay:
            eng.check_finite()

    def train_epoch(self):
        t0 = time.perf_counter()
        with torch.no_grad():
            self._play()
            t1 = time.perf_counter()
            if not self.use_cuda_graph:
                self.update()
            elif self._graph is None and self.epoch_num < 2:
                self.update()                                            # eager warm-up (allocator, NCCL communicators)
            elif self._graph is None:
                torch.cuda.synchronize(self.device)
                l0 = _lib.launch_count()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self.update()
                self.graph_launches["update"] = _lib.launch_count() - l0
                self._graph.replay()
            else:
                self._graph.replay()
        self.epoch_num += 1
        self.frame += self.batch_size * self.world
        return t1 - t0, time.perf_counter() - t1

    def check_peers(self) -> None:
        """Raises when a gradient exchange timed out on this rank (one host read per window; the kernels leave the parameters
        untouched from the failing step on, so ranks cannot silently diverge)."""
        for p in (self.peer, self.peer_step):
            if p is not None:
                p.check()

    def ranks_identical(self) -> bool:
        """True when every rank holds bit-identical parameters, Adam moments and learning rate (all-gather of an exact checksum)."""
        P = self.policy
        return ranks_hold_identical([P.params, P.exp_avg, P.exp_avg_sq, P.lr, P.step], self.world)

    def episode_stats(self):
        """Mean return / length of the episodes finished since the last call, reduced over ranks (one small all-reduce)."""
        out = reduce_episode_stats(self.episode_acc, self.world)
        self.episode_acc.zero_()
        return out

    def episode_infos(self) -> dict:
        """Episode/<key>: the mean of extras['episode'][key] over the steps since the last call (RLGPUAlgoObserver.after_print_stats)."""
        n = float(self.ep_info_n)
        out = {k: float(v) / n for k, v in self.ep_info_sum.items()} if n > 0 else {}
        for v in self.ep_info_sum.values():
            v.zero_()
        self.ep_info_n.zero_()
        return out

    def train(self, max_epochs: Optional[int] = None, log_every: int = 10, log=print, writer=None, nn_dir: Optional[str] = None,
              name: str = "USV", save_freq: int = 0, save_best_after: int = 100, score_to_win: Optional[float] = None):
        """The outer loop of ContinuousA2CBase.train [ref: RLG/common/a2c_common.py:1336-1486] with its scalar stream and checkpoint cadence.

        `writer`: anything with add_scalar(tag, value, step) -- a tensorboard SummaryWriter, or `ScalarLog` below (JSON lines; the image has
        no tensorboardX); tags are the reference's (write_stats :343-362 and the rewards / episode_lengths block :1399-1418).
        `nn_dir`: checkpoints in the reference's .pth schema and naming: `last_<name>_ep_<epoch>_rew_<mean>.pth` every `save_freq` epochs when
        the reward did not improve, `<name>.pth` whenever the mean reward of the last `games_to_track` episodes beats the best so far
        (after `save_best_after` epochs), `last_<name>_ep_<epoch>_rew_<mean>.pth` at max_epochs.
        Everything host-side happens at the log points only (every `log_every` epochs): one synchronize, a handful of scalar reads."""
        import os
        max_epochs = max_epochs or self.cfg.max_epochs
        if nn_dir:
            os.makedirs(nn_dir, exist_ok=True)
        self.last_mean_rewards = getattr(self, "last_mean_rewards", -1e9)
        t_start, t_last, frames_last, play_acc, upd_acc = time.perf_counter(), time.perf_counter(), self.frame, 0.0, 0.0
        while self.epoch_num < max_epochs:
            play, update = self.train_epoch()
            play_acc += play
            upd_acc += update
            at_end = self.epoch_num >= max_epochs
            if self.epoch_num % log_every == 0 or at_end:
                torch.cuda.synchronize(self.device)
                self.check_peers()
                now = time.perf_counter()
                rew, length, cnt = self.episode_stats()
                if cnt:
                    self.mean_reward = rew
                st = self.policy.stats()
                frame, epoch, total_time = self.frame, self.epoch_num, now - t_start
                have_games = len(self.game_rewards) > 0
                mean_rewards, mean_lengths = self.game_rewards.get_mean(), self.game_lengths.get_mean()
                if self.rank == 0 and writer is not None:
                    dt, df = max(now - t_last, 1e-9), frame - frames_last
                    writer.add_scalar("performance/step_inference_rl_update_fps", df / dt, frame)
                    writer.add_scalar("performance/rl_update_time", upd_acc, frame)
                    writer.add_scalar("performance/step_inference_time", play_acc, frame)
                    writer.add_scalar("losses/a_loss", st["a_loss"], frame)
                    writer.add_scalar("losses/c_loss", st["c_loss"], frame)
                    writer.add_scalar("losses/entropy", st["entropy"], frame)
                    writer.add_scalar("losses/bounds_loss", st["b_loss"], frame)
                    writer.add_scalar("info/last_lr", st["lr"], frame)
                    writer.add_scalar("info/lr_mul", 1.0, frame)
                    writer.add_scalar("info/e_clip", self.cfg.e_clip, frame)
                    writer.add_scalar("info/kl", st["kl"], frame)
                    writer.add_scalar("info/epochs", epoch, frame)
                    if have_games:
                        for tag, val in (("rewards", mean_rewards), ("shaped_rewards", mean_rewards * self.cfg.reward_scale), ("episode_lengths", mean_lengths)):
                            writer.add_scalar(tag + "/step", val, frame)
                            writer.add_scalar(tag + "/iter", val, epoch)
                            writer.add_scalar(tag + "/time", val, total_time)
                t_last, frames_last, play_acc, upd_acc = now, frame, 0.0, 0.0
                infos = self.episode_infos()
                if self.rank == 0 and writer is not None:
                    for k, v in infos.items():                 # RLGPUAlgoObserver.after_print_stats
                        writer.add_scalar("Episode/" + k, v, frame)
                if self.rank == 0 and log:
                    log(f"epoch {epoch} frames {frame} reward {rew:.3f} len {length:.1f} ({cnt} eps) "
                        f"meter[{len(self.game_rewards)}] {mean_rewards:.3f}/{mean_lengths:.1f} "
                        f"kl {st['kl']:.5f} lr {st['lr']:.2e} a_loss {st['a_loss']:.4f} c_loss {st['c_loss']:.4f}")
                    if infos:
                        log("  Episode/ " + " ".join(f"{k}={v:.4g}" for k, v in infos.items()))
                if self.rank == 0 and nn_dir and have_games:
                    ck = f"{name}_ep_{epoch}_rew_{mean_rewards}"
                    if save_freq > 0 and epoch % save_freq == 0 and mean_rewards <= self.last_mean_rewards:
                        self.save(os.path.join(nn_dir, "last_" + ck + ".pth"))
                    if mean_rewards > self.last_mean_rewards and epoch >= save_best_after:
                        self.last_mean_rewards = mean_rewards
                        self.save(os.path.join(nn_dir, name + ".pth"))
                        if score_to_win is not None and mean_rewards > score_to_win:
                            self.save(os.path.join(nn_dir, ck + ".pth"))
                            max_epochs = self.epoch_num                 # "Network won": leave the loop like the reference
                if self.rank == 0 and nn_dir and self.epoch_num >= max_epochs:
                    tag = mean_rewards if have_games else float("-inf")
                    self.save(os.path.join(nn_dir, f"last_{name}_ep_{epoch}_rew_{tag}.pth"))
        return self.mean_reward

    # ---- checkpoints: the reference's .pth schema  [ref: a2c_common.py:590-654] --------------------------
    def get_full_state_weights(self):
        P = self.policy
        off, state = 0, {}
        from .policy import param_shapes
        import math
        for idx, shp in enumerate(param_shapes(P.D)):
            n = math.prod(shp)
            state[idx] = {"step": P.step.float().cpu().reshape(()), "exp_avg": P.exp_avg[off:off + n].view(shp).clone(),
                          "exp_avg_sq": P.exp_avg_sq[off:off + n].view(shp).clone()}
            off += n
        opt = {"state": state, "param_groups": [{"lr": float(P.lr), "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": 0,
                                                 "amsgrad": False, "params": list(range(9))}]}
        return {"model": P.state_dict(), "epoch": self.epoch_num, "optimizer": opt, "frame": self.frame,
                "last_mean_rewards": self.mean_reward, "env_state": None}

    def save(self, path: str):
        torch.save(self.get_full_state_weights(), path)

    def restore(self, path: str):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        self.policy.load_state_dict(ck["model"])
        self.epoch_num, self.frame = int(ck.get("epoch", 0)), int(ck.get("frame", 0))
        opt = ck.get("optimizer")
        if opt and opt.get("state"):
            from .policy import param_shapes
            import math
            off = 0
            for idx, shp in enumerate(param_shapes(self.policy.D)):
                n = math.prod(shp)
                st = opt["state"][idx]
                self.policy.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.policy.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                off += n
            self.policy.step.fill_(int(opt["state"][0]["step"]))
            self.policy.lr.fill_(float(opt["param_groups"][0]["lr"]))
