#!/usr/bin/env python
"""The loopz training loop on the GPU learner [ref: OIGE/scripts/rlgames_train_loopz.py:700-1440]: MLPEncode actor / critic with a
squashed Gaussian, horizon-length rollouts, reward scale 0.01, 4 epochs x 4 in-order minibatches, lr 5e-4, std floor 0.05 after
every update.  python scripts/train_loopz.py --num-envs 4096 --updates 200 [--classic]"""
import argparse
import os
import sys
import time
from collections import deque

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.algo.ppo import PPO, module as ppo_module
from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg, load_task_yaml
from omniisaacgymenvs_loop_b200.envs.usv_raisim_vecenv import USVRaisimVecEnv
from omniisaacgymenvs_loop_b200.envs.vec_env_rlgames import VecEnvRLGames
from omniisaacgymenvs_loop_b200.tasks.USV_Virtual import SimConfig, USVVirtual


def make_env(task_cfg: dict, device: str, seed: int):
    env = VecEnvRLGames(headless=True)
    sim = SimConfig({"sim_device": device, "rl_device": device, "seed": seed, "env_id_offset": 0, "task": task_cfg})
    env.set_task(USVVirtual("USVVirtual", sim, env, collect_stats=False), backend="torch")
    return USVRaisimVecEnv(env)


def build_learner(env, device, horizon, seed, mass_dim=8, sampling="in_order", use_cuda_graph=True):
    """rlgames_train_loopz.py:784-842"""
    arch = dict(speed_dim=3, mass_dim=mass_dim, mass_latent_dim=8, mass_encoder_shape=[64, 16])
    actor = ppo_module.Actor(ppo_module.MLPEncode_wrap([128, 128], "LeakyReLU", env.num_obs, env.num_acts, "Tanh", False, **arch),
                             ppo_module.SquashedGaussianDiagonalCovariance(env.num_acts, 0.3, action_scale=1.0), device, seed=seed)
    critic = ppo_module.Critic(ppo_module.MLPEncode_wrap([128, 128], "LeakyReLU", env.num_obs, 1, **arch), device)
    return PPO(actor=actor, critic=critic, num_envs=env.num_envs, num_transitions_per_env=horizon, num_learning_epochs=4, gamma=0.997,
               lam=0.95, num_mini_batches=4, device=device, mini_batch_sampling=sampling, learning_rate=5e-4, use_cuda_graph=use_cuda_graph)


def train(env, ppo, updates, horizon=16, reward_scale=0.01, log_every=10, quiet=False):
    """-> list of (update, mean episode return over the last 100 finished episodes, frames/s)."""
    dev = ppo.device
    ep_ret = torch.zeros(env.num_envs, device=dev)
    window = deque(maxlen=100)
    hist = []
    min_std = torch.full((env.num_acts,), 0.05, device=dev)
    env.reset()
    for update in range(updates):
        t0 = time.time()
        for _ in range(horizon):
            obs = env.observe(as_numpy=False)
            action = ppo.observe(obs)
            reward, dones = env.step(action)
            ep_ret += reward
            ppo.step(value_obs=obs, rews=reward * reward_scale, dones=dones, infos=[])
            if log_every and update % log_every == 0:          # episode-return monitor (host read only on logging updates)
                d = dones.bool()
                if bool(d.any()):
                    window.extend(ep_ret[d].tolist())
            ep_ret.masked_fill_(dones.bool(), 0.0)
        ppo.update(actor_obs=env.observe(as_numpy=False), value_obs=env.observe(as_numpy=False), log_this_iteration=False, update=update)
        ppo.actor.distribution.enforce_minimum_std(min_std)
        if log_every and update % log_every == 0:
            torch.cuda.synchronize(dev)
            fps = horizon * env.num_envs / (time.time() - t0)
            mean_ret = sum(window) / len(window) if window else float("nan")
            hist.append((update, mean_ret, fps))
            if not quiet:
                print(f"update {update:5d}  return(100) {mean_ret:9.3f}  std {ppo.actor.distribution.std.tolist()}  "
                      f"v_loss {ppo.last_stats['mean_value_loss']:.4f}  fps {fps:,.0f}", flush=True)
    return hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task-yaml", default=None)
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--updates", type=int, default=200)
    ap.add_argument("--horizon", type=int, default=16)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--save", default=None)
    ap.add_argument("--checkpoint", default=None)
    args = ap.parse_args()
    device = "cuda:0"
    task_cfg = load_task_yaml(args.task_yaml, num_envs=args.num_envs) if args.task_yaml else live_task_cfg(live_default_config(num_envs=args.num_envs))
    env = make_env(task_cfg, device, args.seed)
    ppo = build_learner(env, device, args.horizon, args.seed)
    start = 0
    if args.checkpoint:
        start = ppo.load_state_dict(torch.load(args.checkpoint, map_location=device, weights_only=False))
        print(f"[loopz] resumed from {args.checkpoint} (start_update={start})")
    train(env, ppo, args.updates, args.horizon)
    if args.save:
        torch.save(ppo.state_dict(start + args.updates - 1), args.save)


if __name__ == "__main__":
    main()
