#!/usr/bin/env python
"""USV SysID distillation loop [ref: omniisaacgymenvs/scripts/dagger_usv_sysid_loopz.py:140-420]: the frozen loopz teacher's mass encoder
labels every step with z* = mass_encoder(privileged tail), the student (StateHistoryEncoder over the last 50 non-privileged observations)
drives the boat through the frozen action head and is regressed onto z* every `horizon` steps.  Everything stays on the device: fused live
env step, history window, student inference, teacher labels, MSE / backward / Adam kernels; one host read per update (the metrics).

    python scripts/dagger_usv_sysid_loopz.py --envs 4096 --updates 50 [--checkpoint teacher.pt]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.algo.ppo.dagger import USVSysIDAgent, USVSysIDTrainer
from omniisaacgymenvs_loop_b200.algo.ppo.module import StateHistoryEncoder
from omniisaacgymenvs_loop_b200.config import UsvLiveConfig, live_default_config, live_task_cfg
from omniisaacgymenvs_loop_b200.envs.usv_raisim_vecenv import USVSysIDVecEnv
from scripts.train_loopz import build_learner, make_env


def build(envs: int, device: str, seed: int, history_len: int = 50, priv_dim: int = 8, horizon: int = 16, checkpoint=None):
    base = make_env(live_task_cfg(live_default_config(num_envs=envs), UsvLiveConfig(priv_dim=priv_dim)), device, seed=seed)
    env = USVSysIDVecEnv(base._env, history_len=history_len, priv_dim=priv_dim, device=device)
    ppo = build_learner(base, device, horizon, seed=seed, mass_dim=priv_dim)         # the teacher policy (random unless a checkpoint is given)
    if checkpoint:
        ppo.load_state_dict(torch.load(checkpoint, map_location=device, weights_only=False))
    arch = ppo.actor.architecture.architecture
    student = StateHistoryEncoder("LeakyReLU", env.obs_nonpriv_dim, history_len, 8, device, seed=seed)
    agent = USVSysIDAgent(teacher_mass_encoder=arch.mass_encoder, id_encoder=student, frozen_action_head=arch.action_mlp,
                          history_len=history_len, obs_nonpriv_dim=env.obs_nonpriv_dim, device=device)
    trainer = USVSysIDTrainer(actor=agent, num_envs=env.num_envs, num_transitions_per_env=horizon, history_dim=history_len * env.obs_nonpriv_dim,
                              latent_dim=8, num_learning_epochs=4, num_mini_batches=4, device=device, learning_rate=5e-4)
    return env, trainer


def run(env, trainer, updates: int, horizon: int = 16, log_every: int = 10, quiet: bool = False):
    env.reset()
    out = []
    for upd in range(updates):
        for _ in range(horizon):
            sysid_obs = env.observe_sysid_obs(as_numpy=False)
            trainer.step(sysid_obs, env.get_priv_tail())
            env.step(trainer.observe(sysid_obs))
        m = trainer.update()
        out.append(m)
        if not quiet and (upd % log_every == 0 or upd == updates - 1):
            print(f"[sysid] update {upd}: mse {m['mse']:.5f} r2_total {m['r2_total']:.4f} zstar_var {m['zstar_var_mean']:.4f} zhat_var {m['zhat_var_mean']:.4f}",
                  flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--updates", type=int, default=50)
    ap.add_argument("--horizon", type=int, default=16)
    ap.add_argument("--history-len", type=int, default=50)
    ap.add_argument("--priv-dim", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    torch.cuda.set_device(a.device)
    torch.manual_seed(a.seed)
    env, trainer = build(a.envs, a.device, a.seed, a.history_len, a.priv_dim, a.horizon, a.checkpoint)
    run(env, trainer, 2, a.horizon, quiet=True)                      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms = run(env, trainer, a.updates, a.horizon)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"[sysid] {a.envs} envs x horizon {a.horizon}: {a.updates} updates in {dt:.2f} s = {a.updates * a.horizon * a.envs / dt:.3e} frames/s "
          f"({dt / a.updates * 1e3:.1f} ms per collect + update); final r2_total {ms[-1]['r2_total']:.4f}")
