#!/usr/bin/env python
"""north_star: "end-to-end PPO frames/sec on CaptureXY matching reference reward curves".  One seed, classic CaptureXY, 512 envs, horizon 16,
one 8192-row minibatch x 8 mini-epochs per epoch, `--epochs` epochs, three learners:

  tf32    rl/a2c.A2CAgent on the tcgen05 TF32 kernels (the default path)
  fp32    the same agent on the fp32 SIMT kernels (the 1e-5 numerics reference)
  oracle  the CPU oracle of the rl_games loop (oracle/ppo_oracle.train_epoch + oracle/usv_oracle.ClassicEnvOracle; test infrastructure --
          this script is a measurement tool, like bench.py's cpu_baseline leg)

Metric per epoch: mean undiscounted return of the episodes that finished during the epoch.  The three runs do NOT share random streams
(Philox in the kernels, torch's generator in the oracle), so curves are compared in distribution: mean return over windows of epochs.
Writes one JSON object (curves + window means) to stdout / --out."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig
from scripts.train_usv import make_env


def gpu_curve(envs, epochs, seed, tensor_cores, device):
    cfg = UsvEnvConfig(num_envs=envs)
    env = make_env(cfg.to_task_cfg(), device, seed=seed, collect_stats=False)
    agent = A2CAgent(env, PPOConfig(seed=seed, minibatch_size=8192), device)
    agent.policy.tensor_cores = agent.policy.tensor_cores and tensor_cores
    curve, step_rew = [], []
    for _ in range(epochs):
        agent.train_epoch()
        r, _l, c = agent.episode_stats()
        curve.append(r if c else float("nan"))
        step_rew.append(float(agent.buf["rewards"].mean()) / agent.cfg.reward_scale)
    return curve, step_rew


def oracle_curve(envs, epochs, seed, horizon=16, minibatch=8192, mini_epochs=8):
    from oracle import ppo_oracle as P
    from oracle import usv_oracle as O
    from tests.util import oracle_cfg

    D = 13
    env = O.ClassicEnvOracle(oracle_cfg(UsvEnvConfig(num_envs=envs)), envs)
    lay = P.param_layout(D)
    g = torch.Generator().manual_seed(seed)
    params = torch.zeros(lay["P"])
    for name in ("w1", "w2", "wv", "wmu"):                      # nn.Linear default init, zero biases, logstd 0 (as PolicyMLP.reset_parameters)
        a, b = lay[name]
        fan = D if name == "w1" else 128
        params[a:b] = (torch.rand(b - a, generator=g) * 2 - 1) / fan ** 0.5
    m, v, step, lr = torch.zeros_like(params), torch.zeros_like(params), 0, 1e-4
    obs_rms, val_rms = P.RunningMeanStd((D,)), P.RunningMeanStd((1,))
    obs, _, done = env.step(torch.zeros((envs, 2)))
    ret_run = torch.zeros(envs)
    curve, step_rew = [], []
    for _ in range(epochs):
        roll = {k: [] for k in ("obses", "actions", "neglogpacs", "values", "mus", "sigmas", "rewards", "dones")}
        fin_sum, fin_n = 0.0, 0
        with torch.no_grad():
            for _t in range(horizon):
                r = P.policy_inference(params, obs, D, obs_rms, val_rms, eps=torch.randn((envs, 2), generator=g))
                roll["obses"].append(obs.clone()); roll["dones"].append(done.to(torch.uint8))
                for k in ("actions", "neglogpacs", "values", "mus", "sigmas"):
                    roll[k].append(r[k])
                obs, rew, done = env.step(torch.clamp(r["actions"], -1.0, 1.0))
                ret_run += rew
                d = done.bool()
                fin_sum += float(ret_run[d].sum()); fin_n += int(d.sum())
                ret_run[d] = 0.0
                roll["rewards"].append(rew * 0.01)
            roll = {k: torch.stack(x) for k, x in roll.items()}
            last_v = P.policy_inference(params, obs, D, obs_rms, val_rms)["values"][:, 0]
            adv = P.discount_values(done.float(), last_v, roll["dones"].float(), roll["values"][..., 0], roll["rewards"])
            ds = P.prepare_dataset(roll, (adv + roll["values"][..., 0]).unsqueeze(-1), val_rms)
        params, m, v, step, lr = P.train_epoch(params, m, v, step, lr, ds, D, obs_rms, minibatch_size=min(minibatch, horizon * envs),
                                               mini_epochs=mini_epochs)
        curve.append(fin_sum / fin_n if fin_n else float("nan"))
        step_rew.append(float(roll["rewards"].mean()) / 0.01)
    return curve, step_rew


def windows(curve, k=4):
    import math
    n = len(curve) // k
    out = []
    for i in range(k):
        xs = [x for x in curve[i * n:(i + 1) * n] if not math.isnan(x)]
        out.append(sum(xs) / len(xs) if xs else float("nan"))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--seeds", default="11,12,13", help="comma-separated seeds (GPU learners run all of them, the oracle the first --oracle-seeds)")
    ap.add_argument("--oracle-seeds", type=int, default=2)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    seeds = [int(x) for x in a.seeds.split(",")]
    res = {"envs": a.envs, "epochs": a.epochs, "seeds": seeds, "metric": "per epoch: mean undiscounted return of the episodes finished in the epoch (NaN: none finished) and mean reward per env-step of the rollout; *_last_half = mean step reward over the second half of the epochs",
           "runs": []}
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    for name, fn in (("tf32", lambda sd: gpu_curve(a.envs, a.epochs, sd, True, "cuda:0")), ("fp32", lambda sd: gpu_curve(a.envs, a.epochs, sd, False, "cuda:0")),
                     ("oracle", lambda sd: oracle_curve(a.envs, a.epochs, sd))):
        for sd in (seeds if name != "oracle" else seeds[:a.oracle_seeds]):
            t0 = time.perf_counter()
            c, sr = fn(sd)
            run = {"learner": name, "seed": sd, "episode_return_quarters": windows(c), "step_reward_quarters": windows(sr),
                   "last_half_mean": sum(windows(sr)[2:]) / 2, "seconds": time.perf_counter() - t0, "episode_return_curve": c, "step_reward_curve": sr}
            res["runs"].append(run)
            print(name, sd, "episode return", [round(x, 2) for x in run["episode_return_quarters"]], "step reward", [round(x, 3) for x in run["step_reward_quarters"]],
                  f"{run['seconds']:.1f} s", file=sys.stderr, flush=True)
    for name in ("tf32", "fp32", "oracle"):
        xs = [r["last_half_mean"] for r in res["runs"] if r["learner"] == name]
        mu = sum(xs) / len(xs)
        res[name + "_last_half"] = {"mean": mu, "min": min(xs), "max": max(xs), "n": len(xs)}
    txt = json.dumps(res)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt + "\n")
    print(json.dumps({k: v for k, v in res.items() if k != "runs"}))
