"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the rl_games PPO path.

torch-on-CPU restatement of the reference's algorithm, one function per reference function with the
file:line it follows (RLG = rl_games/rl_games/).  Pinned against outputs of the reference's own
rl_games classes (tests/golden/ppo.npz, produced by oracle/make_golden.py under oracle/ref_shim.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import philox

H, A = 128, 2


def param_layout(D: int) -> dict:
    """Flat parameter vector in model.parameters() order [RLG/algos_torch/network_builder.py:1488-1575]:
    sigma | actor_mlp.0.weight | .bias | actor_mlp.2.weight | .bias | value.weight | value.bias | mu.weight | mu.bias"""
    o, off = {}, 0
    for name, n in (("sigma", A), ("w1", H * D), ("b1", H), ("w2", H * H), ("b2", H), ("wv", H), ("bv", 1), ("wmu", A * H), ("bmu", A)):
        o[name] = (off, off + n)
        off += n
    o["P"] = off
    return o


def unpack(params: torch.Tensor, D: int) -> dict:
    L = param_layout(D)
    g = lambda k: params[L[k][0]:L[k][1]]
    return dict(sigma=g("sigma"), w1=g("w1").view(H, D), b1=g("b1"), w2=g("w2").view(H, H), b2=g("b2"), wv=g("wv").view(1, H),
                bv=g("bv"), wmu=g("wmu").view(A, H), bmu=g("bmu"))


# RunningMeanStd  [RLG/algos_torch/running_mean_std.py:69-117]
class RunningMeanStd:
    def __init__(self, shape, epsilon=1e-5):
        self.mean = torch.zeros(shape, dtype=torch.float64)
        self.var = torch.ones(shape, dtype=torch.float64)
        self.count = torch.ones((), dtype=torch.float64)
        self.eps = epsilon

    def update(self, x: torch.Tensor):
        bm, bv, bc = x.mean(0), x.var(0), x.shape[0]                     # :86-87 (fp32 batch moments, unbiased var)
        delta = bm - self.mean                                           # :70-79 (Chan merge, promoted to fp64)
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        M2 = self.var * self.count + bv * bc + delta ** 2 * self.count * bc / tot
        self.mean, self.var, self.count = new_mean, M2 / tot, tot

    def normalize(self, x):                                              # :113-116
        y = (x - self.mean.float()) / torch.sqrt(self.var.float() + self.eps)
        return torch.clamp(y, min=-5.0, max=5.0)

    def denormalize(self, x):                                            # :108-110
        y = torch.clamp(x, min=-5.0, max=5.0)
        return torch.sqrt(self.var.float() + self.eps) * y + self.mean.float()


def mlp_forward(params, obs_n, D):
    """a2c_network forward on already-normalised obs -> mu (M,2), value (M,1), logstd (M,2)
    [RLG/algos_torch/network_builder.py:1577-1629]"""
    p = unpack(params, D)
    h1 = torch.tanh(torch.nn.functional.linear(obs_n, p["w1"], p["b1"]))
    h2 = torch.tanh(torch.nn.functional.linear(h1, p["w2"], p["b2"]))
    mu = torch.nn.functional.linear(h2, p["wmu"], p["bmu"])
    value = torch.nn.functional.linear(h2, p["wv"], p["bv"])
    logstd = mu * 0.0 + p["sigma"]                                       # fixed_sigma: state-independent parameter
    return mu, value, logstd


def neglogp(x, mean, std, logstd):                                       # [RLG/algos_torch/models.py:398-401]
    return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * np.log(2.0 * np.pi) * x.size()[-1] + logstd.sum(dim=-1)


def normal_eps(seed: int, rows, counter: int) -> torch.Tensor:
    """The N(0,1) pair the CUDA rollout kernel draws for each row: Box-Muller on Philox stream 100."""
    rows = np.asarray(rows, dtype=np.uint64)
    hi = (rows >> np.uint64(32)) & philox.MASK
    c3 = (np.uint64(100) ^ (hi << np.uint64(8))) & philox.MASK
    r = philox.philox4x32_10(rows & philox.MASK, np.uint64(counter & 0xFFFFFFFF), np.uint64((counter >> 32) & 0xFFFFFFFF), c3,
                             seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = ((r[0] >> np.uint32(8)).astype(np.float64) + 1.0) / 16777216.0
    u2 = (r[1] >> np.uint32(8)).astype(np.float64) / 16777216.0
    rad = np.sqrt(-2.0 * np.log(u1))
    return torch.from_numpy(np.stack([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)], 1).astype(np.float32))


def policy_inference(params, obs, D, obs_rms: RunningMeanStd, val_rms: RunningMeanStd, eps=None):
    """Network.forward(is_train=False)  [RLG/algos_torch/models.py:366-397]"""
    mu, value, logstd = mlp_forward(params, obs_rms.normalize(obs), D)
    sigma = torch.exp(logstd)
    out = dict(mus=mu, sigmas=sigma, values=val_rms.denormalize(value))
    if eps is not None:
        act = mu + sigma * eps
        out.update(actions=act, neglogpacs=neglogp(act, mu, sigma, logstd))
    return out


def actor_loss(old_nlp, nlp, adv, e_clip):                               # [RLG/common/common_losses.py:39-48]
    ratio = torch.exp(old_nlp - nlp)
    s1 = adv * ratio
    s2 = adv * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip)
    return torch.max(-s1, -s2)


def critic_loss(old_v, v, e_clip, ret, clip_value=True):                 # [RLG/common/common_losses.py:10-20]
    if clip_value:
        vpc = old_v + (v - old_v).clamp(-e_clip, e_clip)
        return torch.max((v - ret) ** 2, (vpc - ret) ** 2)
    return (ret - v) ** 2


def bound_loss(mu, soft_bound=1.1):                                      # [RLG/algos_torch/a2c_continuous.py:209-217]
    return (torch.clamp_max(mu + soft_bound, 0.0) ** 2 + torch.clamp_min(mu - soft_bound, 0.0) ** 2).sum(axis=-1)


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma):                         # [RLG/algos_torch/torch_ext.py:27-36]
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma ** 2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma ** 2 + 1e-5))
    return (c1 + c2 - 0.5).sum(dim=-1).mean()


def minibatch_loss(params, batch, D, obs_rms, *, e_clip=0.2, critic_coef=0.5, entropy_coef=0.0, bounds_loss_coef=1e-4,
                   clip_value=True):
    """calc_gradients up to the scalar loss  [RLG/algos_torch/a2c_continuous.py:78-159]"""
    mu, value, logstd = mlp_forward(params, obs_rms.normalize(batch["obs"]), D)
    sigma = torch.exp(logstd)
    nlp = neglogp(batch["actions"], mu, sigma, logstd)
    entropy = torch.distributions.Normal(mu, sigma).entropy().sum(dim=-1)
    a = actor_loss(batch["old_logp_actions"], nlp, batch["advantages"], e_clip)
    c = critic_loss(batch["old_values"], value, e_clip, batch["returns"], clip_value)
    b = bound_loss(mu)
    a_m, c_m, e_m, b_m = a.mean(), c.mean(), entropy.mean(), b.mean()
    loss = a_m + 0.5 * c_m * critic_coef - e_m * entropy_coef + b_m * bounds_loss_coef
    kl = policy_kl(mu.detach(), sigma.detach(), batch["mu"], batch["sigma"])
    return loss, dict(a_loss=a_m, c_loss=c_m, entropy=e_m, b_loss=b_m, kl=kl, mu=mu.detach(), sigma=sigma.detach())


def adam_step(params, grads, m, v, step, lr, *, beta1=0.9, beta2=0.999, eps=1e-8, grad_norm=1.0):
    """clip_grad_norm_ + torch.optim.Adam single step on flat tensors  [RLG/common/a2c_common.py:325-330]"""
    norm = grads.norm()
    coef = torch.clamp(grad_norm / (norm + 1e-6), max=1.0) if grad_norm > 0 else torch.tensor(1.0)
    g = grads * coef
    m = torch.lerp(m, g, 1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return params - (lr / bc1) * (m / denom), m, v, norm


def adaptive_lr(lr, kl, kl_threshold=0.016, min_lr=1e-6, max_lr=1e-2):   # [RLG/common/schedulers.py:26-32]
    if kl > 2.0 * kl_threshold:
        lr = max(lr / 1.5, min_lr)
    if kl < 0.5 * kl_threshold:
        lr = min(lr * 1.5, max_lr)
    return lr


def discount_values(fdones, last_values, mb_fdones, mb_values, mb_rewards, gamma=0.99, tau=0.95):
    """GAE  [RLG/common/a2c_common.py:525-540]; tensors shaped (T,N) / (N,)"""
    T = mb_rewards.shape[0]
    last = 0
    adv = torch.zeros_like(mb_rewards)
    for t in reversed(range(T)):
        nnt = 1.0 - (fdones if t == T - 1 else mb_fdones[t + 1])
        nv = last_values if t == T - 1 else mb_values[t + 1]
        delta = mb_rewards[t] + gamma * nv * nnt - mb_values[t]
        adv[t] = last = delta + gamma * tau * nnt * last
    return adv


def swap_and_flatten01(x):                                               # [RLG/common/a2c_common.py:30-37]
    return x.transpose(0, 1).reshape(x.shape[0] * x.shape[1], *x.shape[2:])


def prepare_dataset(roll: dict, returns, val_rms: RunningMeanStd, *, normalize_value=True, normalize_advantage=True) -> dict:
    """play_steps' hand-over + A2CBase.prepare_dataset  [RLG/common/a2c_common.py:760-774,1257-1320]: (T,N,..) rollout tensors ->
    env-major flat dataset; the value normaliser is updated with the values and applied to them, THEN updated with the returns and
    applied to those (two training-mode forwards); advantages come from the un-normalised pair and are standardised with the
    unbiased std over the whole batch."""
    fl = swap_and_flatten01
    values, rets = fl(roll["values"]), fl(returns)
    adv = rets - values
    if normalize_value:
        val_rms.update(values)
        values = val_rms.normalize(values)
        val_rms.update(rets)
        rets = val_rms.normalize(rets)
    adv = adv.sum(dim=1)
    if normalize_advantage:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    return dict(old_values=values, old_logp_actions=fl(roll["neglogpacs"]), advantages=adv, returns=rets, actions=fl(roll["actions"]),
                obs=fl(roll["obses"]), mu=fl(roll["mus"]).clone(), sigma=fl(roll["sigmas"]).clone())


def train_epoch(params, m, v, step, lr, ds: dict, D, obs_rms: RunningMeanStd, *, minibatch_size, mini_epochs, normalize_input=True,
                kl_threshold=0.016, grad_norm=1.0, trace=None, **loss_kw):
    """The update half of ContinuousA2CBase.train_epoch  [RLG/common/a2c_common.py:1197-1245 ; common/datasets.py:25-77]: in-order
    contiguous minibatches (no shuffling), the obs normaliser updated by the minibatches of mini-epoch 0 only (training-mode forward,
    then .eval()), mu / sigma of every minibatch written back into the dataset, 'legacy' adaptive-KL lr after every minibatch.
    Returns (params, m, v, step, lr)."""
    n_mb = ds["obs"].shape[0] // minibatch_size
    for mini_ep in range(mini_epochs):
        for i in range(n_mb):
            sl = slice(i * minibatch_size, (i + 1) * minibatch_size)
            batch = {k: t[sl] for k, t in ds.items()}
            batch["old_values"], batch["returns"] = batch["old_values"].reshape(-1, 1), batch["returns"].reshape(-1, 1)
            if normalize_input and mini_ep == 0:
                obs_rms.update(batch["obs"])
            prm = params.clone().requires_grad_(True)
            loss, st = minibatch_loss(prm, batch, D, obs_rms, **loss_kw)
            loss.backward()
            step += 1
            params, m, v, norm = adam_step(params, prm.grad, m, v, step, lr, grad_norm=grad_norm)
            ds["mu"][sl], ds["sigma"][sl] = st["mu"], st["sigma"]                   # dataset.update_mu_sigma
            if trace is not None:
                trace.append(dict(a_loss=st["a_loss"].detach(), c_loss=st["c_loss"].detach(), kl=st["kl"], lr=lr, mu=st["mu"], sigma=st["sigma"],
                                  params=params.clone(), obs_mean=obs_rms.mean.clone(), obs_count=obs_rms.count.clone()))
            lr = adaptive_lr(lr, float(st["kl"]), kl_threshold)
    return params, m, v, step, lr
