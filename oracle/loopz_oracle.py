"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (plain torch, fp32) of the loopz PPO learner of the reference, one function per reference function.  Pinned
against `tests/golden/loopz_ppo.npz`, which `oracle/make_golden.py:loopz` produces by running the UNMODIFIED reference classes
(`tests/test_loopz_oracle_cpu.py`); the CUDA kernels of `csrc/ppo_loopz.cu` are then checked against this file on other seeds
and sizes (`tests/test_gpu_loopz.py`).

Flat parameter vector: [actor: mass_encoder.{0,2,4}.{w,b}, action_mlp.{0,2,4}.{w,b} | std[2] | critic: the same twelve]
= `[*actor.parameters(), *critic.parameters()]` of the reference optimiser (OIGE/algo/ppo/ppo.py:60).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

H, E1, E2, E3 = 128, 64, 16, 8


@dataclass
class LoopzCfg:
    obs_dim: int = 33
    speed_dim: int = 3
    mass_dim: int = 8
    tanh_out: bool = True            # architecture.activation: tanh  (OIGE/cfg/task/USV/IROS2024/cfg.yaml:37)
    action_scale: float = 1.0        # task.env.clipActions           (rlgames_train_loopz.py:778-781)
    eps: float = 1e-6
    clip_param: float = 0.2          # PPO defaults (ppo.py:19-27) as used by rlgames_train_loopz.py:828-842
    value_loss_coef: float = 0.5
    entropy_coef: float = 0.0
    use_clipped_value_loss: bool = True
    max_grad_norm: float = 0.5
    learning_rate: float = 5e-4
    gamma: float = 0.997
    lam: float = 0.95


def net_shapes(cfg: LoopzCfg, out: int):
    IN = cfg.obs_dim - cfg.mass_dim + E3
    return [(E1, cfg.mass_dim), (E1,), (E2, E1), (E2,), (E3, E2), (E3,), (H, IN), (H,), (H, H), (H,), (out, H), (out,)]


def split(flat: torch.Tensor, cfg: LoopzCfg):
    """flat -> (actor tensors[12], std[2], critic tensors[12]) as views."""
    off, nets = 0, []
    for out in (2, 1):
        ts = []
        for shp in net_shapes(cfg, out):
            n = math.prod(shp)
            ts.append(flat[off:off + n].view(shp))
            off += n
        nets.append(ts)
        if out == 2:
            std = flat[off:off + 2]
            off += 2
    assert off == flat.numel()
    return nets[0], std, nets[1]


def mlp_encode(p, x: torch.Tensor, cfg: LoopzCfg, tanh_out: bool) -> torch.Tensor:
    """MLPEncode.forward  [ref OIGE/algo/ppo/module.py:340-361]: obs = [speed | task | mass]; the mass tail goes through its own
    encoder (Linear + LeakyReLU three times, :250-270) and the latent is concatenated behind speed and task."""
    sd, md = cfg.speed_dim, cfg.mass_dim
    td = cfg.obs_dim - sd - md
    speed, task, mass = x[:, :sd], x[:, sd:sd + td], x[:, sd + td:sd + td + md]
    z = F.leaky_relu(F.linear(mass, p[0], p[1]))
    z = F.leaky_relu(F.linear(z, p[2], p[3]))
    z = F.leaky_relu(F.linear(z, p[4], p[5]))
    h = torch.cat([speed, task, z], dim=1)
    h = F.leaky_relu(F.linear(h, p[6], p[7]))
    h = F.leaky_relu(F.linear(h, p[8], p[9]))
    o = F.linear(h, p[10], p[11])
    return torch.tanh(o) if tanh_out else o


def log_prob_from_u(mean, std, u, cfg: LoopzCfg):
    """SquashedGaussianDiagonalCovariance._log_prob_from_u  [ref module.py:555-566]."""
    var = std ** 2
    logp_u = (-((u - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(dim=1)
    scale = torch.full((2,), cfg.action_scale)
    log_det = torch.log(scale + cfg.eps).sum() + torch.log(1.0 - torch.tanh(u).pow(2) + cfg.eps).sum(dim=1)
    return logp_u - log_det


def sample_from_noise(mean, std, noise, cfg: LoopzCfg):
    """sample() with the standard-normal draw injected  [ref module.py:568-583]."""
    u = mean + std * noise
    return torch.tanh(u) * cfg.action_scale, log_prob_from_u(mean, std, u, cfg)


def evaluate(mean, std, actions, cfg: LoopzCfg):
    """evaluate(): log-prob of stored actions, "entropy" := -log_prob  [ref module.py:586-637]."""
    mean = torch.nan_to_num(mean, nan=0.0, posinf=0.0, neginf=0.0) if not torch.isfinite(mean).all() else mean
    std = torch.nan_to_num(std, nan=1.0, posinf=1.0, neginf=1.0) if not torch.isfinite(std).all() else std
    a = torch.clamp(actions / (cfg.action_scale + cfg.eps), -1.0 + cfg.eps, 1.0 - cfg.eps)
    u = 0.5 * (torch.log1p(a) - torch.log1p(-a))
    lp = log_prob_from_u(mean, std, u, cfg)
    return lp, -lp


def compute_returns(rewards, values, dones, last_values, gamma, lam):
    """RolloutStorage.compute_returns  [ref OIGE/algo/ppo/storage.py:92-124]; tensors are [T, N, 1], dones uint8."""
    nn = dict(nan=0.0, posinf=0.0, neginf=0.0)
    rewards, values, last_values = torch.nan_to_num(rewards, **nn), torch.nan_to_num(values, **nn), torch.nan_to_num(last_values, **nn)
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        next_is_not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + next_is_not_terminal * gamma * next_values - values[step]
        advantage = delta + next_is_not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    advantages = returns - values
    adv_std = torch.nan_to_num(advantages.std(), **nn)
    advantages = (advantages - advantages.mean()) / (adv_std + 1e-8)
    return torch.nan_to_num(returns, **nn), torch.nan_to_num(advantages, **nn)


def minibatch_loss(flat, cfg: LoopzCfg, actor_obs, critic_obs, actions, target_values, advantages, returns, old_log_prob):
    """The loss of one minibatch of PPO._train_step  [ref OIGE/algo/ppo/ppo.py:245-284]; column tensors are [M, 1]."""
    pa, std, pc = split(flat, cfg)
    nn = dict(nan=0.0, posinf=0.0, neginf=0.0)
    mean = mlp_encode(pa, torch.nan_to_num(actor_obs, **nn), cfg, cfg.tanh_out)
    logp, entropy = evaluate(mean, std.reshape(2), actions, cfg)
    value = mlp_encode(pc, torch.nan_to_num(critic_obs, **nn), cfg, False)
    ratio = torch.exp(logp - old_log_prob.squeeze(-1))
    surrogate = -advantages.squeeze(-1) * ratio
    surrogate_clipped = -advantages.squeeze(-1) * torch.clamp(ratio, 1.0 - cfg.clip_param, 1.0 + cfg.clip_param)
    surrogate_loss = torch.max(surrogate, surrogate_clipped)
    if cfg.use_clipped_value_loss:
        value_clipped = target_values + (value - target_values).clamp(-cfg.clip_param, cfg.clip_param)
        value_loss = torch.max((value - returns).pow(2), (value_clipped - returns).pow(2))
    else:
        value_loss = (returns - value).pow(2)
    loss = (surrogate_loss + cfg.value_loss_coef * value_loss.squeeze(-1) - cfg.entropy_coef * entropy).mean()
    return loss, surrogate_loss.mean(), value_loss.mean(), logp


class Adam:
    """torch.optim.Adam (amsgrad False, weight decay 0) on the flat vector, preceded by clip_grad_norm_  [ref ppo.py:291-299]."""

    def __init__(self, n, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.m, self.v, self.t = torch.zeros(n), torch.zeros(n), 0
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps

    def step(self, flat: torch.Tensor, grad: torch.Tensor, max_grad_norm: float):
        norm = grad.norm(2)
        grad = grad * torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0)
        self.t += 1
        self.m.lerp_(grad, 1 - self.b1)
        self.v.mul_(self.b2).addcmul_(grad, grad, value=1 - self.b2)
        bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        denom = (self.v.sqrt() / math.sqrt(bc2)).add_(self.eps)
        flat.addcdiv_(self.m, denom, value=-(self.lr / bc1))
        return norm


def minibatch_grad(flat, cfg, *batch):
    p = flat.detach().clone().requires_grad_(True)
    loss, surr, vloss, logp = minibatch_loss(p, cfg, *batch)
    loss.backward()
    return p.grad.detach(), float(loss.detach()), float(surr.detach()), float(vloss.detach())


def train_step(flat, cfg: LoopzCfg, storage: dict, num_learning_epochs=4, num_mini_batches=4, opt: Adam | None = None, index_lists=None):
    """PPO._train_step with in-order minibatches (or the given index lists per epoch)  [ref ppo.py:232-321, storage.py:138-148].
    `storage`: actor_obs, critic_obs, actions, values, advantages, returns, actions_log_prob as [T, N, k].  Updates `flat` in place."""
    opt = opt or Adam(flat.numel(), cfg.learning_rate)
    rows = lambda k: storage[k].reshape(-1, storage[k].shape[-1])
    cols = [rows(k) for k in ("actor_obs", "critic_obs", "actions", "values", "advantages", "returns", "actions_log_prob")]
    B = cols[0].shape[0]
    mb = B // num_mini_batches
    sum_v = sum_s = 0.0
    n_valid = 0
    for ep in range(num_learning_epochs):
        batches = index_lists[ep] if index_lists is not None else [slice(b * mb, (b + 1) * mb) for b in range(num_mini_batches)]
        for sel in batches:
            grad, loss, surr, vloss = minibatch_grad(flat, cfg, *[c[sel] for c in cols])
            if not math.isfinite(loss):
                continue
            opt.step(flat, grad, cfg.max_grad_norm)
            sum_v, sum_s, n_valid = sum_v + vloss, sum_s + surr, n_valid + 1
    return (sum_v / n_valid, sum_s / n_valid) if n_valid else (0.0, 0.0)


def enforce_minimum_std(std, min_std):
    """[ref module.py:649-659]"""
    cur = torch.where(torch.isfinite(std), std, min_std)
    return torch.max(cur, min_std)
