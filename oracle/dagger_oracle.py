"""TEST INFRASTRUCTURE ONLY (CPU restatement; not product code, nothing in the package imports it).

The USV SysID distillation step of the fork's DAgger stack -- SURVEY 8(f) row 4, second half.  The CUDA path is csrc/dagger_sysid.cu behind
algo/ppo/dagger.py and module.StateHistoryEncoder; this oracle and its goldens are the parity gate (tests/test_gpu_dagger.py):

  StateHistoryEncoder.forward           OIGE/algo/ppo/module.py:392-448   (student: history of T non-privileged observations -> latent)
  USVSysIDAgent.evaluate                OIGE/algo/ppo/dagger.py:50-66     (action head on [current obs | student latent])
  ObsStorage.mini_batch_generator_inorder  OIGE/algo/ppo/storage.py:36-42 (time-major flattening, contiguous row blocks)
  USVSysIDTrainer._train_step / update  OIGE/algo/ppo/dagger.py:125-196   (MSE to the frozen teacher latent, Adam(5e-4), StepLR(200, 0.1),
                                                                            R^2 / variance diagnostics)

Plain functional torch (F.linear / F.conv1d, explicit Adam), parameters as a flat list in the reference's `parameters()` order.
Pinned against the reference's own classes by oracle/make_golden_dagger.py -> tests/golden/dagger_sysid.npz (tests/test_dagger_oracle_cpu.py).
Reference quirk kept: the per-step projection (bs*T, 32) is RESHAPED to (bs, 32, T), not transposed (module.py:445-446), so the
"channels" the convolutions see interleave time steps and features.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F

# (kernel, stride) of the Conv1d stack per history length  [module.py:404-436]
CONV_STACKS = {50: ((8, 4), (5, 1), (5, 1)), 20: ((6, 2), (4, 2)), 10: ((4, 2), (2, 1))}
LEAKY = 0.01                      # nn.LeakyReLU default negative slope


def history_encoder_shapes(input_size: int, tsteps: int, output_size: int) -> List[Tuple[int, ...]]:
    """Parameter shapes in `StateHistoryEncoder.parameters()` order: encoder, conv layers, linear_output."""
    shapes: List[Tuple[int, ...]] = [(32, input_size), (32,)]
    for k, _ in CONV_STACKS[tsteps]:
        shapes += [(32, 32, k), (32,)]
    return shapes + [(output_size, 32 * 3), (output_size,)]


def history_encoder_forward(params: Sequence[torch.Tensor], hist: torch.Tensor, tsteps: int, act=F.leaky_relu) -> torch.Tensor:
    """hist (bs, tsteps * input_size) -> latent (bs, output_size).  `act` is the constructor's activation_fn (the USV script passes
    LeakyReLU); the conv stack always uses LeakyReLU."""
    bs = hist.shape[0]
    p = list(params)
    x = act(F.linear(hist.reshape(bs * tsteps, -1), p[0], p[1]))
    x = x.reshape(bs, -1, tsteps)                                   # the reference's reshape (not a transpose)
    i = 2
    for _, stride in CONV_STACKS[tsteps]:
        x = F.leaky_relu(F.conv1d(x, p[i], p[i + 1], stride=stride), LEAKY)
        i += 2
    return act(F.linear(x.flatten(1), p[i], p[i + 1]))


def student_action(params, head, sysid_obs: torch.Tensor, tsteps: int, obs_nonpriv_dim: int) -> torch.Tensor:
    """USVSysIDAgent.evaluate: sysid_obs = [history_flat | current non-privileged obs]; `head` maps [current | z_hat] to the action."""
    hd = tsteps * obs_nonpriv_dim
    zhat = history_encoder_forward(params, sysid_obs[:, :hd], tsteps)
    return head(torch.cat([sysid_obs[:, hd:hd + obs_nonpriv_dim], zhat], dim=1))


class SysIDTrainerOracle:
    """USVSysIDTrainer with explicit Adam (betas 0.9 / 0.999, eps 1e-8, no weight decay) and StepLR(step_size=200, gamma=0.1)."""

    def __init__(self, params: Sequence[torch.Tensor], tsteps: int, latent_dim: int, num_learning_epochs: int = 4, num_mini_batches: int = 4,
                 learning_rate: float = 5e-4):
        self.p = [q.detach().clone().requires_grad_(True) for q in params]
        self.m = [torch.zeros_like(q) for q in self.p]
        self.v = [torch.zeros_like(q) for q in self.p]
        self.t, self.itr = 0, 0
        self.tsteps, self.latent_dim = int(tsteps), int(latent_dim)
        self.epochs, self.mbs, self.lr0 = int(num_learning_epochs), int(num_mini_batches), float(learning_rate)

    @property
    def lr(self) -> float:
        return self.lr0 * (0.1 ** (self.itr // 200))                # scheduler.step() once per update

    def _adam(self, grads, lr):
        self.t += 1
        bc1, bc2 = 1.0 - 0.9 ** self.t, 1.0 - 0.999 ** self.t
        with torch.no_grad():
            for q, g, m, v in zip(self.p, grads, self.m, self.v):
                m.lerp_(g, 0.1)
                v.mul_(0.999).addcmul_(g, g, value=0.001)
                q.addcdiv_(m, (v.sqrt() / math.sqrt(bc2)).add_(1e-8), value=-lr / bc1)

    def update(self, hist: torch.Tensor, zstar: torch.Tensor) -> dict:
        """hist (T, N, history_dim), zstar (T, N, latent_dim): one `USVSysIDTrainer.update()` on a full storage."""
        lr = self.lr
        X, Z = hist.reshape(-1, hist.shape[-1]), zstar.reshape(-1, self.latent_dim)
        mb = X.shape[0] // self.mbs
        avg = 0.0
        for _ in range(self.epochs):
            tot = 0.0
            for b in range(self.mbs):
                xb, zb = X[b * mb:(b + 1) * mb], Z[b * mb:(b + 1) * mb]
                loss = F.mse_loss(history_encoder_forward(self.p, xb, self.tsteps), zb)
                grads = torch.autograd.grad(loss, self.p)
                self._adam(grads, lr)
                tot += float(loss.detach())
            avg = tot / max(1, self.mbs)
        self.itr += 1
        out = {"mse": avg}
        with torch.no_grad():
            zh = history_encoder_forward(self.p, X, self.tsteps)
            out["zstar_var_mean"] = float(torch.var(Z, dim=0, unbiased=False).mean())
            out["zhat_var_mean"] = float(torch.var(zh, dim=0, unbiased=False).mean())
            sse, sst = ((zh - Z) ** 2).sum(0), ((Z - Z.mean(0, keepdim=True)) ** 2).sum(0)
            r2 = 1.0 - sse / (sst + 1e-8)
            for i in range(self.latent_dim):
                out[f"r2_dim{i}"] = float(r2[i])
            out["r2_total"] = float(1.0 - sse.sum() / (sst.sum() + 1e-8))
        return out
