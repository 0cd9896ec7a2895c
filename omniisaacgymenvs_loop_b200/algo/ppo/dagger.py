"""USV SysID distillation (the DAgger stack's USV path) on the C-ABI kernels of csrc/dagger_sysid.cu
[ref: omniisaacgymenvs/algo/ppo/dagger.py:13-196 -- USVSysIDAgent, USVSysIDTrainer].

  teacher latent   z* = mass_encoder(priv_tail)                (frozen; `MLPEncode.mass_encoder`, dagger_mlp_forward_f32)
  student latent   z^ = id_encoder(history_flat)               (`StateHistoryEncoder`, dagger_history_encoder_forward_f32)
  action           a  = frozen_action_head([obs_nonpriv, z^])  (frozen; `MLPEncode.action_mlp`)
  update           4 epochs x 4 in-order minibatches of  MSE(z^, z*) -> Adam(5e-4), StepLR(200, 0.1) once per update
                   (dagger_sysid_minibatch_step_f32: forward + backward + Adam in two launches per minibatch, no host sync inside an update)

Same class / method surface as the reference; numpy in -> numpy out where the reference does that, CUDA tensors stay on the device."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ... import _lib
from .module import StateHistoryEncoder
from .storage import ObsStorage


class USVSysIDAgent:
    def __init__(self, *, teacher_mass_encoder, id_encoder: StateHistoryEncoder, frozen_action_head, history_len: int, obs_nonpriv_dim: int,
                 device: str) -> None:
        self.teacher_mass_encoder, self.id_encoder, self.frozen_action_head = teacher_mass_encoder, id_encoder, frozen_action_head
        self.history_len, self.obs_nonpriv_dim = int(history_len), int(obs_nonpriv_dim)
        self.history_dim = self.history_len * self.obs_nonpriv_dim
        self.device = torch.device(device)
        if id_encoder.tsteps != self.history_len or id_encoder.input_size != self.obs_nonpriv_dim:
            raise ValueError("id_encoder was built for another history shape")

    def set_itr(self, _itr) -> None:
        return

    def get_history_encoding(self, sysid_obs_torch: torch.Tensor) -> torch.Tensor:
        # sysid_obs = [history_flat | current non-privileged obs]: the kernel reads the history columns of the wide rows in place
        return self.id_encoder(sysid_obs_torch)

    def evaluate(self, sysid_obs_torch: torch.Tensor) -> torch.Tensor:
        zhat = self.get_history_encoding(sysid_obs_torch)
        cur = sysid_obs_torch[:, self.history_dim:self.history_dim + self.obs_nonpriv_dim]
        return self.frozen_action_head(torch.cat([cur.to(zhat.device, torch.float32), zhat], dim=1))

    def get_student_action(self, sysid_obs_torch: torch.Tensor) -> torch.Tensor:
        return self.evaluate(sysid_obs_torch)

    def teacher_latent(self, priv_tail_torch: torch.Tensor) -> torch.Tensor:
        return self.teacher_mass_encoder(priv_tail_torch)


class USVSysIDTrainer:
    """MSE(id_encoder(history), mass_encoder(priv tail)) with Adam + StepLR(200, 0.1), in-order minibatches of the time-major storage."""

    def __init__(self, *, actor: USVSysIDAgent, num_envs: int, num_transitions_per_env: int, history_dim: int, latent_dim: int,
                 num_learning_epochs: int = 4, num_mini_batches: int = 4, device: str, learning_rate: float = 5e-4) -> None:
        self.actor, self.device = actor, torch.device(device)
        self.history_dim, self.latent_dim = int(history_dim), int(latent_dim)
        enc = actor.id_encoder
        if self.history_dim != enc.tsteps * enc.input_size or self.latent_dim != enc.output_size:
            raise ValueError("history_dim / latent_dim do not match the id_encoder")
        self.storage = ObsStorage(int(num_envs), int(num_transitions_per_env), [self.history_dim], [self.latent_dim], self.device)
        self.num_transitions_per_env, self.num_envs = int(num_transitions_per_env), int(num_envs)
        self.num_learning_epochs, self.num_mini_batches = int(num_learning_epochs), int(num_mini_batches)
        self.itr = 0
        self.base_lr = float(learning_rate)
        f32 = dict(dtype=torch.float32, device=self.device)
        P = enc.flat.numel()
        self.exp_avg, self.exp_avg_sq = torch.zeros(P, **f32), torch.zeros(P, **f32)
        self.grads = torch.zeros(P + 1, **f32)
        L = _lib.lib()
        self.scratch = torch.empty(int(L.dagger_train_scratch_floats(ctypes.c_int32(enc.input_size), ctypes.c_int32(enc.tsteps),
                                                                     ctypes.c_int32(enc.output_size))), **f32)
        self.lr = torch.full((1,), self.base_lr, **f32)
        self.adam_step = torch.zeros(2, dtype=torch.int32, device=self.device)
        self._parity = 0
        self._mse = torch.zeros(1, **f32)

    # ---- rollout side ------------------------------------------------------------------------------------------------------
    def observe(self, sysid_obs):
        """Student action for the wrapper's sysid_obs; numpy in -> numpy out (the reference's contract), CUDA tensor in -> CUDA tensor out."""
        as_numpy = isinstance(sysid_obs, np.ndarray)
        obs = torch.from_numpy(sysid_obs).to(self.device) if as_numpy else sysid_obs
        act = self.actor.get_student_action(obs)
        return act.cpu().numpy() if as_numpy else act

    def step(self, sysid_obs, priv_tail_torch: torch.Tensor) -> None:
        """Stores one transition's supervision pair: history_flat and the teacher latent z*."""
        z_star = self.actor.teacher_latent(priv_tail_torch.to(self.device, dtype=torch.float32))[:, :self.latent_dim]
        hist = sysid_obs[:, :self.history_dim]
        self.storage.add_obs(hist.astype(np.float32, copy=False) if isinstance(hist, np.ndarray) else hist, z_star)

    # ---- update --------------------------------------------------------------------------------------------------------------
    def _minibatch(self, hist: torch.Tensor, zstar: torch.Tensor) -> None:
        enc = self.actor.id_encoder
        rc = _lib.lib().dagger_sysid_minibatch_step_f32(
            _lib.ptr(enc.flat), _lib.ptr(hist), ctypes.c_int64(hist.shape[1]), _lib.ptr(zstar), ctypes.c_int32(enc.input_size),
            ctypes.c_int32(enc.tsteps), ctypes.c_int32(enc.output_size), _lib.ptr(self.grads), _lib.ptr(self.scratch), _lib.ptr(self.exp_avg),
            _lib.ptr(self.exp_avg_sq), _lib.ptr(self.lr), _lib.ptr(self.adam_step), ctypes.c_int32(self._parity), _lib.ptr(self._mse),
            ctypes.c_int64(hist.shape[0]), _lib.stream())
        _lib.check(rc, "dagger_sysid_minibatch_step_f32")
        self._parity ^= 1

    def _train_step(self) -> float:
        self.itr += 1
        self.actor.set_itr(self.itr)
        for _epoch in range(self.num_learning_epochs):
            self._mse.zero_()
            n = 0
            for hist_batch, zstar_batch in self.storage.mini_batch_generator_inorder(self.num_mini_batches):
                self._minibatch(hist_batch, zstar_batch)
                n += 1
        avg = float(self._mse.item()) / max(1, n)                   # the last epoch's mean minibatch loss: the update's one host read
        self.lr.fill_(self.base_lr * (0.1 ** (self.itr // 200)))    # scheduler.step(): StepLR(step_size=200, gamma=0.1)
        return avg

    def update(self) -> dict:
        mse = self._train_step()
        metrics = {"mse": mse}
        obs_all = self.storage.obs.view(-1, self.history_dim)
        zstar_all = self.storage.expert.view(-1, self.latent_dim)
        zhat_all = self.actor.id_encoder(obs_all)
        # R^2 / variance guardrails of the collected batch [ref: dagger.py:150-176]: a few reductions over (B, 8) tensors, once per update
        metrics["zstar_var_mean"] = float(torch.var(zstar_all, dim=0, unbiased=False).mean())
        metrics["zhat_var_mean"] = float(torch.var(zhat_all, dim=0, unbiased=False).mean())
        sse = torch.sum((zhat_all - zstar_all) ** 2, dim=0)
        sst = torch.sum((zstar_all - zstar_all.mean(dim=0, keepdim=True)) ** 2, dim=0)
        r2 = (1.0 - sse / (sst + 1e-8)).tolist()
        for i in range(self.latent_dim):
            metrics[f"r2_dim{i}"] = r2[i]
        metrics["r2_total"] = float(1.0 - sse.sum() / (sst.sum() + 1e-8))
        self.storage.clear()
        return metrics

    # ---- checkpoints: the student's weights + Adam state ----------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"id_encoder_state_dict": self.actor.id_encoder.state_dict(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "step": int(self.adam_step[self._parity].item()), "itr": self.itr}

    def load_state_dict(self, sd: dict) -> None:
        self.actor.id_encoder.load_state_dict(sd["id_encoder_state_dict"])
        if "exp_avg" in sd:
            self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.adam_step.fill_(int(sd.get("step", 0)))
        self.itr = int(sd.get("itr", 0))
        self.lr.fill_(self.base_lr * (0.1 ** (self.itr // 200)))
