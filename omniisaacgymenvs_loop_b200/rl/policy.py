"""USV_PPOcontinuous_MLP policy/value network on the C-ABI kernels: flat parameter vector in rl_games'
model.parameters() order, fp64 running normalisers, fused forward / fused loss+backward / fused clip+Adam+lr.

Surface mirrors what A2CAgent uses of its model [ref: RLG/algos_torch/models.py:366-401,
RLG/common/a2c_common.py:385-420]; checkpoints use the reference's state_dict key names so
`last_USV_ep_*.pth` files load."""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch

from .. import _lib

H, A = 128, 2
STAT = {k[len("PPO_STAT_"):].lower(): v for k, v in _lib.ENUMS.items() if k.startswith("PPO_STAT_") and k != "PPO_STAT_COUNT"}
N_STAT = _lib.ENUMS["PPO_STAT_COUNT"]

_NAMES = ["a2c_network.sigma", "a2c_network.actor_mlp.0.weight", "a2c_network.actor_mlp.0.bias", "a2c_network.actor_mlp.2.weight",
          "a2c_network.actor_mlp.2.bias", "a2c_network.value.weight", "a2c_network.value.bias", "a2c_network.mu.weight",
          "a2c_network.mu.bias"]


def param_shapes(D: int):
    return [(A,), (H, D), (H,), (H, H), (H,), (1, H), (1,), (A, H), (A,)]


class RunningMeanStd:
    """fp64 running moments (Chan merge) + fp32 copies the kernels read  [ref: RLG/algos_torch/running_mean_std.py:65-117]."""

    def __init__(self, size: int, device, epsilon: float = 1e-5):
        self.mean = torch.zeros(size, dtype=torch.float64, device=device)
        self.var = torch.ones(size, dtype=torch.float64, device=device)
        self.count = torch.ones((), dtype=torch.float64, device=device)
        self.eps = epsilon
        self.mean32 = self.mean.float()
        self.var32 = self.var.float()

    def _sync(self):
        self.mean32.copy_(self.mean)
        self.var32.copy_(self.var)

    def update(self, x: torch.Tensor) -> None:
        """One fused kernel: batch moments + Chan merge + fp32 copies (in place: CUDA-graph friendly)."""
        x = x.reshape(x.shape[0], -1)
        if not x.is_contiguous():
            x = x.contiguous()
        rc = _lib.lib().ppo_rms_update_f64(_lib.ptr(x, torch.float32), ctypes.c_int64(x.shape[0]), ctypes.c_int32(x.shape[1]),
                                           _lib.ptr(self.mean), _lib.ptr(self.var), _lib.ptr(self.count.view(1)), _lib.ptr(self.mean32),
                                           _lib.ptr(self.var32), _lib.stream())
        _lib.check(rc, "ppo_rms_update_f64")

    def update_from_moments(self, bm, bv, bc) -> None:
        # in place: the buffers keep their addresses, so the update can live inside a captured CUDA graph
        delta = bm - self.mean
        tot = self.count + bc
        new_mean = self.mean + delta * bc / tot
        new_var = (self.var * self.count + bv * bc + delta ** 2 * self.count * bc / tot) / tot
        self.mean.copy_(new_mean)
        self.var.copy_(new_var)
        self.count.copy_(tot)
        self._sync()

    def normalize(self, x):
        return torch.clamp((x - self.mean32) / torch.sqrt(self.var32 + self.eps), -5.0, 5.0)

    def denormalize(self, x):
        return torch.sqrt(self.var32 + self.eps) * torch.clamp(x, -5.0, 5.0) + self.mean32

    def load(self, mean, var, count):
        self.mean.copy_(torch.as_tensor(mean, dtype=torch.float64))
        self.var.copy_(torch.as_tensor(var, dtype=torch.float64))
        self.count.copy_(torch.as_tensor(count, dtype=torch.float64))
        self._sync()


class PolicyMLP:
    def __init__(self, obs_dim: int, device="cuda:0", seed: int = 0, lr: float = 1e-4, e_clip: float = 0.2, critic_coef: float = 0.5,
                 entropy_coef: float = 0.0, bounds_loss_coef: float = 1e-4, clip_value: bool = True, grad_norm: float = 1.0,
                 kl_threshold: float = 0.016, adaptive_lr: bool = True, world_size: int = 1, tensor_cores: bool = True):
        self.lib = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.UsvLibraryError("PolicyMLP runs on CUDA only (no CPU fallback)")
        self.D = int(obs_dim)
        self.P = int(self.lib.ppo_param_count(ctypes.c_int32(self.D)))
        f32 = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.P, **f32)
        self.exp_avg = torch.zeros(self.P, **f32)
        self.exp_avg_sq = torch.zeros(self.P, **f32)
        self.grads = torch.zeros(self.P + N_STAT, **f32)          # gradient followed by the statistics: one all-reduce span
        self.scratch = torch.empty(int(self.lib.ppo_train_scratch_floats(ctypes.c_int32(self.D))), **f32)
        self._lr2 = torch.full((2,), lr, **f32)             # [0] current, [1] scratch of the Adam kernel
        self._step2 = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.lr, self.step = self._lr2[:1], self._step2[:1]
        self.obs_rms = RunningMeanStd(self.D, self.device)
        self.val_rms = RunningMeanStd(1, self.device)
        self.seed = int(seed)
        self.sample_counter = 0
        # device-side part of the sampling counter (advanced inside captured rollout graphs, see A2CAgent)
        self.counter_offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._counter_offset_host = 0
        # tcgen05/TMEM kernels (TF32) when the obs fit their padded K tile; the fp32 SIMT kernels are the numerics reference
        self.tensor_cores = bool(tensor_cores) and self.D <= 47     # padded K of the first GEMM: 16 (D <= 15) or 48 (D <= 47)
        self._ws = None
        self.packed = torch.zeros(int(self.lib.ppo_packed_weight_floats()), **f32) if self.tensor_cores else None
        self._packed_dirty = True
        self.loss_params = _lib.PpoLossParams(e_clip, critic_coef, entropy_coef, bounds_loss_coef, 1.1, int(clip_value))
        self.adam_params = _lib.PpoAdamParams(0.9, 0.999, 1e-8, grad_norm, 1.0 / world_size, int(adaptive_lr), kl_threshold, 1e-6, 1e-2)
        self.reset_parameters(seed)

    # ---- parameters -------------------------------------------------------------------------
    def views(self) -> Dict[str, torch.Tensor]:
        out, off = {}, 0
        for name, shp in zip(_NAMES, param_shapes(self.D)):
            n = math.prod(shp)
            out[name] = self.params[off:off + n].view(shp)
            off += n
        return out

    def reset_parameters(self, seed: int) -> None:
        """nn.Linear default init (kaiming-uniform a=sqrt(5) -> U(+-1/sqrt(fan_in))) with ZERO biases, logstd = 0
        [ref: RLG/algos_torch/network_builder.py:1562-1568 ; sigma_init const 0]."""
        g = torch.Generator().manual_seed(int(seed))
        v = self.views()
        for name, t in v.items():
            if name.endswith("weight"):
                bound = 1.0 / math.sqrt(t.shape[1])
                t.copy_(((torch.rand(t.shape, generator=g) * 2 - 1) * bound).to(self.device))
            else:
                t.zero_()
        self._packed_dirty = True

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """rl_games 'model' state_dict (same keys / dtypes as the reference checkpoints)."""
        sd = {"value_mean_std.running_mean": self.val_rms.mean.clone(), "value_mean_std.running_var": self.val_rms.var.clone(),
              "value_mean_std.count": self.val_rms.count.clone(),
              "running_mean_std.running_mean_std.state.running_mean": self.obs_rms.mean.clone(),
              "running_mean_std.running_mean_std.state.running_var": self.obs_rms.var.clone(),
              "running_mean_std.running_mean_std.state.count": self.obs_rms.count.clone()}
        sd.update({k: v.clone() for k, v in self.views().items()})
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k, v in self.views().items():
            v.copy_(sd[k].to(self.device))
        self._packed_dirty = True
        self.val_rms.load(sd["value_mean_std.running_mean"], sd["value_mean_std.running_var"], sd["value_mean_std.count"])
        self.obs_rms.load(sd["running_mean_std.running_mean_std.state.running_mean"],
                          sd["running_mean_std.running_mean_std.state.running_var"],
                          sd["running_mean_std.running_mean_std.state.count"])

    # ---- inference (is_train=False) ------------------------------------------------------------
    def act(self, obs: torch.Tensor, out: Optional[dict] = None, row_offset: int = 0) -> dict:
        """get_action_values: samples a ~ N(mu, sigma); returns actions, neglogpacs, values (de-normalised), mus, sigmas."""
        M = obs.shape[0]
        f32 = dict(dtype=torch.float32, device=self.device)
        o = out or dict(actions=torch.empty((M, A), **f32), neglogpacs=torch.empty(M, **f32), values=torch.empty((M, 1), **f32),
                        mus=torch.empty((M, A), **f32), sigmas=torch.empty((M, A), **f32))
        self._forward(obs, o["actions"], o["neglogpacs"], o["values"], o["mus"], o["sigmas"], row_offset)
        self.sample_counter += 1
        return o

    def advance_counter_offset(self, calls: int) -> None:
        """Inside a CUDA-graph capture containing `calls` act() launches: the next replay draws fresh Philox samples."""
        self.counter_offset += calls

    def note_graph_replay(self, calls: int) -> None:
        self.sample_counter += calls
        self._counter_offset_host += calls

    def values(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """get_values: de-normalised value of `obs`  [ref: RLG/common/a2c_common.py:407-437]."""
        v = out if out is not None else torch.empty((obs.shape[0], 1), dtype=torch.float32, device=self.device)
        self._forward(obs, None, None, v, None, None, 0)
        return v

    def pack(self) -> None:
        """Re-arranges the parameters into the tensor-core kernels' operand tiles (after every parameter change)."""
        if self.tensor_cores:
            _lib.check(self.lib.ppo_pack_weights_tc(_lib.ptr(self.params), ctypes.c_int32(self.D), _lib.ptr(self.packed), _lib.stream()),
                       "ppo_pack_weights_tc")
        self._packed_dirty = False

    def _forward(self, obs, actions, neglogp, values, mus, sigmas, row_offset):
        if self._packed_dirty:
            self.pack()
        fn = self.lib.ppo_policy_forward_tc if self.tensor_cores else self.lib.ppo_policy_forward_f32
        extra = (_lib.ptr(self.packed),) if self.tensor_cores else ()
        rc = fn(
            _lib.ptr(self.params), *extra, _lib.ptr(obs, torch.float32), ctypes.c_int32(self.D), _lib.ptr(self.obs_rms.mean32),
            _lib.ptr(self.obs_rms.var32), _lib.ptr(self.val_rms.mean32), _lib.ptr(self.val_rms.var32), ctypes.c_uint64(self.seed),
            ctypes.c_uint64(self.sample_counter - self._counter_offset_host), _lib.ptr(self.counter_offset), ctypes.c_int64(row_offset),
            _lib.ptr(actions), _lib.ptr(neglogp), _lib.ptr(values),
            _lib.ptr(mus), _lib.ptr(sigmas), ctypes.c_int64(obs.shape[0]), _lib.stream())
        _lib.check(rc, "ppo_policy_forward_f32")

    # ---- training ------------------------------------------------------------------------------
    def minibatch_grad(self, obs, actions, old_neglogp, advantages, old_values, returns, mu, sigma) -> torch.Tensor:
        """calc_gradients up to backward(): fills self.grads = [dLoss/dparams | stats]; mu/sigma are updated in place
        (dataset.update_mu_sigma)."""
        M = obs.shape[0]
        common = (_lib.ptr(self.params), _lib.ptr(obs, torch.float32), ctypes.c_int32(self.D), _lib.ptr(self.obs_rms.mean32),
                  _lib.ptr(self.obs_rms.var32), _lib.ptr(actions), _lib.ptr(old_neglogp), _lib.ptr(advantages), _lib.ptr(old_values),
                  _lib.ptr(returns), _lib.ptr(mu), _lib.ptr(sigma), ctypes.byref(self.loss_params), _lib.ptr(self.grads),
                  _lib.ptr(self.scratch))
        if self.tensor_cores:
            need = int(self.lib.ppo_train_tc_workspace_floats(ctypes.c_int64(M)))
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.float32, device=self.device)
            if self._packed_dirty:
                self.pack()
            rc = self.lib.ppo_minibatch_grad_tc(common[0], _lib.ptr(self.packed), *common[1:], _lib.ptr(self._ws), ctypes.c_int64(M),
                                                _lib.stream())
        else:
            rc = self.lib.ppo_minibatch_grad_f32(*common, ctypes.c_int64(M), _lib.stream())
        _lib.check(rc, "ppo_minibatch_grad_f32")
        return self.grads

    def minibatch_step(self, obs, actions, old_neglogp, advantages, old_values, returns, mu, sigma, peer=None) -> None:
        """minibatch_grad + (gradient all-reduce) + optimizer_step + pack on the tensor-core path: T1, T2 and ONE cooperative kernel that
        reduces the partial gradients, exchanges them with the peer ranks over NVLink (`peer`: a rl.peer.PeerStepExchange), clips,
        applies Adam / the adaptive lr and re-packs the operand tiles  [ref: RLG/common/a2c_common.py:308-330]."""
        M = obs.shape[0]
        need = int(self.lib.ppo_train_tc_workspace_floats(ctypes.c_int64(M)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float32, device=self.device)
        if self._packed_dirty:
            self.pack()
        args = (_lib.ptr(self.params), _lib.ptr(self.packed), _lib.ptr(obs, torch.float32), ctypes.c_int32(self.D), _lib.ptr(self.obs_rms.mean32),
                _lib.ptr(self.obs_rms.var32), _lib.ptr(actions), _lib.ptr(old_neglogp), _lib.ptr(advantages), _lib.ptr(old_values),
                _lib.ptr(returns), _lib.ptr(mu), _lib.ptr(sigma), ctypes.byref(self.loss_params), _lib.ptr(self.grads), _lib.ptr(self.scratch),
                _lib.ptr(self._ws), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self._lr2), _lib.ptr(self._step2),
                ctypes.byref(self.adam_params))
        if peer is not None and peer.world > 1:
            rc = self.lib.ppo_minibatch_step_peer_tc(*args, ctypes.byref(peer.comm), _lib.ptr(peer.seq), _lib.ptr(peer.err),
                                                     ctypes.c_int64(M), _lib.stream())
        else:
            rc = self.lib.ppo_minibatch_step_tc(*args, ctypes.c_int64(M), _lib.stream())
        _lib.check(rc, "ppo_minibatch_step_tc")
        self._packed_dirty = False          # the fused tail re-packed the tiles from the updated parameters

    def optimizer_step(self) -> None:
        """trancate_gradients_and_step (+ the adaptive-KL lr update), on device."""
        rc = self.lib.ppo_adam_step_f32(_lib.ptr(self.params), _lib.ptr(self.grads), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                        _lib.ptr(self._lr2), _lib.ptr(self._step2), ctypes.c_int64(self.P), ctypes.byref(self.adam_params),
                                        _lib.stream())
        _lib.check(rc, "ppo_adam_step_f32")
        self.pack()          # same stream, right behind the update: the packed tiles are never stale

    def stats(self) -> Dict[str, float]:
        s = self.grads[self.P:].tolist()
        return {k: s[i] for k, i in STAT.items()}
