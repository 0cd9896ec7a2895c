"""CPU suite: the loopz-PPO oracle restatement vs goldens produced by the reference's own OIGE/algo/ppo classes
(oracle/make_golden.py:loopz), plus the host-side surface of the drop-in package."""
import numpy as np
import pytest
import torch

from oracle import loopz_oracle as Z

T = torch.from_numpy
CFG = Z.LoopzCfg()


def _storage(G):
    return {"actor_obs": T(G["roll_obs"][:-1]), "critic_obs": T(G["roll_obs"][:-1]), "actions": T(G["roll_actions"]),
            "values": T(G["roll_values"]), "advantages": T(G["roll_advantages"]), "returns": T(G["roll_returns"]),
            "actions_log_prob": T(G["roll_log_prob"])}


def test_param_layout_matches_reference_names(golden):
    G = golden("loopz_ppo")
    names = [str(n) for n in G["param_names"]]
    net = [f"architecture.{m}.{i}.{k}" for m in ("mass_encoder", "action_mlp") for i in (0, 2, 4) for k in ("weight", "bias")]
    assert names == net + ["distribution.std"] + net
    sizes = [int(np.prod(s)) for out in (2, 1) for s in Z.net_shapes(CFG, out)]
    assert sizes[:12] + [2] + sizes[12:] == [int(x) for x in G["param_sizes"]]
    assert G["params0"].shape[0] == 45621


def test_forward_sample_evaluate_vs_reference(golden):
    G = golden("loopz_ppo")
    pa, std, pc = Z.split(T(G["params0"]), CFG)
    obs = T(G["inf_obs"])
    means = Z.mlp_encode(pa, obs, CFG, True)
    values = Z.mlp_encode(pc, obs, CFG, False)
    assert torch.allclose(means, T(G["inf_means"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(values, T(G["inf_values"]), rtol=1e-6, atol=1e-6)
    acts, logp = Z.sample_from_noise(T(G["inf_means"]), std, T(G["inf_noise"]), CFG)
    assert torch.allclose(acts, T(G["inf_actions"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(logp, T(G["inf_logp_u"]), rtol=1e-6, atol=1e-6)
    lp, ent = Z.evaluate(T(G["eval_means"]), std, T(G["eval_actions"]), CFG)
    assert torch.allclose(lp, T(G["eval_logp"]), rtol=1e-6, atol=1e-6) and torch.allclose(ent, T(G["eval_entropy"]), rtol=1e-6, atol=1e-6)


def test_compute_returns_vs_reference(golden):
    G = golden("loopz_ppo")
    rewards = T(G["roll_rewards"]).unsqueeze(-1)
    ret, adv = Z.compute_returns(rewards, T(G["roll_values"]), T(G["roll_dones"]).unsqueeze(-1), T(G["roll_last_values"]),
                                 float(G["gamma"]), float(G["lam"]))
    assert torch.equal(ret, T(G["roll_returns"]))
    assert torch.equal(adv, T(G["roll_advantages"]))
    assert torch.isfinite(ret).all() and bool(torch.isnan(rewards).any())     # the NaN reward was sanitised


def test_minibatch_gradient_vs_reference(golden):
    G = golden("loopz_ppo")
    st = _storage(G)
    rows = lambda k: st[k].reshape(-1, st[k].shape[-1])
    cols = [rows(k) for k in ("actor_obs", "critic_obs", "actions", "values", "advantages", "returns", "actions_log_prob")]
    for tag, sel in (("full", slice(None)), ("quarter", slice(0, 96))):
        g, loss, surr, vloss = Z.minibatch_grad(T(G["params0"]), CFG, *[c[sel] for c in cols])
        want = T(G[f"grad_{tag}"])
        assert torch.allclose(g, want, rtol=1e-4, atol=1e-7), float((g - want).abs().max())
        assert abs(vloss - float(G[f"grad_{tag}_value_loss"])) < 1e-6 and abs(surr - float(G[f"grad_{tag}_surrogate"])) < 1e-6


def test_full_update_vs_reference(golden):
    """4 epochs x 4 in-order minibatches, grad-norm clip 0.5, Adam lr 5e-4: parameters after the update and the mean losses."""
    G = golden("loopz_ppo")
    flat = T(G["params0"]).clone()
    vl, sl = Z.train_step(flat, CFG, _storage(G), 4, 4)
    want = T(G["update_params_after"])
    assert torch.allclose(flat, want, rtol=0, atol=2e-6), float((flat - want).abs().max())
    assert float((flat - T(G["params0"])).abs().max()) > 1e-3           # the update moved the parameters
    assert abs(vl - float(G["update_value_loss"])) < 1e-5 and abs(sl - float(G["update_surrogate"])) < 1e-5
    _, std, _ = Z.split(flat, CFG)
    assert torch.allclose(Z.enforce_minimum_std(std, torch.tensor([0.05, 0.5])), T(G["min_std_after"]), atol=2e-6)


def test_host_surface_without_gpu():
    """The drop-in classes validate shapes like the reference and refuse to run without CUDA (no CPU fallback)."""
    from omniisaacgymenvs_loop_b200 import _lib
    from omniisaacgymenvs_loop_b200.algo.ppo import module as M

    w = M.MLPEncode_wrap([128, 128], "LeakyReLU", 33, 2, "Tanh", False, speed_dim=3, mass_dim=8, mass_latent_dim=8, mass_encoder_shape=[64, 16])
    assert w.input_shape == [33] and w.output_shape == [2]
    sd = w.state_dict()
    assert list(sd)[:2] == ["architecture.mass_encoder.0.weight", "architecture.mass_encoder.0.bias"] and len(sd) == 12
    assert sum(v.numel() for v in sd.values()) == 22874
    wt = sd["architecture.action_mlp.2.weight"]
    assert torch.allclose(wt @ wt.T, 2.0 * torch.eye(128), atol=1e-4)            # orthogonal_(gain=sqrt 2)
    with pytest.raises(ValueError):
        M.MLPEncode_wrap([128, 128], "LeakyReLU", 10, 2, None, False, speed_dim=3, mass_dim=8)
    with pytest.raises(NotImplementedError):
        M.MLPEncode_wrap([256, 128], "LeakyReLU", 33, 2)
    dist = M.SquashedGaussianDiagonalCovariance(2, 0.3, action_scale=1.0)
    assert torch.equal(dist.std, torch.tensor([0.3, 0.3])) and set(dist.state_dict()) == {"std", "action_scale"}
    with pytest.raises(_lib.UsvLibraryError):
        M.Actor(w, dist, "cpu")
    with pytest.raises(_lib.UsvLibraryError):
        M.Critic(M.MLPEncode_wrap([128, 128], "LeakyReLU", 33, 1, speed_dim=3, mass_dim=8), "cpu")


def test_loopz_abi_argument_checks_without_gpu():
    """The loopz entry points validate their arguments before touching the device: sizes, NULL pointers, unsupported shapes."""
    import ctypes
    from omniisaacgymenvs_loop_b200 import _lib

    L = _lib.lib()
    E = _lib.ENUMS
    Net = _lib.STRUCTS["PpoLoopzNet"]
    for f in ("ppo_loopz_param_count", "ppo_loopz_actor_param_count", "ppo_loopz_train_scratch_floats", "ppo_loopz_returns_scratch_bytes"):
        getattr(L, f).restype = ctypes.c_int64
    net = Net(33, 8, 1, 1.0, 1e-6)
    assert L.ppo_loopz_param_count(ctypes.byref(net)) == 45621 and L.ppo_loopz_actor_param_count(ctypes.byref(net)) == 22874
    assert L.ppo_loopz_param_count(ctypes.byref(Net(29, 4, 1, 1.0, 1e-6))) == 2 * (64 * 4 + 64 + 1040 + 136 + 33 * 128 + 128 + 16512) + 258 + 129 + 2
    assert L.ppo_loopz_param_count(ctypes.byref(Net(33, 9, 1, 1.0, 1e-6))) == -1          # mass_dim > 8
    assert L.ppo_loopz_param_count(ctypes.byref(Net(8, 8, 1, 1.0, 1e-6))) == -1           # no task columns left
    assert L.ppo_loopz_train_scratch_floats(ctypes.byref(net)) > 2 * 45621 and L.ppo_loopz_returns_scratch_bytes() > 0
    null, one = ctypes.c_void_p(0), ctypes.c_void_p(16)
    i64, i32, f32 = ctypes.c_int64, ctypes.c_int32, ctypes.c_float
    act = lambda M, params, obs, actions: L.ppo_loopz_act_f32(params, ctypes.byref(net), obs, null, ctypes.c_uint64(0), ctypes.c_uint64(0), null,
                                                              i64(0), null, actions, null, null, null, i64(M), null)
    assert act(0, null, null, null) == E["USV_OK"]                                         # empty batch: nothing to do
    assert act(-1, one, one, one) == E["USV_E_SIZE"]
    assert act(8, null, one, one) == E["USV_E_NULL"] and act(8, one, null, one) == E["USV_E_NULL"] and act(8, one, one, null) == E["USV_E_NULL"]
    ret = lambda T, n, p: L.ppo_loopz_returns_f32(p, p, p, p, f32(0.99), f32(0.95), p, p, p, i32(T), i64(n), null)
    assert ret(0, 5, null) == E["USV_OK"] and ret(16, 0, null) == E["USV_OK"] and ret(-1, 5, one) == E["USV_E_SIZE"] and ret(16, 5, null) == E["USV_E_NULL"]
    lp = _lib.STRUCTS["PpoLoopzLossParams"](0.2, 0.5, 0.0, 1)
    grad = lambda M, p: L.ppo_loopz_minibatch_grad_f32(p, ctypes.byref(net), p, p, p, p, p, p, p, null, ctypes.byref(lp), p, p, i64(M), null)
    assert grad(0, one) == E["USV_E_SIZE"] and grad(64, null) == E["USV_E_NULL"]
    assert L.ppo_loopz_enforce_min_std_f32(null, null, i32(2), null) == E["USV_E_NULL"]
    assert L.ppo_loopz_enforce_min_std_f32(one, one, i32(0), null) == E["USV_E_SIZE"]


def test_four_wide_privileged_tail_vs_reference(golden):
    """mass_dim = 4 / obs 29 (environment.mass_dim of the older configs): oracle forward, evaluate and minibatch gradient vs the reference."""
    import dataclasses
    G = golden("loopz_ppo_md4")
    cfg = dataclasses.replace(Z.LoopzCfg(), obs_dim=29, mass_dim=4)
    flat = T(G["params0"])
    assert flat.numel() == 2 * (64 * 4 + 64 + 1040 + 136 + 33 * 128 + 128 + 16512) + 258 + 129 + 2
    pa, std, pc = Z.split(flat, cfg)
    obs = T(G["st_actor_obs"]).reshape(-1, 29)
    assert torch.allclose(Z.mlp_encode(pa, obs, cfg, True), T(G["means"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(Z.mlp_encode(pc, obs, cfg, False), T(G["values"]), rtol=1e-6, atol=1e-6)
    lp, _ = Z.evaluate(T(G["means"]), std, T(G["st_actions"]).reshape(-1, 2), cfg)
    assert torch.allclose(lp, T(G["eval_logp"]), rtol=1e-6, atol=1e-6)
    rows = [T(G["st_" + k]).reshape(obs.shape[0], -1) for k in ("actor_obs", "actor_obs", "actions", "values", "advantages", "returns", "actions_log_prob")]
    g, loss, surr, vloss = Z.minibatch_grad(flat, cfg, *rows)
    want = T(G["grad"])
    assert torch.allclose(g, want, rtol=1e-4, atol=1e-7), float((g - want).abs().max())
    assert abs(vloss - float(G["value_loss"])) < 1e-6 and abs(surr - float(G["surrogate"])) < 1e-6
