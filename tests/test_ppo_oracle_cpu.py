"""CPU suite: the PPO oracle restatement vs goldens produced by the reference's own rl_games classes."""
import numpy as np
import torch

from oracle import ppo_oracle as P

T = torch.from_numpy
D = 13


def _rms(G):
    o = P.RunningMeanStd((D,)); o.mean, o.var, o.count = T(G["obs_mean"]), T(G["obs_var"]), T(G["obs_count"])
    v = P.RunningMeanStd((1,)); v.mean, v.var, v.count = T(G["val_mean"]), T(G["val_var"]), T(G["val_count"])
    return o, v


def test_param_layout_matches_reference_names(golden):
    G = golden("ppo")
    names = [str(n) for n in G["param_names"]]
    assert names == ["a2c_network.sigma", "a2c_network.actor_mlp.0.weight", "a2c_network.actor_mlp.0.bias",
                     "a2c_network.actor_mlp.2.weight", "a2c_network.actor_mlp.2.bias", "a2c_network.value.weight",
                     "a2c_network.value.bias", "a2c_network.mu.weight", "a2c_network.mu.bias"]
    assert P.param_layout(D)["P"] == 18693 == G["params0"].shape[0] and P.param_layout(33)["P"] == 21253


def test_running_mean_std_update(golden):
    G = golden("ppo")
    r = P.RunningMeanStd((D,))
    r.update(T(G["rms_update_obs"]))
    assert torch.equal(r.mean, T(G["obs_mean"])) and torch.equal(r.var, T(G["obs_var"])) and float(r.count) == float(G["obs_count"])
    assert r.mean.dtype == torch.float64


def test_inference_vs_reference(golden):
    G = golden("ppo")
    o, v = _rms(G)
    params = T(G["params0"])
    out = P.policy_inference(params, T(G["obs"]), D, o, v)
    assert torch.allclose(out["mus"], T(G["inf_mus"]), rtol=1e-6, atol=1e-6)
    assert torch.equal(out["sigmas"], T(G["inf_sigmas"]))
    assert torch.allclose(out["values"], T(G["inf_values"]), rtol=1e-6, atol=1e-6)
    mu, sg = T(G["inf_mus"]), T(G["inf_sigmas"])
    nlp = P.neglogp(T(G["inf_actions"]), mu, sg, torch.log(sg))
    assert torch.allclose(nlp, T(G["inf_neglogpacs"]), rtol=1e-5, atol=1e-5)


def test_minibatch_loss_grads_adam_vs_reference(golden):
    G = golden("ppo")
    o, _ = _rms(G)
    params = T(G["params0"]).clone()
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    batch = dict(obs=T(G["obs"]), actions=T(G["mb_actions"]), old_logp_actions=T(G["mb_old_neglogp"]), advantages=T(G["mb_adv"]),
                 old_values=T(G["mb_old_values"]), returns=T(G["mb_returns"]), mu=T(G["mb_old_mu"]), sigma=T(G["mb_old_sigma"]))
    lr = 3e-4
    for it in range(3):
        prm = params.clone().requires_grad_(True)
        loss, st = P.minibatch_loss(prm, batch, D, o)
        loss.backward()
        assert torch.allclose(loss.detach(), T(G[f"it{it}_loss"]), rtol=1e-5, atol=1e-6), it
        for k in ("a_loss", "c_loss", "entropy", "b_loss", "kl"):
            assert torch.allclose(st[k].detach(), T(G[f"it{it}_{k}"]), rtol=2e-5, atol=1e-6), (it, k)
        assert torch.allclose(prm.grad, T(G[f"it{it}_grads"]), rtol=1e-4, atol=1e-7), it
        assert abs(lr - float(G[f"it{it}_lr"])) < 1e-9
        params, m, v, norm = P.adam_step(params, prm.grad, m, v, it + 1, lr)
        assert torch.allclose(norm, T(G[f"it{it}_grad_norm"]), rtol=1e-5)
        assert torch.allclose(params, T(G[f"it{it}_params_after"]), rtol=1e-5, atol=1e-7), it
        lr = P.adaptive_lr(lr, float(st["kl"]))
        assert abs(lr - float(G[f"it{it}_new_lr"])) < 1e-9
        batch["mu"], batch["sigma"] = st["mu"], st["sigma"]


def test_gae_oracle_vs_reference(golden):
    G = golden("gae")
    for tag in "abcd":
        f = lambda k: T(G[f"{tag}_{k}"])
        adv = P.discount_values(f("last_dones").float(), f("last_values").squeeze(-1), f("dones").float(),
                                f("values").squeeze(-1), f("rewards").squeeze(-1))
        assert torch.equal(adv, f("adv").squeeze(-1)), tag


def test_normal_eps_statistics():
    e = P.normal_eps(1234, np.arange(200000), 7)
    assert abs(float(e.mean())) < 5e-3 and abs(float(e.std()) - 1.0) < 5e-3 and torch.isfinite(e).all()


def test_average_meter_matches_reference_rule():
    """A2CAgent.game_rewards / game_lengths: rl_games' AverageMeter arithmetic [RLG/algos_torch/torch_ext.py:281-307] with the
    window size kept on the device (no `dones.nonzero()` sync)."""
    import numpy as np
    import torch
    from omniisaacgymenvs_loop_b200.rl.a2c import AverageMeter
    meter, mean, cur = AverageMeter(100, "cpu"), torch.zeros(1), 0
    g = torch.Generator().manual_seed(0)
    for k in range(300):
        n = int(torch.randint(0, 70, (1,), generator=g))
        v = torch.randn(n, generator=g)
        meter.update(v.sum(), torch.tensor(float(n)))
        if n:                                        # the reference's update(), restated
            size = int(np.clip(n, 0, 100)); old = min(100 - size, cur); cur = old + size
            mean = (mean * old + v.mean() * size) / cur
        assert abs(meter.get_mean() - float(mean)) < 1e-5 and len(meter) == cur
    meter.clear()
    assert len(meter) == 0 and meter.get_mean() == 0.0


def epoch_golden_state(G):
    o = P.RunningMeanStd((D,)); o.mean, o.var, o.count = T(G["obs_mean0"]), T(G["obs_var0"]), T(G["obs_count0"])
    v = P.RunningMeanStd((1,)); v.mean, v.var, v.count = T(G["val_mean0"]), T(G["val_var0"]), T(G["val_count0"])
    return o, v


def test_epoch_prepare_dataset_and_train_epoch_vs_reference(golden):
    """Rows P2 / P3: the oracle's prepare_dataset + train_epoch against two whole epochs of the reference's own
    ContinuousA2CBase.train_epoch / PPODataset / calc_gradients (oracle/make_golden.py:ppo_epoch)."""
    G = golden("ppo_epoch")
    Dd, Th, NA, MB, ME = (int(x) for x in G["shape"])
    assert Dd == D
    o, vr = epoch_golden_state(G)
    params = T(G["params0"]).clone()
    m, v, step, lr = torch.zeros_like(params), torch.zeros_like(params), 0, 1e-4       # G["lr0"], stored as fp32
    for ep in range(2):
        roll = {k: T(G[f"ep{ep}_{k}"]) for k in ("obses", "actions", "neglogpacs", "values", "mus", "sigmas", "rewards", "dones")}
        adv = P.discount_values(T(G[f"ep{ep}_last_dones"]).float(), T(G[f"ep{ep}_last_values"])[:, 0], roll["dones"].float(),
                                roll["values"][..., 0], roll["rewards"][..., 0])
        assert torch.equal((adv + roll["values"][..., 0]).unsqueeze(-1), T(G[f"ep{ep}_returns"]))
        ds = P.prepare_dataset(roll, T(G[f"ep{ep}_returns"]), vr)
        for k in ("obs", "actions", "old_logp_actions"):
            assert torch.equal(ds[k], T(G[f"ep{ep}_ds_{k}"])), k                                   # env-major flatten, bit-exact
        for k in ("old_values", "returns", "advantages"):
            assert torch.allclose(ds[k], T(G[f"ep{ep}_ds_{k}"]), rtol=1e-6, atol=1e-6), k
        assert torch.allclose(vr.mean, T(G[f"ep{ep}_val_mean"]), rtol=1e-12) and torch.allclose(vr.var, T(G[f"ep{ep}_val_var"]), rtol=1e-12)
        assert float(vr.count) == float(G[f"ep{ep}_val_count"])
        trace = []
        params, m, v, step, lr = P.train_epoch(params, m, v, step, lr, ds, D, o, minibatch_size=MB, mini_epochs=ME, trace=trace)
        assert len(trace) == ME * (Th * NA // MB)
        for i, t in enumerate(trace):
            assert abs(t["lr"] - float(G[f"ep{ep}_mb_lr"][i])) <= 1e-6 * t["lr"], (ep, i)         # the adaptive schedule, step by step
            for k in ("a_loss", "c_loss", "kl"):
                assert torch.allclose(t[k], torch.tensor(G[f"ep{ep}_mb_{k}"][i]), rtol=2e-5, atol=1e-6), (ep, i, k)
            assert torch.allclose(t["mu"], T(G[f"ep{ep}_mb_mu"][i]), rtol=1e-5, atol=1e-5), (ep, i)
            assert torch.allclose(t["obs_mean"], T(G[f"ep{ep}_mb_obs_mean"][i]), rtol=1e-12) and float(t["obs_count"]) == float(G[f"ep{ep}_mb_obs_count"][i])
            if ep == 0:
                assert torch.allclose(t["params"], T(G["ep0_mb_params"][i]), rtol=1e-5, atol=1e-6), i
        # the obs normaliser only moved during mini-epoch 0 (count: + one minibatch per minibatch of that mini-epoch)
        assert float(o.count) == float(G[f"ep{ep}_obs_count"]) and torch.allclose(o.var, T(G[f"ep{ep}_obs_var"]), rtol=1e-12)
        assert torch.allclose(params, T(G[f"ep{ep}_params"]), rtol=1e-5, atol=1e-6)
        assert torch.allclose(m, T(G[f"ep{ep}_exp_avg"]), rtol=1e-4, atol=1e-7) and torch.allclose(v, T(G[f"ep{ep}_exp_avg_sq"]), rtol=1e-4, atol=1e-10)
        assert abs(lr - float(G[f"ep{ep}_last_lr"])) <= 1e-6 * lr
        assert torch.allclose(ds["mu"], T(G[f"ep{ep}_ds_mu_final"]), rtol=1e-5, atol=1e-5)           # update_mu_sigma write-back
        assert torch.equal(ds["sigma"], T(G[f"ep{ep}_ds_sigma_final"]))
