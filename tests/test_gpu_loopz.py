"""GPU parity of the loopz PPO kernels (csrc/ppo_loopz.cu, through the C ABI) against goldens produced by the reference's own
OIGE/algo/ppo classes and against the CPU oracle on other seeds / sizes / the 4-wide privileged tail."""
import dataclasses
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.algo.ppo import PPO, module as M  # noqa: E402
from oracle import loopz_oracle as Z  # noqa: E402

DEV = "cuda:0"
T = torch.from_numpy


def build(D=33, md=8, n=48, horizon=8, sampling="in_order", graph=False, seed=0, **kw):
    arch = dict(speed_dim=3, mass_dim=md, mass_latent_dim=8, mass_encoder_shape=[64, 16])
    actor = M.Actor(M.MLPEncode_wrap([128, 128], "LeakyReLU", D, 2, "Tanh", False, **arch),
                    M.SquashedGaussianDiagonalCovariance(2, 0.3, action_scale=1.0), DEV, seed=seed)
    critic = M.Critic(M.MLPEncode_wrap([128, 128], "LeakyReLU", D, 1, **arch), DEV)
    args = dict(num_learning_epochs=4, gamma=0.997, lam=0.95, num_mini_batches=4, learning_rate=5e-4)
    args.update(kw)
    ppo = PPO(actor=actor, critic=critic, num_envs=n, num_transitions_per_env=horizon, device=DEV, mini_batch_sampling=sampling,
              use_cuda_graph=graph, **args)
    return actor, critic, ppo


def fill_from_golden(ppo, G):
    st = ppo.storage
    obs = T(G["roll_obs"]).to(DEV)
    st.actor_obs.copy_(obs[:-1]); st.critic_obs.copy_(obs[:-1])
    st.actions.copy_(T(G["roll_actions"])); st.actions_log_prob.copy_(T(G["roll_log_prob"])); st.values.copy_(T(G["roll_values"]))
    st.rewards.copy_(T(G["roll_rewards"]).unsqueeze(-1)); st.dones.copy_(T(G["roll_dones"]).unsqueeze(-1))
    st.step = st.num_transitions_per_env


def test_inference_and_evaluate_vs_reference(golden):
    G = golden("loopz_ppo")
    actor, critic, ppo = build()
    ppo.params.copy_(T(G["params0"]))
    obs = T(G["inf_obs"]).to(DEV)
    means = actor.noiseless_action(obs)
    values = critic.predict(obs)
    # fp32 forward of a 5-layer net vs torch CPU: 1e-5 relative (north_star), atol for outputs near zero
    assert torch.allclose(means.cpu(), T(G["inf_means"]), rtol=1e-5, atol=2e-6)
    assert torch.allclose(values.cpu(), T(G["inf_values"]), rtol=1e-5, atol=2e-6)
    (lp, ent), m2 = actor.evaluate(obs, T(G["eval_actions"]).to(DEV))
    assert torch.equal(m2, means)
    assert torch.allclose(lp.cpu(), T(G["eval_logp"]), rtol=1e-5, atol=2e-5) and torch.equal(ent, -lp)
    # state_dict uses the reference's key names and round-trips
    sd = actor.architecture.state_dict()
    assert list(sd)[0] == "architecture.mass_encoder.0.weight" and sd["architecture.action_mlp.4.weight"].shape == (2, 128)
    a2, c2, p2 = build(seed=5)
    p2.load_state_dict(ppo.state_dict(update=7))
    assert torch.equal(p2.params, ppo.params)


def test_four_wide_privileged_tail_vs_reference(golden):
    """mass_dim = 4 / obs 29: kernels vs the reference's own forward, evaluate and raw minibatch gradient (IN = 33 > obs_dim: the
    latent spills past the observation width in the staged tile)."""
    G = golden("loopz_ppo_md4")
    actor, critic, ppo = build(D=29, md=4, n=40, horizon=4, max_grad_norm=1e9, learning_rate=0.0, num_learning_epochs=1, num_mini_batches=1)
    ppo.params.copy_(T(G["params0"]))
    obs = T(G["st_actor_obs"]).to(DEV)
    flat = obs.view(-1, 29)
    assert torch.allclose(actor.noiseless_action(flat).cpu(), T(G["means"]), rtol=1e-5, atol=2e-6)
    assert torch.allclose(critic.predict(flat).cpu(), T(G["values"]), rtol=1e-5, atol=2e-6)
    (lp, _), _ = actor.evaluate(flat, T(G["st_actions"]).to(DEV).view(-1, 2))
    assert torch.allclose(lp.cpu(), T(G["eval_logp"]), rtol=1e-5, atol=2e-5)
    st = ppo.storage
    st.actor_obs.copy_(obs); st.critic_obs.copy_(obs)
    for k in ("actions", "actions_log_prob", "values", "returns", "advantages"):
        getattr(st, k).copy_(T(G["st_" + k]))
    ppo._minibatch(0, 160)
    g, want = ppo.grads[:ppo.P].cpu(), T(G["grad"])
    assert torch.allclose(g, want, rtol=1e-4, atol=1e-5 * float(want.abs().max())), float((g - want).abs().max())
    s_ = ppo.minibatch_statistics()
    assert abs(s_["value_loss"] - float(G["value_loss"])) < 1e-5 and abs(s_["surrogate"] - float(G["surrogate"])) < 1e-5


def test_sampling_is_a_squashed_gaussian(golden):
    G = golden("loopz_ppo")
    actor, critic, ppo = build()
    ppo.params.copy_(T(G["params0"]))
    obs = T(G["inf_obs"]).to(DEV).repeat(200, 1)                 # 40 000 rows
    acts, logp = actor.sample(obs)
    acts2, _ = actor.sample(obs)
    assert not torch.equal(acts, acts2)                          # the Philox counter advances
    assert float(acts.abs().max()) <= 1.0
    means = actor.noiseless_action(obs)
    std = actor.distribution.std
    u = torch.atanh(acts.double().clamp(-1 + 1e-12, 1 - 1e-12)).float()
    z = (u - means) / std
    ok = acts.abs().max(dim=1).values < 0.999                    # atanh is ill-conditioned at saturation
    assert abs(float(z[ok].mean())) < 0.02 and abs(float(z[ok].std()) - 1.0) < 0.02
    # log-prob of the sample == evaluate() of the squashed action (same formula through atanh)
    (lp, _), _ = actor.evaluate(obs, acts)
    assert torch.allclose(lp[ok], logp[ok], rtol=1e-3, atol=2e-3)
    # and == the oracle's log-prob for the recovered pre-squash sample
    want = Z.log_prob_from_u(means.cpu(), std.cpu(), u.cpu(), Z.LoopzCfg())
    assert torch.allclose(logp.cpu()[ok.cpu()], want[ok.cpu()], rtol=1e-3, atol=2e-3)


def test_returns_vs_reference(golden):
    G = golden("loopz_ppo")
    _, _, ppo = build()
    fill_from_golden(ppo, G)
    ppo.storage.compute_returns(T(G["roll_last_values"]).to(DEV), float(G["gamma"]), float(G["lam"]))
    assert torch.equal(ppo.storage.returns.cpu(), T(G["roll_returns"]))                       # op-by-op like torch: bit-exact
    assert torch.allclose(ppo.storage.advantages.cpu(), T(G["roll_advantages"]), rtol=1e-5, atol=1e-6)
    assert torch.isfinite(ppo.storage.returns).all()


def test_returns_large_vs_oracle():
    g = torch.Generator().manual_seed(5)
    Tn, n = 16, 5003
    _, _, ppo = build(n=n, horizon=Tn)
    st = ppo.storage
    rew, val = torch.randn((Tn, n, 1), generator=g), torch.randn((Tn, n, 1), generator=g)
    rew[3, 7] = float("inf"); val[5, 9] = float("nan")
    dn = (torch.rand((Tn, n, 1), generator=g) < 0.1).to(torch.uint8)
    lv = torch.randn((n, 1), generator=g)
    st.rewards.copy_(rew); st.values.copy_(val); st.dones.copy_(dn)
    st.compute_returns(lv.to(DEV), 0.997, 0.95)
    ret, adv = Z.compute_returns(rew, val, dn, lv, 0.997, 0.95)
    assert torch.equal(st.returns.cpu(), ret)
    assert torch.allclose(st.advantages.cpu(), adv, rtol=1e-5, atol=1e-5)


def test_minibatch_gradient_vs_reference(golden):
    G = golden("loopz_ppo")
    _, _, ppo = build(max_grad_norm=1e9, learning_rate=0.0)
    ppo.params.copy_(T(G["params0"]))
    fill_from_golden(ppo, G)
    ppo.storage.returns.copy_(T(G["roll_returns"])); ppo.storage.advantages.copy_(T(G["roll_advantages"]))
    for tag, hi in (("full", 384), ("quarter", 96)):
        ppo._minibatch(0, hi)
        g = ppo.grads[:ppo.P].cpu()
        want = T(G[f"grad_{tag}"])
        # fp32 sums over <= 384 samples in a different order than autograd: 1e-4 relative + 1e-5 of the largest entry
        assert torch.allclose(g, want, rtol=1e-4, atol=1e-5 * float(want.abs().max())), float((g - want).abs().max())
        st = ppo.minibatch_statistics()
        assert abs(st["value_loss"] - float(G[f"grad_{tag}_value_loss"])) < 1e-5
        assert abs(st["surrogate"] - float(G[f"grad_{tag}_surrogate"])) < 1e-5
        assert st["skipped"] == 0.0 and st["grad_norm"] == pytest.approx(float(want.norm()), rel=1e-4)
    assert torch.equal(ppo.params.cpu(), T(G["params0"]))          # lr = 0


@pytest.mark.parametrize("graph", [False, True])
def test_full_update_vs_reference(golden, graph):
    """PPO.update semantics: 4 epochs x 4 in-order minibatches, clip 0.5, Adam 5e-4 -- parameters after 16 optimiser steps."""
    G = golden("loopz_ppo")
    actor, _, ppo = build(graph=graph)
    ppo.params.copy_(T(G["params0"]))
    fill_from_golden(ppo, G)
    ppo.storage.compute_returns(T(G["roll_last_values"]).to(DEV), float(G["gamma"]), float(G["lam"]))
    vl, sl, info = ppo._train_step()
    want = T(G["update_params_after"])
    err = float((ppo.params.cpu() - want).abs().max())
    assert err < 2e-5, err                                          # 16 Adam steps of 5e-4 each; the reference moved by > 1e-3
    assert abs(vl - float(G["update_value_loss"])) < 1e-4 and abs(sl - float(G["update_surrogate"])) < 1e-4
    assert info["num_valid_updates"] == 16 and int(ppo.adam_step[ppo._parity]) == 16
    actor.distribution.enforce_minimum_std(torch.tensor([0.05, 0.5]))
    assert torch.allclose(actor.distribution.std.cpu(), T(G["min_std_after"]), atol=2e-5)


@pytest.mark.parametrize("md,D", [(4, 29), (8, 33), (8, 40)])
def test_gradient_vs_oracle_other_shapes(md, D):
    """Other privileged-tail widths / obs sizes, 1500 rows (ragged last tile), saturated actions, shuffled index path."""
    g = torch.Generator().manual_seed(md * 100 + D)
    n, Tn = 375, 4
    _, _, ppo = build(D=D, md=md, n=n, horizon=Tn, max_grad_norm=1e9, learning_rate=0.0)
    cfg = dataclasses.replace(Z.LoopzCfg(), obs_dim=D, mass_dim=md)
    flat = ppo.params.cpu().clone()
    flat += 0.03 * torch.randn(flat.shape, generator=g)
    ppo.params.copy_(flat)
    st = ppo.storage
    B = n * Tn
    obs = torch.randn((Tn, n, D), generator=g)
    act = torch.tanh(torch.randn((Tn, n, 2), generator=g) * 1.5)
    act[0, 0] = torch.tensor([1.0, -1.0])
    cols = {"actor_obs": obs, "critic_obs": obs * 0.5 + 0.1, "actions": act, "values": torch.randn((Tn, n, 1), generator=g),
            "advantages": torch.randn((Tn, n, 1), generator=g), "returns": torch.randn((Tn, n, 1), generator=g),
            "actions_log_prob": torch.randn((Tn, n, 1), generator=g) * 0.3 - 1.0}
    for k, v in cols.items():
        getattr(st, k).copy_(v)
    rows = [cols[k].reshape(B, -1) for k in ("actor_obs", "critic_obs", "actions", "values", "advantages", "returns", "actions_log_prob")]
    want, loss, surr, vloss = Z.minibatch_grad(flat, cfg, *rows)
    ppo._minibatch(0, B)
    got = ppo.grads[:ppo.P].cpu()
    assert torch.allclose(got, want, rtol=2e-4, atol=2e-5 * float(want.abs().max())), float((got - want).abs().max())
    s = ppo.minibatch_statistics()
    assert s["loss"] == pytest.approx(loss, rel=1e-4, abs=1e-5) and s["value_loss"] == pytest.approx(vloss, rel=1e-4)
    # shuffled minibatch: an index list into the storage
    idx = torch.randperm(B, generator=g)[:701]
    want_i, *_ = Z.minibatch_grad(flat, cfg, *[r[idx] for r in rows])
    ppo._minibatch(0, 0, idx.to(DEV))
    got_i = ppo.grads[:ppo.P].cpu()
    assert torch.allclose(got_i, want_i, rtol=2e-4, atol=2e-5 * float(want_i.abs().max()))


def _fill_random(ppo, g, n, Tn, D):
    obs = torch.randn((Tn, n, D), generator=g)
    act = torch.tanh(torch.randn((Tn, n, 2), generator=g) * 1.2)
    cols = {"actor_obs": obs, "critic_obs": obs, "actions": act, "values": torch.randn((Tn, n, 1), generator=g),
            "advantages": torch.randn((Tn, n, 1), generator=g), "returns": torch.randn((Tn, n, 1), generator=g),
            "actions_log_prob": torch.randn((Tn, n, 1), generator=g) * 0.3 - 1.0}
    for k, v in cols.items():
        getattr(ppo.storage, k).copy_(v)


@pytest.mark.parametrize("n,Tn", [(48, 8), (1500, 4), (4096, 16)])
def test_tensor_core_gradient_vs_fp32_kernels(n, Tn):
    """tcgen05 (TF32) minibatch gradient vs the fp32 SIMT kernels on the same minibatch: whole gradient at cosine > 0.998, statistics at
    5e-3, and every weight matrix no further from the fp32 result than 3x what the fp32 kernels THEMSELVES move when only their weights
    are truncated to TF32 (+ 5 %).  That yardstick matters: LeakyReLU's kink makes the gradient discontinuous, so a pre-activation
    within rounding noise of zero flips a whole term, and with cancelling per-sample terms (the critic's) even weight truncation
    alone moves a matrix by 10-40 % in relative L2 -- a fixed tolerance would measure the data, not the kernels."""
    torch.manual_seed(n)
    g = torch.Generator().manual_seed(n)
    kw = dict(n=n, horizon=Tn, max_grad_norm=1e9, learning_rate=0.0)
    (_, _, ref), (_, _, trunc), (_, _, tc) = build(**kw), build(**kw), build(tensor_cores=True, **kw)
    flat = ref.params.cpu() + 0.03 * torch.randn(ref.P, generator=g)
    ref.params.copy_(flat); tc.params.copy_(flat)
    ft = flat.clone().view(torch.int32)
    ft &= ~0x1fff                                             # 10-bit mantissa
    trunc.params.copy_(ft.view(torch.float32))
    _fill_random(ref, g, n, Tn, 33)
    # a rollout-like minibatch: old log-probs / values are the current networks' own outputs plus a little drift, so that ratios and
    # value deltas sit inside the clip ranges (the surrogate's gradient is discontinuous at the clip boundaries as well)
    st = ref.storage
    flat_obs = st.actor_obs.view(-1, 33)
    (lp0, _), _ = ref.actor.evaluate(flat_obs, st.actions.view(-1, 2))
    st.actions_log_prob.copy_((lp0 + 0.03 * torch.randn(lp0.shape, generator=g).to(DEV)).view(Tn, n, 1))
    v0 = ref.critic.predict(flat_obs)
    st.values.copy_((v0 + 0.03 * torch.randn(v0.shape, generator=g).to(DEV)).view(Tn, n, 1))
    st.returns.copy_(st.values + 0.5 * torch.randn(st.values.shape, generator=g).to(DEV))
    for other in (tc, trunc):
        for k in ("actor_obs", "critic_obs", "actions", "values", "advantages", "returns", "actions_log_prob"):
            getattr(other.storage, k).copy_(getattr(ref.storage, k))
    B = n * Tn
    sizes = [int(np.prod(s)) for out in (2, 1) for s in Z.net_shapes(Z.LoopzCfg(), out)]
    sizes = sizes[:12] + [2] + sizes[12:]
    for lo, hi in ((0, B), (B // 4, B // 2)):
        for learner in (ref, tc, trunc):
            learner._minibatch(lo, hi)
        a, b, c = ref.grads[:ref.P].cpu(), tc.grads[:tc.P].cpu(), trunc.grads[:ref.P].cpu()
        assert torch.isfinite(b).all()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        assert cos > 0.998, cos
        off = 0
        for i, sz in enumerate(sizes):
            if sz >= 1024:
                ea, eb, ec = a[off:off + sz], b[off:off + sz], c[off:off + sz]
                err_tc, err_tr = float((ea - eb).norm() / ea.norm()), float((ea - ec).norm() / ea.norm())
                assert err_tc <= 3.0 * err_tr + 0.05, (i, err_tc, err_tr)
            off += sz
        sa, sb = ref.minibatch_statistics(), tc.minibatch_statistics()
        for k in ("surrogate", "value_loss", "log_prob", "loss"):
            assert sb[k] == pytest.approx(sa[k], rel=5e-3, abs=5e-3), (k, sa[k], sb[k])


def test_tensor_core_update_tracks_fp32_update(golden):
    """A full 16-step update on the tensor-core path stays close to the fp32 one (reference golden), eager == graph replay."""
    G = golden("loopz_ppo")
    out = []
    for graph in (False, True):
        _, _, ppo = build(graph=graph, tensor_cores=True)
        ppo.params.copy_(T(G["params0"]))
        fill_from_golden(ppo, G)
        ppo.storage.compute_returns(T(G["roll_last_values"]).to(DEV), float(G["gamma"]), float(G["lam"]))
        vl, sl, info = ppo._train_step()
        assert info["num_valid_updates"] == 16
        assert abs(vl - float(G["update_value_loss"])) < 5e-3 and abs(sl - float(G["update_surrogate"])) < 5e-3
        want = T(G["update_params_after"])
        # Adam normalises every step, so a parameter whose gradient is at the TF32 noise level can move by +-lr per step either way:
        # compare the MOVEMENT of the whole vector (direction and length), not the worst single parameter
        d_ref, d_tc = want - T(G["params0"]), ppo.params.cpu() - T(G["params0"])
        cos = float(torch.dot(d_ref, d_tc) / (d_ref.norm() * d_tc.norm()))
        rel = float((d_tc - d_ref).norm() / d_ref.norm())
        assert cos > 0.98 and rel < 0.2, (cos, rel)
        out.append(ppo.params.clone())
    assert torch.equal(out[0], out[1])


def test_nonfinite_loss_skips_the_step():
    _, _, ppo = build(n=64, horizon=4)
    g = torch.Generator().manual_seed(1)
    st = ppo.storage
    st.actor_obs.copy_(torch.randn(st.actor_obs.shape, generator=g)); st.critic_obs.copy_(st.actor_obs)
    st.actions.copy_(torch.rand(st.actions.shape, generator=g) - 0.5)
    st.advantages.fill_(float("nan"))
    p0, m0 = ppo.params.clone(), ppo.exp_avg.clone()
    ppo._minibatch(0, 256)
    assert ppo.minibatch_statistics()["skipped"] == 1.0
    assert torch.equal(ppo.params, p0) and torch.equal(ppo.exp_avg, m0) and int(ppo.adam_step[ppo._parity]) == 0
    st.advantages.normal_(generator=None)
    ppo._minibatch(0, 256)
    assert ppo.minibatch_statistics()["skipped"] == 0.0 and not torch.equal(ppo.params, p0) and int(ppo.adam_step[ppo._parity]) == 1


def test_shuffle_mode_runs_and_covers_batch():
    _, _, ppo = build(n=100, horizon=8, sampling="shuffle")
    parts = ppo.storage.shuffled_indices(4, ppo._gen)
    assert len(parts) == 4 and sorted(torch.cat(parts).tolist()) == list(range(800))
    g = torch.Generator().manual_seed(2)
    st = ppo.storage
    st.actor_obs.copy_(torch.randn(st.actor_obs.shape, generator=g)); st.critic_obs.copy_(st.actor_obs)
    st.actions.copy_(torch.rand(st.actions.shape, generator=g) - 0.5)
    st.rewards.copy_(torch.randn(st.rewards.shape, generator=g) * 0.1)
    st.compute_returns(torch.zeros(100, 1, device=DEV), 0.997, 0.95)
    p0 = ppo.params.clone()
    vl, sl, info = ppo._train_step()
    assert info["num_valid_updates"] == 16 and math.isfinite(vl) and math.isfinite(sl) and not torch.equal(p0, ppo.params)


def test_rollout_storage_surface():
    """RolloutStorage keeps the reference's surface: add_transitions from numpy or tensors, overflow assertion, both generators."""
    from omniisaacgymenvs_loop_b200.algo.ppo import RolloutStorage
    st = RolloutStorage(10, 3, [33], [33], [2], DEV)
    g = torch.Generator().manual_seed(0)
    for k in range(3):
        obs = torch.randn((10, 33), generator=g)
        args = (obs.numpy(), obs.numpy(), torch.rand((10, 2), generator=g), torch.randn(10, generator=g).numpy(),
                (torch.rand(10, generator=g) < 0.3).numpy(), torch.randn((10, 1), generator=g), torch.randn(10, generator=g))
        if k == 1:                                  # CUDA tensors are accepted as well
            args = tuple(torch.as_tensor(a).to(DEV) for a in args)
        st.add_transitions(*args)
    assert st.step == 3 and st.dones.dtype == torch.uint8 and st.actor_obs.shape == (3, 10, 33)
    with pytest.raises(AssertionError, match="overflow"):
        st.add_transitions(*args)
    st.compute_returns(torch.zeros((10, 1)), 0.99, 0.95)
    assert torch.isfinite(st.returns).all() and abs(float(st.advantages.mean())) < 1e-5
    ordered = list(st.mini_batch_generator_inorder(3))
    assert len(ordered) == 3 and ordered[0][0].shape == (10, 33) and torch.equal(ordered[1][0], st.actor_obs[1])   # time-major row blocks
    shuffled = list(st.mini_batch_generator_shuffle(3))
    assert len(shuffled) == 3 and all(b[2].shape == (10, 2) for b in shuffled)
    st.clear()
    assert st.step == 0


def test_numpy_boundary_and_fused_value_path():
    """The reference's host contract (numpy observations in, numpy actions out, numpy rewards / dones into step) and the device fast
    path (CUDA tensors, critic evaluated in observe()'s launch) fill the storage identically."""
    from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg
    from omniisaacgymenvs_loop_b200.envs.usv_raisim_vecenv import USVRaisimVecEnv
    from scripts.train_loopz import build_learner, make_env

    out = []
    for host in (True, False):
        torch.manual_seed(9)
        env = make_env(live_task_cfg(live_default_config(num_envs=256, max_episode_length=6)), DEV, seed=9)
        assert isinstance(env, USVRaisimVecEnv) and env.num_obs == 33 and env.num_acts == 2
        ppo = build_learner(env, DEV, 8, seed=9, use_cuda_graph=False)
        env.reset()
        for _ in range(8):
            obs = env.observe() if host else env.observe(as_numpy=False)
            action = ppo.observe(obs)
            reward, dones = env.step(action)
            if host:
                assert isinstance(obs, np.ndarray) and isinstance(action, np.ndarray) and action.shape == (256, 2)
                assert isinstance(reward, np.ndarray) and reward.dtype == np.float32 and dones.dtype == np.bool_
                assert env.get_reward_info().shape == (256, 16) and isinstance(env.get_extras(), dict)
            else:
                assert action.is_cuda and reward.is_cuda
            ppo.step(value_obs=obs, rews=reward * 0.01, dones=dones, infos=[])
        out.append(ppo.storage)
        ppo.update(actor_obs=env.observe(as_numpy=False), value_obs=env.observe(as_numpy=False), log_this_iteration=False, update=0)
        assert ppo.storage.step == 0 and torch.isfinite(ppo.params).all()
    a, b = out
    for k in ("actor_obs", "critic_obs", "actions", "actions_log_prob", "values", "rewards", "dones"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert int(a.dones.sum()) > 0


def test_training_loop_on_the_live_task_learns():
    """The loopz loop (scripts/train_loopz.py) on the live CaptureXY task: finite throughout, graph replay == eager, return improves."""
    from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg
    from scripts.train_loopz import build_learner, make_env, train

    def run(graph, updates):
        torch.manual_seed(3)                    # network initialisation draws from torch's global generator, like the reference
        env = make_env(live_task_cfg(live_default_config(num_envs=1024, max_episode_length=300)), DEV, seed=3)
        ppo = build_learner(env, DEV, 16, seed=3, use_cuda_graph=graph)
        hist = train(env, ppo, updates, 16, log_every=5, quiet=True, use_cuda_graph=graph)
        return ppo, hist

    pe, _ = run(False, 5)
    pg, _ = run(True, 5)                                             # cycles 0-1 eager, 2 captured, 3-4 replayed
    assert torch.equal(pe.params, pg.params)                       # the captured rollout + update is the eager loop, bit for bit
    assert torch.equal(pe.storage.actor_obs, pg.storage.actor_obs) and torch.equal(pe.storage.actions, pg.storage.actions)
    ppo, hist = run(True, 120)
    assert torch.isfinite(ppo.params).all() and float(ppo.actor.distribution.std.min()) >= 0.05
    rets = [h[1] for h in hist if h[1] == h[1]]
    assert len(rets) >= 6 and np.mean(rets[-3:]) > np.mean(rets[:3]), rets


def test_checkpoint_optimizer_state_interchanges_with_torch_adam(golden):
    """The .pt dictionary carries 'optimizer_state_dict' in torch.optim.Adam's own layout [ref: rlgames_train_loopz.py:855,1295]: what
    this learner writes loads into a real torch Adam over tensors of the reference's shapes, and what torch Adam writes loads back."""
    G = golden("loopz_ppo")
    actor, critic, ppo = build()
    ppo.params.copy_(T(G["params0"]))
    fill_from_golden(ppo, G)
    ppo.storage.compute_returns(critic.predict(T(G["roll_obs"]).to(DEV)[-1]), ppo.gamma, ppo.lam)
    ppo._train_step()
    ck = ppo.state_dict(update=3)
    assert "optimizer_state_dict" in ck and "optimizer_state" not in ck
    osd = ck["optimizer_state_dict"]
    shapes = [shp for shp, _ in ppo._optimizer_tensors()]
    assert len(osd["state"]) == len(shapes) == 25 and osd["param_groups"][0]["params"] == list(range(25))
    prm = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]           # [*actor.parameters(), *critic.parameters()] stand-ins
    opt = torch.optim.Adam(prm, lr=5e-4)
    opt.load_state_dict({"state": {k: {a: b.cpu() for a, b in v.items()} for k, v in osd["state"].items()}, "param_groups": osd["param_groups"]})
    got = torch.cat([opt.state[p]["exp_avg"].reshape(-1) for p in prm])
    assert torch.equal(got, ppo.exp_avg.cpu()) and float(opt.state[prm[0]]["step"]) == float(ppo.adam_step[ppo._parity])
    # and back: a torch-written state dict restores moments, step count and lr
    a2, c2, p2 = build(seed=5)
    back = opt.state_dict()
    back["param_groups"][0]["lr"] = 1.25e-4
    assert p2.load_state_dict(dict(ck, optimizer_state_dict=back)) == 4
    assert torch.equal(p2.exp_avg, ppo.exp_avg) and torch.equal(p2.exp_avg_sq, ppo.exp_avg_sq) and torch.equal(p2.params, ppo.params)
    assert int(p2.adam_step[0]) == int(ppo.adam_step[ppo._parity]) and abs(float(p2.lr) - 1.25e-4) < 1e-10
    # round-1 files of this repo (flat blobs) still load; a checkpoint with another action scale is refused, not ignored
    legacy = {k: v for k, v in ck.items() if k != "optimizer_state_dict"}
    legacy["optimizer_state"] = {"exp_avg": ppo.exp_avg.clone(), "exp_avg_sq": ppo.exp_avg_sq.clone(), "step": 16, "lr": 5e-4}
    a3, c3, p3 = build(seed=6)
    p3.load_state_dict(legacy)
    assert torch.equal(p3.exp_avg, ppo.exp_avg) and int(p3.adam_step[0]) == 16
    bad = dict(ck, actor_distribution_state_dict=dict(ck["actor_distribution_state_dict"], action_scale=torch.tensor([2.0, 2.0])))
    with pytest.raises(ValueError):
        p3.load_state_dict(bad)
