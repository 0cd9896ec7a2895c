"""GPU tests of the reference-shaped surfaces: VecEnvRLGames / USVVirtual / RLGPUEnv over the fused step, and the PPO loop."""
import dataclasses
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig  # noqa: E402
from omniisaacgymenvs_loop_b200.rl.a2c import A2CAgent, PPOConfig  # noqa: E402
from oracle import usv_oracle as O  # noqa: E402
from scripts.train_usv import make_env  # noqa: E402
from tests.util import assert_close, oracle_cfg  # noqa: E402

DEV = "cuda:0"


def test_vecenv_step_matches_oracle_and_contract():
    cfg = UsvEnvConfig(num_envs=600, max_episode_length=9, seed=77).full_dr()
    env = make_env(cfg.to_task_cfg(), DEV, seed=77)
    info = env.get_env_info()
    assert info["observation_space"]["state"].shape == (13,) and info["action_space"].shape == (2,)
    assert env.get_number_of_agents() == 1
    orc = O.ClassicEnvOracle(oracle_cfg(cfg), 600)
    obs = env.reset()
    o_obs, _, _ = orc.step(torch.zeros((600, 2)))                    # VecEnvRLGames.reset == one zero-action step
    assert set(obs.keys()) == {"obs", "states"} and obs["obs"]["state"].shape == (600, 13) and obs["states"].shape == (600, 0)
    assert_close(obs["obs"]["state"], o_obs, 1e-5, 2e-5, "reset obs")
    g = torch.Generator().manual_seed(0)
    for k in range(20):
        act = torch.rand((600, 2), generator=g) * 3 - 1.5              # beyond clipActions: the vec-env clamps
        od, rew, done, extras = env.step(act.to(DEV))
        o_obs, o_rew, o_done = orc.step(act)
        assert rew.shape == (600,) and done.dtype == torch.int64 and rew.dtype == torch.float32
        assert_close(od["obs"]["state"], o_obs, 1e-4, 2e-3, f"obs {k}"); assert_close(rew, o_rew, 1e-4, 2e-3, f"rew {k}")
        assert torch.equal(done.cpu(), o_done), k
    ep = extras["episode"]
    assert {"distance_reward", "alignment_reward", "position_error", "energy_penalty", "angular_vel_variation_penalty", "actions_sum"} <= set(ep)
    assert "linear_vel_penalty" not in ep                             # disabled penalties have no stat [ref: USV_task_rewards.py:508-523]
    assert all(torch.isfinite(v) for v in ep.values()) and float(ep["position_error"]) > 0
    task = env.env._task
    task.update_state()
    assert task.current_state["position"].shape == (600, 2) and task.progress_buf.dtype == torch.int64
    assert task.hydrodynamics.drag_scale.shape == (600, 1) and task.thrusters_dynamics.thruster_multiplier.shape == (600, 1)


def test_nan_probe_raises_like_reference(monkeypatch):
    monkeypatch.setenv("USV_NAN_PROBE", "1")
    cfg = UsvEnvConfig(num_envs=64)
    tc = cfg.to_task_cfg()
    env = make_env(tc, DEV, seed=1)
    env.env._task._nan_probe_interval = 1
    env.reset()
    bad = torch.zeros((64, 2), device=DEV); bad[3, 1] = float("nan")
    with pytest.raises(RuntimeError, match="USV_NAN_PROBE"):
        env.step(bad)


def test_ppo_loop_runs_and_learns(tmp_path):
    """A short CaptureXY training run: finite losses, KL-adaptive lr moves, reward improves, checkpoint round-trips in the
    reference's .pth schema."""
    cfg = UsvEnvConfig(num_envs=2048, max_episode_length=400)
    env = make_env(cfg.to_task_cfg(), DEV, seed=3, collect_stats=False)
    agent = A2CAgent(env, PPOConfig(seed=3, minibatch_size=8192), DEV)
    assert agent.batch_size == 32768 and agent.num_minibatches == 4
    p0 = agent.policy.params.clone()
    rewards = []
    for chunk in range(4):
        for _ in range(15):
            agent.train_epoch()
        r, l, c = agent.episode_stats()
        rewards.append(r)
        st = agent.policy.stats()
        assert all(map(lambda v: v == v, st.values())), st                 # no NaN
    assert not torch.equal(p0, agent.policy.params) and int(agent.policy.step) == 60 * 8 * 4
    assert torch.isfinite(agent.policy.params).all()
    assert float(agent.policy.obs_rms.count) == 1 + 60 * 32768              # obs normaliser sees each frame once per epoch
    assert float(agent.policy.val_rms.count) == 1 + 2 * 60 * 32768           # values + returns (SURVEY 4.3 accounting)
    assert rewards[-1] > rewards[0], rewards                                # learning signal
    path = os.path.join(tmp_path, "last.pth")
    agent.save(path)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"model", "epoch", "optimizer", "frame", "last_mean_rewards", "env_state"} and ck["frame"] == 60 * 32768
    assert ck["model"]["a2c_network.actor_mlp.0.weight"].shape == (128, 13) and ck["optimizer"]["state"][0]["exp_avg"].shape == (2,)
    other = A2CAgent(make_env(cfg.to_task_cfg(), DEV, seed=4, collect_stats=False), PPOConfig(seed=9), DEV)
    other.restore(path)
    assert torch.equal(other.policy.params, agent.policy.params) and torch.equal(other.policy.exp_avg_sq, agent.policy.exp_avg_sq)
    assert float(other.policy.lr) == pytest.approx(float(agent.policy.lr)) and other.epoch_num == 60


def test_reference_trained_policy_captures_in_fused_env():
    """Closed loop with the reference's own trained classic policy (5450 epochs in Isaac Sim / PhysX, filename reward 38.55):
    the deterministic policy must capture the goal in essentially every episode of the fused env and earn the reward the
    checkpoint was saved at -- the end-to-end check that obs layout, reward, kills and the planar integrator reproduce the task."""
    from scripts.eval_reference_policy import run
    for tc in (False, True):
        r = run(n=2048, steps=1200, tensor_cores=tc)
        assert r["episodes"] > 10000 and r["killed"] == 0 and r["timeouts"] == 0
        assert r["captured"] >= 0.999 * r["episodes"], r
        assert 30.0 < r["mean_return"] < 48.0, r            # checkpoint: rew_38.55 at save time


def test_player_restores_checkpoint_and_plays_reference_policy(tmp_path, capsys):
    """rl_games' test=True path [ref: RLG/algos_torch/players.py:116-219, RLG/common/player.py:319-422]: PpoPlayerContinuous restores the
    reference's .pth schema, get_action(deterministic) is the clamped mu, and run() over the reference's own trained classic policy
    reports the reward the checkpoint was saved at (file name: rew_38.55)."""
    import numpy as np
    from omniisaacgymenvs_loop_b200.rl.players import PpoPlayerContinuous
    n = 1024
    env = make_env(UsvEnvConfig(num_envs=n).to_task_cfg(), DEV, seed=5)
    G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "classic_policy.npz")))
    player = PpoPlayerContinuous(env, {"games_num": 3000, "deterministic": True, "print_stats": False}, DEV, tensor_cores=False)
    player.set_weights({"model": G})
    # checkpoint round trip through the reference schema: a trainer with these weights saves, a fresh player restores
    agent = A2CAgent(env, PPOConfig(seed=1), DEV)
    agent.policy.load_state_dict(player.get_weights()["model"])
    agent.save(str(tmp_path / "ck.pth"))
    other = PpoPlayerContinuous(env, {"games_num": 10}, DEV, tensor_cores=False)
    other.restore(str(tmp_path / "ck.pth"))
    a, b = player.get_weights()["model"], other.get_weights()["model"]
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    obs = env.reset()["obs"]
    det = player.get_action(obs, is_deterministic=True)
    ref = player.model.act(obs["state"])
    assert torch.equal(det, torch.clamp(ref["mus"], -1.0, 1.0)) and det.shape == (n, 2)
    sto = player.get_action(obs, is_deterministic=False)
    assert float(sto.abs().max()) <= 1.0 and not torch.equal(sto, det)
    av_reward, av_steps, games = player.run()
    out = capsys.readouterr().out
    assert "av reward:" in out and "av steps:" in out
    assert 3000 <= games < 3000 + n
    assert 25.0 < av_reward < 50.0 and 20.0 < av_steps < 400.0, (av_reward, av_steps)   # first-finished episodes are biased short


def test_live_vecenv_matches_oracle_and_contract():
    """The live USVVirtual (Variant B: 33-dim obs, obstacles, potential field) behind VecEnvRLGames, from the live YAML tree."""
    from omniisaacgymenvs_loop_b200.config import UsvLiveConfig, live_default_config, live_task_cfg
    from oracle import usv_oracle_b as B
    from tests.test_gpu_live import oracle_live, oracle_task
    n = 80
    cfg = live_default_config(num_envs=n, max_episode_length=8, seed=11, action_bias_steps=6)
    env = make_env(live_task_cfg(cfg), DEV, seed=11)
    info = env.get_env_info()
    assert info["observation_space"]["state"].shape == (33,) and info["action_space"].shape == (2,)
    orc = B.LiveEnvOracle(oracle_cfg(cfg), oracle_task(cfg), oracle_live(UsvLiveConfig()), n)
    obs = env.reset()
    o_obs, _, _ = orc.step(torch.zeros((n, 2)))
    assert obs["obs"]["state"].shape == (n, 33)
    assert_close(obs["obs"]["state"], o_obs, 1e-5, 2e-5, "reset obs")
    g = torch.Generator().manual_seed(1)
    for k in range(18):
        act = torch.rand((n, 2), generator=g) * 3 - 1.5
        od, rew, done, extras = env.step(act.to(DEV))
        o_obs, o_rew, o_done = orc.step(act)
        assert_close(od["obs"]["state"], o_obs, 1e-4, 2e-3, f"obs {k}"); assert_close(rew, o_rew, 1e-4, 5e-3, f"rew {k}")
        assert torch.equal(done.cpu(), o_done), k
    ep = extras["episode"]
    assert {"total_reward", "potential_shaping_reward", "collision_reward", "success", "collision", "danger_mean", "u_mean",
            "energy_penalty", "g_safe_mean"} <= set(ep)
    assert "linear_vel_penalty" not in ep and all(torch.isfinite(v) for v in ep.values())
    assert 0.0 <= float(ep["success"]) <= 1.0 and 0.0 <= float(ep["collision"]) <= 1.0


def test_live_ppo_loop_runs():
    """PPO on the 33-dim live task (fp32 SIMT policy kernels: the tcgen05 path packs obs_dim <= 15): finite and moving."""
    from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg
    env = make_env(live_task_cfg(live_default_config(num_envs=1024)), DEV, seed=5, collect_stats=False)
    agent = A2CAgent(env, PPOConfig(seed=5, minibatch_size=8192), DEV)
    p0 = agent.policy.params.clone()
    for _ in range(6):
        agent.train_epoch()
    st = agent.policy.stats()
    assert all(v == v for v in st.values()), st
    assert torch.isfinite(agent.policy.params).all() and not torch.equal(p0, agent.policy.params)


def test_graphed_rollout_and_update_equal_eager():
    """Rollout and update phases replayed as CUDA graphs (device-side Philox step / sample offsets) reproduce the eager loop bit for
    bit: same rollout buffers, same parameters, same host-side counters after every epoch."""
    def build(graph):
        cfg = UsvEnvConfig(num_envs=1024, max_episode_length=40).full_dr()
        env = make_env(cfg.to_task_cfg(), DEV, seed=21)
        return A2CAgent(env, PPOConfig(seed=21, minibatch_size=4096), DEV, use_cuda_graph=graph)
    a, b = build(False), build(True)
    for ep in range(6):
        a.train_epoch()
        b.train_epoch()
        for k in ("obses", "actions", "rewards", "dones", "values", "neglogpacs"):
            assert torch.equal(a.buf[k], b.buf[k]), (ep, k)
        assert torch.equal(a.policy.params, b.policy.params), ep
        ea, eb = a.vec_env.env._task.engine, b.vec_env.env._task.engine
        assert ea.step_counter == eb.step_counter and a.policy.sample_counter == b.policy.sample_counter
        assert torch.equal(ea.state, eb.state) and torch.equal(ea.reset_buf, eb.reset_buf)
    assert b._graph_play is not None and b._graph is not None and a._graph_play is None
    xa, xb = a.vec_env.env._task.extras["episode"], b.vec_env.env._task.extras["episode"]
    assert set(xa) == set(xb) and all(torch.equal(xa[k], xb[k]) for k in xa)
    ra, rb = a.episode_stats(), b.episode_stats()
    assert ra == rb and ra[2] > 0
    # eager stepping continues seamlessly after graph replays (host and device parts of the counters stay consistent)
    act = torch.zeros((1024, 2), device=DEV)
    oa, ob = a.vec_env.step(act), b.vec_env.step(act)
    assert torch.equal(oa[0]["obs"]["state"], ob[0]["obs"]["state"]) and torch.equal(oa[1], ob[1])


def test_graphed_live_rollout_equals_eager():
    """The LIVE task's rollout (scene rebuilds + fused live step, device-side step offset) replayed as a CUDA graph reproduces the eager
    loop bit for bit, across the boundary where the host switches the initial action bias off (graphs are keyed by that parameter;
    the straddling rollout runs eagerly)."""
    import dataclasses
    from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg

    def build(graph):
        cfg = dataclasses.replace(live_default_config(num_envs=512, max_episode_length=30), action_bias_steps=72)
        env = make_env(live_task_cfg(cfg), DEV, seed=5, collect_stats=False)
        return A2CAgent(env, PPOConfig(seed=5, minibatch_size=4096), DEV, use_cuda_graph=graph)
    a, b = build(False), build(True)
    keys = []
    for ep in range(8):
        a.train_epoch()
        b.train_epoch()
        keys.append(b._graph_play_key)
        for k in ("obses", "actions", "rewards", "dones", "values"):
            assert torch.equal(a.buf[k], b.buf[k]), (ep, k)
        ea, eb = a.vec_env.env._task.engine, b.vec_env.env._task.engine
        assert ea.step_counter == eb.step_counter == 16 * (ep + 1) + 1        # + the reset step
        assert torch.equal(ea.state, eb.state) and torch.equal(ea.bstate, eb.bstate) and torch.equal(ea.bconsts, eb.bconsts)
        assert torch.equal(ea.potential, eb.potential) and torch.equal(ea.reset_epoch, eb.reset_epoch)
        assert torch.equal(a.policy.params, b.policy.params), ep
    assert "bias" in keys and keys[-1] == "steady" and b._graph_play is not None and a._graph_play is None


def test_vecenv_curriculum_free_running_vs_oracle():
    """The task-level `step` (control steps / horizon_length) drives the spawn / kill curriculum through USVVirtual."""
    cfg = UsvEnvConfig(num_envs=512, max_episode_length=6, seed=9, spawn_curriculum=True, spawn_curriculum_min_dist=0.2,
                       spawn_curriculum_max_dist=2.0, spawn_curriculum_kill_dist=3.0, spawn_curriculum_warmup=1, spawn_curriculum_end=2,
                       spawn_min_dist=3.0, spawn_max_dist=9.0, kill_dist=15.0)
    env = make_env(cfg.to_task_cfg(), DEV, seed=9)
    orc = O.ClassicEnvOracle(oracle_cfg(cfg), 512)
    obs = env.reset()
    o_obs, _, _ = orc.step(torch.zeros((512, 2)))
    assert_close(obs["obs"]["state"], o_obs, 1e-5, 2e-5, "reset obs")
    g = torch.Generator().manual_seed(4)
    for k in range(48):                                               # crosses warm-up (step 1) and end (step 2) of the curriculum
        act = torch.rand((512, 2), generator=g) * 2 - 1
        od, rew, done, _ = env.step(act.to(DEV))
        o_obs, o_rew, o_done = orc.step(act)
        assert torch.equal(done.cpu(), o_done), k
        assert_close(od["obs"]["state"], o_obs, 1e-4, 2e-3, f"obs {k}")
    assert abs(env.env._task.step - orc.curriculum_step) < 1e-9 and orc.curriculum_step > 3.0


def test_reward_curves_tf32_vs_fp32_vs_recorded_oracle():
    """north_star's "matching reference reward curves": same seeds, the tcgen05 TF32 path and the fp32 path (1e-5 parity with rl_games) train
    CaptureXY to the same place, and that place is where the CPU oracle of the rl_games loop ends up (profiles/r02_reward_curves.md: mean step
    reward over the second half of 150 epochs, 2048 envs: oracle 0.164 (0.104 .. 0.210 over 3 seeds), recorded with scripts/reward_curves.py)."""
    from scripts.reward_curves import gpu_curve, windows
    last = {}
    for name, tc in (("tf32", True), ("fp32", False)):
        xs = []
        for seed in (12, 14, 15, 16):
            _, step_rew = gpu_curve(2048, 150, seed, tc, DEV)
            q = windows(step_rew)
            assert q[3] > q[0] + 0.15, (name, seed, q)                  # every seed learns: from about -0.1 to +0.1 .. +0.25
            xs.append(0.5 * (q[2] + q[3]))
        last[name] = sum(xs) / len(xs)
    assert abs(last["tf32"] - last["fp32"]) < 0.05, last             # seed-to-seed spread is ~0.1; the learners differ by < 0.01 in the record
    for name, v in last.items():
        assert 0.10 < v < 0.26, (name, v)                            # the oracle's recorded range, widened by the seed spread


def test_train_loop_checkpoint_cadence_and_scalar_stream(tmp_path):
    """A2CAgent.train with the reference's checkpoint cadence and scalar tags [ref: RLG/common/a2c_common.py:343-362,1399-1470]: the best
    checkpoint `<name>.pth` appears once a mean reward beats the running best after `save_best_after`, `last_<name>_ep_<n>_rew_<r>.pth` at
    max_epochs, every file loads back through the .pth schema, and the scalar stream carries the reference's tags."""
    import json
    from omniisaacgymenvs_loop_b200.rl.a2c import ScalarLog
    cfg = UsvEnvConfig(num_envs=1024, max_episode_length=60)
    env = make_env(cfg.to_task_cfg(), DEV, seed=5, collect_stats=False)
    agent = A2CAgent(env, PPOConfig(seed=5, minibatch_size=8192), DEV)
    nn_dir, log_path = os.path.join(tmp_path, "nn"), os.path.join(tmp_path, "scalars.jsonl")
    w = ScalarLog(log_path)
    agent.train(max_epochs=24, log_every=4, log=None, writer=w, nn_dir=nn_dir, name="USV", save_freq=8, save_best_after=8)
    w.close()
    files = sorted(os.listdir(nn_dir))
    assert "USV.pth" in files and any(f.startswith("last_USV_ep_24_rew_") for f in files), files
    ck = torch.load(os.path.join(nn_dir, "USV.pth"), weights_only=False)
    assert set(ck) == {"model", "epoch", "optimizer", "frame", "last_mean_rewards", "env_state"} and 8 <= ck["epoch"] <= 24
    tags = {json.loads(l)["tag"] for l in open(log_path)}
    assert {"losses/a_loss", "losses/c_loss", "losses/entropy", "info/last_lr", "info/kl", "info/epochs", "rewards/step", "rewards/iter",
            "rewards/time", "episode_lengths/step", "performance/step_inference_rl_update_fps"} <= tags, tags
    fresh = A2CAgent(make_env(cfg.to_task_cfg(), DEV, seed=6, collect_stats=False), PPOConfig(seed=6, minibatch_size=8192), DEV)
    fresh.restore(os.path.join(nn_dir, "USV.pth"))
    assert fresh.epoch_num == ck["epoch"] and torch.equal(fresh.policy.views()["a2c_network.mu.weight"].cpu(), ck["model"]["a2c_network.mu.weight"].cpu())
