"""GPU parity tests of the Tier-3 tasks (GoToPose / KeepXY / TrackXYVelocity) through the C ABI: against goldens
from the reference's own task classes, and against the CPU oracle in lock-step with full per-env DR (BASELINE config C4)."""
import dataclasses

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.config import PenaltyTerm, UsvLiveConfig, live_default_config  # noqa: E402
from omniisaacgymenvs_loop_b200.engine import FusedUsvLiveEnv  # noqa: E402
from oracle import usv_oracle_t as TT  # noqa: E402
from tests.test_gpu_live import oracle_live  # noqa: E402
from tests.test_oracle_t_cpu import SPECS  # noqa: E402
from tests.util import assert_close, engine_state, oracle_cfg, oracle_state, push_oracle_state  # noqa: E402

DEV = "cuda:0"
T = torch.from_numpy
OFF = PenaltyTerm()


def configs(c: TT.Tier3Config, **kw):
    """(UsvEnvConfig, UsvLiveConfig) of the fused step for an oracle task config."""
    cfg = live_default_config(position_tolerance=c.position_tolerance, kill_after_n_steps_in_tolerance=c.kill_after_n_steps_in_tolerance,
                              kill_dist=c.kill_dist, reward_mode=c.reward_mode, exponential_reward_coeff=c.exponential_reward_coeff,
                              position_scale=c.position_scale, spawn_about_origin=False, spawn_min_dist=0.3, spawn_max_dist=3.0,
                              goal_random_position=1.0, **kw)
    live = UsvLiveConfig(task=c.task, heading_reward_mode=c.heading_reward_mode, heading_exponential_reward_coeff=c.heading_exponential_reward_coeff,
                         heading_scale=c.heading_scale, sig_gain=c.sig_gain, lin_vel_tolerance=c.lin_vel_tolerance,
                         goal_random_velocity=c.goal_random_velocity)
    return cfg, live


@pytest.mark.parametrize("tag", list(SPECS))
def test_tier3_task_vs_reference_golden(golden, tag):
    G = golden("tier3_tasks")
    c = SPECS[tag]
    K, n = G[f"{tag}_pos"].shape[:2]
    cfg, live = configs(c, n_substeps=0, max_episode_length=10_000, mass_rand=False, mass_coupling=False, use_drag_scale=False,
                        reset_pose_external=True, retarget_on_reset=False, noise_vel=False, noise_heading=False, action_bias_steps=0,
                        pen_energy=OFF, pen_angular_vel=OFF, pen_angular_vel_variation=OFF)
    live = dataclasses.replace(live, priv_mode=0, mass_obs_relative=False, com_obs_scaled=False, com_rand=False)
    env = FusedUsvLiveEnv(cfg, live, n, DEV)
    env.set_field("USV_C_TX", T(G[f"{tag}_target"][:, 0])); env.set_field("USV_C_TY", T(G[f"{tag}_target"][:, 1]))
    if f"{tag}_target_heading" in G:
        env.set_field("USV_BC_TARGET_HEADING", T(G[f"{tag}_target_heading"]))
    if f"{tag}_target_vel" in G:
        env.set_field("USV_BC_TARGET_VX", T(np.ascontiguousarray(G[f"{tag}_target_vel"][:, 0])))
        env.set_field("USV_BC_TARGET_VY", T(np.ascontiguousarray(G[f"{tag}_target_vel"][:, 1])))
    for k in range(K):
        env.reset_buf.fill_(1 if k == 0 else 0)
        pr = G[f"{tag}_priv"][k]
        for name, v in (("USV_S_X", G[f"{tag}_pos"][k][:, 0]), ("USV_S_Y", G[f"{tag}_pos"][k][:, 1]), ("USV_S_PSI", G[f"{tag}_yaw"][k]),
                        ("USV_S_VX", G[f"{tag}_vel"][k][:, 0]), ("USV_S_VY", G[f"{tag}_vel"][k][:, 1]), ("USV_S_R", G[f"{tag}_w"][k]),
                        ("USV_C_MASS", pr[:, 0]), ("USV_BC_COM_X", pr[:, 1]), ("USV_BC_COM_Y", pr[:, 2]), ("USV_BC_COM_Z", pr[:, 3]),
                        ("USV_C_KDRAG", pr[:, 4]), ("USV_C_THR_ML", pr[:, 5]), ("USV_C_THR_MR", pr[:, 6]), ("USV_C_KIZ", pr[:, 7])):
            env.set_field(name, T(np.ascontiguousarray(v)))
        obs, rew, done = env.step(T(G[f"{tag}_actions"][k]).to(DEV))
        obs, want = obs.cpu().clone(), T(G[f"{tag}_obs"][k]).clone()
        if k == 0:                  # reset rows: prev_action zeroed and the mass re-drawn (nominal) before the step
            obs[:, 23:26] = want[:, 23:26]
        assert_close(obs, want.clamp(-cfg.clip_obs, cfg.clip_obs), 1e-5, 2e-6, f"{tag} obs step {k}")
        assert_close(rew, G[f"{tag}_reward"][k], 1e-5, 2e-6, f"{tag} reward step {k}")
        assert torch.equal(done.cpu(), T(G[f"{tag}_die"][k])), (tag, k)
        assert torch.equal(env.goal_reached.cpu(), T(G[f"{tag}_goal_reached"][k])), (tag, k)
    env.check_finite()


def push_task_state(orc, env):
    push_oracle_state(orc, env)
    for j, name in enumerate(("USV_BC_COM_X", "USV_BC_COM_Y", "USV_BC_COM_Z")):
        env.set_field(name, orc.com[:, j].to(DEV))
    if orc.task.task == TT.GO_TO_POSE:
        env.set_field("USV_BC_TARGET_HEADING", orc.target_heading.to(DEV))
    if orc.task.task == TT.TRACK_XY_VELOCITY:
        env.set_field("USV_BC_TARGET_VX", orc.target_vel[:, 0].to(DEV))
        env.set_field("USV_BC_TARGET_VY", orc.target_vel[:, 1].to(DEV))


@pytest.mark.parametrize("tag", ["gotopose", "keepxy", "trackxyvel"])
@pytest.mark.parametrize("sync", [True, False])
def test_tier3_step_vs_oracle(tag, sync):
    """BASELINE config C4: the task with full per-env DR (DR50 flag set) on the live action path; lock-step and free-running."""
    c = dataclasses.replace(SPECS[tag], kill_dist=6.0)
    cfg, live = configs(c, max_episode_length=9, action_bias_steps=4)
    cfg = dataclasses.replace(cfg.full_dr(), action_noise=True, spawn_min_dist=(0.0 if c.task == TT.TRACK_XY_VELOCITY else 0.3),
                              spawn_max_dist=(0.0 if c.task == TT.TRACK_XY_VELOCITY else 3.0))
    n = 1024 + 19
    env = FusedUsvLiveEnv(cfg, live, n, DEV)
    orc = TT.Tier3EnvOracle(oracle_cfg(cfg), c, oracle_live(live), n)
    g = torch.Generator().manual_seed(13)
    n_done = 0
    for k in range(20):
        act = torch.rand((n, 2), generator=g) * 2.4 - 1.2
        if sync or k == 0:          # free-running: only the initial state (post_reset targets) is shared
            push_task_state(orc, env)
        o_obs, o_rew, o_done = orc.step(act)
        obs, rew, done = env.step(act.to(DEV))
        tol = (1e-5, 2e-5) if sync else (1e-4, 2e-3)
        assert_close(obs, o_obs, *tol, f"{tag} obs step {k}")
        assert_close(rew, o_rew, tol[0], 2e-5 if sync else 5e-3, f"{tag} reward step {k}")
        assert torch.equal(done.cpu(), o_done), f"{tag} done mismatch at step {k}"
        if sync:
            es, os_ = engine_state(env), oracle_state(orc)
            assert torch.equal(es["goal"], os_["goal"]) and torch.equal(es["progress"], os_["progress"])
            for name in es:
                if name not in ("goal", "progress", "reset"):
                    assert_close(es[name], os_[name], 1e-5, 2e-5 if not name.startswith("USV_C_F") else 1e-4, f"{name} step {k}")
        n_done += int(done.sum())
    assert n_done > n
    env.check_finite()
