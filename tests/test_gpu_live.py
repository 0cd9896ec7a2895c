"""GPU parity tests of Variant B (live CaptureXY with static obstacles), through the C ABI:
(i) against the goldens produced by the reference's own CaptureXYTask / BatchedMapGPU (tests/golden/capture_xy_live.npz) and
(ii) against the CPU oracle (oracle/usv_oracle_b.py) on the same seeded inputs, resets and scene rebuilds included.
Bars: obs / reward terms 1e-5 relative (atol written at each assert), kills / outcome latches / obstacle placement and the
cost-to-go field bit-exact."""
import dataclasses
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.config import PenaltyTerm, UsvEnvConfig, UsvLiveConfig  # noqa: E402
from omniisaacgymenvs_loop_b200.engine import FusedUsvLiveEnv  # noqa: E402
from oracle import usv_oracle_b as B  # noqa: E402
from tests.util import assert_close, engine_state, oracle_cfg, oracle_state, push_oracle_state  # noqa: E402

DEV = "cuda:0"
T = torch.from_numpy
OFF = PenaltyTerm()

LIVE_CFG = UsvEnvConfig(
    dt=0.01, n_substeps=10, max_episode_length=200, action_affine=True, penalties_use_u=True, action_noise=False,
    spawn_about_origin=True, retarget_on_reset=True, spawn_min_dist=9.0, spawn_max_dist=12.0,
    position_tolerance=1.0, kill_after_n_steps_in_tolerance=1, kill_dist=20.0, goal_reward=20.0, time_reward=-0.05,
    position_scale=1.5, align_la1=0.04, mass_rand=True, mass_min=34.96, mass_max=54.96, mass_base=34.96,
    mass_coupling=True, couple_mass_max=54.96, couple_thr_a=0.5, kdrag_min=1.0, kdrag_max=1.5, use_drag_scale=True,
    pen_energy=PenaltyTerm(1, 0.005, 0.0, 0.0), pen_angular_vel=PenaltyTerm(4, 0.02, 0.0, 0.4),
    pen_angular_vel_variation=PenaltyTerm(4, 0.005, 0.0, 0.10))


def oracle_live(live: UsvLiveConfig):
    return B.LivePrivConfig(priv_mode=live.priv_mode, mass_obs_relative=live.mass_obs_relative, com_obs_scaled=live.com_obs_scaled,
                            com_scale=live.com_scale, priv_a=live.priv_a, priv_b=live.priv_b, priv_active=live.priv_active,
                            com_rand=live.com_rand, com_base=live.com_base, com_disp=live.com_disp,
                            collision_threshold=live.collision_threshold, fixed_horizon_eval=live.fixed_horizon_eval)


def oracle_task(cfg: UsvEnvConfig):
    return B.LiveTaskConfig(position_tolerance=cfg.position_tolerance, kill_after_n_steps_in_tolerance=cfg.kill_after_n_steps_in_tolerance,
                            kill_dist=cfg.kill_dist, boundary_cost=cfg.boundary_cost, goal_reward=cfg.goal_reward,
                            time_reward=cfg.time_reward, position_scale=cfg.position_scale, align_la1=cfg.align_la1,
                            align_la2=cfg.align_la2, align_la3=cfg.align_la3, spawn_min_dist=cfg.spawn_min_dist,
                            spawn_max_dist=cfg.spawn_max_dist, goal_random_position=cfg.goal_random_position)


# ------------------------------------------------------------------------------------------
def test_potential_field_builder_vs_reference_golden(golden):
    """B6: occupancy/SDF -> wavefront cost-to-go (bit-exact, +inf cells included) -> potential field, batch-global maxima."""
    G = golden("capture_xy_live")
    env = FusedUsvLiveEnv(LIVE_CFG, UsvLiveConfig(), 32, DEV)
    field, cost = env.build_fields(T(G["obstacles0"]), T(G["target"]), want_cost=True)
    assert torch.equal(cost.cpu(), T(G["cost0"]))
    assert_close(field, G["field0"], 1e-6, 1e-6, "potential field")
    # a batch of one: the global maxima change, the raw cost does not
    f1, c1 = env.build_fields(T(G["obstacles0"][3:4]), T(G["target"][3:4]), want_cost=True)
    assert torch.equal(c1.cpu()[0], T(G["cost0"][3]))
    want = B.potential_field(T(G["cost0"][3:4]), T(G["sdf0"][3:4]))
    assert_close(f1, want, 1e-6, 1e-6, "potential field (batch of one)")


def test_field_builder_statistics_path_vs_oracle_on_walled_scenes():
    """The cost kernel hands the field kernel per-scene extrema (finite-cost and unreachable cells apart, the latter scaled by their common
    rep factor afterwards) and takes the batch maxima itself: scenes with a pocket the wavefront cannot enter (unreachable FREE cells next to
    obstacles), scenes without any obstacle in reach (no inside cell, J = 0 everywhere) and ordinary ones in ONE batch, against the oracle's
    BatchedMapGPU restatement (d_multi_gemini.py:194-271) -- and with more scenes than the env owns, so the workspace has to grow."""
    env = FusedUsvLiveEnv(LIVE_CFG, UsvLiveConfig(), 8, DEV)
    g = torch.Generator().manual_seed(11)
    m = 24
    obst = torch.full((m, 16, 2), 999.0)
    tgt = torch.zeros((m, 2))
    for k in range(m):
        if k % 3 == 0:        # a ring of 16 overlapping discs of radius 0.5 at 1.9 m around (6, 6): the inside is free but unreachable
            ang = torch.arange(16) * (2 * math.pi / 16)
            obst[k, :, 0] = 6.0 + 1.9 * torch.cos(ang)
            obst[k, :, 1] = 6.0 + 1.9 * torch.sin(ang)
            tgt[k] = torch.tensor([-5.0, -4.0]) + torch.rand(2, generator=g)
        elif k % 3 == 1:      # nothing on the map
            tgt[k] = torch.rand(2, generator=g) * 20 - 10
        else:                 # scattered
            obst[k] = torch.rand((16, 2), generator=g) * 24 - 12
            tgt[k] = torch.tensor([13.0, 13.0]) - torch.rand(2, generator=g)
    field, cost = env.build_fields(obst, tgt, want_cost=True)
    occ, sdf = B.occupancy_and_sdf(obst)
    want_cost = B.cost_to_go(occ, tgt)
    assert torch.equal(cost.cpu(), want_cost)
    pocket = want_cost[0].isinf() & (occ[0] < 0.5)
    assert int(pocket.sum()) > 20                                   # unreachable free cells exist
    assert_close(field, B.potential_field(want_cost, sdf), 1e-6, 1e-6, "potential field, mixed batch")
    # the same scenes one by one: other batch maxima, same machinery
    for k in (0, 1, 2):
        f1 = env.build_fields(obst[k:k + 1], tgt[k:k + 1])
        assert_close(f1, B.potential_field(want_cost[k:k + 1], sdf[k:k + 1]), 1e-6, 1e-6, f"potential field, scene {k} alone")


def test_cost_to_go_in_place_relaxation_equals_jacobi_on_odd_scenes():
    """The in-place asynchronous relaxation reaches the Jacobi fixed point bit for bit; scenes where the 225th Jacobi iterate is NOT
    the fixed point of an asynchronous order (target cell occupied / on the border wall) take the literal-Jacobi path."""
    g = torch.Generator().manual_seed(31)
    m = 12
    obst = (torch.rand((m, 16, 2), generator=g) * 24 - 12)
    target = torch.rand((m, 2), generator=g) * 20 - 10
    obst[0, 0] = target[0]                                   # target inside an obstacle
    target[1] = torch.tensor([14.95, -14.95])                # target in the corner cell (border wall: occupied)
    target[2] = torch.tensor([-14.7, -14.7])                 # target next to a corner: longest paths (~209 < 224)
    target[3] = torch.tensor([40.0, 3.0])                    # outside the map: clamped onto the border
    obst[4, :8, 0] = torch.linspace(-3.5, 3.5, 8); obst[4, :8, 1] = 2.0      # a wall of touching discs across the map centre
    target[4] = torch.tensor([0.0, 6.0])
    obst[5] = 999.0                                          # all obstacles in limbo
    env = FusedUsvLiveEnv(LIVE_CFG, UsvLiveConfig(), 16, DEV)
    field, cost = env.build_fields(obst, target, want_cost=True)
    occ, sdf = B.occupancy_and_sdf(obst)
    want = B.cost_to_go(occ, target)
    assert torch.equal(cost.cpu(), want)
    assert float(want[2][torch.isfinite(want[2])].max()) > 200.0
    assert_close(field, B.potential_field(want, sdf), 1e-6, 1e-6, "potential field (odd scenes)")


def test_live_task_vs_reference_golden(golden):
    """B1-B4 on the states the reference task was driven with (n_substeps=0: the kernel's physics is a no-op, the host writes
    pose / velocity every step = the reference's scene-replay mode): obs(33), reward, kills, outcome latches."""
    G = golden("capture_xy_live")
    K, n = G["pos"].shape[:2]
    cfg = dataclasses.replace(LIVE_CFG, n_substeps=0, max_episode_length=10_000, mass_rand=False, mass_coupling=False, use_drag_scale=False,
                              reset_pose_external=True, retarget_on_reset=False, noise_vel=False, noise_heading=False, noise_pos=False,
                              pen_energy=OFF, pen_angular_vel=OFF, pen_angular_vel_variation=OFF)
    live = UsvLiveConfig(priv_mode=0, mass_obs_relative=False, com_obs_scaled=False, com_rand=False)
    env = FusedUsvLiveEnv(cfg, live, n, DEV)
    env.set_obstacles(T(G["obstacles0"]))
    env.potential.copy_(T(G["field0"]).to(DEV))
    env.set_field("USV_C_TX", T(G["target"][:, 0])); env.set_field("USV_C_TY", T(G["target"][:, 1]))
    reset_ids = T(G["reset_ids"])
    for k in range(K):
        was_reset = torch.zeros(n, dtype=torch.bool)
        env.reset_buf.zero_()
        env.reset_epoch.fill_(-1)
        if k == 0:
            was_reset[:] = True
        elif k == int(G["reset_step"]):
            was_reset[reset_ids] = True
            env.set_obstacles(T(G["obstacles1"]))
            env.potential.copy_(T(G["field1"]).to(DEV))
        if was_reset.any():
            env.reset_buf.copy_(was_reset.long().to(DEV))
            env.mark_host_reset()
        for name, v in (("USV_S_X", G["pos"][k][:, 0]), ("USV_S_Y", G["pos"][k][:, 1]), ("USV_S_PSI", G["yaw"][k]),
                        ("USV_S_VX", G["vel"][k][:, 0]), ("USV_S_VY", G["vel"][k][:, 1]), ("USV_S_R", G["w"][k]),
                        ("USV_C_MASS", G["priv"][k][:, 0]), ("USV_BC_COM_X", G["priv"][k][:, 1]), ("USV_BC_COM_Y", G["priv"][k][:, 2]),
                        ("USV_BC_COM_Z", G["priv"][k][:, 3]), ("USV_C_KDRAG", G["priv"][k][:, 4]), ("USV_C_THR_ML", G["priv"][k][:, 5]),
                        ("USV_C_THR_MR", G["priv"][k][:, 6]), ("USV_C_KIZ", G["priv"][k][:, 7])):
            env.set_field(name, T(np.ascontiguousarray(v)))
        obs, rew, done = env.step(T(G["prev_action"][k]).to(DEV), rebuild_scene=False)
        obs, want = obs.cpu().clone(), T(G["obs"][k]).clone()
        # a resetting env sees prev_action = 0 and the nominal mass (reset_idx zeroes / re-draws them before the step; the
        # golden drove get_state_observations directly with random values): those three columns are checked separately
        assert torch.all(obs[was_reset][:, 23:25] == 0) and torch.all(obs[was_reset][:, 25] == min(np.float32(cfg.mass_base), np.float32(cfg.clip_obs)))
        obs[was_reset, 23:26] = want[was_reset, 23:26]
        want = want.clamp(-cfg.clip_obs, cfg.clip_obs)          # VecEnvRLGames._process_data; the golden is the task's raw buffer
        assert_close(obs, want, 1e-5, 2e-6, f"obs step {k}")
        # the shaping term is 2 x 100 x (difference of two nearly equal potentials): 1e-7 relative on the potential -> 1e-4
        assert_close(rew, G["reward"][k], 1e-5, 1.5e-4, f"reward step {k}")
        assert torch.equal(done.cpu(), T(G["die"][k])), k
        assert torch.equal(env.goal_reached.cpu(), T(G["goal_reached"][k]))
        succ, coll = env.episode_outcomes()
        assert torch.equal(succ.cpu().int(), T(G["done_success"][k])) and torch.equal(coll.cpu().int(), T(G["done_collision"][k]))
    env.check_finite()


# ------------------------------------------------------------------------------------------
LIVE_STATE = [("USV_BS_PREV_H", lambda o: o.S.prev_h), ("USV_BS_PREV_POT", lambda o: o.S.prev_pot)]


def push_live_state(orc, env):
    push_oracle_state(orc, env)
    z = torch.zeros(orc.n)
    for name, get in LIVE_STATE:
        v = get(orc)
        env.set_field(name, (z if v is None else v).to(DEV))
    env.set_field("USV_BS_OUTCOME", (orc.S.done_success + 2 * orc.S.done_collision).to(torch.int32).to(DEV))
    for j in range(3):
        env.set_field(("USV_BC_COM_X", "USV_BC_COM_Y", "USV_BC_COM_Z")[j], orc.com[:, j].to(DEV))
    env.set_obstacles(orc.obstacles)
    env.potential.copy_(orc.field.to(DEV))
    env.reset_epoch.fill_(-1)
    if bool(orc.reset_buf.any()):
        env.mark_host_reset()


def _lockstep(cfg, live, n, steps, seed=7, sync=True, collect_stats=False):
    env = FusedUsvLiveEnv(cfg, live, n, DEV, collect_stats=collect_stats)
    orc = B.LiveEnvOracle(oracle_cfg(cfg), oracle_task(cfg), oracle_live(live), n)
    g = torch.Generator().manual_seed(seed)
    for k in range(steps):
        act = torch.rand((n, 2), generator=g) * 2.4 - 1.2
        if sync:
            push_live_state(orc, env)
        o_out = orc.step(act)
        obs, rew, done = env.step(act.to(DEV))
        yield k, env, orc, (obs.cpu(), rew.cpu(), done.cpu()), o_out


def test_live_step_vs_oracle_lockstep():
    """Full control steps (dynamics + live task + resets with scene rebuild) from identical state each step."""
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=7, kill_dist=12.5, action_bias=-0.6, action_bias_steps=5)
    live = UsvLiveConfig()
    n = 96 + 5
    n_done = 0
    for k, env, orc, (obs, rew, done), (o_obs, o_rew, o_done) in _lockstep(cfg, live, n, 12):
        assert obs.shape == (n, 33) and done.dtype == torch.int64
        # B5: the same Philox draws -> the same obstacles, bit for bit; B6: the fields rebuilt for exactly the envs that reset
        assert torch.equal(env.obstacles.cpu(), orc.obstacles), f"obstacles step {k}"
        assert_close(env.potential, orc.field, 1e-6, 1e-6, f"fields step {k}")
        assert_close(obs, o_obs, 1e-5, 2e-5, f"obs step {k}")
        assert_close(rew, o_rew, 1e-5, 2e-4, f"reward step {k}")
        assert torch.equal(done, o_done), f"done mismatch at step {k}"
        es, os_ = engine_state(env), oracle_state(orc)
        assert torch.equal(es["goal"], os_["goal"]) and torch.equal(es["progress"], os_["progress"])
        for name in es:
            if name not in ("goal", "progress", "reset"):
                assert_close(es[name], os_[name], 1e-5, 2e-5, f"{name} step {k}")
        succ, coll = env.episode_outcomes()
        assert torch.equal(succ.cpu().int(), orc.S.done_success) and torch.equal(coll.cpu().int(), orc.S.done_collision)
        assert_close(env.field("USV_BS_PREV_POT"), orc.S.prev_pot, 1e-6, 1e-6, f"prev_potential step {k}")
        n_done += int(done.sum())
    assert n_done > n
    env.check_finite()


def test_live_legacy_com_disc_vs_oracle():
    """The configuration branches no shipped YAML takes, through the live step: legacy disc-shaped CoM re-draw (its oracle half is
    pinned to MDD._randomize_com in tests/test_config_branches_cpu.py), a partial coupling target list with the independent log-space
    k_Iz draw beside it, and a water current  [ref: OIGE/tasks/USV/USV_disturbances.py:108-124 ; OIGE/tasks/USV_Virtual.py:153-170,988-1040]."""
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=6, kill_dist=12.5, couple_targets=3, kiz_rand=True, kiz_log=True,
                              couple_kiz_min=0.8, couple_kiz_max=1.7, use_water_current=True, flow_vel_xy=(0.25, -0.15))
    live = UsvLiveConfig(com_rand=2, com_base=(0.02, -0.01, 0.03), com_disp=(0.12, 0.0, 0.0))
    n = 64 + 3
    n_done = 0
    for k, env, orc, (obs, rew, done), (o_obs, o_rew, o_done) in _lockstep(cfg, live, n, 14):
        assert_close(obs, o_obs, 1e-5, 2e-5, f"obs step {k}")
        assert_close(rew, o_rew, 1e-5, 2e-4, f"reward step {k}")
        assert torch.equal(done, o_done), f"done mismatch at step {k}"
        com = torch.stack([env.field(f"USV_BC_COM_{a}") for a in "XYZ"], 1).cpu()
        assert_close(com, orc.com, 1e-6, 1e-7, f"CoM step {k}")
        assert torch.equal(com[:, 2], torch.full((n,), 0.03))                       # the disc leaves z alone
        es, os_ = engine_state(env), oracle_state(orc)
        for name in ("USV_C_MASS", "USV_C_KDRAG", "USV_C_THR_ML", "USV_C_THR_MR", "USV_C_KIZ", "USV_S_VX", "USV_S_VY", "USV_S_R"):
            assert_close(es[name], os_[name], 1e-5, 2e-5, f"{name} step {k}")
        n_done += int(done.sum())
    assert n_done > n
    r = torch.linalg.vector_norm(com[:, :2] - torch.tensor([0.02, -0.01]), dim=1)
    assert float(r.max()) <= 0.12 + 1e-6 and float(r.std()) > 0.01
    kiz = env.field("USV_C_KIZ").cpu()
    assert float(kiz.min()) >= 0.8 - 1e-6 and float(kiz.max()) <= 1.7 + 1e-6 and float(kiz.std()) > 0.05   # independent draw, not the coupled one
    env.check_finite()


def test_live_step_free_running_vs_oracle():
    """No re-sync: resets, obstacle re-draws and field rebuilds happen on the same steps for the same envs."""
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=9)
    n = 64
    for k, env, orc, (obs, rew, done), (o_obs, o_rew, o_done) in _lockstep(cfg, UsvLiveConfig(), n, 22, sync=False):
        assert torch.equal(done, o_done), f"done mismatch at step {k}"
        assert torch.equal(env.obstacles.cpu(), orc.obstacles), f"obstacles step {k}"
        assert_close(obs, o_obs, 1e-4, 2e-3, f"free-running obs step {k}")
        assert_close(rew, o_rew, 1e-4, 5e-3, f"free-running reward step {k}")


def test_live_stats_vs_oracle():
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=6)
    n = 64
    E = __import__("omniisaacgymenvs_loop_b200._lib", fromlist=["ENUMS"]).ENUMS
    sums = None
    for k, env, orc, _, _ in _lockstep(cfg, UsvLiveConfig(), n, 9, sync=True, collect_stats=True):
        L = orc.last
        st = L["state"]
        thr = orc.current_forces
        rows = {
            "TOTAL_REWARD": L["reward"], "DISTANCE_REWARD": L["distance_reward"], "ALIGNMENT_REWARD": L["alignment_reward"],
            "HEADING_IMPROVE_REWARD": L["heading_improve"], "POTENTIAL_SHAPING_REWARD": L["potential_shaping"],
            "SPEED_REWARD": L["speed_reward"], "ANGULAR_REWARD": L["angular_reward"], "TURN_HAZARD_PENALTY": L["turn_hazard"],
            "GOAL_REWARD": L["goal_reward"], "COLLISION_REWARD": L["collision_penalty"], "TIME_REWARD": torch.full((n,), cfg.time_reward),
            "POSITION_ERROR": L["d"], "BOUNDARY_PENALTY": L["boundary_penalty"], "DANGER_MEAN": L["danger"], "DANGER_HI_RATE": L["danger_hi"],
            "G_GATE_MEAN": L["g_gate"], "LINEAR_VEL_PENALTY": L["pen_lin"], "ANGULAR_VEL_PENALTY": L["pen_ang"],
            "ANGULAR_VEL_VARIATION_PENALTY": L["pen_angvar"], "ENERGY_PENALTY": L["pen_energy"], "ACTION_VARIATION_PENALTY": L["pen_actvar"],
            "NORMED_LINEAR_VEL": torch.norm(st["linear_velocity"], dim=-1), "NORMED_ANGULAR_VEL": st["angular_velocity"].abs(),
            "ACTIONS_SUM": L["raw_actions"].sum(-1), "CMD_NEG_RATE": (L["before_rect"] < 0).float().mean(1),
            "THRUSTER_FORCE_NEG_RATE": (thr < 0).float().mean(1), "U_MEAN": L["unit"].mean(1), "U_LOW_RATE": (L["unit"] < 0.05).float().mean(1),
            "U_SUM": L["unit"].sum(1)}
        terms = torch.zeros((E["USV_BST_COUNT"], n))
        for name, v in rows.items():
            terms[E["USV_BST_" + name]] = v
        assert len(rows) == E["USV_BST_COUNT"]
        if sums is None:
            sums = torch.zeros_like(terms)
        sums[:, L["reset_ids"]] = 0
        sums += terms
        assert_close(env.bstats_matrix().cpu(), sums, 1e-5, 5e-4, f"episode sums step {k}")


def test_live_full_size_properties():
    """4096 envs, 40 free-running steps with random actions: finite, bounded, outcomes consistent, obstacles legal."""
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=30)
    n = 4096
    env = FusedUsvLiveEnv(cfg, UsvLiveConfig(), n, DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    resets = 0
    for k in range(40):
        obs, rew, done = env.step(torch.rand((n, 2), generator=g, device=DEV) * 2 - 1)
        resets += int(done.sum())
        assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
        assert float(obs.abs().max()) <= cfg.clip_obs
    env.check_finite()
    assert resets >= n
    ob = env.obstacles                                        # (n,16,2)
    valid = ob[..., 0] < 900
    tgt = torch.stack([env.field("USV_C_TX"), env.field("USV_C_TY")], 1)
    assert bool((((ob - tgt[:, None]).norm(dim=-1) >= 3.0) | ~valid).all())
    d2 = ((ob[:, :, None] - ob[:, None]) ** 2).sum(-1)
    pair = valid[:, :, None] & valid[:, None] & ~torch.eye(16, dtype=torch.bool, device=DEV)
    assert bool(((d2 >= 2.5 * 2.5 - 1e-4) | ~pair).all())
    assert bool((((ob - tgt[:, None]).abs() <= 12.0 + 1e-4) | ~valid[..., None]).all())
    f = env.potential
    assert bool(torch.isfinite(f).all()) and float(f.min()) >= 0.0 and float(f.max()) <= 1.5 + 1e-5
    succ, coll = env.episode_outcomes()
    assert bool(((succ + coll) <= 1).all())


# ------------------------------------------------------------------------------------------
# the live USVVirtual's host chains around the task, against goldens produced by the reference class itself
def _golden_live_cfg(G, **kw):
    return dataclasses.replace(
        LIVE_CFG, n_lut=int(G["n_lut"]), lut_points_left=tuple(G["lut_points_left"].tolist()), lut_points_right=tuple(G["lut_points_right"].tolist()),
        reset_pose_external=True, retarget_on_reset=False, mass_rand=False, mass_coupling=False, use_drag_scale=False, noise_vel=False,
        noise_heading=False, pen_energy=OFF, pen_angular_vel=OFF, pen_angular_vel_variation=OFF, **kw)


def test_live_action_path_vs_reference_golden(golden):
    """A12: the thrust target the kernel feeds the lag filter and the prev_action observation, with / without the initial bias."""
    from oracle.usv_oracle import lag_alpha, thruster_lag
    G = golden("live_virtual")
    ids = T(G["act_reset_ids"])
    for tag, steps in (("bias", 5), ("nobias", 0)):
        n = G[f"act_{tag}_in"].shape[0]
        cfg = _golden_live_cfg(G, n_substeps=1, action_bias=float(G["act_bias"]), action_bias_steps=steps)
        env = FusedUsvLiveEnv(cfg, UsvLiveConfig(), n, DEV)
        env.reset_buf.zero_()
        env.reset_buf[ids.to(DEV)] = 1
        env.mark_host_reset()
        obs, _, _ = env.step(T(G[f"act_{tag}_in"]).to(DEV), rebuild_scene=False)
        assert torch.equal(obs[:, 23:25].cpu(), T(G[f"act_{tag}_prev"]))
        want = thruster_lag(torch.zeros((n, 2)), T(G[f"act_{tag}_target"]), lag_alpha(cfg.dt, cfg.time_constant))
        got = torch.stack([env.field("USV_S_THR_L"), env.field("USV_S_THR_R")], 1).cpu()
        assert torch.equal(got, want), tag


def test_live_priv_tail_vs_reference_golden(golden):
    """B1 privileged tail [m_rel, com/scale, enc(k_drag), enc(thr_L), enc(thr_R), enc(k_Iz)] in the three encodings."""
    G = golden("live_virtual")
    n = G["cpl_mass"].shape[0]
    scale = tuple(float(x) for x in G["cpl_com_scale"])
    for mode, code, a, b in (("minmax", 2, (1.0, 0.5, 0.5, 1.0), (0.5, 0.5, 0.5, 0.5)), ("centered", 1, (1.0,) * 4, (0.5,) * 4),
                             ("raw", 0, (0.0,) * 4, (1.0,) * 4)):
        env = FusedUsvLiveEnv(_golden_live_cfg(G, n_substeps=0), UsvLiveConfig(priv_mode=code, priv_a=a, priv_b=b, com_scale=scale, com_rand=False), n, DEV)
        env.reset_buf.zero_()
        env.reset_epoch.fill_(-1)
        for name, v in (("USV_C_MASS", G["cpl_mass"]), ("USV_C_KDRAG", G["cpl_kdrag"]), ("USV_C_THR_ML", G["cpl_thr"]), ("USV_C_THR_MR", G["cpl_thr"]),
                        ("USV_C_KIZ", G["cpl_kiz"]), ("USV_BC_COM_X", G["cpl_com"][:, 0]), ("USV_BC_COM_Y", G["cpl_com"][:, 1]),
                        ("USV_BC_COM_Z", G["cpl_com"][:, 2])):
            env.set_field(name, T(np.ascontiguousarray(v)))
        obs, _, _ = env.step(torch.zeros((n, 2), device=DEV), rebuild_scene=False)
        assert_close(obs[:, 25:33], G[f"priv_{mode}"], 1e-6, 1e-7, f"priv tail ({mode})")
        # mass.masscom_obs_source == "base": same envs, the tail shows base / neutral values  [ref: USV_Virtual.py:840-880]
        env.live = dataclasses.replace(env.live, masscom_obs_base=True)
        env._live_params = env.live.to_params()
        obs, _, _ = env.step(torch.zeros((n, 2), device=DEV), rebuild_scene=False)
        assert torch.equal(obs[:, 25:33].cpu(), T(G[f"priv_base_{mode}"])), mode


def test_scene_replay_npz_vs_oracle(tmp_path):
    """(f)3: the reference's NPZ scene format drives resets (goal, <= 16 obstacles with a count, start pose / velocity); the oracle
    applies the same scenes the way CaptureXYTask.apply_scene / _scene_replay_apply do."""
    from omniisaacgymenvs_loop_b200.scene_replay import SceneReplay
    S, n = 7, 20
    g = torch.Generator().manual_seed(17)
    goal = torch.rand((S, 2), generator=g) * 2 - 1
    ang = torch.rand(S, generator=g) * 6.28
    start = goal + torch.stack([torch.cos(ang), torch.sin(ang)], 1) * (9 + 3 * torch.rand((S, 1), generator=g))
    obst = torch.cat([goal.unsqueeze(1) + torch.rand((S, 10, 2), generator=g) * 24 - 12, torch.full((S, 10, 1), 2.0)], dim=-1)   # (S,k,3): z ignored
    count = torch.randint(3, 11, (S,), generator=g)
    path = str(tmp_path / "scenes.npz")
    np.savez(path, obstacles_xy=obst.numpy(), obstacles_count=count.numpy(), start_pos=start.numpy(), start_yaw=(torch.rand(S, generator=g) * 6 - 3).numpy(),
             start_vel=(torch.rand((S, 2), generator=g) - 0.5).numpy(), goal_pos=goal.numpy())
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=4, reset_pose_external=True, retarget_on_reset=False)
    live = UsvLiveConfig()
    env = FusedUsvLiveEnv(cfg, live, n, DEV)
    rp = SceneReplay(env, path, cycle=True)
    orc = B.LiveEnvOracle(oracle_cfg(cfg), oracle_task(cfg), oracle_live(live), n)
    # oracle-side scene application (the oracle's own reset_idx must not place obstacles: it is bypassed for the scene part)
    base_reset = B.ClassicEnvOracle.reset_idx
    nxt = torch.zeros(n, dtype=torch.long)

    def reset_with_scene(self, ids, step):
        if ids.numel() == 0:
            return
        idx = nxt[ids] % S
        nxt[ids] += 1
        pos, yaw, vel, gl, ob = rp.scenes(idx)
        self.outcome_at_reset = {}
        self.S.reset(ids)
        gids = self.env_ids[ids.numpy()]
        rc = torch.from_numpy(B.philox.uniform4(self.cfg.seed, gids, step, B.philox.RS_RESET_COM))
        self.com[ids] = torch.tensor(self.priv.com_base) + (rc[:, 0:3] * 2 - 1) * torch.tensor(self.priv.com_disp)
        self.target[ids], self.obstacles[ids] = gl, ob
        self.field[ids] = B.build_field(ob, gl)[0]
        base_reset(self, ids, step)                                    # dynamics DR + bookkeeping (pose is external)
        self.pos[ids], self.psi[ids], self.vel[ids], self.r[ids] = pos, yaw, vel, 0.0

    orc.reset_idx = reset_with_scene.__get__(orc)
    for k in range(11):
        act = torch.rand((n, 2), generator=g) * 2 - 1
        o_obs, o_rew, o_done = orc.step(act)
        obs, rew, done = rp.step(act.to(DEV))
        assert torch.equal(env.obstacles.cpu(), orc.obstacles), k
        assert_close(env.potential, orc.field, 1e-6, 1e-6, f"fields step {k}")
        assert_close(obs, o_obs, 1e-4, 2e-3, f"obs step {k}")
        assert_close(rew, o_rew, 1e-4, 5e-3, f"reward step {k}")
        assert torch.equal(done.cpu(), o_done), k
    assert int(rp.next_scene_idx.min()) >= 3 and torch.equal(rp.last_scene_idx, (nxt - 1) % S)


def test_live_capture_steps_equals_eager():
    """K live control steps (scene rebuilds included) replayed from one CUDA graph == the eager steps, bit for bit, over several replays."""
    n, K = 300, 6
    cfg = dataclasses.replace(LIVE_CFG, max_episode_length=9, action_bias_steps=0)
    g = torch.Generator().manual_seed(4)
    acts = (torch.rand((3 * K + 1, n, 2), generator=g) * 2 - 1).to(DEV)
    a, b = FusedUsvLiveEnv(cfg, UsvLiveConfig(), n, DEV), FusedUsvLiveEnv(cfg, UsvLiveConfig(), n, DEV)
    a.step(acts[0]); b.step(acts[0])
    slot = torch.zeros((K, n, 2), device=DEV)
    obs, rew, done = torch.zeros((K, n, 33), device=DEV), torch.zeros((K, n), device=DEV), torch.zeros((K, n), dtype=torch.long, device=DEV)
    replay = b.capture_steps(slot, obs, rew, done)
    for r in range(3):
        slot.copy_(acts[1 + r * K:1 + (r + 1) * K])
        replay()
        for k in range(K):
            o, w, d = a.step(acts[1 + r * K + k])
            assert torch.equal(o, obs[k]) and torch.equal(w, rew[k]) and torch.equal(d, done[k]), (r, k)
        assert a.step_counter == b.step_counter and torch.equal(a.state, b.state) and torch.equal(a.potential, b.potential)
    assert int(done.sum()) > 0


def test_live_four_wide_privileged_tail_vs_reference_golden(golden):
    """env.mass_dim / priv_dim = 4 [ref: OIGE/tasks/USV_Virtual.py:484-488,854-856 ; USV_core.py:23-52,127-170]: the observation is 29 wide
    (3 + 20 + 2 + [mass, CoM]); vs the reference's live CaptureXYTask built with priv_dim=4 on the states of
    tests/golden/capture_xy_live_pd4.npz (same scene as capture_xy_live.npz), and the drop-in surface reports the 29-wide space."""
    G, G8 = golden("capture_xy_live_pd4"), golden("capture_xy_live")
    K, n = G["pos"].shape[:2]
    assert G["obs"].shape == (K, n, 29)
    cfg = dataclasses.replace(LIVE_CFG, n_substeps=0, max_episode_length=10_000, mass_rand=False, mass_coupling=False, use_drag_scale=False,
                              reset_pose_external=True, retarget_on_reset=False, noise_vel=False, noise_heading=False, noise_pos=False,
                              pen_energy=OFF, pen_angular_vel=OFF, pen_angular_vel_variation=OFF)
    live = UsvLiveConfig(priv_mode=0, mass_obs_relative=False, com_obs_scaled=False, com_rand=False, priv_dim=4)
    env = FusedUsvLiveEnv(cfg, live, n, DEV)
    assert env.obs.shape == (n, 29)
    env.set_obstacles(T(G8["obstacles0"]))
    env.potential.copy_(T(G8["field0"]).to(DEV))
    env.set_field("USV_C_TX", T(G8["target"][:, 0])); env.set_field("USV_C_TY", T(G8["target"][:, 1]))
    env.step(torch.zeros((n, 2), device=DEV), rebuild_scene=False)        # the initial reset of every env (prev_action = 0 on that step)
    for k in range(K):
        env.reset_buf.zero_()
        env.reset_epoch.fill_(-1)
        for name, v in (("USV_S_X", G["pos"][k][:, 0]), ("USV_S_Y", G["pos"][k][:, 1]), ("USV_S_PSI", G["yaw"][k]),
                        ("USV_S_VX", G["vel"][k][:, 0]), ("USV_S_VY", G["vel"][k][:, 1]), ("USV_S_R", G["w"][k]),
                        ("USV_C_MASS", G["priv"][k][:, 0]), ("USV_BC_COM_X", G["priv"][k][:, 1]), ("USV_BC_COM_Y", G["priv"][k][:, 2]),
                        ("USV_BC_COM_Z", G["priv"][k][:, 3])):
            env.set_field(name, T(np.ascontiguousarray(v)))
        obs, rew, done = env.step(T(G["prev_action"][k]).to(DEV), rebuild_scene=False)
        assert obs.shape == (n, 29)
        assert_close(obs, T(G["obs"][k]).clamp(-cfg.clip_obs, cfg.clip_obs), 1e-5, 2e-6, f"29-wide obs step {k}")
    env.check_finite()
    # the drop-in surface: a task YAML with mass_dim 4 builds a 29-wide observation space (round 1 raised NotImplementedError here)
    from omniisaacgymenvs_loop_b200.config import live_default_config, live_task_cfg
    from scripts.train_loopz import make_env
    venv = make_env(live_task_cfg(live_default_config(num_envs=64), UsvLiveConfig(priv_dim=4)), DEV, seed=3)
    assert venv.num_obs == 29
    obs = venv.observe(as_numpy=False) if hasattr(venv, "observe") else None
    assert obs is None or obs.shape == (64, 29)
