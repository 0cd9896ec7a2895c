"""CPU suite: the Variant-B (live CaptureXY, static obstacles) oracle vs goldens produced by the reference's own task and
BatchedMapGPU (oracle/make_golden.py:variant_b)."""
import numpy as np
import torch

from oracle import usv_oracle_b as B

T = torch.from_numpy


def test_potential_field_builder_vs_reference(golden):
    G = golden("capture_xy_live")
    obst, target = T(G["obstacles0"]), T(G["target"])
    occ, sdf = B.occupancy_and_sdf(obst)
    assert torch.equal(occ.to(torch.uint8), T(G["occupancy0"])) and torch.equal(sdf, T(G["sdf0"]))
    cost = B.cost_to_go(occ, target)
    assert torch.equal(cost, T(G["cost0"]))                       # bit-exact incl. the +inf cells
    field = B.potential_field(cost, sdf)
    assert torch.allclose(field, T(G["field0"]), rtol=1e-6, atol=1e-7)


def test_grid_sample_restatement(golden):
    G = golden("capture_xy_live")
    field = T(G["field0"])
    for k in range(3):
        pos = T(G["pos"][k])
        got = B.sample_potential(field, pos)
        want = torch.nn.functional.grid_sample(field.unsqueeze(1), (2.0 * pos / 30.0).view(-1, 1, 1, 2), align_corners=False,
                                               padding_mode="border").view(-1)
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
        assert torch.allclose(got, T(G["potential"][k]), rtol=1e-6, atol=1e-7)


def test_live_task_vs_reference(golden):
    """obs (33) / reward terms / kills / outcome latches over 6 steps with a reset batch (prev_potential=None quirk)."""
    G = golden("capture_xy_live")
    c = B.LiveTaskConfig()
    K, n = G["pos"].shape[:2]
    S = B.LiveRewardState(n)
    obst, field, target = T(G["obstacles0"]).clone(), T(G["field0"]).clone(), T(G["target"])
    for k in range(K):
        if k == int(G["reset_step"]):
            ids = T(G["reset_ids"])
            S.reset(ids)
            obst, field = T(G["obstacles1"]).clone(), T(G["field1"]).clone()
        yaw = T(G["yaw"][k])
        state = {"position": T(G["pos"][k]), "orientation": torch.stack([torch.cos(yaw), torch.sin(yaw)], 1),
                 "linear_velocity": T(G["vel"][k]), "angular_velocity": T(G["w"][k])}
        obs, aux = B.live_observation(state, target, obst, T(G["prev_action"][k]), T(G["priv"][k]))
        assert torch.allclose(obs, T(G["obs"][k]), rtol=1e-6, atol=1e-6), k
        out = B.live_reward(c, S, aux, state, obst, field)
        for name in ("distance_reward", "alignment_reward", "potential_shaping", "turn_hazard", "speed_reward", "angular_reward",
                     "heading_improve", "collision_penalty", "goal_reward", "danger"):
            # the shaping term is 100 x (difference of two nearly equal potentials): 1e-7 relative on the potential -> 1e-5 absolute
            atol = 5e-5 if name == "potential_shaping" else 1e-6
            assert torch.allclose(out[name], T(G[name][k]), rtol=1e-5, atol=atol), (k, name)
        assert torch.allclose(out["reward"], T(G["reward"][k]), rtol=1e-5, atol=1e-4), k
        die = B.live_kills(c, S, aux["d"], state["position"], obst)
        assert torch.equal(die, T(G["die"][k])) and torch.equal(S.goal_reached, T(G["goal_reached"][k]))
        assert torch.equal(S.done_success, T(G["done_success"][k])) and torch.equal(S.done_collision, T(G["done_collision"][k]))
    # the crafted rows did their job: a collision, a goal capture, a distance kill
    assert G["done_collision"][0][0] == 1 and G["done_success"][0][1] == 1 and G["die"][0][2] == 1
    # quirk 8: on the step after the reset batch the potential shaping is exactly zero for EVERY env
    assert np.all(G["potential_shaping"][int(G["reset_step"])] == 0.0)


# ---- the live USVVirtual's own host chains (goldens: oracle/make_golden.py:live_virtual) -------------------------------------
def _live_oracle_cfg(G, **kw):
    import dataclasses
    from oracle.usv_oracle import EnvConfig
    return dataclasses.replace(
        EnvConfig(), action_affine=True, penalties_use_u=True, action_noise=False, noise_vel=False, noise_heading=False, n_substeps=0,
        reset_pose_external=True, n_lut=int(G["n_lut"]), lut_points_left=tuple(G["lut_points_left"].tolist()),
        lut_points_right=tuple(G["lut_points_right"].tolist()), mass_base=34.96, couple_mass_max=54.96, couple_thr_a=0.5,
        kdrag_min=1.0, kdrag_max=1.5, couple_kiz_min=1.0, couple_kiz_max=1.5, **kw)


def test_live_action_path_vs_reference(golden):
    """A12: bias (first N steps) -> clamp -> affine map to [0,1] -> LUT, resets zero the command and prev_action."""
    from oracle.usv_oracle import ClassicEnvOracle
    G = golden("live_virtual")
    ids = T(G["act_reset_ids"])
    for tag, steps in (("bias", 5), ("nobias", 0)):
        n = G[f"act_{tag}_in"].shape[0]
        orc = ClassicEnvOracle(_live_oracle_cfg(G, action_bias=float(G["act_bias"]), action_bias_steps=steps), n)
        orc.reset_buf[:] = 0
        orc.reset_buf[ids] = 1
        _, dyn = orc.dynamics(T(G[f"act_{tag}_in"]))
        prev = dyn["raw_actions"].clone()
        prev[ids] = 0
        assert torch.equal(prev, T(G[f"act_{tag}_prev"]))
        assert torch.equal(dyn["before_rect"], T(G[f"act_{tag}_before_rect"]))
        assert torch.equal(dyn["unit"], T(G[f"act_{tag}_unit"]))
        assert torch.equal(dyn["target"], T(G[f"act_{tag}_target"]))
    # a reset env commands u = 0, i.e. the LUT entry of the mid-range command (not exactly 0 N)
    assert len(np.unique(G["act_bias_target"][G["act_reset_ids"]], axis=0)) == 1


def test_mass_coupling_and_priv_tail_vs_reference(golden):
    from oracle.usv_oracle import mass_coupling
    G = golden("live_virtual")
    n = G["cpl_mass"].shape[0]
    cfg = _live_oracle_cfg(G)
    kd, sthr, kiz = mass_coupling(cfg, T(G["cpl_mass"]))
    assert torch.equal(kd, T(G["cpl_kdrag"])) and torch.equal(sthr, T(G["cpl_thr"])) and torch.equal(kiz, T(G["cpl_kiz"]))
    scale = tuple(float(x) for x in G["cpl_com_scale"])
    for mode, code, a, b in (("minmax", 2, (1.0, 0.5, 0.5, 1.0), (0.5, 0.5, 0.5, 0.5)), ("centered", 1, (1.0,) * 4, (0.5,) * 4),
                             ("raw", 0, (0.0,) * 4, (1.0,) * 4)):
        orc = B.LiveEnvOracle(cfg, B.LiveTaskConfig(), B.LivePrivConfig(priv_mode=code, priv_a=a, priv_b=b, com_scale=scale), n)
        orc.mass, orc.com = T(G["cpl_mass"]).clone(), T(G["cpl_com"]).clone()
        orc.drag_scale[:, 0], orc.thr_mult_left, orc.thr_mult_right, orc.k_iz = kd, sthr.clone(), sthr.clone(), kiz
        assert torch.allclose(orc.priv_tail(), T(G[f"priv_{mode}"]), rtol=1e-6, atol=1e-7), mode
        # mass.masscom_obs_source == "base": base encodings / neutral parameters for every env  [ref: USV_Virtual.py:840-880]
        orc.priv = B.LivePrivConfig(priv_mode=code, priv_a=a, priv_b=b, com_scale=scale, masscom_obs_base=True)
        assert torch.equal(orc.priv_tail(), T(G[f"priv_base_{mode}"])), mode
    # a non-zero base CoM and raw encodings: the tail is the constants themselves
    orc.priv = B.LivePrivConfig(priv_mode=2, priv_a=(1.0, 0.5, 0.5, 1.0), priv_b=(0.5, 0.5, 0.5, 0.5), mass_obs_relative=False, com_obs_scaled=False,
                                com_base=(0.1, -0.02, 0.03), masscom_obs_base=True)
    want = torch.tensor([cfg.mass_base, 0.1, -0.02, 0.03, 0.0, 0.0, 0.0, 0.0]).repeat(n, 1)
    assert torch.equal(orc.priv_tail(), want)
