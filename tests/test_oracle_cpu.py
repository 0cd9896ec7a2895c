"""CPU suite: the oracle restatement vs the golden vectors produced by the reference itself
(oracle/make_golden.py), plus the pins of the third-party boundaries (Philox, quaternion_to_matrix)."""
import math

import numpy as np
import pytest
import torch

from oracle import philox, usv_oracle as O, integrator64

T = torch.from_numpy


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_philox_uniform_range():
    u = philox.uniform4(99, np.arange(100000), 3, 1)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(float(u.mean()) - 0.5) < 5e-3


def test_quaternion_to_matrix_pins():
    # unit and non-unit quaternions: R must be orthonormal with det +1 and match the axis-angle form
    g = torch.Generator().manual_seed(0)
    q = torch.randn((64, 4), generator=g, dtype=torch.float64)
    q[0] = torch.tensor([1.0, 0, 0, 0], dtype=torch.float64)
    R = O.quaternion_to_matrix(q)
    eye = torch.eye(3, dtype=torch.float64).expand(64, 3, 3)
    assert torch.allclose(R @ R.mT, eye, atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64, dtype=torch.float64), atol=1e-12)
    assert torch.equal(R[0], torch.eye(3, dtype=torch.float64))
    yaw = 0.7
    Rz = O.quaternion_to_matrix(torch.tensor([[math.cos(yaw / 2), 0, 0, math.sin(yaw / 2)]], dtype=torch.float64))[0]
    want = torch.tensor([[math.cos(yaw), -math.sin(yaw), 0], [math.sin(yaw), math.cos(yaw), 0], [0, 0, 1]], dtype=torch.float64)
    assert torch.allclose(Rz, want, atol=1e-15)


def test_hydrostatics_vs_reference(golden):
    G = golden("force_modules")
    out, fg, tg = O.hydrostatics_local(T(G["vol"]), T(G["rpy"]), T(G["quat"]), water_density=1000, gravity=-9.81,
                                       metacentric_width=0.5, metacentric_length=0.65,
                                       average_hydrostatics_force_value=275, amplify_torque=1.0)
    assert torch.equal(out, T(G["hs_out"]))
    assert torch.equal(fg, T(G["hs_force_global"])) and torch.equal(tg, T(G["hs_torque_global"]))
    out2, _, _ = O.hydrostatics_local(T(G["vol"]), T(G["rpy"]), T(G["quat"]), water_density=1025.0, gravity=-9.80665,
                                      metacentric_width=0.4, metacentric_length=0.7,
                                      average_hydrostatics_force_value=300.0, amplify_torque=2.5)
    assert torch.equal(out2, T(G["hs_out_alt"]))


def test_hydrostatics_known_answer():
    # SURVEY appendix D.2
    q = torch.tensor([[1.0, 0, 0, 0], [math.cos(math.pi / 6), 0, 0, math.sin(math.pi / 6)]])
    out, _, _ = O.hydrostatics_local(torch.tensor([0.035, 0.02]), torch.tensor([[0.05, -0.02, 0], [0, 0, math.pi / 3]]), q,
                                     water_density=1000, gravity=-9.81, metacentric_width=0.5, metacentric_length=0.65,
                                     average_hydrostatics_force_value=275, amplify_torque=1.0)
    want = torch.tensor([[0, 0, 343.35000610, -6.87213564, 3.57476163, 0], [0, 0, 196.19999695, 0, 0, 0]])
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-6)


HD_BASE = dict(linear_damping_forward_speed=[0.0] * 6, offset_linear_damping=0.0, offset_lin_forward_damping_speed=0.0,
               offset_nonlin_damping=0.0, scaling_damping=1.0, use_drag_scale=False)


def test_hydrodynamics_vs_reference(golden):
    G = golden("force_modules")
    n = G["quat"].shape[0]
    lin = torch.tensor([[0.0, 99.99, 99.99, 13.0, 13.0, 0.82985084]] * n)
    quad = torch.tensor([[17.257603, 99.99, 10.0, 5.0, 5.0, 17.33600724]] * n)
    kd = torch.ones((n, 1))
    drag, vel = O.hydrodynamics(T(G["quat"]), T(G["vel6"]), lin, quad, kd, **HD_BASE)
    assert torch.equal(drag, T(G["hd_drag"])) and torch.equal(vel, T(G["hd_local_vel"]))
    drag, _ = O.hydrodynamics(T(G["quat_planar"]), T(G["vel6_planar"]), lin, quad, kd, **HD_BASE)
    assert torch.equal(drag, T(G["hd_drag_planar"]))
    drag, _ = O.hydrodynamics(T(G["quat"]), T(G["vel6"]), lin, quad, kd, **{**HD_BASE, "use_water_current": True,
                                                                          "flow_vel": [0.3, -0.2, 0.05]})
    assert torch.equal(drag, T(G["hd_drag_current"]))
    P2 = dict(linear_damping_forward_speed=[0.1, 0.2, 0.0, 0.0, 0.0, 0.05], offset_linear_damping=0.5,
              offset_lin_forward_damping_speed=0.25, offset_nonlin_damping=0.125, scaling_damping=1.25, use_drag_scale=True)
    drag, _ = O.hydrodynamics(T(G["quat"]), T(G["vel6"]), T(G["hd2_lin"]), T(G["hd2_quad"]), T(G["hd2_kdrag"]), **P2)
    assert torch.equal(drag, T(G["hd2_drag"]))
    # SURVEY appendix D.1 rows
    assert np.allclose(G["hd_drag"][0], [-17.25760269, -74.99250031, 0, 0, 0, -1.80919611], rtol=1e-6)
    assert np.allclose(G["hd_drag"][1], [-12.15466404, -184.19181824, 0, 0, 0, 9.07553959], rtol=1e-5)


@pytest.mark.parametrize("name", ["classic", "live", "nominal"])
def test_thruster_vs_reference(golden, name):
    G = golden("force_modules")
    lutL = O.build_lut(G[f"lut_{name}_points_left"], 1000)
    lutR = O.build_lut(G[f"lut_{name}_points_right"], 1000)
    assert torch.equal(lutL, T(G[f"lut_{name}_left"])) and torch.equal(lutR, T(G[f"lut_{name}_right"]))
    # the ATen-free restatement of F.interpolate must give the same table bit for bit
    assert np.array_equal(O.build_lut_restated(G[f"lut_{name}_points_left"], 1000), G[f"lut_{name}_left"])
    assert np.array_equal(O.build_lut_restated(G[f"lut_{name}_points_right"], 1000), G[f"lut_{name}_right"])
    before, _ = O.thruster_target(T(G["thr_cmd"]), lutL, lutR)
    assert torch.equal(before, T(G[f"thr_{name}_before"]))
    cur = torch.zeros_like(before)
    alpha = O.lag_alpha(0.02, 0.05)
    assert float(alpha) == float(G["thr_alpha"])
    for k in range(6):
        cur = O.thruster_lag(cur, before, alpha)
        assert torch.equal(cur, T(G[f"thr_{name}_lag6"][k][:, [0, 3]]))


def test_thruster_known_answers(golden):
    G = golden("force_modules")
    lut = G["lut_live_left"]
    assert np.allclose(lut[499:503], [0, 0.08007812, 0.24023438, 0.40039825], atol=1e-7)
    assert np.isclose(lut[864], 58.37837219) and lut[999] == 80.0
    assert np.allclose(G["thr_live_before"][:2], [[58.37837219, 39.95995331], [0.08007812, 80.0]])
    assert np.allclose(G["thr_live_lag6"][1][:2][:, [0, 3]], [[32.14727783, 22.00478935], [0.04409670, 44.05368423]])


def test_thruster_multipliers(golden):
    G = golden("force_modules")
    lutL, lutR = T(G["lut_classic_left"]), T(G["lut_classic_right"])
    _, after = O.thruster_target(T(G["thr_cmd"]), lutL, lutR, T(G["thr_mult_left"]), T(G["thr_mult_right"]))
    assert torch.equal(after, T(G["thr_sep_after"]))


def test_disturbances_vs_reference(golden):
    G = golden("disturbances")
    cfg = O.EnvConfig().full_dr()
    assert np.allclose(cfg.force_ranges, G["ranges"])
    n = G["root_pos"].shape[0]
    env = O.ClassicEnvOracle(cfg, n)
    env.f_const = T(G["f_const"][:, :2]).clone()
    env.f_freq = torch.stack([T(G["fxf"]), T(G["fyf"])], 1)
    env.f_shift = torch.stack([T(G["fxs"]), T(G["fys"])], 1)
    env.f_amp = T(G["famp"]).clone()
    env.t_const = T(G["t_const"][:, 2]).clone()
    env.t_freq, env.t_shift, env.t_amp = T(G["tf"]).clone(), T(G["ts"]).clone(), T(G["tamp"]).clone()
    env.pos = T(G["root_pos"][:, :2]).clone()
    env.vel[:] = 0
    env.current_forces[:] = 0
    _, Fx, Fy, Tz, *_ = env.planar_wrench()
    assert torch.equal(Fx, T(G["forces"][:, 0])) and torch.equal(Fy, T(G["forces"][:, 1]))
    assert torch.equal(Tz, T(G["torques"][:, 2]))


def test_classic_task_vs_reference(golden):
    """obs / reward / penalties / kills of the classic CaptureXYTask over a 6-step trajectory with a reset."""
    G = golden("capture_xy_classic")
    c = O.EnvConfig()
    K, n = G["pos"].shape[:2]
    target = T(G["target"])
    goal = torch.zeros(n, dtype=torch.int32)
    prev_d = prev_w = prev_asum = None
    for k in range(K):
        yaw = T(G["yaw"][k])
        state = {"position": T(G["pos"][k]), "orientation": torch.stack([torch.cos(yaw), torch.sin(yaw)], 1),
                 "linear_velocity": T(G["vel"][k]), "angular_velocity": T(G["w"][k])}
        just_reset = T(G["reset_ids"]) if k == int(G["reset_step"]) else (torch.arange(n) if k == 0 else torch.tensor([], dtype=torch.long))
        if k == int(G["reset_step"]):
            goal[just_reset] = 0
        obs, aux = O.capture_xy_observation(state, target)
        out = O.capture_xy_reward(c, aux, state, goal, aux["d"] if prev_d is None else prev_d, just_reset)
        pen = O.penalties(c, state, T(G["actions"][k]), prev_w, prev_asum, first_call=(k == 0))
        die = O.capture_xy_kills(c, aux["d"], out["speed"], goal)
        assert torch.equal(obs, T(G["obs"][k])), k
        assert torch.equal(out["reward"], T(G["reward"][k])), k
        assert torch.equal(out["distance_reward"], T(G["distance_reward"][k]))
        assert torch.equal(out["alignment_reward"], T(G["alignment_reward"][k]))
        assert torch.equal(out["speed_reward"], T(G["speed_reward"][k]))
        assert torch.equal(pen["total"], T(G["penalty"][k])), k
        assert torch.equal(die, T(G["die"][k])) and die.dtype == torch.long
        assert torch.equal(goal, T(G["goal_reached"][k]))
        prev_d, prev_w, prev_asum = aux["d"], state["angular_velocity"], pen["asum"]
    # crafted rows did what they were crafted for
    assert G["die"][0][0] == 1 and G["goal_reached"][2][0] == 3      # in tolerance and slow: counter runs
    assert G["die"][0][1] == 1                                       # beyond kill_dist
    assert G["goal_reached"][0][2] == 0                              # in tolerance but too fast


def test_classic_known_answer_d4():
    """SURVEY appendix D.4 (hand-checked vectors produced from the reference)."""
    c = O.EnvConfig()
    target = torch.zeros((3, 2))
    goal = torch.zeros(3, dtype=torch.int32)

    def st(pos, yaw, v, w):
        yaw = torch.tensor(yaw)
        return {"position": torch.tensor(pos), "orientation": torch.stack([torch.cos(yaw), torch.sin(yaw)], 1),
                "linear_velocity": torch.tensor(v), "angular_velocity": torch.tensor(w)}

    s0 = st([[5, 2], [0.05, 0.02], [-21, 0.0]], [0.3, -2.0, 1.0], [[1, 0.2], [0.01, 0.02], [0, 0.0]], [0.1, -0.2, 0])
    obs0, aux0 = O.capture_xy_observation(s0, target)
    r0 = O.capture_xy_reward(c, aux0, s0, goal, aux0["d"], torch.tensor([], dtype=torch.long))
    assert torch.allclose(obs0[0], torch.tensor([1.01444054, -0.10445291, 0.1, -0.99676108, -0.08041954, 5.38516474,
                                                  1.0, 0.2, 0, 1.0, 0.2, 0, 0]), atol=1e-6)
    assert torch.allclose(r0["reward"], torch.tensor([-0.09216417, 30.16174507, -0.18176800]), atol=1e-5)
    assert O.capture_xy_kills(c, aux0["d"], r0["speed"], goal).tolist() == [0, 1, 1] and goal.tolist() == [0, 1, 0]
    s1 = st([[4.8, 1.9], [0.04, 0.02], [-21.1, 0.0]], [0.35, -2.0, 1.0], [[1.1, 0.2], [0.01, 0.02], [0, 0.0]], [0.12, -0.2, 0])
    _, aux1 = O.capture_xy_observation(s1, target)
    r1 = O.capture_xy_reward(c, aux1, s1, goal, aux0["d"], torch.tensor([], dtype=torch.long))
    assert torch.allclose(r1["reward"], torch.tensor([0.13038145, 60.17282486, -0.28176838]), atol=1e-5)
    assert goal.tolist() == [0, 2, 0]


def test_heading_error_fmod_quirk():
    """fmod keeps the sign of the dividend: beta-theta < -pi is NOT wrapped (reference quirk)."""
    yaw = torch.tensor([3.0])
    s = {"position": torch.tensor([[1.0, 0.1411]]), "orientation": torch.stack([torch.cos(yaw), torch.sin(yaw)], 1),
         "linear_velocity": torch.zeros((1, 2)), "angular_velocity": torch.zeros(1)}
    _, aux = O.capture_xy_observation(s, torch.zeros((1, 2)))   # beta ~ -3.0
    assert float(aux["herr"]) > math.pi


def test_env_oracle_runs_and_is_deterministic():
    cfg = O.EnvConfig().full_dr()
    a = O.ClassicEnvOracle(cfg, 64)
    b = O.ClassicEnvOracle(cfg, 64)
    g = torch.Generator().manual_seed(3)
    for _ in range(5):
        act = torch.rand((64, 2), generator=g) * 2 - 1
        oa, ra, da = a.step(act)
        ob, rb, db = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    assert torch.isfinite(oa).all() and oa.shape == (64, 13) and da.dtype == torch.long
    # sharding invariance: envs 32..63 of the big batch == a 32-env oracle with env_id_offset=32
    c = O.ClassicEnvOracle(cfg, 32, env_id_offset=32)
    g = torch.Generator().manual_seed(3)
    for _ in range(5):
        act = torch.rand((64, 2), generator=g) * 2 - 1
        oc, rc, dc = c.step(act[32:])
    assert torch.equal(oc, oa[32:]) and torch.equal(rc, ra[32:])


def test_fp32_substeps_vs_float64_integration():
    """The fp32 oracle's sub-step loop against the float64 host integration (north_star)."""
    cfg = O.EnvConfig().full_dr()
    n = 256
    env = O.ClassicEnvOracle(cfg, n)
    g = torch.Generator().manual_seed(11)
    env.step(torch.rand((n, 2), generator=g) * 2 - 1)     # resets + randomises every env
    for _ in range(3):
        env.step(torch.rand((n, 2), generator=g) * 2 - 1)
    cmd = torch.rand((n, 2), generator=g) * 2 - 1
    _, target = O.thruster_target(cmd, env.lut_left, env.lut_right, env.thr_mult_left, env.thr_mult_right)
    d = lambda t: t.double().numpy().copy()
    state = dict(x=d(env.pos[:, 0]), y=d(env.pos[:, 1]), psi=d(env.psi), vx=d(env.vel[:, 0]), vy=d(env.vel[:, 1]),
                 r=d(env.r), thrL=d(env.current_forces[:, 0]), thrR=d(env.current_forces[:, 1]))
    const = dict(mass=d(env.mass), lin=d(env.linear_damping[:, [0, 1, 5]]), quad=d(env.quadratic_damping[:, [0, 1, 5]]),
                 kdrag=d(env.drag_scale[:, 0]), kiz=d(env.k_iz), fcx=d(env.f_const[:, 0]), fcy=d(env.f_const[:, 1]),
                 fxf=d(env.f_freq[:, 0]), fyf=d(env.f_freq[:, 1]), fxs=d(env.f_shift[:, 0]), fys=d(env.f_shift[:, 1]),
                 famp=d(env.f_amp), tc=d(env.t_const), tf=d(env.t_freq), ts=d(env.t_shift), tamp=d(env.t_amp))
    ref = integrator64.substeps(state, const, d(target), dt=cfg.dt, alpha=float(env.alpha.double()), n_substeps=50,
                                izz=cfg.izz, thr_y_left=cfg.thr_y_left, thr_y_right=cfg.thr_y_right,
                                use_const_force=True, use_sin_force=True, use_const_torque=True, use_sin_torque=True)
    for _ in range(50):
        env.substep(target)
    for name, got in (("x", env.pos[:, 0]), ("y", env.pos[:, 1]), ("psi", env.psi), ("vx", env.vel[:, 0]),
                      ("vy", env.vel[:, 1]), ("r", env.r)):
        err = np.abs(got.double().numpy() - ref[name])
        assert err.max() < 2e-4, (name, err.max())


def test_spawn_and_kill_curriculum_vs_reference(golden):
    """A18/A19 optional curriculum: spawn annulus and kill distance as functions of the task `step`, against the reference's
    get_spawns / update_kills driven at 9 steps around the warm-up / end knees; the product's host function must agree too."""
    import dataclasses
    from omniisaacgymenvs_loop_b200.config import UsvEnvConfig
    G = golden("classic_curriculum")
    cmin, cmax, ckill, warm, end, rmin, rmax, kill = G["params"].tolist()
    kw = dict(spawn_curriculum=True, spawn_curriculum_min_dist=cmin, spawn_curriculum_max_dist=cmax, spawn_curriculum_kill_dist=ckill,
              spawn_curriculum_warmup=int(warm), spawn_curriculum_end=int(end), spawn_min_dist=rmin, spawn_max_dist=rmax, kill_dist=kill)
    c, pc = dataclasses.replace(O.EnvConfig(), **kw), UsvEnvConfig(**kw)
    for k, st in enumerate(G["steps"].tolist()):
        lo, hi, kd = O.curriculum(c, st)
        assert (lo, hi, kd) == pc.curriculum(st)
        r = T(G["u"][k]) * (hi - lo) + lo
        assert torch.allclose(r, T(G["r"][k]), rtol=1e-5, atol=1e-5), (st, lo, hi)
        assert torch.equal((T(G["dist"]) > kd).long(), T(G["die"][k])), (st, kd)
    assert O.curriculum(dataclasses.replace(c, spawn_curriculum=False), 10.0) == (rmin, rmax, kill)
