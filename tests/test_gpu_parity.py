"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against (i) the golden vectors produced by the reference itself and (ii) the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): forces / observations / rewards within 1e-5 relative
in fp32 (atol scaled to the quantity, written at each assert); done / capture / reset-index outputs
bit-exact; the integrator against a float64 host integration."""
import ctypes
import dataclasses
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200 import _lib  # noqa: E402
from omniisaacgymenvs_loop_b200.config import UsvEnvConfig  # noqa: E402
from omniisaacgymenvs_loop_b200.engine import FusedUsvEnv  # noqa: E402
from omniisaacgymenvs_loop_b200.envs.USV.Hydrodynamics import HydrodynamicsObject  # noqa: E402
from omniisaacgymenvs_loop_b200.envs.USV.Hydrostatics import HydrostaticsObject  # noqa: E402
from omniisaacgymenvs_loop_b200.envs.USV.ThrusterDynamics import DynamicsFirstOrder  # noqa: E402
from oracle import integrator64, usv_oracle as O  # noqa: E402
from tests.util import assert_close, engine_state, oracle_cfg, oracle_state, push_oracle_state  # noqa: E402

DEV = "cuda:0"
RTOL = 1e-5
cu = lambda a: torch.as_tensor(a).to(DEV)

DRAG_CFG = dict(use_drag_randomization=False, u_linear_rand=0.1, v_linear_rand=0.1, w_linear_rand=0.0, p_linear_rand=0.0,
                q_linear_rand=0.0, r_linear_rand=0.1, u_quad_rand=0.1, v_quad_rand=0.1, w_quad_rand=0.0, p_quad_rand=0.0,
                q_quad_rand=0.0, r_quad_rand=0.1)
THR_CFG = dict(use_thruster_randomization=False, thruster_rand=0.5, use_separate_randomization=False, left_rand=0.5, right_rand=0.5)
LIN = [0.0, 99.99, 99.99, 13.0, 13.0, 0.82985084]
QUAD = [17.257603, 99.99, 10.0, 5.0, 5.0, 17.33600724]


# ------------------------------------------------------------------------------------------
# force layer vs the reference's goldens (through the reference-shaped classes)
def test_hydrostatics_vs_reference_golden(golden):
    G = golden("force_modules")
    n = G["vol"].shape[0]
    H = HydrostaticsObject(n, DEV, 1000, -9.81, 0.5, 0.65, 275, 1.0, 0.0, 1.0, 0.3, -10.0)
    out = H.compute_archimedes_metacentric_local(cu(G["vol"]), cu(G["rpy"]), cu(G["quat"]))
    assert out.shape == (n, 6) and out.dtype == torch.float32
    assert_close(out, G["hs_out"], RTOL, 1e-4, "hydrostatics local")         # forces O(100 N): atol 1e-4 = 1e-6 relative
    assert_close(H.archimedes_force_global, G["hs_force_global"], RTOL, 1e-5)
    assert_close(H.archimedes_torque_global, G["hs_torque_global"], RTOL, 1e-5)
    out = H.compute_archimedes_metacentric_local(cu(G["vol"]), cu(G["rpy"]) * 0, cu(G["quat_planar"]))
    assert_close(out, G["hs_out_planar"], RTOL, 1e-4)
    H2 = HydrostaticsObject(n, DEV, 1025.0, -9.80665, 0.4, 0.7, 300.0, 2.5, 0.0, 1.0, 0.3, -10.0)
    assert_close(H2.compute_archimedes_metacentric_local(cu(G["vol"]), cu(G["rpy"]), cu(G["quat"])), G["hs_out_alt"], RTOL, 1e-4)
    fg, tg = H.compute_archimedes_metacentric_global(cu(G["vol"]), cu(G["rpy"]))
    assert_close(fg, G["hs_force_global"], RTOL, 1e-5)


def _hydro(n, cfg_extra=None, **kw):
    cfg = dict(DRAG_CFG)
    cfg.update(cfg_extra or {})
    args = dict(task_cfg=cfg, num_envs=n, device=DEV, water_density=1000, gravity=-9.81, linear_damping=LIN, quadratic_damping=QUAD,
                linear_damping_forward_speed=[0.0] * 6, offset_linear_damping=0.0, offset_lin_forward_damping_speed=0.0,
                offset_nonlin_damping=0.0, scaling_damping=1.0, offset_added_mass=0.0, scaling_added_mass=1.0, alpha=0.3, last_time=-10.0)
    args.update(kw)
    return HydrodynamicsObject(**args)


def test_hydrodynamics_vs_reference_golden(golden):
    G = golden("force_modules")
    n = G["quat"].shape[0]
    D = _hydro(n)
    # drag is a product of O(100) coefficients and O(1) velocities that are themselves differences
    # (R^T v): atol 2e-4 N ~ 1e-6 of the force scale
    assert_close(D.ComputeHydrodynamicsEffects(0.01, cu(G["quat"]), cu(G["vel6"]), False, [0.0, 0.0, 0.0]), G["hd_drag"], RTOL, 2e-4, "drag")
    assert_close(D.local_velocities, G["hd_local_vel"], RTOL, 1e-6, "local vel")
    assert_close(D.ComputeHydrodynamicsEffects(0.01, cu(G["quat_planar"]), cu(G["vel6_planar"]), False, [0.0, 0.0, 0.0]),
                 G["hd_drag_planar"], RTOL, 2e-4, "drag planar")
    assert_close(D.ComputeHydrodynamicsEffects(0.01, cu(G["quat"]), cu(G["vel6"]), True, [0.3, -0.2, 0.05]), G["hd_drag_current"], RTOL, 2e-4)
    D2 = _hydro(n, dict(use_drag_scale_randomization=True, k_drag_min=1.0, k_drag_max=1.5),
                linear_damping_forward_speed=[0.1, 0.2, 0.0, 0.0, 0.0, 0.05], offset_linear_damping=0.5,
                offset_lin_forward_damping_speed=0.25, offset_nonlin_damping=0.125, scaling_damping=1.25)
    assert float(D2.drag_scale.min()) >= 1.0 and float(D2.drag_scale.max()) <= 1.5      # ctor sampled k_drag in range
    D2.linear_damping[:] = cu(G["hd2_lin"]); D2.quadratic_damping[:] = cu(G["hd2_quad"]); D2.drag_scale[:] = cu(G["hd2_kdrag"])
    assert_close(D2.ComputeHydrodynamicsEffects(0.01, cu(G["quat"]), cu(G["vel6"]), False, [0.0, 0.0, 0.0]), G["hd2_drag"], RTOL, 3e-4)
    assert_close(D2.ComputeHydrodynamicsEffects(0.01, cu(G["quat_planar"]), cu(G["vel6_planar"]), False, [0, 0, 0]), G["hd2_drag_planar"], RTOL, 3e-4)
    # ComputeDampingMatrix(vel) * vel * -1 == drag for identity attitude
    v = cu(G["vel6"])
    Dm = D2.ComputeDampingMatrix(v)
    ident = torch.zeros((n, 4), device=DEV); ident[:, 0] = 1
    assert_close(-Dm * v, D2.ComputeHydrodynamicsEffects(0.01, ident, v, False, [0, 0, 0]), 1e-6, 1e-6)


def test_hydrodynamics_reset_coefficients_ranges():
    n = 4096
    D = _hydro(n, dict(use_drag_randomization=True, use_drag_scale_randomization=True, k_drag_min=0.5, k_drag_max=2.0,
                       k_drag_sample_space="log"))
    base_l, base_q = torch.tensor(LIN, device=DEV), torch.tensor(QUAD, device=DEV)
    ids = torch.arange(0, n, 2, device=DEV)
    before = D.linear_damping.clone()
    D.reset_coefficients(ids, ids.numel())
    frac = torch.tensor([0.1, 0.1, 0, 0, 0, 0.1], device=DEV)
    assert ((D.linear_damping - base_l).abs() <= frac * base_l + 1e-6).all()
    assert ((D.quadratic_damping - base_q).abs() <= frac * base_q + 1e-6).all()
    assert torch.equal(D.linear_damping[1::2], before[1::2]) and not torch.equal(D.linear_damping[0::2], before[0::2])
    assert float(D.drag_scale.min()) >= 0.5 and float(D.drag_scale.max()) <= 2.0
    assert abs(float(torch.log(D.drag_scale).mean())) < 0.05       # log-uniform on [0.5, 2] is centred on 0


@pytest.mark.parametrize("name", ["classic", "live", "nominal"])
def test_thruster_vs_reference_golden(golden, name):
    G = golden("force_modules")
    n = G["thr_cmd"].shape[0]
    T = DynamicsFirstOrder(dict(THR_CFG), n, DEV, 0.05, 0.02, 1000, G[f"lut_{name}_points_left"].tolist(),
                           G[f"lut_{name}_points_right"].tolist(), [0.0] * 5, [0.0] * 5, -1.0, 1.0)
    # LUT and LUT lookup are bit-exact (index work), lag is evaluated op by op like torch -> bit-exact too
    assert np.array_equal(T.y_linear_interp_left.cpu().numpy(), G[f"lut_{name}_left"])
    assert np.array_equal(T.y_linear_interp_right.cpu().numpy(), G[f"lut_{name}_right"])
    T.set_target_force(cu(G["thr_cmd"]))
    assert np.array_equal(T.thruster_forces_before_dynamics.cpu().numpy(), G[f"thr_{name}_before"])
    for k in range(6):
        thr = T.update_forces()
        assert np.array_equal(thr.cpu().numpy(), G[f"thr_{name}_lag6"][k]), k


def test_thruster_multipliers_vs_reference_golden(golden):
    G = golden("force_modules")
    n = G["thr_cmd"].shape[0]
    T = DynamicsFirstOrder(dict(THR_CFG, use_thruster_randomization=True, use_separate_randomization=True), n, DEV, 0.05, 0.02, 1000,
                           G["lut_classic_points_left"].tolist(), G["lut_classic_points_right"].tolist(), [0.0] * 5, [0.0] * 5, -1.0, 1.0)
    assert float(T.thruster_left_multiplier.min()) >= 0.5 and float(T.thruster_left_multiplier.max()) <= 1.5
    T.thruster_left_multiplier[:] = cu(G["thr_mult_left"]); T.thruster_right_multiplier[:] = cu(G["thr_mult_right"])
    T.set_target_force(cu(G["thr_cmd"]))
    assert np.array_equal(T.thruster_forces_after_randomization.cpu().numpy(), G["thr_sep_after"])
    for k in range(3):
        assert np.array_equal(T.update_forces().cpu().numpy(), G["thr_sep_lag3"][k])


# ------------------------------------------------------------------------------------------
# fused env step vs the oracle
def _lockstep(cfg, n, steps, seed=5, sync=True, collect_stats=False):
    env = FusedUsvEnv(cfg, n, DEV, collect_stats=collect_stats)
    orc = O.ClassicEnvOracle(oracle_cfg(cfg), n)
    g = torch.Generator().manual_seed(seed)
    for k in range(steps):
        act = torch.rand((n, 2), generator=g) * 2.4 - 1.2          # beyond +-1: exercises both clamps
        if sync:
            push_oracle_state(orc, env)
        o_obs, o_rew, o_done = orc.step(act)
        obs, rew, done = env.step(act.to(DEV))
        yield k, env, orc, (obs.cpu(), rew.cpu(), done.cpu()), (o_obs, o_rew, o_done)


# configuration branches no shipped YAML takes (their oracle halves are pinned to the reference in tests/test_config_branches_cpu.py)
_BRANCHES = {
    "water_current": dict(use_water_current=True, flow_vel_xy=(0.3, -0.2)),
    "full_dr_water_current": dict(use_water_current=True, flow_vel_xy=(-0.45, 0.15)),
    "couple_drag_only": dict(mass_rand=True, mass_min=30.0, mass_max=54.96, mass_base=34.96, mass_coupling=True, couple_targets=1,
                             use_drag_scale=True, kdrag_min=1.0, kdrag_max=1.5, thr_rand=True, thr_separate=True, kiz_rand=True,
                             couple_kiz_min=0.8, couple_kiz_max=1.7),
    "couple_thruster_kiz": dict(mass_rand=True, mass_min=30.0, mass_max=54.96, mass_base=34.96, mass_coupling=True, couple_targets=6,
                                use_drag_scale=True, kdrag_rand=True, kdrag_min=0.7, kdrag_max=1.6, kdrag_log=True, thr_rand=True,
                                kiz_rand=True, couple_kiz_min=1.0, couple_kiz_max=1.5),
    "kiz_linear": dict(kiz_rand=True, couple_kiz_min=0.8, couple_kiz_max=1.7),
    "kiz_log": dict(kiz_rand=True, kiz_log=True, couple_kiz_min=0.8, couple_kiz_max=1.7, mass_rand=True),
}


@pytest.mark.parametrize("variant", ["classic", "full_dr", "curriculum"] + list(_BRANCHES))
def test_fused_step_vs_oracle_lockstep(variant):
    """Every step starts from identical state (oracle state pushed to the GPU): obs/reward 1e-5, done bit-exact,
    reset index set bit-exact, next state 1e-5."""
    cfg = UsvEnvConfig(max_episode_length=12, kill_dist=12.5)
    cfg = cfg.full_dr() if variant.startswith("full_dr") else cfg
    if variant in _BRANCHES:
        cfg = dataclasses.replace(cfg, **_BRANCHES[variant])
    if variant == "curriculum":      # spawn annulus and kill distance move every control step (knees at step 0.25 and 0.75)
        cfg = dataclasses.replace(cfg, spawn_curriculum=True, spawn_curriculum_min_dist=0.2, spawn_curriculum_max_dist=3.0,
                                  spawn_curriculum_kill_dist=4.0, spawn_curriculum_warmup=0.25, spawn_curriculum_end=0.75,
                                  spawn_min_dist=2.0, spawn_max_dist=11.0, max_episode_length=5)
    n = 4096 + 37                                                   # ragged tail block
    n_done = 0
    for k, env, orc, (obs, rew, done), (o_obs, o_rew, o_done) in _lockstep(cfg, n, 16):
        assert done.dtype == torch.int64 and obs.shape == (n, 13)
        # obs: velocities O(1), distance O(10): atol 2e-5 is 1e-6..1e-5 of scale; angles enter through sin/cos of differences
        assert_close(obs, o_obs, RTOL, 2e-5, f"obs step {k}")
        # reward = sum of terms of scale 1; distance reward is a difference of two O(10) distances -> atol 2e-5
        assert_close(rew, o_rew, RTOL, 2e-5, f"reward step {k}")
        assert torch.equal(done, o_done), f"done mismatch at step {k}: {(done != o_done).sum()} envs"
        es, os_ = engine_state(env), oracle_state(orc)
        assert torch.equal(es["goal"], os_["goal"]) and torch.equal(es["progress"], os_["progress"])
        for name in es:
            if name in ("goal", "progress", "reset"):
                continue
            atol = 2e-5 if not name.startswith("USV_C_F") else 1e-4   # shifts are O(100)
            assert_close(es[name], os_[name], RTOL, atol, f"{name} step {k}")
        n_done += int(done.sum())
    assert n_done > n          # every env was reset at least once beyond the initial reset (max_episode_length=12)
    env.check_finite()


def test_fused_step_free_running_vs_oracle():
    """No re-sync for 40 steps: trajectories stay together (chaos-free horizon), resets happen at the same steps."""
    cfg = UsvEnvConfig(max_episode_length=25).full_dr()
    n = 2048
    mism = 0
    for k, env, orc, (obs, rew, done), (o_obs, o_rew, o_done) in _lockstep(cfg, n, 40, sync=False):
        assert_close(obs, o_obs, 1e-4, 2e-3, f"free-running obs step {k}")
        mism += int((done != o_done).sum())
    assert mism == 0


def test_stats_accumulators_vs_oracle():
    cfg = UsvEnvConfig(max_episode_length=9).full_dr()
    n = 1024
    sums = None
    for k, env, orc, _, _ in _lockstep(cfg, n, 12, sync=True, collect_stats=True):
        L = orc.last
        terms = torch.stack([L["distance_reward"], L["alignment_reward"], L["speed_reward"], L["d"], L["speed"],
                             L["boundary_penalty"], L["boundary_dist"], L["pen_lin"], L["pen_ang"], L["pen_angvar"],
                             L["pen_energy"], L["pen_actvar"], L["speed"], orc.prev_w.abs(), L["asum"]])
        if sums is None:
            sums = torch.zeros_like(terms)
        sums[:, L["reset_ids"]] = 0           # reset_idx clears episode_sums [ref: SNAP/USV_Virtual.py:812-817]
        sums += terms
        got = env.stats_matrix().cpu()
        bp = 5   # USV_ST_BOUNDARY_PENALTY = -exp(-(d-kill_dist)/0.25)*25 ~ 1e16: exp() amplifies the 1e-7 relative error of d by
        #          |x| ~ 40, so this (diagnostic-only, never added to the reward) row is compared at 1e-4 relative
        rows = [i for i in range(got.shape[0]) if i != bp]
        assert_close(got[rows], sums[rows], 1e-5, 1e-3 if k > 0 else 1e-4, f"episode sums step {k}")
        assert_close(got[bp], sums[bp], 1e-4, 1e-3, f"boundary-penalty sum step {k}")


def test_rollout_kernel_equals_step_kernel():
    cfg = UsvEnvConfig(max_episode_length=10).full_dr()
    n, T = 3000, 24
    a = FusedUsvEnv(cfg, n, DEV)
    b = FusedUsvEnv(cfg, n, DEV)
    act = (torch.rand((T, n, 2), generator=torch.Generator().manual_seed(1)) * 2 - 1).to(DEV)
    obs = torch.empty((T, n, 13), device=DEV); rew = torch.empty((T, n), device=DEV)
    done = torch.empty((T, n), dtype=torch.long, device=DEV)
    b.rollout(act, obs, rew, done)
    # same device code, but two kernels: the compiler may contract a*b+c differently in each, so compare at 1e-5
    for t in range(T):
        o, r, d = a.step(act[t])
        assert_close(o, obs[t], 1e-5, 2e-5, f"obs t={t}"); assert_close(r, rew[t], 1e-5, 2e-5, f"rew t={t}")
        assert torch.equal(d, done[t]), t
    assert_close(a.state[:, :11], b.state[:, :11], 1e-5, 2e-5, "final state")
    assert torch.equal(a.state[:, 11:].view(torch.int32), b.state[:, 11:].view(torch.int32))      # goal counter, progress
    assert_close(a.consts, b.consts, 1e-5, 1e-4, "final consts") and torch.equal(a.reset_buf, b.reset_buf)
    # outputs are optional
    c = FusedUsvEnv(cfg, n, DEV)
    c.rollout(act)
    assert torch.equal(c.state, b.state)


def test_planar_forces_match_6dof_force_modules():
    """The fused kernel's planar specialisation == the 6-DOF force kernels == oracle on planar states."""
    cfg = UsvEnvConfig().full_dr()
    n = 2048
    env = FusedUsvEnv(cfg, n, DEV)
    orc = O.ClassicEnvOracle(oracle_cfg(cfg), n)
    g = torch.Generator().manual_seed(2)
    for _ in range(3):
        orc.step(torch.rand((n, 2), generator=g) * 2 - 1)
    push_oracle_state(orc, env)
    out = env.planar_forces().cpu()
    drag, Fx, Fy, Tz, ax, ay, _ = orc.planar_wrench()
    assert_close(out[:, 0], drag[:, 0], RTOL, 2e-4, "du"); assert_close(out[:, 1], drag[:, 1], RTOL, 2e-4, "dv")
    assert_close(out[:, 2], drag[:, 5], RTOL, 2e-5, "dr")
    assert_close(out[:, 3], Fx, RTOL, 3e-4, "Fx"); assert_close(out[:, 4], Fy, RTOL, 3e-4, "Fy"); assert_close(out[:, 5], Tz, RTOL, 5e-5, "Tz")
    assert_close(out[:, 6], ax, RTOL, 1e-5, "ax"); assert_close(out[:, 7], ay, RTOL, 1e-5, "ay")
    # and against the stand-alone 6-DOF hydrodynamics kernel fed the planar-embedded state
    D = _hydro(n)
    D.linear_damping[:] = orc.linear_damping.to(DEV); D.quadratic_damping[:] = orc.quadratic_damping.to(DEV)
    half = orc.psi * 0.5
    quat = torch.stack([torch.cos(half), torch.zeros(n), torch.zeros(n), torch.sin(half)], 1).to(DEV)
    vel6 = torch.zeros((n, 6)); vel6[:, :2] = orc.vel; vel6[:, 5] = orc.r
    d6 = D.ComputeHydrodynamicsEffects(0.01, quat, vel6.to(DEV), False, [0, 0, 0]).cpu()
    assert_close(out[:, 0], d6[:, 0], RTOL, 2e-4); assert_close(out[:, 1], d6[:, 1], RTOL, 2e-4); assert_close(out[:, 2], d6[:, 5], RTOL, 2e-5)


def test_integrator_vs_float64_host_integration():
    """50 physics sub-steps (10 control steps, constant command, no reset) against the float64 integration."""
    cfg = dataclasses.replace(UsvEnvConfig().full_dr(), action_noise=False, max_episode_length=10_000, kill_dist=1e9,
                              position_tolerance=0.0)
    n = 4096
    env = FusedUsvEnv(cfg, n, DEV)
    orc = O.ClassicEnvOracle(oracle_cfg(cfg), n)
    g = torch.Generator().manual_seed(4)
    orc.step(torch.rand((n, 2), generator=g) * 2 - 1)
    orc.step(torch.rand((n, 2), generator=g) * 2 - 1)
    push_oracle_state(orc, env)
    cmd = torch.rand((n, 2), generator=g) * 2 - 1
    _, target = O.thruster_target(cmd, orc.lut_left, orc.lut_right, orc.thr_mult_left, orc.thr_mult_right)
    d = lambda t: t.double().numpy().copy()
    state = dict(x=d(orc.pos[:, 0]), y=d(orc.pos[:, 1]), psi=d(orc.psi), vx=d(orc.vel[:, 0]), vy=d(orc.vel[:, 1]), r=d(orc.r),
                 thrL=d(orc.current_forces[:, 0]), thrR=d(orc.current_forces[:, 1]))
    const = dict(mass=d(orc.mass), lin=d(orc.linear_damping[:, [0, 1, 5]]), quad=d(orc.quadratic_damping[:, [0, 1, 5]]),
                 kdrag=d(orc.drag_scale[:, 0]), kiz=d(orc.k_iz), fcx=d(orc.f_const[:, 0]), fcy=d(orc.f_const[:, 1]),
                 fxf=d(orc.f_freq[:, 0]), fyf=d(orc.f_freq[:, 1]), fxs=d(orc.f_shift[:, 0]), fys=d(orc.f_shift[:, 1]),
                 famp=d(orc.f_amp), tc=d(orc.t_const), tf=d(orc.t_freq), ts=d(orc.t_shift), tamp=d(orc.t_amp))
    ref = integrator64.substeps(state, const, d(target), dt=cfg.dt, alpha=float(np.float32(cfg.lag_alpha)), n_substeps=50,
                                izz=cfg.izz, thr_y_left=cfg.thr_y_left, thr_y_right=cfg.thr_y_right, use_const_force=True,
                                use_sin_force=True, use_const_torque=True, use_sin_torque=True)
    act = cmd.to(DEV)
    for _ in range(10):
        env.step(act)
    for name, key in (("USV_S_X", "x"), ("USV_S_Y", "y"), ("USV_S_VX", "vx"), ("USV_S_VY", "vy"), ("USV_S_R", "r"),
                      ("USV_S_THR_L", "thrL"), ("USV_S_THR_R", "thrR")):
        err = np.abs(env.field(name).double().cpu().numpy() - ref[key])
        assert err.max() < 3e-4, (name, float(err.max()))      # fp32 round-off over 50 steps, positions O(10) m
    dpsi = env.field("USV_S_PSI").double().cpu().numpy() - ref["psi"]
    dpsi = (dpsi + math.pi) % (2 * math.pi) - math.pi
    assert np.abs(dpsi).max() < 3e-4


# ------------------------------------------------------------------------------------------
# edge cases and full-size properties
def test_edge_sizes_and_error_codes():
    cfg = UsvEnvConfig()
    L = _lib.lib()
    for n in (1, 31, 256, 257):
        env = FusedUsvEnv(cfg, n, DEV)
        orc = O.ClassicEnvOracle(oracle_cfg(cfg), n)
        act = torch.zeros((n, 2))
        obs, rew, done = env.step(act.to(DEV))
        o_obs, o_rew, o_done = orc.step(act)
        assert_close(obs, o_obs, RTOL, 2e-5); assert_close(rew, o_rew, RTOL, 2e-5); assert torch.equal(done.cpu(), o_done)
    env = FusedUsvEnv(cfg, 8, DEV)
    p = env.params()
    z = ctypes.c_void_p(0)
    # n = 0 is a no-op, NULL / negative / misaligned arguments are rejected with the documented codes
    assert L.usv_step_fused_f32(ctypes.byref(env._buffers), z, z, z, ctypes.c_int64(0), ctypes.byref(p), _lib.stream()) == 0
    assert L.usv_step_fused_f32(ctypes.byref(env._buffers), z, z, z, ctypes.c_int64(8), ctypes.byref(p), _lib.stream()) == _lib.ENUMS["USV_E_NULL"]
    assert L.usv_step_fused_f32(ctypes.byref(env._buffers), z, z, z, ctypes.c_int64(-1), ctypes.byref(p), _lib.stream()) == _lib.ENUMS["USV_E_SIZE"]
    assert L.usv_step_fused_f32(z, z, z, z, ctypes.c_int64(8), ctypes.byref(p), _lib.stream()) == _lib.ENUMS["USV_E_NULL"]
    act = torch.zeros(17, device=DEV)
    assert L.usv_step_fused_f32(ctypes.byref(env._buffers), ctypes.c_void_p(act.data_ptr() + 4), _lib.ptr(env.obs), _lib.ptr(env.rew),
                                ctypes.c_int64(8), ctypes.byref(p), _lib.stream()) == _lib.ENUMS["USV_E_ALIGN"]
    p.n_lut = 1
    assert L.usv_step_fused_f32(ctypes.byref(env._buffers), _lib.ptr(torch.zeros((8, 2), device=DEV)), _lib.ptr(env.obs), _lib.ptr(env.rew),
                                ctypes.c_int64(8), ctypes.byref(p), _lib.stream()) == _lib.ENUMS["USV_E_PARAM"]
    assert L.ppo_gae_f32(z, z, z, z, z, ctypes.c_float(0.99), ctypes.c_float(0.95), z, z, ctypes.c_int32(16), ctypes.c_int64(0), _lib.stream()) == 0
    with pytest.raises(_lib.UsvLibraryError):
        env.step(torch.zeros((8, 2)))                               # CPU tensor: no silent fallback


def test_nan_probe_flag():
    env = FusedUsvEnv(UsvEnvConfig(), 64, DEV)
    act = torch.zeros((64, 2), device=DEV)
    env.step(act)
    env.check_finite()
    act[5, 0] = float("nan")
    env.step(act)
    with pytest.raises(RuntimeError, match="USV_NAN_PROBE"):
        env.check_finite()


def test_full_size_properties():
    """2^20 envs (BASELINE sweep top point): determinism, rank-sharding invariance, invariants of the outputs."""
    cfg = UsvEnvConfig(max_episode_length=6).full_dr()
    n = 1 << 20
    g = torch.Generator(device=DEV).manual_seed(0)
    acts = [torch.rand((n, 2), device=DEV, generator=g) * 2 - 1 for _ in range(8)]
    a, b = FusedUsvEnv(cfg, n, DEV), FusedUsvEnv(cfg, n, DEV)
    half = n // 2
    lo, hi = FusedUsvEnv(cfg, half, DEV, env_id_offset=0), FusedUsvEnv(cfg, half, DEV, env_id_offset=half)
    checksum = 0.0
    for t, act in enumerate(acts):
        prev_reset = a.reset_buf.clone()
        oa, ra, da = a.step(act)
        ob, rb, db = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)          # deterministic
        ol, rl, dl = lo.step(act[:half].contiguous())
        oh, rh, dh = hi.step(act[half:].contiguous())
        assert torch.equal(torch.cat([ol, oh]), oa) and torch.equal(torch.cat([rl, rh]), ra) and torch.equal(torch.cat([dl, dh]), da)
        assert torch.isfinite(oa).all() and torch.isfinite(ra).all()
        assert float(oa.abs().max()) <= cfg.clip_obs
        assert ((da == 0) | (da == 1)).all()
        prog = a.progress_buf
        assert (prog[prev_reset == 1] == 1).all()                    # a flagged env was reset, then stepped once
        assert (da[prog >= cfg.max_episode_length - 1] == 1).all()   # time-out always sets done
        assert (oa[:, 8] == 0).all() and (oa[:, 11:] == 0).all()     # unwritten obs columns stay 0 [ref: SNAP/USV_core.py:53-54]
        assert torch.allclose(oa[:, 3] ** 2 + oa[:, 4] ** 2, torch.ones(n, device=DEV), atol=1e-5)   # (cos a, sin a)
        checksum += float(ra.double().sum())
    assert math.isfinite(checksum)
    a.check_finite()


# ------------------------------------------------------------------------------------------
# GAE
def _gae(rew, val, dones, last_v, last_d, gamma=0.99, tau=0.95):
    T, n = rew.shape
    adv = torch.empty((T, n), device=DEV); ret = torch.empty((T, n), device=DEV)
    rc = _lib.lib().ppo_gae_f32(_lib.ptr(rew), _lib.ptr(val), _lib.ptr(dones), _lib.ptr(last_v), _lib.ptr(last_d), ctypes.c_float(gamma),
                                ctypes.c_float(tau), _lib.ptr(adv), _lib.ptr(ret), ctypes.c_int32(T), ctypes.c_int64(n), _lib.stream())
    _lib.check(rc, "ppo_gae_f32")
    return adv, ret


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_gae_vs_reference_golden(golden, tag):
    G = golden("gae")
    f = lambda k: cu(G[f"{tag}_{k}"])
    adv, ret = _gae(f("rewards").squeeze(-1).contiguous(), f("values").squeeze(-1).contiguous(), f("dones"), f("last_values").squeeze(-1).contiguous(),
                    f("last_dones"))
    # evaluated op by op in torch's order -> expected bit-exact; the bar is 1e-5 relative
    assert_close(adv, G[f"{tag}_adv"].squeeze(-1), 1e-6, 1e-7, "advantages")
    assert_close(ret, G[f"{tag}_returns"].squeeze(-1), 1e-6, 1e-7, "returns")


def test_gae_full_size_properties():
    """C3 size and beyond: linearity in the rewards and the all-done / zero-reward identities."""
    T, n = 16, 1 << 20
    g = torch.Generator(device=DEV).manual_seed(0)
    rew = torch.randn((T, n), device=DEV, generator=g); val = torch.randn((T, n), device=DEV, generator=g)
    dones = (torch.rand((T, n), device=DEV, generator=g) < 0.1).to(torch.uint8)
    last_v = torch.randn(n, device=DEV, generator=g); last_d = (torch.rand(n, device=DEV, generator=g) < 0.1).to(torch.uint8)
    a1, r1 = _gae(rew, val, dones, last_v, last_d)
    assert torch.equal(r1, a1 + val)
    zero_v = torch.zeros_like(val); zero_lv = torch.zeros_like(last_v)
    a_r, _ = _gae(rew, zero_v, dones, zero_lv, last_d)              # A is linear in (r, V): A(r,V) = A(r,0) + A(0,V)
    a_v, _ = _gae(torch.zeros_like(rew), val, dones, last_v, last_d)
    assert_close(a1, a_r + a_v, 1e-5, 1e-5, "linearity")
    ones = torch.ones_like(dones); one_l = torch.ones_like(last_d)
    a_d, _ = _gae(rew, val, ones, last_v, one_l)                    # every transition terminal -> A = r - V
    assert torch.equal(a_d, rew - val)
    # python double reference on a slice
    sl = slice(0, 4096)
    R, V, D = rew[:, sl].double().cpu(), val[:, sl].double().cpu(), dones[:, sl].double().cpu()
    last, nv, nnt = torch.zeros(4096, dtype=torch.float64), last_v[sl].double().cpu(), 1 - last_d[sl].double().cpu()
    for t in reversed(range(T)):
        delta = R[t] + 0.99 * nv * nnt - V[t]
        last = delta + 0.99 * 0.95 * nnt * last
        assert_close(a1[t, sl], last, 1e-5, 1e-5, f"gae t={t}")
        nv, nnt = V[t], 1 - D[t]


def test_host_stepper_equals_plain_steps():
    """engine.HostStepper (pinned host buffers, copies overlapped on 3 streams) delivers exactly what FusedUsvEnv.step produces."""
    from omniisaacgymenvs_loop_b200.engine import HostStepper
    cfg = UsvEnvConfig(max_episode_length=6).full_dr()
    n, K = 50000 + 3, 9
    a, b = FusedUsvEnv(cfg, n, DEV), FusedUsvEnv(cfg, n, DEV)
    hs = HostStepper(b, depth=2)
    g = torch.Generator().manual_seed(2)
    acts = [(torch.rand((n, 2), generator=g) * 2 - 1).pin_memory() for _ in range(K)]
    outs = [(torch.empty((n, 13)).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()) for _ in range(K)]
    tickets = [hs.submit(acts[k], *outs[k]) for k in range(K)]
    hs.wait(tickets[-1])
    hs.synchronize()
    for k in range(K):
        obs, rew, done = a.step(acts[k].to(DEV))
        assert torch.equal(outs[k][0], obs.cpu()) and torch.equal(outs[k][1], rew.cpu()) and torch.equal(outs[k][2], done.cpu().to(torch.uint8)), k
    assert torch.equal(a.state, b.state)


def test_captured_steps_equal_eager_steps():
    """FusedUsvEnv.capture_steps: K control steps replayed from one CUDA graph continue the eager Philox sequence bit for bit."""
    cfg = UsvEnvConfig(max_episode_length=7).full_dr()
    n, K = 4096, 12
    a, b = FusedUsvEnv(cfg, n, DEV), FusedUsvEnv(cfg, n, DEV)
    g = torch.Generator().manual_seed(5)
    warm = (torch.rand((n, 2), generator=g) * 2 - 1).to(DEV)
    a.step(warm); b.step(warm)
    acts = (torch.rand((K, n, 2), generator=g) * 2 - 1).to(DEV)
    obs, rew, done = torch.empty((K, n, 13), device=DEV), torch.empty((K, n), device=DEV), torch.empty((K, n), dtype=torch.long, device=DEV)
    replay = b.capture_steps(acts, obs, rew, done)
    for rep in range(3):
        replay()
        for k in range(K):
            o, r, d = a.step(acts[k])
            assert torch.equal(o, obs[k]) and torch.equal(r, rew[k]) and torch.equal(d, done[k]), (rep, k)
    assert a.step_counter == b.step_counter and torch.equal(a.state, b.state)
    o1, o2 = a.step(warm)[0].clone(), b.step(warm)[0].clone()       # eager steps after replays stay in sequence
    assert torch.equal(o1, o2)


def test_reference_style_apply_forces_on_the_planar_view_vs_float64_integration(golden):
    """The simulator surface (robots.PlanarHeronView / PlanarWorld) driven the way the reference's USVVirtual.apply_forces drives Isaac Sim
    [ref: OIGE/tasks/USV_Virtual.py:1103-1133 ; envs/vec_env_rlgames.py:154-171]: per physics sub-step the stand-alone force modules
    (HydrodynamicsObject.ComputeHydrodynamicsEffects, DynamicsFirstOrder.update_forces) are evaluated on the view's poses / velocities,
    pushed through base / thruster_left / thruster_right .apply_forces_and_torques_at_pos, and world.step() integrates -- against the float64
    host integration of the same model (50 sub-steps)."""
    from omniisaacgymenvs_loop_b200.robots import PlanarHeronView, PlanarWorld
    n, dt = 512, 0.01
    g = torch.Generator().manual_seed(6)
    view = PlanarHeronView(n, DEV, mass=34.96, izz=10.0)
    world = PlanarWorld(dt)
    world.add(view)
    pos = torch.zeros((n, 3)); pos[:, :2] = torch.rand((n, 2), generator=g) * 20 - 10
    yaw = (torch.rand(n, generator=g) * 2 - 1) * math.pi
    quat = torch.stack([torch.cos(yaw / 2), torch.zeros(n), torch.zeros(n), torch.sin(yaw / 2)], 1)
    vel6 = torch.zeros((n, 6)); vel6[:, :2] = torch.rand((n, 2), generator=g) * 3 - 1.5; vel6[:, 5] = torch.rand(n, generator=g) * 2 - 1
    view.set_world_poses(pos.to(DEV), quat.to(DEV))
    view.set_velocities(vel6.to(DEV))
    p, q = view.get_world_poses()
    assert_close(p, pos, 0, 1e-6); assert_close(q, quat, 0, 1e-6); assert_close(view.get_velocities(), vel6, 0, 0)
    D = _hydro(n)
    GF = golden("force_modules")
    Tm = DynamicsFirstOrder(dict(THR_CFG), n, DEV, 0.05, dt, 1000, GF["lut_classic_points_left"].tolist(), GF["lut_classic_points_right"].tolist(),
                            [0.0] * 5, [0.0] * 5, -1.0, 1.0)
    cmd = (torch.rand((n, 2), generator=g) * 2 - 1).to(DEV)
    Tm.set_target_force(cmd)
    target = Tm.thruster_forces_before_dynamics.double().cpu().numpy()
    d = lambda t: t.double().cpu().numpy().copy()
    lin, quad = torch.tensor(LIN)[[0, 1, 5]].double().numpy(), torch.tensor(QUAD)[[0, 1, 5]].double().numpy()
    state = dict(x=d(pos[:, 0]), y=d(pos[:, 1]), psi=d(yaw), vx=d(vel6[:, 0]), vy=d(vel6[:, 1]), r=d(vel6[:, 5]), thrL=np.zeros(n), thrR=np.zeros(n))
    const = dict(mass=np.full(n, 34.96), lin=np.tile(lin, (n, 1)), quad=np.tile(quad, (n, 1)), kdrag=np.ones(n), kiz=np.ones(n))
    ref = integrator64.substeps(state, const, target, dt=dt, alpha=float(torch.exp(torch.tensor(-dt / 0.05))), n_substeps=50, izz=10.0,
                                thr_y_left=0.377654, thr_y_right=-0.377654)
    for _ in range(50):                                        # == USVVirtual.apply_forces() + world.step(), once per sub-step
        _, quats = view.get_world_poses()
        drag = D.ComputeHydrodynamicsEffects(dt, quats, view.get_velocities(), False, [0, 0, 0])
        thr = Tm.update_forces()                               # [n,6]: left thrust in [:, :3], right in [:, 3:], along the thruster x axis
        view.base.apply_forces_and_torques_at_pos(forces=drag[:, :3].contiguous(), torques=drag[:, 3:].contiguous(), is_global=False)
        view.thruster_left.apply_forces_and_torques_at_pos(forces=thr[:, :3].contiguous(), is_global=False)
        view.thruster_right.apply_forces_and_torques_at_pos(forces=thr[:, 3:].contiguous(), is_global=False)
        world.step(render=False)
    assert world.current_time_step_index == 50
    for col, key in ((0, "x"), (1, "y")):
        assert np.abs(view.pose[:, col].double().cpu().numpy() - ref[key]).max() < 3e-4, key
    for col, key in ((0, "vx"), (1, "vy"), (2, "r")):
        assert np.abs(view.vel[:, col].double().cpu().numpy() - ref[key]).max() < 3e-4, key
    dpsi = view.pose[:, 2].double().cpu().numpy() - ref["psi"]
    assert np.abs((dpsi + math.pi) % (2 * math.pi) - math.pi).max() < 3e-4
    # a world-frame force is rotated into the body frame: pushing along +x_world accelerates vx only, whatever the heading
    view.vel.zero_(); view.wrench.zero_()
    f = torch.zeros((n, 3), device=DEV); f[:, 0] = 34.96
    view.base.apply_forces_and_torques_at_pos(forces=f, is_global=True)
    world.step()
    assert_close(view.vel[:, 0], torch.full((n,), dt), 1e-5, 1e-7); assert float(view.vel[:, 1].abs().max()) < 1e-6 and float(view.wrench.abs().max()) == 0.0
