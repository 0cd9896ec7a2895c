"""The reference's task-class calls, fed the reference's own state tensors, against the reference's own outputs -- no oracle in between.

tests/golden/capture_xy_classic.npz was produced by the classic snapshot's CaptureXYTask.get_state_observations / compute_reward /
update_kills and Penalties.compute_penalty (oracle/make_golden.py:classic_task) on a 6-step trajectory with a reset batch; the same
tensors go through tasks/USV/USV_capture_xy.py + USV_task_rewards.py (usv_capturexy_obs_reward_done_f32: the device function of the
fused step's task part) and, in the second test, through the fused step kernel itself with zero physics sub-steps.
Bar (north_star): observations / rewards 1e-5 relative in fp32, die / goal counter bit-exact."""
import dataclasses

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniisaacgymenvs_loop_b200.config import UsvEnvConfig  # noqa: E402
from omniisaacgymenvs_loop_b200.tasks.USV.USV_task_factory import task_factory  # noqa: E402
from omniisaacgymenvs_loop_b200.tasks.USV.USV_core import parse_data_dict  # noqa: E402
from omniisaacgymenvs_loop_b200.tasks.USV.USV_task_rewards import Penalties  # noqa: E402
from tests.util import assert_close  # noqa: E402

DEV = "cuda:0"
cu = lambda a: torch.as_tensor(a).to(DEV).contiguous()


def _state(G, k):
    yaw = cu(G["yaw"][k])
    return {"position": cu(G["pos"][k]), "orientation": torch.stack([torch.cos(yaw), torch.sin(yaw)], 1).contiguous(),
            "linear_velocity": cu(G["vel"][k]), "angular_velocity": cu(G["w"][k])}


def test_classic_task_classes_vs_reference_golden(golden):
    G = golden("capture_xy_classic")
    env = UsvEnvConfig().to_task_cfg()["env"]              # the classic snapshot's YAML sections (product defaults)
    K, n = G["pos"].shape[:2]
    task = task_factory.get(env["task_parameters"], env["reward_parameters"], n, DEV)
    pen = parse_data_dict(Penalties(), env["penalties_parameters"])
    task._target_positions.copy_(cu(G["target"]))
    assert torch.equal(task.just_had_been_reset.cpu(), torch.arange(n))
    for k in range(K):
        if k == int(G["reset_step"]):
            task.reset(cu(G["reset_ids"]))
        st, act = _state(G, k), cu(G["actions"][k])
        obs = task.get_state_observations(st, "local")
        rew = task.compute_reward(st, act)
        p = pen.compute_penalty(st, act, 0.0)
        die = task.update_kills(0)
        # atol: fp32 rounding of cos / sin of the bearing error near their zeros, distances up to 21 m
        assert_close(obs, G["obs"][k], 1e-5, 2e-6, f"obs step {k}")
        assert_close(task.distance_reward, G["distance_reward"][k], 1e-5, 2e-6, f"distance reward step {k}")
        assert_close(task.alignment_reward, G["alignment_reward"][k], 1e-5, 1e-7, f"alignment reward step {k}")
        assert_close(task.a, G["speed_reward"][k], 1e-5, 1e-7, f"speed reward step {k}")
        assert_close(rew, G["reward"][k], 1e-5, 2e-6, f"reward step {k}")
        assert_close(p, G["penalty"][k], 1e-5, 1e-7, f"penalty step {k}")
        assert die.dtype == torch.int64 and torch.equal(die.cpu(), torch.from_numpy(G["die"][k])), f"die step {k}"
        assert torch.equal(task._goal_reached.cpu(), torch.from_numpy(G["goal_reached"][k])), f"goal counter step {k}"
        assert task.just_had_been_reset.numel() == 0
    with pytest.raises(NotImplementedError):
        task.get_state_observations(_state(G, 0), "world")
    with pytest.raises(NotImplementedError):
        task_factory.get({"name": "GoToPose"}, {"name": "GoToPose"}, 4, DEV)


def test_fused_step_task_part_vs_reference_golden(golden):
    """The headline kernel itself on the golden states: zero physics sub-steps and no noise leave the pushed state untouched, so
    usv_step_fused_f32 reduces to its task part (post_classic): obs / reward + penalties / kills / goal counter of the reference.
    task.reset(ids) of the reference (counter cleared, distance reward of that step zeroed) is emulated on the pushed state: counter 0
    and prev_position_dist = the current distance (a zero progress term in every reward mode)."""
    from omniisaacgymenvs_loop_b200.engine import FusedUsvEnv
    G = golden("capture_xy_classic")
    K, n = G["pos"].shape[:2]
    cfg = dataclasses.replace(UsvEnvConfig(), n_substeps=0, clip_actions=2.0, clip_obs=1e9, max_episode_length=10 ** 6,
                              action_noise=False, noise_pos=False, noise_vel=False, noise_heading=False)
    env = FusedUsvEnv(cfg, n, DEV)
    env.step(torch.zeros((n, 2), device=DEV))               # the initial reset of every env; then the state is overwritten per step
    env.set_field("USV_C_TX", cu(G["target"][:, 0])); env.set_field("USV_C_TY", cu(G["target"][:, 1]))
    env.set_field("USV_S_GOAL_CNT", torch.zeros(n, dtype=torch.int32, device=DEV))
    dist = lambda k: torch.from_numpy(np.linalg.norm((G["target"] - G["pos"][k]).astype(np.float32), axis=1)).to(DEV)
    for k in range(K):
        env.set_field("USV_S_X", cu(G["pos"][k][:, 0])); env.set_field("USV_S_Y", cu(G["pos"][k][:, 1]))
        env.set_field("USV_S_PSI", cu(G["yaw"][k]))
        env.set_field("USV_S_VX", cu(G["vel"][k][:, 0])); env.set_field("USV_S_VY", cu(G["vel"][k][:, 1])); env.set_field("USV_S_R", cu(G["w"][k]))
        env.reset_buf.zero_()                               # envs the kernel flagged done would reset themselves: the golden state is pushed instead
        if k == 0:
            env.first_call = True                           # Penalties' first call: both variations are zero
            env.set_field("USV_S_PREV_D", dist(0))
        else:
            env.set_field("USV_S_PREV_D", dist(k - 1))
        if k == int(G["reset_step"]):
            ids = cu(G["reset_ids"])
            env.set_field("USV_S_GOAL_CNT", torch.zeros(len(ids), dtype=torch.int32, device=DEV), ids)
            env.set_field("USV_S_PREV_D", dist(k)[ids], ids)
        obs, rew, done = env.step(cu(G["actions"][k]))
        assert_close(obs, G["obs"][k], 1e-5, 2e-6, f"obs step {k}")
        assert_close(rew, G["reward"][k] + G["penalty"][k], 1e-5, 3e-6, f"reward + penalty step {k}")
        assert torch.equal(done.cpu(), torch.from_numpy(G["die"][k])), f"die step {k}"
        assert torch.equal(env.goal_reached.cpu().to(torch.int32), torch.from_numpy(G["goal_reached"][k])), f"goal counter step {k}"
    env.check_finite()


def test_fused_disturbance_wrench_vs_reference_golden(golden):
    """A11 on the device, straight against the reference: ForceDisturbance / TorqueDisturbance.get_disturbance_forces at the golden's
    world positions (tests/golden/disturbances.npz) vs the disturbance terms of the fused kernel's planar wrench -- at rest, heading 0
    and zero thrust the net body wrench of usv_planar_forces_f32 IS the disturbance (no drag, body frame == world frame)."""
    from omniisaacgymenvs_loop_b200.engine import FusedUsvEnv
    G = golden("disturbances")
    n = G["root_pos"].shape[0]
    cfg = dataclasses.replace(UsvEnvConfig().full_dr(), envs_per_row=0)          # the golden positions are world positions already
    env = FusedUsvEnv(cfg, n, DEV)
    env.step(torch.zeros((n, 2), device=DEV))
    z = torch.zeros(n, device=DEV)
    for name, v in (("USV_S_X", cu(G["root_pos"][:, 0])), ("USV_S_Y", cu(G["root_pos"][:, 1])), ("USV_S_PSI", z), ("USV_S_VX", z), ("USV_S_VY", z),
                    ("USV_S_R", z), ("USV_S_THR_L", z), ("USV_S_THR_R", z),
                    ("USV_C_FCX", cu(G["f_const"][:, 0])), ("USV_C_FCY", cu(G["f_const"][:, 1])), ("USV_C_FXF", cu(G["fxf"])), ("USV_C_FYF", cu(G["fyf"])),
                    ("USV_C_FXS", cu(G["fxs"])), ("USV_C_FYS", cu(G["fys"])), ("USV_C_FAMP", cu(G["famp"])), ("USV_C_TC", cu(G["t_const"][:, 2])),
                    ("USV_C_TF", cu(G["tf"])), ("USV_C_TS", cu(G["ts"])), ("USV_C_TAMP", cu(G["tamp"]))):
        env.set_field(name, v)
    out = env.planar_forces()
    # conditioning, not kernel error: the phase  pos * freq + shift  reaches ~200 rad, where one fp32 ulp is 1.5e-5 rad; the reference
    # rounds the product and the sum separately, the kernel contracts them into one FMA, so the two arguments differ by up to an ulp:
    # |d sin| <= 1.5e-5 x amplitude (<= 1.8 N, 0.18 N m).  The sinusoid itself (MUFU.SIN after range reduction) adds 4e-7 x amplitude.
    assert_close(out[:, 3], G["forces"][:, 0], 1e-5, 3e-5, "disturbance Fx"); assert_close(out[:, 4], G["forces"][:, 1], 1e-5, 3e-5, "disturbance Fy")
    assert_close(out[:, 5], G["torques"][:, 2], 1e-5, 3e-6, "disturbance Tz")
    assert float(np.abs(G["forces"][:, :2]).max()) > 0.5 and float(np.abs(G["torques"][:, 2]).max()) > 0.02


def test_kill_curriculum_vs_reference_golden(golden):
    """A18 with the spawn / kill curriculum, straight against the reference: update_kills(step) of the reference task at 9 steps around
    the warm-up / end knees (tests/golden/classic_curriculum.npz) vs CaptureXYTask.update_kills(step) on the kernel, bit-exact."""
    G = golden("classic_curriculum")
    cmin, cmax, ckill, warm, end, rmin, rmax, kill = G["params"].tolist()
    tp = dict(UsvEnvConfig().to_task_cfg()["env"]["task_parameters"], spawn_curriculum=True, spawn_curriculum_min_dist=cmin,
              spawn_curriculum_max_dist=cmax, spawn_curriculum_kill_dist=ckill, spawn_curriculum_warmup=int(warm), spawn_curriculum_end=int(end),
              min_spawn_dist=rmin, max_spawn_dist=rmax, kill_dist=kill)
    n = G["dist"].shape[0]
    task = task_factory.get(tp, UsvEnvConfig().to_task_cfg()["env"]["reward_parameters"], n, DEV)
    st = {"position": torch.stack([cu(G["dist"]), torch.zeros(n, device=DEV)], 1).contiguous(),        # target at the origin: |pos| = dist
          "orientation": torch.tensor([[1.0, 0.0]], device=DEV).repeat(n, 1), "linear_velocity": torch.full((n, 2), 0.5, device=DEV),
          "angular_velocity": torch.zeros(n, device=DEV)}
    task.get_state_observations(st, "local")
    for k, step in enumerate(G["steps"].tolist()):
        assert torch.equal(task.update_kills(step).cpu(), torch.from_numpy(G["die"][k])), f"curriculum step {step}"
    assert int(G["die"].sum()) > 0 and int((1 - G["die"]).sum()) > 0


def test_round2_entry_points_edge_sizes_and_error_codes():
    """Empty / ragged inputs and the documented error codes of the entry points added in round 2 (stand-alone CaptureXY task calls, planar
    rigid-body view, SysID student, fused multi-rank minibatch step): n = 0 is a no-op, NULL / size / parameter errors come back as codes,
    CPU tensors and CPU devices are refused by the Python surfaces (no silent fallback)."""
    import ctypes
    from omniisaacgymenvs_loop_b200 import _lib
    from omniisaacgymenvs_loop_b200.algo.ppo.module import StateHistoryEncoder
    from omniisaacgymenvs_loop_b200.robots import PlanarHeronView
    L, E = _lib.lib(), _lib.ENUMS
    z, st = ctypes.c_void_p(0), _lib.stream()
    p = UsvEnvConfig().to_params()
    io = _lib.UsvCaptureXYIO()
    io.what = E["USV_CXY_OBS"]
    assert L.usv_capturexy_obs_reward_done_f32(ctypes.byref(io), ctypes.c_int64(0), ctypes.byref(p), st) == 0
    assert L.usv_capturexy_obs_reward_done_f32(ctypes.byref(io), ctypes.c_int64(-1), ctypes.byref(p), st) == E["USV_E_SIZE"]
    assert L.usv_capturexy_obs_reward_done_f32(ctypes.byref(io), ctypes.c_int64(4), ctypes.byref(p), st) == E["USV_E_NULL"]
    assert L.usv_capturexy_obs_reward_done_f32(z, ctypes.c_int64(4), ctypes.byref(p), st) == E["USV_E_NULL"]
    io.what = 0
    assert L.usv_capturexy_obs_reward_done_f32(ctypes.byref(io), ctypes.c_int64(4), ctypes.byref(p), st) == E["USV_E_PARAM"]
    # ragged sizes through the class surface: 1, 31, 257 envs against the golden's first rows repeated
    for n in (1, 31, 257):
        env = UsvEnvConfig().to_task_cfg()["env"]
        task = task_factory.get(env["task_parameters"], env["reward_parameters"], n, DEV)
        stt = {"position": torch.full((n, 2), 3.0, device=DEV), "orientation": torch.tensor([[1.0, 0.0]], device=DEV).repeat(n, 1),
               "linear_velocity": torch.zeros((n, 2), device=DEV), "angular_velocity": torch.zeros(n, device=DEV)}
        obs = task.get_state_observations(stt, "local")
        assert obs.shape == (n, 13) and float((obs[:, 5] - 18.0 ** 0.5).abs().max()) < 1e-5
        assert task.update_kills(0).sum() == 0
    with pytest.raises(_lib.UsvLibraryError):
        task_factory.get(env["task_parameters"], env["reward_parameters"], 4, "cpu")
    # planar view
    w = torch.zeros((4, 3), device=DEV)
    assert L.usv_planar_wrench_accumulate_f32(_lib.ptr(w), z, z, z, ctypes.c_float(0), ctypes.c_float(0), ctypes.c_int32(0), ctypes.c_int64(4), st) == E["USV_E_NULL"]
    assert L.usv_planar_wrench_accumulate_f32(_lib.ptr(w), _lib.ptr(w), z, z, ctypes.c_float(0), ctypes.c_float(0), ctypes.c_int32(1), ctypes.c_int64(4), st) == E["USV_E_NULL"]
    assert L.usv_planar_rigid_step_f32(_lib.ptr(w), _lib.ptr(w), _lib.ptr(w), _lib.ptr(w), _lib.ptr(w), ctypes.c_float(0.0), ctypes.c_int64(4), st) == E["USV_E_PARAM"]
    assert L.usv_planar_rigid_step_f32(z, z, z, z, z, ctypes.c_float(0.01), ctypes.c_int64(0), st) == 0
    with pytest.raises(_lib.UsvLibraryError):
        PlanarHeronView(4, "cpu")
    # SysID student
    assert L.dagger_history_encoder_param_count(ctypes.c_int32(25), ctypes.c_int32(50), ctypes.c_int32(8)) == 20136
    assert L.dagger_history_encoder_param_count(ctypes.c_int32(25), ctypes.c_int32(30), ctypes.c_int32(8)) == -1
    assert L.dagger_history_encoder_param_count(ctypes.c_int32(40), ctypes.c_int32(50), ctypes.c_int32(8)) == -1
    assert L.dagger_history_encoder_forward_f32(z, z, ctypes.c_int64(1250), ctypes.c_int32(25), ctypes.c_int32(50), ctypes.c_int32(8), z, ctypes.c_int64(0), st) == 0
    assert L.dagger_history_encoder_forward_f32(z, z, ctypes.c_int64(1250), ctypes.c_int32(25), ctypes.c_int32(50), ctypes.c_int32(8), z, ctypes.c_int64(3), st) == E["USV_E_NULL"]
    assert L.dagger_history_encoder_forward_f32(z, z, ctypes.c_int64(100), ctypes.c_int32(25), ctypes.c_int32(50), ctypes.c_int32(8), z, ctypes.c_int64(3), st) == E["USV_E_SIZE"]
    enc = StateHistoryEncoder("LeakyReLU", 25, 50, 8, DEV, seed=1)
    assert enc(torch.zeros((1, 1250), device=DEV)).shape == (1, 8)                       # a single sample: one warp of one CTA
    with pytest.raises(NotImplementedError):
        StateHistoryEncoder("LeakyReLU", 25, 30, 8, DEV)
    with pytest.raises(_lib.UsvLibraryError):
        StateHistoryEncoder("LeakyReLU", 25, 50, 8, "cpu")
    # fused multi-rank minibatch step: a window that is too small, a rank outside the world
    assert L.ppo_minibatch_step_peer_entries(ctypes.c_int32(13)) == 293 * 64 and L.ppo_minibatch_step_peer_entries(ctypes.c_int32(60)) == -1
    torch.cuda.synchronize()
