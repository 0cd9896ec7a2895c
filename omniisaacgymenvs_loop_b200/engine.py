"""FusedUsvEnv: owner of the structure-of-arrays env state in HBM and the thin caller of the fused
step / rollout kernels.  This is what sits *below* the reference-shaped surfaces (USVVirtual,
VecEnvRLGames) in this package: where the reference crosses into Isaac Sim / PhysX
[ref: OIGE/envs/vec_env_rlgames.py:154-173], this class launches one sm_100a kernel.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from .config import UsvEnvConfig, UsvLiveConfig

E = _lib.ENUMS
OBS_DIM = 13
LIVE_OBS_DIM = E["USV_B_OBS"]
GRID = E["USV_B_GRID"]


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class FusedUsvEnv:
    def __init__(self, cfg: UsvEnvConfig, num_envs: Optional[int] = None, device="cuda:0", env_id_offset: int = 0,
                 collect_stats: bool = False):
        self.lib = _lib.lib()
        self.cfg = cfg
        self.num_envs = n = int(num_envs if num_envs is not None else cfg.num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.UsvLibraryError("FusedUsvEnv runs on CUDA only (no CPU fallback); got device=%s" % device)
        self.env_id_offset = int(env_id_offset)
        self.stride = stride = _round_up(max(n, 1), 32)       # every SoA row starts on a 128 B line
        f32 = dict(dtype=torch.float32, device=self.device)
        # AoSoA: [tile][field][32 lanes] -- a warp owns one tile; every field of an env is an immediate offset
        nt = stride // 32
        self.state = torch.zeros((nt, E["USV_S_COUNT"], 32), **f32)
        self.consts = torch.zeros((nt, E["USV_C_COUNT"], 32), **f32)
        self.stats = torch.zeros((nt, E["USV_ST_COUNT"], 32), **f32) if collect_stats else None
        self.reset_buf = torch.ones(n, dtype=torch.long, device=self.device)     # [ref: SNAP/USV_Virtual.py:342-344]
        self.obs = torch.zeros((n, OBS_DIM), **f32)
        self.rew = torch.zeros(n, **f32)
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.step_counter = 0
        # device-side part of the Philox step index: a captured graph of control steps adds to it at the end of each replay
        # (advance_step_offset); the kernels see  p.step_counter + *step_offset  == self.step_counter  at every launch
        self.step_offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._step_offset_host = 0
        self.curriculum_step = 0.0       # the task's `step` (control steps / horizon_length); drives the optional spawn / kill curriculum
        self.first_call = True
        # per-episode constants start at their nominal values (reset_idx rewrites them)
        c = self.consts
        c[:, E["USV_C_MASS"]] = cfg.mass_base
        for j, (lf, qf) in enumerate(zip(("USV_C_LIN_U", "USV_C_LIN_V", "USV_C_LIN_R"),
                                         ("USV_C_QUAD_U", "USV_C_QUAD_V", "USV_C_QUAD_R"))):
            c[:, E[lf]] = cfg.lin_base[j]
            c[:, E[qf]] = cfg.quad_base[j]
        for name in ("USV_C_KDRAG", "USV_C_THR_ML", "USV_C_THR_MR", "USV_C_KIZ"):
            c[:, E[name]] = 1.0
        # thruster LUTs  [ref: OIGE/envs/USV/ThrusterDynamics.py:152-177]
        self.lut_left = self._build_lut(cfg.lut_points_left)
        self.lut_right = self._build_lut(cfg.lut_points_right)
        # post_reset -> set_targets -> task.get_goals on every env  [ref: SNAP/USV_Virtual.py:652-700]
        if cfg.goal_random_position > 0:
            self.randomize_targets(torch.arange(n, device=self.device))
        self._buffers = self._make_buffers()

    # ---- helpers -----------------------------------------------------------------------
    def _build_lut(self, points) -> torch.Tensor:
        pts = torch.tensor(points, dtype=torch.float32, device=self.device)
        lut = torch.empty(self.cfg.n_lut, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.usv_thruster_build_lut_f32(_lib.ptr(pts), ctypes.c_int32(pts.numel()), _lib.ptr(lut),
                                                       ctypes.c_int32(self.cfg.n_lut), _lib.stream()), "build_lut")
        return lut

    def _make_buffers(self):
        b = _lib.UsvEnvBuffers()
        b.state, b.state_stride = self.state.data_ptr(), self.stride
        b.consts, b.consts_stride = self.consts.data_ptr(), self.stride
        b.stats = self.stats.data_ptr() if self.stats is not None else None
        b.stats_stride = self.stride
        b.reset_buf = self.reset_buf.data_ptr()
        b.lut_left, b.lut_right = self.lut_left.data_ptr(), self.lut_right.data_ptr()
        b.nonfinite_flag = self.nonfinite.data_ptr()
        b.step_offset = self.step_offset.data_ptr()
        return b

    def _src(self, name: str) -> torch.Tensor:
        return self.state if name.startswith("USV_S_") else (self.consts if name.startswith("USV_C_") else self.stats)

    def field_view(self, name: str) -> torch.Tensor:
        """(tiles, 32) strided VIEW of one field (in-place ops write through to the kernel's buffers)."""
        return self._src(name)[:, E[name], :]

    def field(self, name: str) -> torch.Tensor:
        """(N,) COPY of one field in env order, e.g. field('USV_S_X') or field('USV_C_MASS')."""
        return self.field_view(name).reshape(-1)[: self.num_envs]

    def set_field(self, name: str, values, env_ids: Optional[torch.Tensor] = None) -> None:
        """Writes (N,) values (or values for `env_ids`) into a field; int tensors are stored as int32 bit patterns."""
        v = torch.as_tensor(values, device=self.device)
        v = v.to(torch.int32).view(torch.float32) if not v.is_floating_point() else v.to(torch.float32)
        view = self.field_view(name)
        if env_ids is None:
            flat = torch.zeros(self.stride, dtype=torch.float32, device=self.device)
            flat[: self.num_envs] = v
            view.copy_(flat.view(-1, 32))
        else:
            ids = env_ids.to(self.device, torch.long)
            view[ids >> 5, ids & 31] = v

    def int_field(self, name: str) -> torch.Tensor:
        return self.field(name).view(torch.int32)

    def stats_matrix(self) -> torch.Tensor:
        """(USV_ST_COUNT, N) copy of the episode-sum accumulators."""
        return self.stats.permute(1, 0, 2).reshape(E["USV_ST_COUNT"], -1)[:, : self.num_envs]

    @property
    def progress_buf(self) -> torch.Tensor:
        return self.int_field("USV_S_PROGRESS").long()

    @property
    def goal_reached(self) -> torch.Tensor:
        return self.int_field("USV_S_GOAL_CNT")

    def randomize_targets(self, env_ids: torch.Tensor) -> None:
        g = float(self.cfg.goal_random_position)
        tmp = torch.zeros((self.num_envs, 2), dtype=torch.float32, device=self.device)
        base = torch.zeros(2, dtype=torch.float32, device=self.device)
        lo, hi = base - g, base + g
        ids = env_ids.to(torch.long).contiguous()
        _lib.check(self.lib.usv_randomize_rows_f32(_lib.ptr(tmp), ctypes.c_int64(2), _lib.ptr(ids), ctypes.c_int64(ids.numel()),
                                                   ctypes.c_int32(2), _lib.ptr(base), _lib.ptr(lo), _lib.ptr(hi), ctypes.c_int32(0),
                                                   ctypes.c_uint64(self.cfg.seed), ctypes.c_uint64(self.step_counter),
                                                   ctypes.c_uint32(0), _lib.stream()), "randomize_rows")
        self.set_field("USV_C_TX", tmp[ids, 0], ids)
        self.set_field("USV_C_TY", tmp[ids, 1], ids)

    def params(self):
        # the struct is built once (130 fields); per call only the counters change
        p = getattr(self, "_params", None)
        if p is None:
            p = self._params = self.cfg.to_params(0, self.env_id_offset, False)
        p.step_counter = self.step_counter - self._step_offset_host
        p.first_call = int(self.first_call)
        # live action path: the initial bias applies to the first N control steps only  [ref: OIGE/tasks/USV_Virtual.py:1070-1077]
        p.action_bias = self.cfg.action_bias if self.step_counter < self.cfg.action_bias_steps else 0.0
        if self.cfg.spawn_curriculum:
            # get_spawns(step) runs in pre_physics_step, update_kills(step) after calculate_metrics has advanced `step` by 1/horizon
            p.spawn_min_dist, p.spawn_max_dist, _ = self.cfg.curriculum(self.curriculum_step)
            p.kill_dist = self.cfg.curriculum(self.curriculum_step + 1.0 / self.cfg.horizon_length)[2]
        return p

    def graph_key(self, steps: int):
        """Host-side parameters that a captured graph of the next `steps` control steps would bake in (None: the window straddles a
        change and must run eagerly).  The only such parameter is the live task's initial action bias, which is switched off by the
        host after `action_bias_steps` control steps  [ref: OIGE/tasks/USV_Virtual.py:1070-1077]; the live task's evaluation switch
        of the privileged tail's source (`masscom_obs_base`) is part of the key too."""
        lim = self.cfg.action_bias_steps
        src = "/base" if getattr(getattr(self, "live", None), "masscom_obs_base", False) else ""
        if self.cfg.action_bias == 0.0 or self.step_counter >= lim:
            return "steady" + src
        return "bias" + src if self.step_counter + steps <= lim else None

    def advance_step_offset(self, steps: int) -> None:
        """Inside a CUDA-graph capture of `steps` control steps: make the next replay continue the Philox step sequence."""
        self.step_offset += steps

    def note_graph_replay(self, steps: int) -> None:
        """Host bookkeeping after a replay of a graph that contains `steps` control steps and one advance_step_offset(steps)."""
        self.step_counter += steps
        self._step_offset_host += steps

    def capture_steps(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, rew: Optional[torch.Tensor] = None,
                      done: Optional[torch.Tensor] = None):
        """Captures K = actions.shape[0] control steps (actions (K,N,2) read in place at every replay; optional outputs (K,N,13),
        (K,N), (K,N) long) into ONE CUDA graph and returns a `replay()` callable.  The Philox step index advances by K per replay
        through the device-side step offset, so replays continue the same random sequence the eager path would produce.  For
        small batches (BASELINE config[1]: 4096 envs) this removes the per-step launch / Python cost."""
        if self._buffers.step_offset in (None, 0):
            raise NotImplementedError("graph capture of control steps needs the device-side step offset (classic engine)")
        K = int(actions.shape[0])
        if self.first_call:
            raise RuntimeError("take at least one eager step before capturing (the first-call flag is baked into the graph)")
        if self.graph_key(K) is None:
            raise RuntimeError("the next K control steps straddle the end of the initial action bias (a host-side parameter the graph "
                               "would bake in): step eagerly past it, then capture")
        saved = self.step_counter
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(K):
                self.step(actions[k], None if obs is None else obs[k], None if rew is None else rew[k])
                if done is not None:
                    done[k].copy_(self.reset_buf)
            self.advance_step_offset(K)
        self.step_counter = saved                        # the capture ran the host code without executing anything

        def replay():
            graph.replay()
            self.note_graph_replay(K)

        replay.graph = graph
        return replay

    # ---- the hot path ------------------------------------------------------------------
    def step(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, rew: Optional[torch.Tensor] = None):
        """One control step of every env (== VecEnvRLGames.step).  Returns (obs (N,13), rew (N,), reset_buf (N,) long);
        the returned tensors are the env's own buffers (as in the reference, callers clone what they keep)."""
        obs = self.obs if obs is None else obs
        rew = self.rew if rew is None else rew
        p = self.params()
        rc = self.lib.usv_step_fused_f32(ctypes.byref(self._buffers), _lib.ptr(actions, torch.float32), _lib.ptr(obs),
                                         _lib.ptr(rew), ctypes.c_int64(self.num_envs), ctypes.byref(p), _lib.stream())
        _lib.check(rc, "usv_step_fused_f32")
        self.step_counter += 1
        self.first_call = False
        return obs, rew, self.reset_buf

    def rollout(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, rew: Optional[torch.Tensor] = None,
                done: Optional[torch.Tensor] = None):
        """T control steps in one launch; actions (T,N,2), optional outputs (T,N,13)/(T,N)/(T,N) long."""
        T = int(actions.shape[0])
        p = self.params()
        rc = self.lib.usv_rollout_fused_f32(ctypes.byref(self._buffers), _lib.ptr(actions, torch.float32), _lib.ptr(obs),
                                            _lib.ptr(rew), _lib.ptr(done), ctypes.c_int32(T), ctypes.c_int64(self.num_envs),
                                            ctypes.byref(p), _lib.stream())
        _lib.check(rc, "usv_rollout_fused_f32")
        self.step_counter += T
        self.first_call = False

    def planar_forces(self) -> torch.Tensor:
        """(N,8): body drag (u,v,r), net body wrench (Fx,Fy,Tz), world accel (ax,ay) of one sub-step."""
        out = torch.empty((self.num_envs, 8), dtype=torch.float32, device=self.device)
        p = self.params()
        _lib.check(self.lib.usv_planar_forces_f32(ctypes.byref(self._buffers), _lib.ptr(out), ctypes.c_int64(self.num_envs),
                                                  ctypes.byref(p), _lib.stream()), "usv_planar_forces_f32")
        return out

    def check_finite(self) -> None:
        """USV_NAN_PROBE semantics [ref: OIGE/envs/vec_env_rlgames.py:42-80] without a per-step host sync:
        the kernels OR a device flag; this reads it (one sync) and raises like the reference does."""
        flag = int(self.nonfinite.item())
        if flag != 0:
            what = "+".join(w for bit, w in ((1, "obs/reward"), (2, "actions(clamped)")) if flag & bit)
            raise RuntimeError(f"[USV_NAN_PROBE] non-finite detected: {what} of the fused env step")


class FusedUsvLiveEnv(FusedUsvEnv):
    """Variant B: the live CaptureXY task with 16 static obstacles and a per-env potential field
    [ref: OIGE/tasks/USV/USV_capture_xy_static_obs.py ; OIGE/tasks/USV/d_multi_gemini.py].  Per control step: one scene
    rebuild pass over the envs that reset (obstacle placement + potential field, usv_reset_b.cu) and one fused step kernel
    (usv_step_b.cu).  The potential fields cost 90 KB of HBM per env (16 384 envs = 1.5 GB)."""

    def __init__(self, cfg: UsvEnvConfig, live: Optional[UsvLiveConfig] = None, num_envs: Optional[int] = None, device="cuda:0",
                 env_id_offset: int = 0, collect_stats: bool = False):
        super().__init__(cfg, num_envs, device, env_id_offset, collect_stats=False)
        self.live = live if live is not None else UsvLiveConfig()
        n, nt = self.num_envs, self.stride // 32
        f32 = dict(dtype=torch.float32, device=self.device)
        self.bstate = torch.zeros((nt, E["USV_BS_COUNT"], 32), **f32)
        self.bconsts = torch.zeros((nt, E["USV_BC_COUNT"], 32), **f32)
        self.bstats = torch.zeros((nt, E["USV_BST_COUNT"], 32), **f32) if collect_stats else None
        for j in range(3):
            self.bconsts[:, E["USV_BC_COM_X"] + j] = self.live.com_base[j]
        self.task = int(self.live.task)
        # task.global_potential_field (only the obstacle task has one)
        self.potential = torch.zeros((n, GRID, GRID), **f32) if self.task == 0 else torch.zeros((1, GRID, GRID), **f32)
        self.obs_dim = LIVE_OBS_DIM - 8 + int(self.live.priv_dim)          # 33, or 29 with the 4-wide privileged tail
        self.obs = torch.zeros((n, self.obs_dim), **f32)
        # BatchedMapGPU.__init__: cell-centre coordinates  [ref: d_multi_gemini.py:15-19]
        ms = self.live.map_size
        cell = ms / GRID
        self.cell_centres = torch.linspace(-ms / 2 + cell / 2, ms / 2 - cell / 2, GRID, device=self.device)
        self.reset_epoch = torch.zeros(2, dtype=torch.int64, device=self.device)   # every env starts flagged: epoch 0 == step 0
        self.workspace = torch.zeros(int(self.lib.usv_live_scene_workspace_bytes(n)) // 4 + 4, dtype=torch.int32, device=self.device)
        self._live_params = self.live.to_params()
        lb = _lib.UsvLiveBuffers()
        lb.bstate, lb.bstate_stride = self.bstate.data_ptr(), self.stride
        lb.bconsts, lb.bconsts_stride = self.bconsts.data_ptr(), self.stride
        lb.bstats = self.bstats.data_ptr() if self.bstats is not None else None
        lb.bstats_stride = self.stride
        lb.field = self.potential.data_ptr()
        lb.reset_epoch = self.reset_epoch.data_ptr()
        self._live_buffers = lb

    def _src(self, name: str) -> torch.Tensor:
        if name.startswith("USV_BS_"):
            return self.bstate
        if name.startswith("USV_BC_"):
            return self.bconsts
        if name.startswith("USV_BST_"):
            return self.bstats
        return super()._src(name)

    # ---- host access to the live buffers -------------------------------------------------
    @property
    def obstacles(self) -> torch.Tensor:
        """(N,16,2) copy of xunlian_pos[:, :, :2] (env-local frame)."""
        o = self.bconsts[:, E["USV_BC_OBST"]:, :].permute(0, 2, 1).reshape(-1, E["USV_B_OBSTACLES"], 2)
        return o[: self.num_envs].clone()

    def set_obstacles(self, obstacles: torch.Tensor) -> None:
        o = torch.zeros((self.stride, E["USV_B_OBSTACLES"] * 2), dtype=torch.float32, device=self.device)
        o[: self.num_envs] = obstacles.to(self.device, torch.float32).reshape(self.num_envs, -1)
        self.bconsts[:, E["USV_BC_OBST"]:, :] = o.view(-1, 32, E["USV_B_OBSTACLES"] * 2).permute(0, 2, 1)

    def episode_outcomes(self):
        """(success, collision) latches  [ref: USV_capture_xy_static_obs.py:708-713]."""
        oc = self.int_field("USV_BS_OUTCOME")
        return (oc & 1).float(), ((oc >> 1) & 1).float()

    def mark_host_reset(self) -> None:
        """Call after writing reset_buf from the host (keeps the prev_potential=None quirk exact, see usv_b200.h)."""
        self.reset_epoch[self.step_counter & 1] = self.step_counter

    def bstats_matrix(self) -> torch.Tensor:
        return self.bstats.permute(1, 0, 2).reshape(E["USV_BST_COUNT"], -1)[:, : self.num_envs]

    def build_fields(self, obstacles: torch.Tensor, targets: torch.Tensor, want_cost: bool = False):
        """BatchedMapGPU on a dense batch: (m,16,2), (m,2) -> field (m,150,150) [, raw cost-to-go]."""
        m = int(obstacles.shape[0])
        obstacles = obstacles.to(self.device, torch.float32).contiguous()
        targets = targets.to(self.device, torch.float32).contiguous()
        field = torch.empty((m, GRID, GRID), dtype=torch.float32, device=self.device)
        cost = torch.empty_like(field) if want_cost else None
        need = int(self.lib.usv_live_scene_workspace_bytes(m)) // 4 + 4
        if self.workspace.numel() < need:       # a dense batch larger than the env count: the per-scene statistics need the room
            self.workspace = torch.zeros(need, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.usv_live_build_fields_f32(_lib.ptr(obstacles), _lib.ptr(targets), _lib.ptr(self.cell_centres),
                                                      _lib.ptr(field), _lib.ptr(cost), _lib.ptr(self.workspace),
                                                      ctypes.c_int64(m), _lib.stream()), "usv_live_build_fields_f32")
        return (field, cost) if want_cost else field

    # ---- the hot path ------------------------------------------------------------------
    def step(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, rew: Optional[torch.Tensor] = None,
             rebuild_scene: bool = True):
        """One control step of every env (== VecEnvRLGames.step over the live task).  Returns (obs (N, 25 + priv_dim), rew, reset_buf)."""
        obs = self.obs if obs is None else obs
        rew = self.rew if rew is None else rew
        p = self.params()
        n = ctypes.c_int64(self.num_envs)
        if rebuild_scene and self.task == 0:
            _lib.check(self.lib.usv_live_reset_scene_f32(ctypes.byref(self._buffers), ctypes.byref(self._live_buffers),
                                                         _lib.ptr(self.cell_centres), _lib.ptr(self.workspace), n, ctypes.byref(p),
                                                         _lib.stream()), "usv_live_reset_scene_f32")
        _lib.check(self.lib.usv_step_live_f32(ctypes.byref(self._buffers), ctypes.byref(self._live_buffers),
                                              _lib.ptr(actions, torch.float32), _lib.ptr(obs), _lib.ptr(rew), n, ctypes.byref(p),
                                              ctypes.byref(self._live_params), _lib.stream()), "usv_step_live_f32")
        self.step_counter += 1
        self.first_call = False
        return obs, rew, self.reset_buf

    def rollout(self, *a, **k):
        raise NotImplementedError("the multi-step rollout kernel exists for the classic task only (the live task's "
                                  "prev_potential quirk needs a grid-wide flag between control steps)")


class HostStepper:
    """Host-buffer stepping with the PCIe copies overlapped: `submit(k)` takes one control step's actions from pinned host memory
    and delivers that step's observations / rewards / dones into pinned host memory, asynchronously, on three streams (copy-in,
    compute, copy-out) with `depth` device staging slots, so the host->device copy of step k+1 and the device->host copy of step k
    run while the kernel of the step in between executes.  Dones travel as uint8 (what rl_games stores, [ref: RLG/common/
    a2c_common.py:454]) instead of the int64 reset_buf.  Results of a submit are valid after `wait(ticket)` / `synchronize()`."""

    def __init__(self, env: FusedUsvEnv, depth: int = 2):
        self.env, self.depth = env, int(depth)
        n, dev = env.num_envs, env.device
        od = env.obs.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        self.d_act = [torch.empty((n, 2), **f32) for _ in range(depth)]
        self.d_obs = [torch.empty((n, od), **f32) for _ in range(depth)]
        self.d_rew = [torch.empty(n, **f32) for _ in range(depth)]
        self.d_done = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_step = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.k = 0

    def submit(self, h_act: torch.Tensor, h_obs: torch.Tensor, h_rew: torch.Tensor, h_done: torch.Tensor) -> int:
        """All four tensors are pinned host tensors: actions (N,2) fp32 in; obs (N,D) fp32, rew (N,) fp32, done (N,) uint8 out."""
        slot = self.k % self.depth
        cur = torch.cuda.current_stream(self.env.device)
        if self.k >= self.depth:
            self.s_in.wait_event(self.ev_step[slot])      # the step that last read d_act[slot] has run
            cur.wait_event(self.ev_out[slot])             # ... and its outputs have left d_obs / d_rew / d_done[slot]
        with torch.cuda.stream(self.s_in):
            self.d_act[slot].copy_(h_act, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        cur.wait_event(self.ev_in[slot])
        self.env.step(self.d_act[slot], obs=self.d_obs[slot], rew=self.d_rew[slot])
        self.d_done[slot].copy_(self.env.reset_buf)       # int64 -> uint8 on the device
        self.ev_step[slot].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_step[slot])
            h_obs.copy_(self.d_obs[slot], non_blocking=True)
            h_rew.copy_(self.d_rew[slot], non_blocking=True)
            h_done.copy_(self.d_done[slot], non_blocking=True)
            self.ev_out[slot].record(self.s_out)
        self.k += 1
        return self.k - 1

    def wait(self, ticket: int) -> None:
        if self.k - ticket <= self.depth:                 # still tracked by a slot event
            self.ev_out[ticket % self.depth].synchronize()

    def synchronize(self) -> None:
        self.s_out.synchronize()
