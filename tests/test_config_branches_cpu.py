"""Configuration branches of the live USVVirtual that no shipped YAML takes -- partial mass-driven coupling target lists, the
independent k_Iz draw, the legacy disc-shaped CoM randomisation, the water current -- : the oracle's restatement pinned to the
reference's own outputs (tests/golden/config_branches.npz, written by oracle/make_golden.py:config_branches) and the YAML mapping of
the product config.  The CUDA halves are the `_BRANCHES` variants of tests/test_gpu_parity.py::test_fused_step_vs_oracle_lockstep and
tests/test_gpu_live.py::test_live_legacy_com_disc_vs_oracle."""
import copy
import dataclasses
import os

import numpy as np
import pytest
import torch

from omniisaacgymenvs_loop_b200.config import (UsvEnvConfig, UsvLiveConfig, live_default_config, live_env_config, live_task_cfg)
from oracle import philox
from oracle import usv_oracle as O
from oracle import usv_oracle_b as B
from tests.util import oracle_cfg

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "config_branches.npz"))
T = lambda k: torch.from_numpy(G[k])


@pytest.mark.parametrize("tag,bits", [("drag", 1), ("thr_kiz", 6), ("kiz", 4)])
def test_partial_coupling_vs_reference_golden(tag, bits):
    """_apply_mass_driven_coupling with a partial target list: targets follow the mass, the others keep what the independent
    randomisations wrote  [ref: OIGE/tasks/USV_Virtual.py:988-1040]."""
    c = O.EnvConfig(mass_coupling=True, couple_targets=bits, mass_base=float(G["cpl_mass_base"]), couple_mass_max=float(G["cpl_mass_max"]),
                    kdrag_min=float(G["cpl_kdrag_rng"][0]), kdrag_max=float(G["cpl_kdrag_rng"][1]), couple_thr_a=float(G["cpl_thr_a"]),
                    couple_kiz_min=float(G["cpl_kiz_rng"][0]), couple_kiz_max=float(G["cpl_kiz_rng"][1]))
    kd, sthr, kiz = O.mass_coupling(c, T(f"cpl_{tag}_mass"))
    want = {"kdrag": kd if bits & 1 else torch.full_like(kd, 1.25), "thr_l": sthr if bits & 2 else torch.full_like(kd, 0.75),
            "thr_r": sthr if bits & 2 else torch.full_like(kd, 1.125), "kiz": kiz if bits & 4 else torch.full_like(kd, 1.375)}
    for k, v in want.items():
        assert torch.equal(v, T(f"cpl_{tag}_{k}")), k


def test_partial_coupling_inside_the_oracle_reset():
    """reset_idx applies exactly the listed targets: the others equal the reset of the same config without coupling."""
    base = dict(mass_rand=True, mass_min=30.0, mass_max=54.96, mass_base=34.96, use_drag_scale=True, kdrag_rand=True, kdrag_min=0.7,
                kdrag_max=1.6, thr_rand=True, thr_separate=True, kiz_rand=True, couple_kiz_min=0.8, couple_kiz_max=1.7)
    n = 64
    free = O.ClassicEnvOracle(O.EnvConfig(**base), n)
    free.reset_idx(torch.arange(n), 3)
    for bits in (1, 2, 4, 5, 7):
        c = O.EnvConfig(**base, mass_coupling=True, couple_targets=bits)
        orc = O.ClassicEnvOracle(c, n)
        orc.reset_idx(torch.arange(n), 3)
        kd, sthr, kiz = O.mass_coupling(c, orc.mass)
        assert torch.equal(orc.mass, free.mass)
        assert torch.equal(orc.drag_scale[:, 0], kd if bits & 1 else free.drag_scale[:, 0])
        assert torch.equal(orc.thr_mult_left, sthr if bits & 2 else free.thr_mult_left)
        assert torch.equal(orc.thr_mult_right, sthr if bits & 2 else free.thr_mult_right)
        assert torch.equal(orc.k_iz, kiz if bits & 4 else free.k_iz)
        assert not torch.equal(free.thr_mult_left, free.thr_mult_right)


@pytest.mark.parametrize("space", ["linear", "log"])
def test_independent_k_iz_vs_reference_golden(space):
    """_sample_k_iz on the uniforms the reference drew  [ref: OIGE/tasks/USV_Virtual.py:153-170]."""
    lo, hi = float(G["kiz_rng"][0]), float(G["kiz_rng"][1])
    got = O.sample_k_iz(T(f"kiz_{space}_u"), lo, hi, space == "log")
    assert torch.equal(got, T(f"kiz_{space}"))
    # and inside reset_idx: stream RS_RESET_5, word 1
    n = 32
    orc = O.ClassicEnvOracle(O.EnvConfig(kiz_rand=True, kiz_log=space == "log", couple_kiz_min=lo, couple_kiz_max=hi), n)
    orc.reset_idx(torch.arange(n), 9)
    u = torch.from_numpy(philox.uniform4(orc.cfg.seed, orc.env_ids, 9, philox.RS_RESET[5]))[:, 1]
    assert torch.equal(orc.k_iz, O.sample_k_iz(u, lo, hi, space == "log"))
    assert float(orc.k_iz.min()) >= lo and float(orc.k_iz.max()) <= hi and float(orc.k_iz.std()) > 0.05


def test_legacy_com_disc_vs_reference_golden():
    """MDD._randomize_com without com_displacement_xyz  [ref: OIGE/tasks/USV/USV_disturbances.py:108-124]."""
    got = O.com_disc(tuple(float(x) for x in G["com_disc_base"]), T("com_disc_u_r"), T("com_disc_u_theta"), float(G["com_disc_max"]))
    assert torch.equal(got, T("com_disc"))
    assert torch.equal(got[:, 2], torch.full((got.shape[0],), float(G["com_disc_base"][2])))


def test_water_current_in_the_oracle_step_is_the_reference_hydrodynamics():
    """planar_wrench hands the flow to the A2 restatement (pinned with a current in tests/test_oracle_cpu.py): the drag of a hull at
    rest in a current equals the drag of a hull moving at minus the current in still water."""
    n = 16
    still = O.ClassicEnvOracle(O.EnvConfig(), n)
    flow = O.ClassicEnvOracle(O.EnvConfig(use_water_current=True, flow_vel_xy=(0.4, -0.25)), n)
    g = torch.Generator().manual_seed(2)
    psi = torch.rand(n, generator=g) * 6.0 - 3.0
    for o in (still, flow):
        o.psi[:] = psi
        o.r[:] = 0.0
        o.current_forces[:] = 0.0
    still.vel[:, 0], still.vel[:, 1] = -0.4, 0.25
    flow.vel[:] = 0.0
    a, b = still.planar_wrench(), flow.planar_wrench()
    for x, y in zip(a, b):
        assert torch.allclose(torch.as_tensor(x), torch.as_tensor(y), rtol=1e-6, atol=1e-6)


def test_yaml_mapping_of_the_branches():
    t = live_task_cfg()
    dist = t["env"]["disturbances"]
    # shipped: full target list, no independent k_Iz, no current, box-shaped CoM
    cfg = live_env_config(t)
    assert cfg.mass_coupling and cfg.couple_targets == 7 and not cfg.kiz_rand and not cfg.use_water_current
    assert UsvLiveConfig.from_task_cfg(t).com_rand == 1
    # partial targets + independent log-space k_Iz + current + legacy disc
    t2 = copy.deepcopy(t)
    d2 = t2["env"]["disturbances"]
    d2["coupling"]["mass_driven"]["targets"] = ["drag_scale", "yaw_inertia"]
    d2["inertia"].update(use_yaw_inertia_randomization=True, k_Iz_sample_space="log")
    t2["env"]["water_current"] = {"use_water_current": True, "flow_velocity": [0.2, -0.1, 0.0]}
    d2["mass"]["com_displacement_xyz"] = None
    d2["mass"]["CoM_max_displacement"] = 0.12
    cfg2, live2 = live_env_config(t2), UsvLiveConfig.from_task_cfg(t2)
    assert cfg2.couple_targets == 5 and cfg2.kiz_rand and cfg2.kiz_log
    assert cfg2.use_water_current and cfg2.flow_vel_xy == (0.2, -0.1)
    assert live2.com_rand == 2 and live2.com_disp[0] == pytest.approx(0.12)
    # thruster is not a target any more: its privileged range is the independent one (1 - a, 1 + a)
    a = float(d2["thruster"]["thruster_rand"])
    assert live2.priv_a[1] == pytest.approx(1.0 - a) and live2.priv_b[1] == pytest.approx(2 * a)
    # round trip through the writer
    t3 = live_task_cfg(cfg2, live2)
    cfg3, live3 = live_env_config(t3), UsvLiveConfig.from_task_cfg(t3)
    for f in ("couple_targets", "kiz_rand", "kiz_log", "use_water_current", "flow_vel_xy", "mass_coupling"):
        assert getattr(cfg3, f) == getattr(cfg2, f), f
    assert live3.com_rand == 2 and live3.com_disp[0] == pytest.approx(0.12)
    # the struct carries them
    p = cfg2.to_params()
    assert p.couple_targets == 5 and p.kiz_rand == 1 and p.kiz_log == 1 and p.use_water_current == 1
    assert (p.flow_vel_xy[0], p.flow_vel_xy[1]) == (pytest.approx(0.2), pytest.approx(-0.1))
    assert live2.to_params().com_rand == 2
    with pytest.raises(ValueError):
        t4 = copy.deepcopy(t)
        t4["env"]["disturbances"]["coupling"]["mass_driven"]["targets"] = ["drag_scale", "ballast"]
        live_env_config(t4)


def test_oracle_cfg_carries_the_new_fields():
    c = oracle_cfg(dataclasses.replace(UsvEnvConfig(), couple_targets=3, kiz_rand=True, kiz_log=True, use_water_current=True, flow_vel_xy=(0.1, 0.2)))
    assert (c.couple_targets, c.kiz_rand, c.kiz_log, c.use_water_current, tuple(c.flow_vel_xy)) == (3, True, True, True, (0.1, 0.2))
