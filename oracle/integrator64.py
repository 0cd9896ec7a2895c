"""TEST INFRASTRUCTURE ONLY.  float64 host integration of the planar rigid-body stand-in for the
PhysX step (north_star: "The integrator is checked against a float64 host
integration").  numpy, vectorised over envs; no torch, no fp32 anywhere.

Model (DESIGN.md "integrator"): state (x, y, psi, vx, vy, r), thruster lag state (thrL, thrR);
per sub-step of length dt:
    thr   <- thr*alpha + (1-alpha)*target                       [OIGE/envs/USV/ThrusterDynamics.py:129-141]
    u,v   <- R(psi)^T (vx,vy)                                   [OIGE/envs/USV/Hydrodynamics.py:213-222]
    drag  <- -((L + Q|nu|) * s * k) nu   per DOF u,v,r          [Hydrodynamics.py:176-205,243]
    F_b   <- F_dist + drag_uv + (thrL+thrR, 0)                  [SNAP/USV_Virtual.py:621-650]
    tau   <- tau_dist + drag_r - yL*thrL - yR*thrR              (thruster mounts, heron.urdf:167,242)
    v_w   <- v_w + dt * R(psi) F_b / m ;  r <- r + dt * tau/(Izz*kIz)      (semi-implicit Euler)
    x_w   <- x_w + dt * v_w           ;  psi <- psi + dt * r
"""
from __future__ import annotations

import numpy as np


def substeps(state: dict, const: dict, target: np.ndarray, *, dt: float, alpha: float, n_substeps: int,
             izz: float, thr_y_left: float, thr_y_right: float, scaling: float = 1.0, use_drag_scale: bool = False,
             use_const_force=False, use_sin_force=False, use_const_torque=False, use_sin_torque=False,
             origin=None):
    """All arrays float64.  state: x,y,psi,vx,vy,r,thrL,thrR ; const: mass,lin(3),quad(3),kdrag,kiz and
    the disturbance parameters; target: (n,2) thrust targets after multipliers.  Returns the new state."""
    s = {k: np.array(v, dtype=np.float64) for k, v in state.items()}
    c = {k: np.array(v, dtype=np.float64) for k, v in const.items()}
    tgt = np.asarray(target, dtype=np.float64)
    ox, oy = (0.0, 0.0) if origin is None else (np.asarray(origin[0], np.float64), np.asarray(origin[1], np.float64))
    for _ in range(n_substeps):
        s["thrL"] = s["thrL"] * alpha + (1.0 - alpha) * tgt[:, 0]
        s["thrR"] = s["thrR"] * alpha + (1.0 - alpha) * tgt[:, 1]
        cs, sn = np.cos(s["psi"]), np.sin(s["psi"])
        u = cs * s["vx"] + sn * s["vy"]
        v = -sn * s["vx"] + cs * s["vy"]
        w = s["r"]
        k = c["kdrag"] if use_drag_scale else 1.0
        du = -((c["lin"][:, 0] + c["quad"][:, 0] * np.abs(u)) * scaling * k) * u
        dv = -((c["lin"][:, 1] + c["quad"][:, 1] * np.abs(v)) * scaling * k) * v
        dr = -((c["lin"][:, 2] + c["quad"][:, 2] * np.abs(w)) * scaling * k) * w
        fdx = fdy = td = 0.0
        if use_const_force:
            fdx, fdy = c["fcx"], c["fcy"]
        if use_sin_force:
            fdx = c["fcx"] + np.sin((s["x"] + ox) * c["fxf"] + c["fxs"]) * c["famp"]
            fdy = c["fcy"] + np.sin((s["y"] + oy) * c["fyf"] + c["fys"]) * c["famp"]
        if use_const_torque:
            td = c["tc"]
        if use_sin_torque:
            td = c["tc"] + np.sin(((s["x"] + ox) + (s["y"] + oy)) * c["tf"] + c["ts"]) * c["tamp"]
        Fx = fdx + du + s["thrL"] + s["thrR"]
        Fy = fdy + dv
        Tz = td + dr - thr_y_left * s["thrL"] - thr_y_right * s["thrR"]
        ax = (cs * Fx - sn * Fy) / c["mass"]
        ay = (sn * Fx + cs * Fy) / c["mass"]
        rdot = Tz / (izz * c["kiz"])
        s["vx"] = s["vx"] + dt * ax
        s["vy"] = s["vy"] + dt * ay
        s["r"] = s["r"] + dt * rdot
        s["x"] = s["x"] + dt * s["vx"]
        s["y"] = s["y"] + dt * s["vy"]
        s["psi"] = s["psi"] + dt * s["r"]
    return s
