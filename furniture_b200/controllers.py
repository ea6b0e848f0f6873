"""Host side of the torque controllers (NEW_CONTROLLERS of furniture/env/furniture.py:41-47; furniture/env/controllers/arm_controller.py).

`ctl_config(name)` packs struct fe_ctl_config (include/furniture_b200.h) from the parameters of controllers/controller_config.hjson the way
FurnitureEnv._load_controller builds a controller (furniture.py:1664-1701): action ranges, the fixed impedances (impedance_flag is False in
that file for every controller), joint gains, and the length of the goal ramp as the reference computes it.  The arithmetic itself is
csrc/fe_ctl.h; oracle/controller_oracle.py is its float64 numpy counterpart, pinned to the reference's own classes.
"""
from __future__ import annotations

import ctypes as C
import math

i32, f32, f64 = C.c_int32, C.c_float, C.c_double
MODES = {"joint_torque": 0, "joint_velocity": 1, "joint_impedance": 2, "position_orientation": 3, "position": 4}
RAMP_RATIO, CONTROL_FREQ = 0.20, 20  # Controller.__init__ :76, control_freq default :30 (joint_impedance's hjson entry: 20 as well)

CONTROLLER_DEFAULTS = {  # controllers/controller_config.hjson
    "position_orientation": dict(control_range_pos=0.05, control_range_ori=0.2, initial_impedance_pos=150.0, initial_impedance_ori=150.0, initial_damping=1.0),
    "position": dict(control_range_pos=0.05, initial_impedance_pos=150.0, initial_impedance_ori=150.0, initial_damping=1.0),
    "joint_impedance": dict(control_range=[0.2] * 7, kp_max=[100, 100, 100, 100, 50, 30, 10], kp_min=[10, 10, 10, 10, 10, 1, 1], damping_max=[2] * 7, damping_min=[0] * 7),
    "joint_velocity": dict(control_range=[1] * 7, kv=[8.0, 7.0, 6.0, 4.0, 2.0, 0.5, 0.1]),
    "joint_torque": dict(control_range=[0.5, 0.5, 0.5, 0.2, 0.2, 0.1, 0.1]),
}


class FeCtlConfig(C.Structure):
    _fields_ = [("struct_bytes", i32), ("mode", i32), ("control_dim", i32), ("pad_", i32), ("move_speed", f64),
                ("control_max", f64 * 7), ("kp", f64 * 7), ("damping", f64 * 7), ("kv", f64 * 7), ("ramp_steps", f64),
                ("hand_pos", f32 * 3), ("hand_quat", f32 * 4)]


def ctl_config(name, model=None, timestep=None, move_speed=0.1, **overrides) -> FeCtlConfig:
    """`model`: the composed scene (mjcf.Model) the controller will drive -- gives the model timestep and where `right_hand` sits in the link
    that carries it; without it (the arithmetic test hook) the hand frame is the link frame"""
    if name not in MODES:
        raise KeyError("unknown controller %r (one of %s)" % (name, sorted(MODES)))
    p = dict(CONTROLLER_DEFAULTS[name])
    for k, v in overrides.items():
        if k not in p:
            raise KeyError(k)
        p[k] = v
    c = FeCtlConfig()
    c.struct_bytes, c.mode, c.move_speed = C.sizeof(FeCtlConfig), MODES[name], float(move_speed)
    if timestep is None:
        timestep = float(model.opt_timestep) if model is not None else 0.002
    c.ramp_steps = float(math.floor(RAMP_RATIO * CONTROL_FREQ / timestep))
    c.hand_quat[0] = 1.0
    if model is not None:
        from .ik import arm_chain

        ch = arm_chain(model)
        c.hand_pos[:] = [float(x) for x in ch["hand_pos"]]
        c.hand_quat[:] = [float(x) for x in ch["hand_quat"]]
    if name in ("position", "position_orientation"):
        cmax = [p["control_range_pos"]] * 3 + ([p["control_range_ori"]] * 3 if name == "position_orientation" else [])
        kp = [p["initial_impedance_pos"]] * 3 + [p["initial_impedance_ori"]] * 3
        for k in range(6):
            c.kp[k], c.damping[k] = float(kp[k]), float(p["initial_damping"])
    else:
        cmax = list(p["control_range"])
        if name == "joint_impedance":
            for k in range(7):
                c.kp[k] = (p["kp_max"][k] + p["kp_min"][k]) * 0.5
                c.damping[k] = (p["damping_max"][k] + p["damping_min"][k]) * 0.5
        if name == "joint_velocity":
            for k in range(7):
                c.kv[k] = float(p["kv"][k])
    c.control_dim = len(cmax)
    for k, v in enumerate(cmax):
        c.control_max[k] = float(v)
    return c
