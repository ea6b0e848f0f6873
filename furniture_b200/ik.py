"""Host side of control_type="ik" for the Sawyer env (FurnitureEnv._do_ik_step, furniture/env/furniture.py:2899-2996, and
SawyerIKController, furniture/env/controllers/sawyer_ik_controller.py:46-299).

The reference solves the inverse kinematics in a pybullet copy of the arm (`p.calculateInverseKinematics`, 20 calls, joint damping 0.1,
rest poses, joint limits) -- a host round trip per env step and a dependency (pybullet + its URDF) that is absent here.  This build keeps
everything the reference does around that call (action scaling and axis swap, workspace clipping, the accumulated target orientation with
its quaternion conventions, the 0.3 sensitivity, the P controller -5 * (q - q_cmd) clipped to [-1, 1], three closed-loop repeats of
_do_simulation) and replaces the solver by a damped-least-squares IK on the arm's own kinematic chain, run by the warp that owns the env
(csrc/fe_ik.h); the same algorithm in numpy float64 is test infrastructure (oracle/ik_oracle.py: solve_ik).  The joint
targets therefore differ from pybullet's within the arm's one-dimensional null space; the end-effector pose they reach is the same target.

`quaternion_mode=1` is control_type="ik_quaternion" (furniture.py:2998-3058): move 3, a quaternion (w, x, y, z) relative to the hand's
current orientation, gripper, connect -- 9 numbers, nothing accumulated.

Two arms (Baxter, controllers/baxter_ik_controller.py) are two independent chains with that controller's constants.

  arm_chain(model, arm) -> the 7 joint frames from the robot base to `<arm>_hand`, taken from the composed model at zero joint angles
  ik_config(model)   -> struct fe_ik_config (include/furniture_b200.h): gains, limits, rest pose, workspace, the chain
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import mjcf

i32, f32 = C.c_int32, C.c_float
NJ = 7
REST_POSE = [0, -1.18, 0.00, 2.18, 0.00, 0.57, 3.3161]          # sawyer_ik_controller.py:268
LOWER = [-3.05, -3.82, -3.05, -3.05, -2.98, -2.98, -4.71]       # :196-198
UPPER = [3.05, 2.28, 3.05, 3.05, 2.98, 2.98, 4.71]


class FeIkArm(C.Structure):
    _fields_ = [("rest_pose", f32 * NJ), ("lower", f32 * NJ), ("upper", f32 * NJ),
                ("link_pos", (f32 * 3) * NJ), ("link_quat", (f32 * 4) * NJ), ("jaxis", (f32 * 3) * NJ), ("jpos", (f32 * 3) * NJ),
                ("hand_pos", f32 * 3), ("hand_quat", f32 * 4), ("arm_qadr", i32 * NJ), ("pad_", i32)]


class FeIkConfig(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("action_repeat", i32), ("max_iters", i32), ("quaternion_mode", i32),
        ("move_speed", f32), ("rotate_speed", f32), ("user_sensitivity", f32), ("kp", f32), ("damping", f32), ("null_gain", f32),
        ("tol_pos", f32), ("tol_rot", f32), ("max_step_pos", f32), ("max_step_rot", f32),
        ("min_pos", f32 * 3), ("max_pos", f32 * 3), ("base_pos", f32 * 3), ("base_quat", f32 * 4), ("narms", i32), ("pad_", i32),
        ("arm", FeIkArm * 2),
    ]

    # the first arm's chain, under the names a one-arm caller expects
    hand_pos = property(lambda self: self.arm[0].hand_pos)
    hand_quat = property(lambda self: self.arm[0].hand_quat)


def arm_chain(m: mjcf.Model, arm: int = 0):
    """frames of the joint bodies of arm `arm` (0: right, 1: left) relative to one another at zero joint angles: link_pos / link_quat [k] =
    body of joint k in the frame of the body of joint k-1 (k = 0: in the world), joint axis / anchor in the body frame, `<arm>_hand` in the
    last joint body's frame, and the world pose of the robot `base` body (targets are kept in the base frame, furniture.py:3381-3427)"""
    joints = list(m.meta["robot_joints"])[NJ * arm : NJ * arm + NJ]
    q0 = np.array(m.qpos0, dtype=np.float64)
    for jn in m.meta["robot_joints"]:
        q0[int(m.jnt_qposadr[m.names["jnt"].index(jn)])] = 0.0
    kin = mjcf.kinematics_np(m, q0)
    ch = dict(link_pos=[], link_quat=[], jaxis=[], jpos=[], qadr=[])
    prev_p, prev_q = np.zeros(3), np.array([1.0, 0, 0, 0])
    for jn in joints:
        j = m.names["jnt"].index(jn)
        b = int(m.jnt_bodyid[j])
        p, q = kin["xpos"][b], kin["xquat"][b]
        Rp = mjcf.q_to_mat(prev_q)
        ch["link_pos"].append(Rp.T @ (p - prev_p))
        ch["link_quat"].append(mjcf.q_norm(mjcf.q_mul(mjcf.q_conj(prev_q), q)))
        ch["jaxis"].append(np.array(m.jnt_axis[j], dtype=np.float64))
        ch["jpos"].append(np.array(m.jnt_pos[j], dtype=np.float64))
        ch["qadr"].append(int(m.jnt_qposadr[j]))
        prev_p, prev_q = p, q
    hb = m.names["body"].index(m.meta["hand_body" if arm == 0 else "hand_body2"])
    Rp = mjcf.q_to_mat(prev_q)
    ch["hand_pos"] = Rp.T @ (kin["xpos"][hb] - prev_p)
    ch["hand_quat"] = mjcf.q_norm(mjcf.q_mul(mjcf.q_conj(prev_q), kin["xquat"][hb]))
    bb = m.names["body"].index("base")
    ch["base_pos"], ch["base_quat"] = kin["xpos"][bb].copy(), kin["xquat"][bb].copy()
    ch["lower"] = [float(m.jnt_range[m.names["jnt"].index(jn)][0]) for jn in joints]
    ch["upper"] = [float(m.jnt_range[m.names["jnt"].index(jn)][1]) for jn in joints]
    return ch


IK_DEFAULTS = dict(quaternion_mode=0, action_repeat=3, max_iters=20, move_speed=0.1, rotate_speed=22.5, user_sensitivity=0.3, kp=5.0, damping=0.1, null_gain=0.0,
                   tol_pos=1e-4, tol_rot=1e-3, max_step_pos=0.05, max_step_rot=0.2, min_pos=(-1.5, -1.5, 0.0), max_pos=(1.5, 1.5, 1.5))
# BaxterIKController (controllers/baxter_ik_controller.py): user_sensitivity 1.0 (:43), velocities -2 * delta (:96), jointDamping 0.7 (:251)
BAXTER_OVERRIDES = dict(user_sensitivity=1.0, kp=2.0, damping=0.7)


def ik_params(m: mjcf.Model, **kw):
    """plain-Python view of the IK parameters (what `ik_config` packs); config/furniture.py:84-89 for the speeds, furniture.py:166-172 for
    the workspace and the three repeats, sawyer_ik_controller.py / baxter_ik_controller.py for sensitivity, gain, damping, rest pose, limits.
    One chain per arm (`chains`; `chain` = the first); the Sawyer's rest pose and limits are the controller's literals, Baxter's limits the
    joint ranges of the model (its controller reads them from the URDF) and its rest pose the current joints (null_gain is 0 anyway)"""
    narms = 2 if m.meta.get("eef_site2") else 1
    p = dict(IK_DEFAULTS)
    if narms == 2:
        p.update(BAXTER_OVERRIDES)
    for k, v in kw.items():
        if k not in p:
            raise KeyError(k)
        p[k] = v
    chains = [arm_chain(m, a) for a in range(narms)]
    if narms == 1:
        chains[0]["lower"], chains[0]["upper"], chains[0]["rest_pose"] = list(LOWER), list(UPPER), list(REST_POSE)
    else:
        for ch in chains:
            ch["rest_pose"] = [0.0] * NJ
    p.update(narms=narms, chains=chains, chain=chains[0], rest_pose=chains[0]["rest_pose"], lower=chains[0]["lower"], upper=chains[0]["upper"])
    return p


def ik_config(m: mjcf.Model, **kw) -> FeIkConfig:
    p = ik_params(m, **kw)
    c = FeIkConfig()
    c.struct_bytes, c.narms = C.sizeof(FeIkConfig), p["narms"]
    for k in ("action_repeat", "max_iters", "quaternion_mode"):
        setattr(c, k, int(p[k]))
    for k in ("move_speed", "rotate_speed", "user_sensitivity", "kp", "damping", "null_gain", "tol_pos", "tol_rot", "max_step_pos", "max_step_rot"):
        setattr(c, k, float(p[k]))
    for k in ("min_pos", "max_pos"):
        getattr(c, k)[:] = [float(x) for x in p[k]]
    c.base_pos[:] = list(p["chain"]["base_pos"]); c.base_quat[:] = list(p["chain"]["base_quat"])
    for a, ch in enumerate(p["chains"]):
        arm = c.arm[a]
        for k in ("rest_pose", "lower", "upper"):
            getattr(arm, k)[:] = [float(x) for x in ch[k]]
        for k in range(NJ):
            arm.link_pos[k][:] = list(ch["link_pos"][k]); arm.link_quat[k][:] = list(ch["link_quat"][k])
            arm.jaxis[k][:] = list(ch["jaxis"][k]); arm.jpos[k][:] = list(ch["jpos"][k])
            arm.arm_qadr[k] = ch["qadr"][k]
        arm.hand_pos[:] = list(ch["hand_pos"]); arm.hand_quat[:] = list(ch["hand_quat"])
    return c
