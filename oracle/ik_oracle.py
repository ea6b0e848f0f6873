"""CPU ORACLE for control_type="ik" (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

  ik_pre        <- FurnitureEnv._do_ik_step up to the controller call  furniture/env/furniture.py:2909-2931, _bounded_d_pos :1252-1258,
                   _make_input :1332-1343; transform_utils.euler_to_quat :617-630, quat_multiply :33-49, quat_inverse :112-119,
                   quat2mat :207-229 (float32 where the reference computes in float32)
  IKOracle      <- SawyerIKController.get_control / sync_state / joint_positions_for_eef_command
                   furniture/env/controllers/sawyer_ik_controller.py:46-118, :248-281 -- with the pybullet solve replaced by
                   solve_ik below (damped least squares on the arm's own chain, float64), which is the algorithm the device
                   runs in float32 (csrc/fe_ik.h)

ik_pre is pinned against tests/golden/ik_pre.npz (the reference's own _do_ik_step run around stand-ins, tools/make_golden_ik.py).
The solver is PARITY UNPINNED against pybullet (absent here and on the GPU box, profiles/r2_probe_mujoco.log): it is checked for what an
IK has to deliver -- the commanded joints reach the commanded hand pose -- and device against this copy.
"""
import math

import numpy as np

from furniture_b200 import ik as IK
from furniture_b200 import mjcf


# ------------------------------------------------------------------ the solver, numpy float64 (same steps as csrc/fe_ik.h: fe_ik_fk, fe_ik_solve)
def chain_fk(ch, q):
    """world pose of `right_hand` and the world anchors / axes of the 7 joints"""
    p, quat = np.zeros(3), np.array([1.0, 0, 0, 0])
    anchors, axes = [], []
    for k in range(IK.NJ):
        R = mjcf.q_to_mat(quat)
        p0 = p + R @ ch["link_pos"][k]
        q0 = mjcf.q_mul(quat, ch["link_quat"][k])
        R0 = mjcf.q_to_mat(q0)
        anchors.append(p0 + R0 @ ch["jpos"][k])
        axes.append(R0 @ ch["jaxis"][k])
        quat = mjcf.q_norm(mjcf.q_mul(q0, mjcf.q_axis_angle(ch["jaxis"][k], q[k])))
        p = anchors[-1] - mjcf.q_to_mat(quat) @ ch["jpos"][k]
    R = mjcf.q_to_mat(quat)
    return p + R @ ch["hand_pos"], mjcf.q_norm(mjcf.q_mul(quat, ch["hand_quat"])), np.array(anchors), np.array(axes)


def rot_error(q_target, q_cur):
    """rotation vector of q_target * conj(q_cur) (world frame)"""
    e = mjcf.q_mul(q_target, mjcf.q_conj(q_cur))
    if e[0] < 0:
        e = -e
    n = np.linalg.norm(e[1:])
    if n < 1e-9:
        return 2.0 * e[1:]
    return 2.0 * np.arctan2(n, e[0]) * e[1:] / n


def solve_ik(p, q_start, target_pos_world, target_quat_world, ch=None):
    """damped least squares with a null-space pull to the rest pose and joint limits; returns (q, iterations); `ch`: the arm's chain"""
    ch = ch or p["chain"]
    q = np.array(q_start, dtype=np.float64)
    lam2 = p["damping"] ** 2
    rest, lo, hi = np.array(ch.get("rest_pose", p["rest_pose"])), np.array(ch.get("lower", p["lower"])), np.array(ch.get("upper", p["upper"]))
    it = 0
    for it in range(p["max_iters"]):
        hp, hq, anchors, axes = chain_fk(ch, q)
        ep = target_pos_world - hp
        er = rot_error(target_quat_world, hq)
        np_, nr = np.linalg.norm(ep), np.linalg.norm(er)
        if np_ < p["tol_pos"] and nr < p["tol_rot"]:
            break
        if np_ > p["max_step_pos"]:
            ep = ep * (p["max_step_pos"] / np_)
        if nr > p["max_step_rot"]:
            er = er * (p["max_step_rot"] / nr)
        J = np.zeros((6, IK.NJ))
        for k in range(IK.NJ):
            J[:3, k] = np.cross(axes[k], hp - anchors[k])
            J[3:, k] = axes[k]
        A = J @ J.T + lam2 * np.eye(6)
        e = np.concatenate([ep, er])
        z = p["null_gain"] * (rest - q)
        dq = J.T @ np.linalg.solve(A, e - J @ z) + z
        q = np.clip(q + dq, lo, hi)
    return q, it


def _hamilton(a, b):  # (w, x, y, z)
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def euler_to_quat(rot_deg, quat):
    """T.euler_to_quat(rotation, quat): q3 * q2 * q1 about z, y, x, left-multiplied by Quaternion(quat) -- pyquaternion reads the four
    numbers as (w, x, y, z), whatever the caller meant (the env hands it an (x, y, z, w) array)"""
    def ax(axis, deg):
        h = np.deg2rad(deg) / 2.0
        return np.concatenate([[np.cos(h)], np.asarray(axis, dtype=np.float64) * np.sin(h)])

    q = _hamilton(_hamilton(ax([0, 0, 1], rot_deg[2]), ax([0, 1, 0], rot_deg[1])), ax([1, 0, 0], rot_deg[0]))
    return list(_hamilton(np.asarray(quat, dtype=np.float64), q))


def quat_multiply(q1, q0):  # (x, y, z, w), float32 result
    x0, y0, z0, w0 = q0
    x1, y1, z1, w1 = q1
    return np.array((x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0, -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0, x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
                     -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0), dtype=np.float32)


def quat_inverse(q):
    c = np.array((-q[0], -q[1], -q[2], q[3]), dtype=np.float32)
    return c / np.dot(q, q)


def quat2mat(q_xyzw):
    q = np.array(q_xyzw, dtype=np.float32, copy=True)[[3, 0, 1, 2]]
    n = np.dot(q, q)
    if n < np.finfo(float).eps * 4.0:
        return np.identity(3)
    q *= math.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([[1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0]], [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0]],
                     [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2]]])


def mat2quat(R):
    """(x, y, z, w) float32 of a rotation matrix; the reference takes an eigenvector (transform_utils.py:298-375), whose sign is free --
    every use of it here is invariant under q -> -q"""
    w, x, y, z = mjcf.mat_to_q(np.asarray(R, dtype=np.float64))
    return np.array([x, y, z, w], dtype=np.float32)


def ik_pre(action, hand_pos_world, hand_quat_base, s, p):
    """-> (dpos, rotation (3,3) in the base frame, new accumulated target s, gripper action)"""
    a = np.array(action, dtype=np.float64)
    a[:3] = a[:3] * p["move_speed"]
    a[:3] = [-a[1], a[0], a[2]]
    d_pos = np.clip(a[:3], np.asarray(p["min_pos"]) - hand_pos_world, np.asarray(p["max_pos"]) - hand_pos_world)
    s_new = euler_to_quat(a[3:6] * p["rotate_speed"], s)
    d_quat = quat_multiply(quat_inverse(hand_quat_base), s_new)
    rotation = quat2mat(quat_multiply(hand_quat_base, d_quat))
    return d_pos, rotation, s_new, a[-2]


def ik_pre_quaternion(action, hand_pos_world, hand_quat_base, p):
    """control_type="ik_quaternion" (furniture.py:2998-3058): action = move 3, quaternion (w, x, y, z) relative to the hand's current
    orientation, gripper, connect -> (dpos, rotation in the base frame, gripper action)"""
    a = np.array(action, dtype=np.float64)
    d = a[:3] * p["move_speed"]
    d_pos = np.clip([-d[1], d[0], d[2]], np.asarray(p["min_pos"]) - hand_pos_world, np.asarray(p["max_pos"]) - hand_pos_world)
    arm_quat = a[3:7][[1, 2, 3, 0]]  # T.convert_quat(..., to="xyzw")
    rotation = quat2mat(quat_multiply(hand_quat_base, arm_quat))
    return d_pos, rotation, a[7]


class IKOracle:
    """one arm: the accumulated orientation target, the position target in the base frame, the commanded joints"""

    def __init__(self, params, arm=0):
        self.p = params
        ch = self.ch = params["chains"][arm] if "chains" in params else params["chain"]
        self.base_R = mjcf.q_to_mat(ch["base_quat"])
        self.base_p = np.asarray(ch["base_pos"], dtype=np.float64)

    def to_base(self, pos_world, quat_world_wxyz):
        return self.base_R.T @ (pos_world - self.base_p), self.base_R.T @ mjcf.q_to_mat(quat_world_wxyz)

    def sync(self, hand_pos_world, hand_quat_world_wxyz):
        """_reset's tail (furniture.py:1643-1650): _initial_right_hand_quat = _right_hand_quat; controller.sync_state()"""
        pb, Rb = self.to_base(hand_pos_world, hand_quat_world_wxyz)
        self.s = mat2quat(Rb)
        self.target_pos = pb.copy()
        self.q_cmd = None

    def command(self, action, hand_pos_world, hand_quat_world_wxyz, jpos):
        """first get_control of an env step: new targets, IK from the current joints, then the P controller"""
        _, Rb = self.to_base(hand_pos_world, hand_quat_world_wxyz)
        if self.p.get("quaternion_mode"):
            d_pos, rotation, grip = ik_pre_quaternion(action, hand_pos_world, mat2quat(Rb), self.p)
        else:
            d_pos, rotation, self.s, grip = ik_pre(action, hand_pos_world, mat2quat(Rb), self.s, self.p)
        self.target_pos = self.target_pos + d_pos * self.p["user_sensitivity"]
        tp = self.base_p + self.base_R @ self.target_pos
        tq = mjcf.mat_to_q(self.base_R @ rotation)
        self.q_cmd, self.iters = solve_ik(self.p, jpos, tp, tq, self.ch)
        return self.velocities(jpos), grip

    def velocities(self, jpos):
        return np.clip(-self.p["kp"] * (np.asarray(jpos, dtype=np.float64) - self.q_cmd), -1, 1)
