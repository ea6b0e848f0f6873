"""CPU ORACLE for the five torque controllers (NEW_CONTROLLERS) (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

numpy restatement of furniture/env/controllers/arm_controller.py with the parameters of controllers/controller_config.hjson:

  ArmController.torques  <- Controller.transform_action :98-105, linear_interpolate :157-164, calculate_orientation_error :179-201,
                            JointTorqueController.action_to_torques :275-303, JointVelocityController :345-366,
                            JointImpedanceController :432-496, PositionOrientationController :639-739 (+ update_model_opspace :752-799,
                            set_goal_position / set_goal_orientation :801-812), PositionController :925-930
Pinned against tests/golden/controllers.npz (the reference's own classes on a stand-in simulator, tools/make_golden_controllers.py).

One class, five modes, because the reference's five classes share one skeleton: scale the action, at a policy step set a goal and a
linear ramp towards it, at every mj_step advance the ramp and turn the error into torques.  Two reference quirks are kept: the ramp is
`floor(0.2 * control_freq / timestep)` = 2000 mj_steps long although an env step has 50 (control_freq is multiplied, not divided); the
position controller fixes its orientation goal at the first policy step of the controller's life, not of the episode.
"""
import numpy as np

CONFIG = {  # controllers/controller_config.hjson (defaults; impedance_flag False, interpolation "linear" everywhere)
    "position_orientation": dict(control_range_pos=0.05, control_range_ori=0.2, kp=150.0, damping=1.0),
    "position": dict(control_range_pos=0.05, kp=150.0, damping=1.0),
    "joint_impedance": dict(control_range=[0.2] * 7, kp_max=[100, 100, 100, 100, 50, 30, 10], kp_min=[10, 10, 10, 10, 10, 1, 1], damping_max=[2] * 7, damping_min=[0] * 7),
    "joint_velocity": dict(control_range=[1] * 7, kv=[8.0, 7.0, 6.0, 4.0, 2.0, 0.5, 0.1]),
    "joint_torque": dict(control_range=[0.5, 0.5, 0.5, 0.2, 0.2, 0.1, 0.1]),
}
RAMP_RATIO, CONTROL_FREQ = 0.20, 20


def euler2mat(euler):  # transform_utils.py:360-380
    ai, aj, ak = -euler[2], -euler[1], -euler[0]
    si, sj, sk = np.sin(ai), np.sin(aj), np.sin(ak)
    ci, cj, ck = np.cos(ai), np.cos(aj), np.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([[cj * ci, cj * si, -sj], [sj * cs - sc, sj * ss + cc, cj * sk], [sj * cc + ss, sj * sc - cs, cj * ck]])


def orientation_error(desired, current):
    return 0.5 * (np.cross(current[:, 0], desired[:, 0]) + np.cross(current[:, 1], desired[:, 1]) + np.cross(current[:, 2], desired[:, 2]))


def pinv_sym(A, threshold=0.00025):
    u, s, v = np.linalg.svd(A)
    return v.T.dot(np.diag([0 if x < threshold else 1.0 / x for x in s])).dot(u.T)


class ArmController:
    def __init__(self, mode, timestep=0.002):
        c = CONFIG[mode]
        self.mode = mode
        if mode == "position_orientation":
            self.cmax = np.concatenate([np.ones(3) * c["control_range_pos"], np.ones(3) * c["control_range_ori"]])
        elif mode == "position":
            self.cmax = np.ones(3) * c["control_range_pos"]
        else:
            self.cmax = np.array(c["control_range"], dtype=np.float64)
        self.control_dim = len(self.cmax)
        self.steps = np.floor(RAMP_RATIO * CONTROL_FREQ / timestep)
        if mode in ("position", "position_orientation"):
            self.kp = np.concatenate([np.ones(3) * c["kp"], np.ones(3) * c["kp"]])
            self.damping = np.ones(6) * c["damping"]
        elif mode == "joint_impedance":
            self.kp = (np.array(c["kp_max"]) + np.array(c["kp_min"])) * 0.5
            self.damping = (np.array(c["damping_max"]) + np.array(c["damping_min"])) * 0.5
        elif mode == "joint_velocity":
            self.kv = np.array(c["kv"])
        self.ori_goal = None  # position mode: set once
        self.reset()

    def reset(self):
        self.step = 0
        if self.mode in ("position", "position_orientation"):
            self.last_pos, self.last_ori = np.zeros(3), np.eye(3)
        else:
            self.last = np.zeros(self.control_dim)

    def _scale(self, a):
        a = np.clip(np.asarray(a, dtype=np.float64)[: self.control_dim], -1, 1)
        return (a - 0.0) * (abs(self.cmax - (-self.cmax)) / abs(1 - (-1))) + (self.cmax + (-self.cmax)) / 2.0

    def torques(self, action, policy_step, pos, R, velp, velr, q, qvel, Jx, Jr, M):
        a = self._scale(action)
        n = int(self.steps)
        if self.mode in ("joint_torque", "joint_velocity", "joint_impedance"):
            if policy_step:
                self.step = 0
                if self.mode == "joint_impedance":
                    goal = q + a
                    if np.linalg.norm(self.last) == 0:
                        self.last = q
                else:
                    goal = np.array(a)
                self.start, self.delta = self.last, (goal - self.last) / self.steps
            self.last = self.start + (self.step + 1) * self.delta
            if self.step < self.steps - 1:
                self.step += 1
            if self.mode == "joint_torque":
                return np.array(self.last)
            if self.mode == "joint_velocity":
                return np.multiply(self.kv, (self.last - qvel))
            kv = 2 * np.sqrt(self.kp) * self.damping
            v = np.array(qvel, dtype=np.float64)
            norm = np.linalg.norm(v)
            if norm > 7.0:
                v = v / (norm * 7.0)
            return np.dot(M, np.multiply(self.kp, self.last - q) - np.multiply(kv, v))
        # operational space
        if policy_step:
            self.step = 0
            goal_pos = pos + a[0:3]
            if self.mode == "position_orientation":
                self.ori_goal = np.dot(euler2mat(-a[3:6]).T, R)
            elif self.ori_goal is None:
                self.ori_goal = np.array(R)
            if np.linalg.norm(self.last_pos) == 0:
                self.last_pos = pos
            if (self.last_ori == np.eye(3)).all():
                self.last_ori = R
            self.pos_start, self.pos_delta = self.last_pos, (goal_pos - self.last_pos) / self.steps
            self.ori_delta = orientation_error(self.ori_goal, self.last_ori) / self.steps
            self.ori_start = self.last_ori
        self.last_pos = self.pos_start + (self.step + 1) * self.pos_delta
        self.last_ori = np.dot(euler2mat(-((self.step + 1) * self.ori_delta)).T, self.ori_start)
        if self.step < self.steps - 1:
            self.step += 1
        kv = 2 * np.sqrt(self.kp) * self.damping
        force = np.multiply(self.last_pos - pos, self.kp[0:3]) - np.multiply(velp, kv[0:3])
        torque = np.multiply(orientation_error(self.last_ori, R), self.kp[3:6]) - np.multiply(velr, kv[3:6])
        Minv = np.linalg.inv(M)
        lam_x = pinv_sym(np.dot(np.dot(Jx, Minv), Jx.T))
        lam_r = pinv_sym(np.dot(np.dot(Jr, Minv), Jr.T))
        wrench = np.concatenate([np.dot(lam_x, force), np.dot(lam_r, torque)])
        return np.dot(np.vstack([Jx, Jr]).T, wrench)
