"""CPU ORACLE for the phase-based dense reward (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

numpy (float64) restatement of FurnitureSawyerDenseRewardEnv's reward machine, furniture/env/furniture_sawyer_dense.py:

  DenseOracle.begin_episode  <- _reset_reward_variables :128-139
  DenseOracle.begin_subtask  <- _update_reward_variables :150-219 (+ _set_next_subtask :141-148)
  DenseOracle.step           <- _collect_values :225-280 and _compute_reward :282-586
  phase terms                <- _init_eef_reward :588-610 ... _move_leg_fine_reward :853-943, _stable_grip_reward :945-985,
                                _gripper_penalty :987-1003, _ctrl_penalty :1005-1009, _move_other_part_penalty :1011-1022
  forward_rotated            <- FurnitureEnv._project_connector_forward  furniture/env/furniture.py:1178-1199

Pinned against tests/golden/dense_reward.npz, which holds what the reference's own Python answered on a scripted world
(tools/make_golden_dense.py): phases, subtasks, done / success flags bit-exactly, rewards to 1e-12 relative.

The machine is written as a table of phases instead of the reference's if-chain; the arithmetic (operation order, numpy calls) is
the reference's.  One conscious deviation: `prev_grasp` exists from the start in both reward modes -- the reference creates
`_prev_grasp_dist` only when diff_rew is on and raises AttributeError in _grasp_leg_reward otherwise.
"""
import numpy as np

from . import assembly_oracle as A

INIT_EEF, ABOVE_LEG, LOWER_EEF, GRASP, LIFT, ALIGN, MOVE, FINE = range(8)
INFO_KEYS = ["phase_bonus", "ctrl_penalty", "gripper_penalty", "move_other_part_penalty", "drop_penalty", "touch", "drop_leg", "table_moved",
             "stable_grip_succ", "skip_to_lift_leg", "skip_to_move_leg_fine"]

# defaults of furniture/config/furniture_sawyer_dense.py:5-71 (and ctrl_penalty_coef of config/furniture.py:291)
DEFAULTS = dict(
    diff_rew=True, phase_bonus=5000.0, eef_forward_dist_coef=2.0, eef_up_dist_coef=4.0, eef_rot_threshold=0.95, gripper_penalty_coef=1.0,
    move_other_part_penalty_coef=50.0, drop_penalty_coef=20.0, early_termination=False, init_eef_pos_dist_coef=100.0, move_eef_pos_dist_coef=100.0,
    lower_eef_pos_dist_coef=1000.0, grasp_dist_coef=200.0, lift_z_dist_coef=500.0, lift_xy_dist_coef=250.0, lift_z_pos_threshold=0.02,
    lift_xy_pos_threshold=0.05, align_pos_dist_coef=100.0, align_rot_dist_coef=50.0, align_pos_threshold=0.2, align_rot_threshold=0.85,
    move_pos_dist_coef=300.0, move_rot_dist_coef=50.0, move_pos_threshold=0.06, move_rot_threshold=0.85, move_fine_pos_exp_coef=-25.0,
    move_fine_pos_dist_coef=500.0, move_fine_rot_dist_coef=200.0, aligned_bonus_coef=10.0, ctrl_penalty_coef=1e-3, phase_ob=False,
    reset_robot_after_attach=False, alignment_pos_dist=0.02, alignment_rot_dist_up=0.99, alignment_rot_dist_forward=0.99, alignment_project_dist=0.0)


class Subtask:
    """one row of the recipe: who moves (leg), onto whom (table), through which connector sites"""

    def __init__(self, leg, table, leg_site, table_site, angle, waypoint_z, grip_init):
        self.leg, self.table, self.leg_site, self.table_site = leg, table, leg_site, table_site
        self.angle = angle                                                  # site_recipe[i][2] or None
        self.allowed = [float(x) for x in leg_site.split(",")[1:-1] if x]  # the angles in the connector's name
        self.waypoint_z = waypoint_z                                        # waypoints[i][0][2]
        self.grip_init = grip_init                                          # grip_init_pos[i][0] (3 or 4 numbers) or None


def subtasks_of(recipe):
    gi = recipe.get("grip_init_pos")
    out = []
    for i, (leg, table) in enumerate(recipe["recipe"]):
        sr = recipe["site_recipe"][i]
        g = gi[i][0] if (gi is not None and gi[i] is not None) else None
        out.append(Subtask(leg, table, sr[0], sr[1], sr[2] if len(sr) == 3 else None, recipe["waypoints"][i][0][2], g))
    return out


class DenseOracle:
    """`world` answers pos(name), up(name), forward(name) (site rotation columns 2 and 1), finger_contact(leg) -> (left, right)."""

    def __init__(self, world, recipe, cfg=None, success_num_conn=None, preassembled=()):
        self.w = world
        self.c = dict(DEFAULTS)
        self.c.update(cfg or {})
        self.sub = subtasks_of(recipe)
        self.z_finedist = recipe["z_finedist"]
        self.n_goal = len(self.sub) if success_num_conn is None else success_num_conn
        self.preassembled = list(preassembled)
        self.prev = {"grasp": -1}
        self.success = False

    # ------------------------------------------------------------------ episode / subtask set-up
    def _claim_grasp_sites(self, leg):
        for i in range(len(self.sub)):
            names = ("%s_ltgt_site%d" % (leg, i), "%s_rtgt_site%d" % (leg, i))
            if names[0] not in self.used and names[1] not in self.used:
                self.used.update(names)
                return names
        return names  # every pair taken: the reference keeps the last pair it looked at

    def begin_episode(self):
        self.subtask = len(self.preassembled)
        self.used = set()
        for k in range(len(self.preassembled)):
            self._claim_grasp_sites(self.sub[k].leg)
        self.begin_subtask()

    def begin_subtask(self):
        c, w, st = self.c, self.w, self.sub[self.subtask]
        self.dropped = self.table_moved = self.lifted = False
        self.table_site0 = w.pos(st.table_site)
        self.leg0 = w.pos(st.leg)
        self.lift_target = self.leg0 + [0, 0, st.waypoint_z]
        self.fine_aligned = 0
        eef = w.pos("griptip_site")
        self.phase = ABOVE_LEG if c["reset_robot_after_attach"] else INIT_EEF
        if st.grip_init is not None:
            self.init_eef = eef.copy() + st.grip_init[:3]
            if len(st.grip_init) == 4:
                self.init_eef[2] = st.grip_init[3] - 0.085
        else:
            self.phase = ABOVE_LEG
        self.gl, self.gr = self._claim_grasp_sites(st.leg)
        if c["diff_rew"]:
            if self.phase == ABOVE_LEG:
                self.prev["above_leg"] = np.linalg.norm(eef - (self.grasp_pos() + [0, 0, 0.05]))
            else:
                self.prev["init_eef"] = np.linalg.norm(eef - self.init_eef)
            self.prev["grasp"] = -1
            self.prev["lift_z"] = st.waypoint_z
            self.prev["lift_xy"] = 0.0

    def grasp_pos(self):
        return (self.w.pos(self.gl) + self.w.pos(self.gr)) / 2

    def _next_subtask(self):
        self.subtask += 1
        if self.subtask == self.n_goal:
            return True
        self.begin_subtask()
        return False

    def forward_rotated(self, st):
        w = self.w
        up1, f1, f2 = w.up(st.leg_site), w.forward(st.leg_site), w.forward(st.table_site)
        if st.angle is not None:
            return A.rotate_vector(f1, up1, st.angle)
        cs = A.cos_siml(f1, f2)
        plus, minus = A.rotate_vector_cos_siml(f1, up1, cs, 1), A.rotate_vector_cos_siml(f1, up1, cs, -1)
        return plus if A.cos_siml(plus, f2) > A.cos_siml(minus, f2) else minus

    def is_aligned(self, st):
        w, c = self.w, self.c
        m = lambda s: np.stack([np.zeros(3), w.forward(s), w.up(s)], axis=1)
        ok, _ = A.is_aligned(w.pos(st.leg_site), m(st.leg_site), w.pos(st.table_site), m(st.table_site), st.allowed,
                             (c["alignment_pos_dist"], c["alignment_rot_dist_up"], c["alignment_rot_dist_forward"], c["alignment_project_dist"]))
        return ok

    # ------------------------------------------------------------------ the shaped terms
    def _shaped(self, key, value, coef, shape, sign=1.0, mult=10.0, plain=None):
        """diff_rew: coef * mult * sign * (shape(value) - shape(previous)), then remember value; else the plain term."""
        if self.c["diff_rew"]:
            cur, old = shape(value), shape(self.prev[key])
            r = ((cur - old) if sign > 0 else (old - cur)) * coef * mult
            self.prev[key] = value
            return r
        return plain

    def _stable_grip(self):
        c, w = self.c, self.w
        up_d = A.cos_siml(w.up("grip_site"), [0, 0, -1])
        gv = w.pos(self.gr) - w.pos(self.gl)
        fw = w.forward("grip_site")
        fw_d = max(A.cos_siml(fw, gv), A.cos_siml(-fw, gv))
        rew, ok = 0, True
        if self.phase <= LIFT:
            rew += c["eef_up_dist_coef"] * (up_d - 1)
            ok = ok and up_d > c["eef_rot_threshold"]
        if ABOVE_LEG <= self.phase <= LIFT:
            rew += (np.abs(fw_d) - 1) * c["eef_forward_dist_coef"]
            ok = ok and fw_d > c["eef_rot_threshold"]
        return rew, bool(ok)

    def _lower_term(self, v):
        target = v["grasp"] + [0, 0, -0.015]
        d = np.linalg.norm(v["eef"] - target)
        r = self._shaped("eef_leg", d, self.c["lower_eef_pos_dist_coef"], lambda x: min(x, 0.2), sign=-1, plain=-d * self.c["lower_eef_pos_dist_coef"])
        return r, bool(np.linalg.norm(v["eef"][:2] - target[:2]) < 0.02 and np.abs(v["eef"][2] - target[2]) < 0.015)

    # ------------------------------------------------------------------ one env step
    def step(self, ac, connected):
        c, w = self.c, self.w
        st = self.sub[self.subtask]
        P = c["phase_bonus"]
        info = dict.fromkeys(INFO_KEYS, 0)
        self.success = False
        done = False
        bonus = 0
        # ---- what the world looks like
        left, right = w.finger_contact(st.leg)
        touched = int(left and right)
        v = dict(eef=w.pos("griptip_site"), grasp=self.grasp_pos(), leg=w.pos(st.leg), leg_site=w.pos(st.leg_site), table_site=w.pos(st.table_site))
        leg_up, table_up, table_fw = w.up(st.leg_site), w.up(st.table_site), w.forward(st.table_site)
        fw_rot = self.forward_rotated(st) if len(st.allowed) else w.forward(st.leg_site)
        above = v["table_site"] + [0, 0, self.z_finedist]
        safe_grasp = touched and (v["eef"][2] < v["grasp"][2] - 0.000)
        d_site = np.linalg.norm(v["table_site"] - v["leg_site"])
        d_above = np.linalg.norm(above - v["leg_site"])
        up_sim = A.cos_siml(leg_up, table_up)
        fw_sim = A.cos_siml(fw_rot, table_fw)
        proj_t = A.cos_siml(-table_up, v["leg_site"] - v["table_site"])
        proj_l = A.cos_siml(leg_up, v["table_site"] - v["leg_site"])
        disp = np.linalg.norm(v["table_site"] - self.table_site0)
        ctrl = np.linalg.norm(ac[:-2]) * -c["ctrl_penalty_coef"]
        _, grip_ok = self._stable_grip()
        other = -c["move_other_part_penalty_coef"] * disp
        moved = disp > 0.1
        # ---- shortcuts a policy may take (only without the phase observation)
        if not c["phase_ob"]:
            if safe_grasp and grip_ok and self.phase < GRASP:
                info["skip_to_lift_leg"] = 1
                self.phase = LIFT
            if touched and self.phase in (LIFT, ALIGN):
                if (d_site < c["move_pos_threshold"] or d_above < c["move_pos_threshold"]) and up_sim > c["move_rot_threshold"] and fw_sim > c["move_rot_threshold"]:
                    info["skip_to_move_leg_fine"] = 1
                    self.phase = FINE
                    self.prev.update(move_pos=d_site, move_up=up_sim, move_fwd=fw_sim, proj_t=proj_t, proj_l=proj_l)
        info["touch"] = touched
        info["drop_leg"] = int(self.phase > GRASP and not touched and not self.dropped and not connected)
        info["table_moved"] = int(moved and not self.table_moved)
        grip_rew, grip_ok = self._stable_grip()
        open_phase = self.phase <= LOWER_EEF
        hand_ok = ac[-2] < 0 if open_phase else ac[-2] > 0
        hand = (-ac[-2] if open_phase else ac[-2]) * c["gripper_penalty_coef"]

        def mishap(kind, cost):  # a dropped leg / a pushed table: flag it once, end the episode if early_termination
            nonlocal done, bonus
            if kind == "drop":
                self.dropped = True
            else:
                self.table_moved = True
            done = c["early_termination"]
            if c["early_termination"]:
                bonus = bonus - cost

        def attached():  # bonus of a correct connection, then on to the next subtask
            nonlocal done, bonus
            bonus += P * 2
            bonus -= self.fine_aligned * c["aligned_bonus_coef"]
            self.phase = INIT_EEF
            done = self.success = self._next_subtask()

        phase = self.phase
        term = 0
        if phase != FINE and connected:
            if moved:
                mishap("table", P)
            elif self.is_aligned(st):
                attached()
            else:
                self.success = False
                done = True
        elif phase == INIT_EEF:
            d = np.linalg.norm(v["eef"] - self.init_eef)
            term = self._shaped("init_eef", d, c["init_eef_pos_dist_coef"], lambda x: np.exp(-10 * min(x, 0.5)), plain=-d * c["init_eef_pos_dist_coef"])
            if d < 0.03 and grip_ok and hand_ok:
                self.phase += 1
                bonus += P
                self.prev["above_leg"] = np.linalg.norm(v["eef"] - (v["grasp"] + [0, 0, 0.05]))
        elif phase == ABOVE_LEG:
            d = np.linalg.norm(v["eef"] - (v["grasp"] + [0, 0, 0.05]))
            term = self._shaped("above_leg", d, c["move_eef_pos_dist_coef"], lambda x: min(x, 1.0), sign=-1, plain=-d * c["move_eef_pos_dist_coef"])
            if d < 0.03 and grip_ok and hand_ok:
                self.phase += 1
                bonus += P
                self.prev["eef_leg"] = np.linalg.norm(v["eef"] - (v["grasp"] + [0, 0, -0.015]))
        elif phase == LOWER_EEF:
            term, ok = self._lower_term(v)
            if ok and grip_ok and hand_ok:
                bonus += P
                self.phase += 1
        elif phase == GRASP:
            term, _ = self._lower_term(v)
            term += (ac[-2] - self.prev["grasp"]) * c["grasp_dist_coef"]
            self.prev["grasp"] = ac[-2]
            if touched and safe_grasp and grip_ok:
                self.phase += 1
                bonus += P
        elif phase == LIFT:
            xy = np.linalg.norm(self.lift_target[:2] - v["leg"][:2])
            z = np.abs(self.lift_target[2] - v["leg"][2])
            rz = self._shaped("lift_z", z, c["lift_z_dist_coef"], lambda x: min(x, 0.5), sign=-1, plain=-z * c["lift_z_dist_coef"])
            rxy = self._shaped("lift_xy", xy, c["lift_xy_dist_coef"], lambda x: min(x, 0.8), sign=-1, plain=-xy * c["lift_xy_dist_coef"])
            term = rxy + rz
            if touched and v["leg"][2] > (self.leg0[2] + 0.01) and safe_grasp and not self.lifted:
                self.lifted = True
                term += P / 2
            if not touched:
                term = min(term, 0)
            if not touched:
                mishap("drop", P / 2)
            elif moved:
                mishap("table", P / 2)
            elif xy < c["lift_xy_pos_threshold"] and z < c["lift_z_pos_threshold"]:
                self.phase += 1
                bonus += P
                self.prev.update(move_pos=0, move_up=up_sim, move_fwd=fw_sim)
        elif phase in (ALIGN, MOVE):
            if phase == ALIGN:
                d = np.linalg.norm(self.lift_target - w.pos(st.leg))
                rp = self._shaped("move_pos", d, c["align_pos_dist_coef"], lambda x: min(x, 0.4), sign=-1, plain=-d * c["align_pos_dist_coef"])
                ident = lambda x: x
                ru = self._shaped("move_up", up_sim, c["align_rot_dist_coef"], ident, plain=(up_sim - 1) * c["align_rot_dist_coef"])
                rf = self._shaped("move_fwd", fw_sim, c["align_rot_dist_coef"], ident, plain=(fw_sim - 1) * c["align_rot_dist_coef"])
                ok = d < c["align_pos_threshold"] and up_sim > c["align_rot_threshold"] and fw_sim > c["align_rot_threshold"] and touched
            else:
                rp = self._shaped("move_pos", d_above, c["move_pos_dist_coef"], lambda x: min(x, 0.5), sign=-1, plain=-d_site * c["move_pos_dist_coef"])
                pos = lambda x: max(x, 0)
                ru = self._shaped("move_up", up_sim, c["move_rot_dist_coef"], pos, plain=(up_sim - 1) * c["move_rot_dist_coef"])
                rf = self._shaped("move_fwd", fw_sim, c["move_rot_dist_coef"], pos, plain=(fw_sim - 1) * c["move_rot_dist_coef"])
                ok = (d_above < c["move_pos_threshold"] or d_site < c["move_pos_threshold"]) and up_sim > c["move_rot_threshold"] and \
                    fw_sim > c["move_rot_threshold"] and touched
            if not touched:
                rp, ru, rf = min(rp, 0), min(ru, 0), min(rf, 0)
            term = rp + ru + rf
            if not touched:
                mishap("drop", P / 2)
            elif moved:
                mishap("table", P / 2)
            elif ok:
                self.phase += 1
                bonus += P * 2
                if phase == ALIGN:
                    self.prev["move_pos"] = d_above
                else:
                    self.prev.update(move_pos=d_site, proj_t=proj_t, proj_l=proj_l)
        else:  # FINE
            k, thr = c["move_fine_rot_dist_coef"], c["move_rot_threshold"]
            rp = self._shaped("move_pos", d_site, c["move_fine_pos_dist_coef"], lambda x: np.exp(c["move_fine_pos_exp_coef"] * x),
                              plain=-d_site * c["move_fine_pos_dist_coef"])
            ang = lambda x: np.exp(-2 * (1 - max(x, thr - 0.1)))
            ru = self._shaped("move_up", up_sim, k, ang, plain=(up_sim - 1) * k)
            rf = self._shaped("move_fwd", fw_sim, k, ang, plain=(fw_sim - 1) * k)
            prj = lambda x: np.exp(-3 * (1 - max(abs(x), 0.5)))
            rt = self._shaped("proj_t", proj_t, k, prj, mult=5.0, plain=(proj_t - 1) * k / 10)
            rl = self._shaped("proj_l", proj_l, k, prj, mult=5.0, plain=(proj_l - 1) * k / 10)
            aligned = self.is_aligned(st)
            good = connected and aligned
            if not touched:
                rp, ru, rf, rt, rl = min(rp, 0), min(ru, 0), min(rf, 0), min(rt, 0), min(rl, 0)
            term = rp + ru + rf + rt + rl
            if aligned:
                self.fine_aligned += 1
                term += (ac[-1] + 1) * c["aligned_bonus_coef"]
            if connected:
                term = 0
            if moved:
                mishap("table", P)
            elif good:
                attached()
            elif connected:
                done = True
                self.success = False
            if not touched and not good:
                mishap("drop", P)
        reward = 0
        reward += ctrl + term + grip_rew
        reward += hand + bonus + other
        if self.dropped and not c["early_termination"]:
            reward -= c["drop_penalty_coef"]
            info["drop_penalty"] = -c["drop_penalty_coef"]
        info.update(phase_bonus=bonus, ctrl_penalty=ctrl, gripper_penalty=hand, move_other_part_penalty=other, stable_grip_succ=int(grip_ok))
        return reward, bool(done), info
