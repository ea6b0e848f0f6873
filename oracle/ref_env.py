"""CPU ORACLE of the env layer (TEST INFRASTRUCTURE, NOT PRODUCT CODE): a single-env restatement of
FurnitureSawyerEnv and FurnitureBaxterEnv (control_type="impedance") on top of the C physics oracle.  Baxter: two arms
(furniture.py:89-92), 17 actions (furniture_baxter.py:52-58), per-arm finger scans and observations (:98-155), no gripper
discretisation, gravity compensation on the arm and gripper joints only (the head joint is left alone, furniture.py:3372-3377).

Restates, with reference line cites:
  reset      FurnitureEnv._reset                      furniture/env/furniture.py:1406-1663
  step       FurnitureEnv.step/_step/_step_continuous  :364-449, :1260-1330;  FurnitureSawyerEnv._step  furniture_sawyer.py:66-84
  action     _setup_action                             :3332-3379
  connect    _try_connect / _is_aligned / _connect     :926-1153, :847-924  (assembly_oracle.py)
  obs        _get_obs                                  :1344-1387, furniture_sawyer.py:103-155
  reward     _compute_reward / _after_step             :482-541, :451-480
Used as the checker for the device env kernels and as the CPU arm of bench.py (`--impl reference`, cpu_baseline).
Random draws use numpy's RandomState(seed) like the reference (furniture.py:72), in the reference's order; the placement
is pinned to the reference's own sampler code (tests/golden/placement.npz) and the engine runs the same MT19937 stream per
env, so resets are compared draw for draw (tests/test_reset_rng.py).
"""
from __future__ import annotations

import numpy as np

from . import assembly_oracle as A
from .oracle import OracleSim


class Cfg:
    control_freq = 10
    max_episode_steps = 2000
    discrete_grip = True
    rescale_actions = True
    auto_align = True
    alignment_pos_dist, alignment_rot_dist_up, alignment_rot_dist_forward, alignment_project_dist = 0.1, 0.9, 0.9, 0.3
    ctrl_penalty_coef, unstable_penalty_coef, success_reward, touch_reward, pick_reward = 1e-3, 100.0, 100.0, 10.0, 100.0
    furn_xyz_rand, furn_rot_rand, agent_xyz_rand = 0.02, 3.0, 0.001
    furn_size_rand = 0.0
    seed = 123


class OracleFurnitureEnv:
    def __init__(self, model, cfg=None):
        self.m = model
        self.cfg = cfg or Cfg()
        self.sim = OracleSim(model)
        meta = model.meta
        self.parts = list(meta["part_names"])
        self.npart = len(self.parts)
        self.part_body = [model.names["body"].index(n) for n in self.parts]
        self.part_qadr = [int(model.jnt_qposadr[model.names["jnt"].index(n)]) for n in self.parts]
        self.part_dadr = [int(model.jnt_dofadr[model.names["jnt"].index(n)]) for n in self.parts]
        self.narm, self.ngrip = len(meta["robot_joints"]), len(meta["gripper_joints"])
        self.narms = 2 if meta.get("eef_site2") else 1
        jq = lambda n: int(model.jnt_qposadr[model.names["jnt"].index(n)])  # robot joints are scalar: qpos index = dof index
        self.arm_idx = [jq(n) for n in meta["robot_joints"]]
        self.grip_idx = [jq(n) for n in meta["gripper_joints"]]
        self.nr = int(sum(1 for j in range(model.njnt) if model.jnt_type[j] != 0))
        self.dof = self.narm + self.narms + 1
        self.rng = np.random.RandomState(self.cfg.seed)
        if self.cfg.furn_size_rand != 0:  # _load_model_object draws the size factor first (furniture.py:1989-1991)
            self.resize_factor = 1 + self.rng.uniform(-self.cfg.furn_size_rand, self.cfg.furn_size_rand, 1)[0]
        g = model.names["geom"]
        self.lf = [[g.index(n) for n in meta["l_finger_geoms"]]] + ([[g.index(n) for n in meta["l_finger_geoms2"]]] if self.narms == 2 else [])
        self.rf = [[g.index(n) for n in meta["r_finger_geoms"]]] + ([[g.index(n) for n in meta["r_finger_geoms2"]]] if self.narms == 2 else [])
        self.floor = g.index("FLOOR")
        self.robot_geoms = [i for i, n in enumerate(g) if n in set(meta["robot_contact_geoms"])]
        self.part_col_geoms = [i for i, n in enumerate(g) if "collision" in n and model.names["body"][model.geom_bodyid[i]] in self.parts]
        self.conn_sites = [s for s, n in enumerate(model.names["site"]) if "conn_site" in n]
        self.eef_site = [model.names["site"].index(meta[k]) for k in ("eef_site", "eef_site2")[: self.narms]]
        self.hand_body = [model.names["body"].index(meta[k]) for k in ("hand_body", "hand_body2")[: self.narms]]
        self.nsub = int((1.0 / self.cfg.control_freq) / model.opt_timestep)
        self.group = list(range(self.npart))
        self.connected_sites = set()
        self.num_connected = 0

    # ---- helpers
    def _find(self, i):
        while self.group[i] != i:
            i = self.group[i]
        return i

    def _part_of_body(self, b):
        return self.part_body.index(b)

    def _qpos(self, p):
        return self.sim.qpos[self.part_qadr[p] : self.part_qadr[p] + 7].copy()

    def _set_qpos(self, p, pos, quat):
        self.sim.qpos[self.part_qadr[p] : self.part_qadr[p] + 3] = pos
        self.sim.qpos[self.part_qadr[p] + 3 : self.part_qadr[p] + 7] = quat

    def _stop(self, p, gravity):  # furniture.py:2778-2800
        b = self.part_body[p]
        self.sim.xfrc_applied[6 * b : 6 * b + 6] = [0, 0, -gravity * self.m.opt_gravity[2] * self.m.body_mass[b], 0, 0, 0]
        self.sim.qvel[self.part_dadr[p] : self.part_dadr[p] + 6] = 0
        self.sim.qfrc_applied[self.part_dadr[p] : self.part_dadr[p] + 6] = 0

    def _slow(self, p):  # :2821-2842
        b = self.part_body[p]
        self.sim.xfrc_applied[6 * b : 6 * b + 6] = [0, 0, -self.m.opt_gravity[2] * self.m.body_mass[b], 0, 0, 0]
        d = self.part_dadr[p]
        self.sim.qvel[d : d + 6] = np.clip(self.sim.qvel[d : d + 6], -0.2, 0.2)
        self.sim.qfrc_applied[d : d + 6] = 0

    def _fwd_step(self):
        self.sim.forward()
        self.sim.step()

    def _grav_comp(self):  # :3372-3377: arm joints and gripper joints
        idx = self.arm_idx + self.grip_idx
        self.sim.qfrc_applied[idx] = self.sim.qfrc_bias[idx]

    def _init_robot(self):  # :1761-1779
        noise = self.rng.uniform(-self.cfg.agent_xyz_rand, self.cfg.agent_xyz_rand, self.narm)
        self.sim.qpos[self.arm_idx] = self.m.meta["robot_init_qpos"] + noise
        self.sim.qpos[self.grip_idx] = self.m.meta["gripper_init_qpos"]

    def _site_pose(self, s):  # _site_xpos_xquat :1044-1055
        b = self.m.site_bodyid[s]
        bq = self.sim.xquat[4 * b : 4 * b + 4]
        return np.hstack([self.sim.site_xpos[3 * s : 3 * s + 3], A._qmul(bq, self.m.site_quat[s])])

    # ---- reset
    def place(self):
        """UniformRandomSampler.sample (placement_sampler.py:137-190): per part, in XML order, x and y uniform around the XML
        init pose until no horizontal-radius disc overlaps a part placed before, z + 0.01, then one draw for the rotation
        noise whose value is always furn_rot_rand (uniform(high=max, low=max), :127-135).  Pinned draw for draw to the
        reference's own sampler by tests/golden/placement.npz."""
        m, cfg = self.m, self.cfg
        placed, out = [], []
        for p, name in enumerate(self.parts):
            init = m.meta["part_init_qpos"][name]
            r = m.meta["part_radius"][name]
            for _ in range(10000):
                x = init[0] + self.rng.uniform(-cfg.furn_xyz_rand, cfg.furn_xyz_rand)
                y = init[1] + self.rng.uniform(-cfg.furn_xyz_rand, cfg.furn_xyz_rand)
                if all(np.linalg.norm([x - px, y - py], 2) > pr + r for px, py, pr in placed):
                    break
            rot = self.rng.uniform(high=cfg.furn_rot_rand, low=cfg.furn_rot_rand)
            quat = A.euler_to_quat([rot, 0, 0], init[3:7])
            placed.append((x, y, r))
            out.append((np.array([x, y, init[2] + 0.01]), np.asarray(quat, dtype=np.float64)))
        return out

    def reset(self):
        sim, m, cfg = self.sim, self.m, self.cfg
        if cfg.furn_size_rand != 0:  # :1428-1431 (edits the XML tree only)
            self.rng.uniform(-cfg.furn_size_rand, cfg.furn_size_rand, 1)
        sim.reset()
        saved = {g: (sim.geom_contype[g], sim.geom_conaffinity[g]) for g in self.robot_geoms}
        for g in self.robot_geoms:
            sim.geom_contype[g] = 0; sim.geom_conaffinity[g] = 0
        for g in self.part_col_geoms:
            sim.geom_contype[g] = 1; sim.geom_conaffinity[g] = 1
        self.group = list(range(self.npart))
        self.connected_sites = set()
        self.num_connected = self.prev_num_connected = 0
        self.touched = [False] * self.npart
        self.picked = [False] * self.npart
        sim.eq_active[:] = 0
        for p, (pos, quat) in enumerate(self.place()):
            self._set_qpos(p, pos, quat)
        for _ in range(10):
            for p in range(self.npart):
                self._stop(p, 0)
            for _ in range(10):
                self._fwd_step()
                for p in range(self.npart):
                    self._slow(p)
        self._grav_comp()
        self._init_robot()
        self._fwd_step()
        for g, (ct, ca) in saved.items():
            sim.geom_contype[g] = ct; sim.geom_conaffinity[g] = ca
        self._grav_comp()
        for _ in range(100):
            self._init_robot()
            self._fwd_step()
        sim.ctrl[:] = 0; sim.qfrc_applied[:] = 0; sim.xfrc_applied[:] = 0; sim.qacc[:] = 0; sim.qacc_warmstart[:] = 0
        sim.forward()
        self._grav_comp()
        for _ in range(100):
            self._fwd_step()
        self.episode_len = 0
        self.connected_body1 = None
        return self.obs()

    # ---- step
    def touch_bits(self):
        bits = [0] * self.npart
        for c in self.sim.contacts():
            for ga, gb in ((c.geom1, c.geom2), (c.geom2, c.geom1)):
                b = self.m.geom_bodyid[gb]
                if b in self.part_body:
                    p = self.part_body.index(b)
                    bits[p] |= (1 if ga in self.lf[0] else 0) | (2 if ga in self.rf[0] else 0) | (4 if ga == self.floor else 0)
                    if self.narms == 2:
                        bits[p] |= (8 if ga in self.lf[1] else 0) | (16 if ga in self.rf[1] else 0)
        return bits

    def _move_group(self, obj, translation, target_quat, gravity=0):  # :1163-1176
        base = self._qpos(obj)
        g = self._find(obj)
        for i in range(self.npart):
            if self._find(i) == g:
                npos, nq = A.transform_to_target_quat(base, self._qpos(i), np.asarray(target_quat, dtype=np.float64))
                self._set_qpos(i, npos + translation, nq)
                self._stop(i, gravity)

    def _group_min_z(self, obj):  # _get_bounding_box :749-769 (min starts at 0)
        g = self._find(obj)
        mn = 0.0
        for i in range(self.npart):
            if self._find(i) == g:
                for s in range(self.m.nsite):
                    if self.m.site_bodyid[s] == self.part_body[i]:
                        mn = min(mn, self.sim.site_xpos[3 * s + 2])
        return mn

    def _try_connect(self, part1):
        m, cfg = self.m, self.cfg
        g1 = self._find(part1)
        thr = (cfg.alignment_pos_dist, cfg.alignment_rot_dist_up, cfg.alignment_rot_dist_forward, cfg.alignment_project_dist)
        if m.neq == 0:
            return False
        for s1 in self.conn_sites:
            if self._find(self._part_of_body(m.site_bodyid[s1])) != g1:
                continue
            n1 = m.names["site"][s1]
            for s2 in self.conn_sites:
                if s1 in self.connected_sites or s2 in self.connected_sites:
                    continue
                n2 = m.names["site"][s2]
                if n1.split(",")[0].split("-") != n2.split(",")[0].split("-")[::-1]:
                    continue
                angles = [float(x) for x in n1.split(",")[1:-1] if x]
                ok, tq = A.is_aligned(self.sim.site_xpos[3 * s1 : 3 * s1 + 3], self.sim.site_xmat[9 * s1 : 9 * s1 + 9], self.sim.site_xpos[3 * s2 : 3 * s2 + 3],
                                      self.sim.site_xmat[9 * s2 : 9 * s2 + 9], angles, thr)
                if tq is not None:
                    self.target_quat = tq
                if ok:
                    self._connect(s1, s2)
                    return True
        return False

    def _connect(self, s1, s2):  # :847-924
        m, sim = self.m, self.sim
        self.connected_sites |= {s1, s2}
        body1, body2 = self._part_of_body(m.site_bodyid[s1]), self._part_of_body(m.site_bodyid[s2])
        g1, g2 = self._find(body1), self._find(body2)
        for g in range(m.ngeom):
            b = m.geom_bodyid[g]
            if b in self.part_body and self._find(self.part_body.index(b)) in (g1, g2) and sim.geom_contype[g] != 0:
                sim.geom_contype[g], sim.geom_conaffinity[g] = A.connect_masks(g1)
        if self.cfg.auto_align:  # _align_connectors / _move_site_to_target :1224-1250
            target = self._site_pose(s1)
            target[3:] = self.target_quat
            base = self._site_pose(s2)
            bq = self._qpos(body2)
            _, nq = A.transform_to_target_quat(base, bq, target[3:])
            nsp, _ = A.transform_to_target_quat(bq, base, nq)
            self._move_group(body2, target[:3] - nsp, nq, 0)
        self._fwd_step()
        mn = min(self._group_min_z(body1), self._group_min_z(body2))
        if mn < 0:
            for obj in (body1, body2):  # _move_rotate_object + _is_inside (one more step each)
                base = self._qpos(obj)
                g = self._find(obj)
                for i in range(self.npart):
                    if self._find(i) == g:
                        npos, nq = A.transform_to_target_quat(base, self._qpos(i), base[3:])
                        self._set_qpos(i, npos + np.array([0, 0, -mn]), nq)
                self._fwd_step()
        self._fwd_step()
        for e in range(m.neq):  # _activate_weld :2761-2776
            a, b = self._part_of_body(m.eq_obj1id[e]), self._part_of_body(m.eq_obj2id[e])
            if a in (body1, body2) and b in (body1, body2):
                sim.eq_data[7 * e : 7 * e + 7] = A.rel_pose(self._qpos(a), self._qpos(b))
                sim.eq_active[e] = 1
                self.group[self._find(body1)] = self._find(body2)
        self.num_connected += 1
        self.connected_body1 = body1
        self.connected_pose = self._qpos(body1)

    def obs(self):
        sim, m = self.sim, self.m
        ob = []
        for p in range(self.npart):
            b = self.part_body[p]
            ob += list(sim.xpos[3 * b : 3 * b + 3]) + list(sim.xquat[4 * b : 4 * b + 4])
        na, ng = self.narm // self.narms, self.ngrip // self.narms
        for arm in range(self.narms):  # furniture_sawyer.py:103-155, furniture_baxter.py:98-155
            ai, gi = self.arm_idx[arm * na : (arm + 1) * na], self.grip_idx[arm * ng : (arm + 1) * ng]
            ob += list(sim.qpos[ai]) + list(sim.qvel[ai]) + list(sim.qpos[gi])
            s = self.eef_site[arm]
            ob += list(sim.site_xpos[3 * s : 3 * s + 3])
            hq = sim.xquat[4 * self.hand_body[arm] : 4 * self.hand_body[arm] + 4]
            ob += [hq[1], hq[2], hq[3], hq[0]]
            v = sim.site_velocity(s)
            ob += list(v[:3]) + list(v[3:])
        return np.array(ob)

    def set_controls(self, action):
        a = np.asarray(action, dtype=np.float64).copy()
        if self.cfg.discrete_grip and self.narms == 1:  # FurnitureSawyerEnv._step only
            a[-2] = -1 if a[-2] < 0 else 1
        act = np.clip(a[:-1], -1, 1) if self.cfg.rescale_actions else a[:-1]
        full = np.zeros(self.m.nu)  # per actuator: arm joints straight through, each gripper's action as [g, -g] over its two actuators
        seen = {}
        for u in range(self.m.nu):
            jn = self.m.names["jnt"][int(self.m.actuator_jntid[u])]
            if jn in self.m.meta["robot_joints"]:
                full[u] = act[self.m.meta["robot_joints"].index(jn)]
            else:
                gidx = self.m.meta["gripper_joints"].index(jn) // (self.ngrip // self.narms)
                full[u] = act[self.narm + gidx] * (1 if seen.get(gidx, 0) == 0 else -1)
                seen[gidx] = seen.get(gidx, 0) + 1
        cr = self.m.actuator_ctrlrange
        if self.cfg.rescale_actions:
            full = 0.5 * (cr[:, 1] + cr[:, 0]) + 0.5 * (cr[:, 1] - cr[:, 0]) * full
        self.sim.ctrl[:] = full
        self._grav_comp()
        return a[-1]

    def _do_simulation(self):
        """forward + nsub mj_steps with the controls in place; True if the solver blew up (furniture.py:2877-2897)"""
        self.sim.forward()
        self.sim.step(self.nsub)
        fail = bool(self.sim.scalar("warning") & 2)
        if fail:
            self.sim.L.om_clear_warning(self.sim.d)
        return fail

    def _simulate(self, raw):
        """_step_continuous up to the connect scan -> (some _do_simulation failed, the env still has to be reset for it)"""
        self.set_controls(raw)
        fail = self._do_simulation()
        return fail, fail

    def step(self, action):
        raw = np.asarray(action, dtype=np.float64)
        connect = raw[-1]
        fail, reset_now = self._simulate(raw)
        if reset_now:
            self.reset()
        else:
            if connect > 0:  # :1290-1322: per arm the first part both fingers touch; return at the first connection
                bits = self.touch_bits()
                for arm in range(self.narms):
                    both = 3 if arm == 0 else 24
                    hit = [p for p in range(self.npart) if bits[p] & both == both]
                    if hit and self._try_connect(hit[0]):
                        break
            if self.connected_body1 is not None:
                b1 = self.connected_body1
                self.sim.forward()
                self._move_group(b1, self.connected_pose[:3] - self._qpos(b1)[:3], self.connected_pose[3:], 0)
                self.connected_body1 = None
                self._fwd_step()
        ob = self.obs()
        reward, done, success = self._compute_reward(raw, fail)
        self.episode_len += 1
        if self.episode_len == self.cfg.max_episode_steps or fail:
            done = True
            if fail:
                reward -= self.cfg.unstable_penalty_coef
        info = dict(num_connected=self.num_connected, success=int(success), unstable=int(fail), episode_length=self.episode_len)
        return ob, reward, done, info

    def _compute_reward(self, raw, fail):  # FurnitureEnv._compute_reward, furniture.py:482-541
        touch_r = pick_r = 0.0
        if not fail:
            bits = self.touch_bits()
            for arm in range(self.narms):
                both = 3 if arm == 0 else 24
                for p in range(self.npart):
                    if bits[p] & both == both:
                        if not self.touched[p]:
                            self.touched[p] = True; touch_r += self.cfg.touch_reward
                        if not (bits[p] & 4) and not self.picked[p]:
                            self.picked[p] = True; pick_r += self.cfg.pick_reward
        success_r = self.cfg.success_reward * (self.num_connected - self.prev_num_connected)
        self.prev_num_connected = self.num_connected
        reward = success_r + touch_r + pick_r - self.cfg.ctrl_penalty_coef * float(np.square(raw).sum())
        success = self.num_connected == self.npart - 1 and self.npart > 1
        return reward, success, success


class IKMixin:
    """control_type="ik" / "ik_quaternion" (FurnitureEnv._do_ik_step, furniture.py:2899-3063) over the oracle simulator, one or two arms; the
    solver is the damped-least-squares IK of oracle/ik_oracle.py (pybullet's is not available)"""

    def _ik_setup(self, **ik_kw):
        from furniture_b200 import ik as IK
        from .ik_oracle import IKOracle

        self.ikp = IK.ik_params(self.m, **ik_kw)
        na = self.ikp["narms"]
        self.iks = [IKOracle(self.ikp, a) for a in range(na)]
        self.ik = self.iks[0]
        self.hands = [self.m.names["body"].index(n) for n in ("right_hand", "left_hand")[:na]]
        self.per = 7 if self.ikp.get("quaternion_mode") else 6
        self.dof = na * self.per + na + 1  # per arm: move 3, rotate 3 (or a quaternion); a gripper per arm; connect (furniture_sawyer.py:60-63, furniture_baxter.py:52-62)

    def _hand(self, arm=0):
        b = self.hands[arm]
        return np.array(self.sim.xpos[3 * b : 3 * b + 3]), np.array(self.sim.xquat[4 * b : 4 * b + 4])

    def _ik_sync(self):  # _reset's tail, furniture.py:1643-1650
        for a, ik in enumerate(self.iks):
            ik.sync(*self._hand(a))

    def _simulate(self, raw):
        a = raw.copy()
        na, per = len(self.iks), self.per
        if self.cfg.discrete_grip and na == 1:
            a[-2] = -1 if a[-2] < 0 else 1
        jpos = lambda arm: np.array(self.sim.qpos[self.arm_idx[7 * arm : 7 * arm + 7]])
        vel, grips = [], []
        for arm, ik in enumerate(self.iks):
            arm_action = np.concatenate([a[arm * per : (arm + 1) * per], [a[na * per + arm], a[-1]]])
            v, g = ik.command(arm_action, *self._hand(arm), jpos(arm))
            vel.append(v)
            grips.append(g)
        fail = reset_now = False
        R = self.ikp["action_repeat"]
        for r in range(R):
            if r > 0:
                vel = [ik.velocities(jpos(arm)) for arm, ik in enumerate(self.iks)]
            self.low_action = np.concatenate(vel + [grips])
            self.set_controls(np.concatenate([self.low_action, [raw[-1]]]))
            if self._do_simulation():
                fail = True
                if r + 1 < R:
                    self.reset()
                else:
                    reset_now = True
        return fail, reset_now


class OracleIKEnv(IKMixin, OracleFurnitureEnv):
    def __init__(self, model, cfg=None, **ik_kw):
        super().__init__(model, cfg)
        self._ik_setup(**ik_kw)

    def reset(self):
        ob = super().reset()
        self._ik_sync()
        return ob


class DenseCfg(Cfg):  # what config/furniture_sawyer_dense.py:5-14 changes in the base env
    max_episode_steps = 150
    auto_align = False
    alignment_pos_dist, alignment_rot_dist_up, alignment_rot_dist_forward, alignment_project_dist = 0.02, 0.99, 0.99, 0.0


class SimWorld:
    """the oracle simulator seen through the names the dense reward asks for (FurnitureEnv._get_pos / _get_up_vector /
    _get_forward_vector, furniture.py:3121-3200; FurnitureSawyerEnv._finger_contact, furniture_sawyer.py:220-243)"""

    def __init__(self, env):
        self.e = env

    def pos(self, name):
        m, sim = self.e.m, self.e.sim
        if name in m.names["body"]:
            b = m.names["body"].index(name)
            return np.array(sim.xpos[3 * b : 3 * b + 3])
        s = m.names["site"].index(name)
        return np.array(sim.site_xpos[3 * s : 3 * s + 3])

    def _mat(self, name):
        s = self.e.m.names["site"].index(name)
        return np.array(self.e.sim.site_xmat[9 * s : 9 * s + 9]).reshape(3, 3)

    def up(self, name):
        return self._mat(name)[:, 2].copy()

    def forward(self, name):
        return self._mat(name)[:, 1].copy()

    def finger_contact(self, leg):
        e = self.e
        bits = e.touch_bits()[e.parts.index(leg)]
        return bool(bits & 1), bool(bits & 2)


class OracleDenseEnv(OracleFurnitureEnv):
    """FurnitureSawyerDenseRewardEnv (furniture_sawyer_dense.py): the Sawyer env with the reward machine of oracle/dense_oracle.py"""

    def __init__(self, model, cfg=None, dense_cfg=None):
        import json

        from .dense_oracle import DenseOracle

        super().__init__(model, cfg or DenseCfg())
        c = self.cfg
        dc = dict(dense_cfg or {})
        dc.update(ctrl_penalty_coef=c.ctrl_penalty_coef, alignment_pos_dist=c.alignment_pos_dist, alignment_rot_dist_up=c.alignment_rot_dist_up,
                  alignment_rot_dist_forward=c.alignment_rot_dist_forward, alignment_project_dist=c.alignment_project_dist)
        self.dense = DenseOracle(SimWorld(self), json.loads(model.meta["recipe_json"]), dc, success_num_conn=self.npart - 1)

    def obs(self):
        ob = super().obs()
        if self.dense.c["phase_ob"] and hasattr(self.dense, "phase"):
            ob = np.concatenate([ob, np.eye(8)[self.dense.phase]])
        return ob

    def reset(self):
        super().reset()
        self.dense.begin_episode()
        return self.obs()

    def step(self, action):
        ob, reward, done, info = super().step(action)
        return self.obs(), reward, done, info  # the phase one-hot follows the phase after the reward (furniture_sawyer_dense.py:119-126)

    def _compute_reward(self, raw, fail):
        connected = self.num_connected != self.prev_num_connected  # _connected: _connect ran during this step
        self.prev_num_connected = self.num_connected
        reward, done, self.dense_info = self.dense.step(np.asarray(raw, dtype=np.float64), connected)
        base_done = self.num_connected == self.npart - 1 and self.npart > 1  # FurnitureEnv._step, furniture.py:438-445
        return reward, bool(done or base_done), self.dense.success


class OracleDenseIKEnv(IKMixin, OracleDenseEnv):
    """the dense-reward env driven through control_type="ik" (the combination the reference's training scripts use: env id
    IKEASawyerDense-v0 with the default control type)"""

    def __init__(self, model, cfg=None, dense_cfg=None, **ik_kw):
        OracleDenseEnv.__init__(self, model, cfg, dense_cfg)
        self._ik_setup(**ik_kw)

    def reset(self):
        ob = OracleDenseEnv.reset(self)
        self._ik_sync()
        return ob


class OracleControllerEnv(OracleFurnitureEnv):
    """the one-arm env under one of the NEW_CONTROLLERS (FurnitureEnv._do_controller_step, furniture.py:3065-3093, with _pre_action
    :1706-1759 before every mj_step) on the torque-actuated robot; the controller is oracle/controller_oracle.py"""

    def __init__(self, model, name, cfg=None, move_speed=0.1):
        from .controller_oracle import ArmController

        super().__init__(model, cfg)
        self.ctl = ArmController(name, timestep=float(model.opt_timestep))
        self.move_speed = move_speed
        self.hand = model.names["body"].index("right_hand")
        self.arm_jnt = [model.names["jnt"].index(n) for n in model.meta["robot_joints"]]
        self.dof = self.ctl.control_dim + 2

    def reset(self):
        ob = super().reset()
        self.ctl.reset()  # furniture.py:1885-1887
        return ob

    def readings(self):
        sim, b = self.sim, self.hand
        pos = np.array(sim.xpos[3 * b : 3 * b + 3])
        R = np.array(sim.xmat[9 * b : 9 * b + 9]).reshape(3, 3)
        Jx, Jr = np.zeros((3, 7)), np.zeros((3, 7))
        for k, j in enumerate(self.arm_jnt):
            axis, anchor = np.array(sim.xaxis[3 * j : 3 * j + 3]), np.array(sim.xanchor[3 * j : 3 * j + 3])
            Jx[:, k], Jr[:, k] = np.cross(axis, pos - anchor), axis
        q, qv = np.array(sim.qpos[self.arm_idx]), np.array(sim.qvel[self.arm_idx])
        M = np.array(sim.qM).reshape(self.m.nv, self.m.nv)[np.ix_(self.arm_idx, self.arm_idx)]
        # body_xvelp / body_xvelr come from the velocities of the last forward pass (cvel), i.e. from the joint velocities before the
        # last integration, like the Jacobian; qpos / qvel themselves are the integrated ones
        return pos, R, Jx @ self._qvel_fwd, Jr @ self._qvel_fwd, q, qv, Jx, Jr, M

    def _simulate(self, raw):
        a = raw.copy()
        if self.cfg.discrete_grip:
            a[-2] = -1 if a[-2] < 0 else 1
        n = self.ctl.control_dim
        arm = np.zeros(7)
        arm[:n] = a[:n]
        arm[:3] = arm[:3] * self.move_speed
        arm[:3] = [-arm[1], arm[0], arm[2]]
        m, sim = self.m, self.sim
        cr = m.actuator_ctrlrange
        sim.forward()
        self.torques = []
        for i in range(self.nsub):
            if i == 0:
                self._qvel_fwd = np.array(sim.qvel[self.arm_idx])
            tau = self.ctl.torques(arm, i == 0, *self.readings())
            self.torques.append(tau)
            seen = 0
            for u in range(m.nu):
                jn = m.names["jnt"][int(m.actuator_jntid[u])]
                if jn in m.meta["robot_joints"]:
                    k = m.meta["robot_joints"].index(jn)
                    sim.ctrl[u] = sim.qfrc_bias[self.arm_idx[k]] + tau[k]
                else:
                    g = a[-2] * (1 if seen == 0 else -1)
                    seen += 1
                    sim.ctrl[u] = 0.5 * (cr[u, 1] + cr[u, 0]) + 0.5 * (cr[u, 1] - cr[u, 0]) * g
            self._qvel_fwd = np.array(sim.qvel[self.arm_idx])
            sim.step(1)
        fail = bool(sim.scalar("warning") & 2)
        if fail:
            sim.L.om_clear_warning(sim.d)
        return fail, fail
