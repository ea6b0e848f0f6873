"""Host-side lowering of a compiled MJCF ``Model`` into the engine's ``fe_model`` table (csrc/fe_model.h).

* bodies without joints are fused into the link they are welded to (mass / CoM / inertia combined; geoms and
  sites re-expressed in the link frame) -- the reference's Sawyer+table_lack scene has 36 bodies but only 14 links;
* only geoms that can ever collide are kept; the static filters of mj_collision are already resolved in
  ``Model.collision_pairs`` and are re-indexed here with type(g1) <= type(g2);
* per-geom / per-weld ``invweight0`` values are carried over so the constraint regularisation R matches MuJoCo's
  diagApprox.

The ctypes ``FeModel`` mirrors ``struct fe_model`` field by field; ``Engine`` checks sizeof/offsets against the
loaded library before use.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import mjcf

MAXLINK, MAXRDOF, MAXPART, MAXDOF, MAXGEOM, MAXPAIR, MAXSITE, MAXEQ, MAXU, MAXMESHVERT = 32, 20, 16, 116, 96, 2048, 256, 40, 20, 512
MAGIC = 0x46453035
TAG_FLOOR, TAG_LFINGER, TAG_RFINGER, TAG_ROBOT, TAG_LFINGER2, TAG_RFINGER2, TAG_PART_SHIFT = 1, 2, 4, 8, 16, 32, 8

i32, f32 = C.c_int32, C.c_float


class FeModel(C.Structure):
    _fields_ = [
        ("magic", i32), ("struct_bytes", i32),
        ("nq", i32), ("nv", i32), ("nu", i32), ("nlink", i32), ("nrlink", i32), ("nr", i32), ("npart", i32), ("ngeom", i32),
        ("npair", i32), ("nsite", i32), ("neq", i32), ("maxdepth", i32), ("has_margin", i32), ("has_gap", i32), ("nmov", i32),
        ("timestep", f32), ("gravity", f32 * 3), ("impratio", f32), ("meaninertia", f32), ("robot_ref", f32 * 3),
        ("link_parent", i32 * MAXLINK), ("link_jtype", i32 * MAXLINK), ("link_qadr", i32 * MAXLINK), ("link_dadr", i32 * MAXLINK),
        ("link_depth", i32 * MAXLINK), ("link_ancmask", i32 * MAXLINK),
        ("link_pos", (f32 * 3) * MAXLINK), ("link_quat", (f32 * 4) * MAXLINK), ("link_jaxis", (f32 * 3) * MAXLINK), ("link_jpos", (f32 * 3) * MAXLINK),
        ("link_mass", f32 * MAXLINK), ("link_com", (f32 * 3) * MAXLINK), ("link_inertia_c", (f32 * 6) * MAXLINK), ("link_inertia_o", (f32 * 6) * MAXLINK),
        ("dof_damping", f32 * MAXDOF), ("rdof_armature", f32 * MAXRDOF),
        ("rdof_limited", i32 * MAXRDOF), ("rdof_range", (f32 * 2) * MAXRDOF), ("rdof_invweight", f32 * MAXRDOF),
        ("rdof_solref", (f32 * 2) * MAXRDOF), ("rdof_solimp", (f32 * 3) * MAXRDOF),
        ("act_type", i32 * MAXU), ("act_dof", i32 * MAXU), ("act_qadr", i32 * MAXU), ("act_ctrllimited", i32 * MAXU), ("act_forcelimited", i32 * MAXU),
        ("act_gear", f32 * MAXU), ("act_gain", f32 * MAXU), ("act_bias", (f32 * 3) * MAXU), ("act_ctrlrange", (f32 * 2) * MAXU), ("act_forcerange", (f32 * 2) * MAXU),
        ("geom_type", i32 * MAXGEOM), ("geom_link", i32 * MAXGEOM), ("geom_contype0", i32 * MAXGEOM), ("geom_conaffinity0", i32 * MAXGEOM), ("geom_tag", i32 * MAXGEOM),
        ("geom_pos", (f32 * 3) * MAXGEOM), ("geom_mat", (f32 * 9) * MAXGEOM), ("geom_size", (f32 * 3) * MAXGEOM), ("geom_rbound", f32 * MAXGEOM),
        ("geom_friction", f32 * MAXGEOM), ("geom_solref", (f32 * 2) * MAXGEOM), ("geom_solimp", (f32 * 3) * MAXGEOM), ("geom_invweight", f32 * MAXGEOM),
        ("geom_margin", f32 * MAXGEOM), ("geom_gap", f32 * MAXGEOM), ("geom_mov", i32 * MAXGEOM),
        ("geom_meshadr", i32 * MAXGEOM), ("geom_meshnum", i32 * MAXGEOM), ("mesh_vert", (f32 * 3) * MAXMESHVERT),
        ("pair_g1", i32 * MAXPAIR), ("pair_g2", i32 * MAXPAIR),
        ("site_link", i32 * MAXSITE), ("site_pos", (f32 * 3) * MAXSITE), ("site_quat", (f32 * 4) * MAXSITE),
        ("eq_link1", i32 * MAXEQ), ("eq_link2", i32 * MAXEQ), ("eq_active0", i32 * MAXEQ),
        ("eq_solref", (f32 * 2) * MAXEQ), ("eq_solimp", (f32 * 3) * MAXEQ), ("eq_invw_t", f32 * MAXEQ), ("eq_invw_r", f32 * MAXEQ), ("eq_data0", (f32 * 7) * MAXEQ),
    ]


def _sym6(I):
    return [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]


class EngineModel:
    """fe_model + the index maps the env layer needs (names -> engine ids)."""

    def __init__(self, m: mjcf.Model):
        self.src = m
        fm = FeModel()
        fm.magic, fm.struct_bytes = MAGIC, C.sizeof(FeModel)
        kin = mjcf.kinematics_np(m, m.qpos0)
        link_bodies = [b for b in range(1, m.nbody) if m.body_jntnum[b] > 0]
        robot = [b for b in link_bodies if m.jnt_type[m.body_jntadr[b]] != mjcf.JNT_FREE]
        parts = [b for b in link_bodies if m.jnt_type[m.body_jntadr[b]] == mjcf.JNT_FREE]
        assert all(m.body_jntnum[b] == 1 for b in link_bodies), "one joint per moving body"
        order = robot + parts
        assert order == sorted(robot) + sorted(parts)
        assert all(r < p for r in robot for p in parts) or not robot or not parts
        self.link_body = order
        self.body2link = {b: i for i, b in enumerate(order)}
        nl, nrl, npart = len(order), len(robot), len(parts)
        assert nl <= MAXLINK and nrl <= MAXRDOF and npart <= MAXPART
        fm.nq, fm.nv, fm.nu = m.nq, m.nv, m.nu
        fm.nlink, fm.nrlink, fm.nr, fm.npart = nl, nrl, nrl, npart
        fm.timestep, fm.impratio, fm.meaninertia = m.opt_timestep, m.opt_impratio, m.stat_meaninertia
        fm.gravity[:] = list(m.opt_gravity)

        def weld_link(b):
            w = int(m.body_weldid[b])
            return self.body2link[w] if w != 0 else -1

        self.weld_link = weld_link
        maxdepth = 0
        for i, b in enumerate(order):
            ja = int(m.body_jntadr[b])
            p = weld_link(int(m.body_parentid[b]))
            fm.link_parent[i] = p
            fm.link_jtype[i] = int(m.jnt_type[ja])
            fm.link_qadr[i] = int(m.jnt_qposadr[ja])
            fm.link_dadr[i] = int(m.jnt_dofadr[ja])
            depth = 0 if p < 0 else fm.link_depth[p] + 1
            fm.link_depth[i] = depth
            maxdepth = max(maxdepth, depth)
            if i < nrl:
                assert fm.link_dadr[i] == i and fm.link_qadr[i] == i, "robot dofs must be 0..nr-1 in link order"
                fm.link_ancmask[i] = (1 << i) | (fm.link_ancmask[p] if p >= 0 else 0)
            else:
                assert p < 0, "parts must hang off the world"
                assert fm.link_dadr[i] == nrl + 6 * (i - nrl) and fm.link_qadr[i] == nrl + 7 * (i - nrl)
            # frame relative to the parent link frame at qpos0 (robot joints have ref = 0)
            if p >= 0:
                pb = order[p]
                Rp, xp = kin["xmat"][pb], kin["xpos"][pb]
                rel_pos = Rp.T @ (kin["xpos"][b] - xp)
                rel_quat = mjcf.q_mul(mjcf.q_conj(kin["xquat"][pb]), kin["xquat"][b])
            else:
                rel_pos, rel_quat = kin["xpos"][b], kin["xquat"][b]
            fm.link_pos[i][:] = list(rel_pos)
            fm.link_quat[i][:] = list(mjcf.q_norm(rel_quat))
            fm.link_jaxis[i][:] = list(m.jnt_axis[ja])
            fm.link_jpos[i][:] = list(m.jnt_pos[ja])
            # fused inertia of every body welded to this link, in the link frame
            Rl, xl = kin["xmat"][b], kin["xpos"][b]
            members = [bb for bb in range(1, m.nbody) if int(m.body_weldid[bb]) == b]
            mass = sum(m.body_mass[bb] for bb in members)
            com = sum(m.body_mass[bb] * (Rl.T @ (kin["xipos"][bb] - xl)) for bb in members) / mass
            Ic = np.zeros((3, 3))
            for bb in members:
                Rb = Rl.T @ kin["ximat"][bb]
                d = Rl.T @ (kin["xipos"][bb] - xl) - com
                Ic += Rb @ np.diag(m.body_inertia[bb]) @ Rb.T + m.body_mass[bb] * (d @ d * np.eye(3) - np.outer(d, d))
            Io = Ic + mass * (com @ com * np.eye(3) - np.outer(com, com))
            fm.link_mass[i] = mass
            fm.link_com[i][:] = list(com)
            fm.link_inertia_c[i][:] = _sym6(Ic)
            fm.link_inertia_o[i][:] = _sym6(Io)
        fm.maxdepth = maxdepth
        fm.robot_ref[:] = list(kin["xpos"][robot[0]]) if robot else [0, 0, 0]
        for d in range(m.nv):
            fm.dof_damping[d] = m.dof_damping[d]
        for i in range(nrl):
            j = int(m.body_jntadr[order[i]])
            fm.rdof_armature[i] = m.dof_armature[i]
            fm.rdof_limited[i] = int(m.jnt_limited[j])
            fm.rdof_range[i][:] = list(m.jnt_range[j])
            fm.rdof_invweight[i] = m.dof_invweight0[i]
            fm.rdof_solref[i][:] = list(m.jnt_solref[j])
            assert abs(m.jnt_solimp[j][3] - 0.5) < 1e-12 and abs(m.jnt_solimp[j][4] - 2) < 1e-12
            fm.rdof_solimp[i][:] = list(m.jnt_solimp[j][:3])
        assert m.nu <= MAXU
        for u in range(m.nu):
            j = int(m.actuator_jntid[u])
            fm.act_type[u] = int(m.actuator_type[u])
            fm.act_dof[u] = int(m.jnt_dofadr[j])
            fm.act_qadr[u] = int(m.jnt_qposadr[j])
            fm.act_ctrllimited[u] = int(m.actuator_ctrllimited[u])
            fm.act_forcelimited[u] = int(m.actuator_forcelimited[u])
            fm.act_gear[u] = m.actuator_gear[u]
            fm.act_gain[u] = m.actuator_gainprm[u]
            fm.act_bias[u][:] = list(m.actuator_biasprm[u])
            fm.act_ctrlrange[u][:] = list(m.actuator_ctrlrange[u])
            fm.act_forcerange[u][:] = list(m.actuator_forcerange[u])
        # geoms that can ever collide
        meta = m.meta or {}
        lf, rf = set(meta.get("l_finger_geoms", [])), set(meta.get("r_finger_geoms", []))
        lf2, rf2 = set(meta.get("l_finger_geoms2", [])), set(meta.get("r_finger_geoms2", []))  # second arm (Baxter's left gripper)
        robot_geoms = set(meta.get("robot_contact_geoms", []))
        part_names = list(meta.get("part_names", [m.names["body"][b] for b in parts]))
        movable = list(meta.get("movable_geoms", []))  # Cursor agent: the cursors are static bodies repositioned through model.body_pos
        fm.nmov = len(movable)
        keep = [g for g in range(m.ngeom) if m.geom_contype[g] != 0 or m.geom_conaffinity[g] != 0 or "collision" in m.names["geom"][g]]
        self.geom_src = keep
        self.geom_map = {g: i for i, g in enumerate(keep)}
        assert len(keep) <= MAXGEOM
        fm.ngeom = len(keep)
        for i, g in enumerate(keep):
            b = int(m.geom_bodyid[g])
            l = weld_link(b)
            name = m.names["geom"][g]
            fm.geom_type[i] = int(m.geom_type[g])
            assert fm.geom_type[i] in (0, 2, 3, 5, 6, 7), "geom type not supported by the engine: %s" % name
            fm.geom_meshadr[i], fm.geom_meshnum[i] = int(m.geom_meshadr[g]), int(m.geom_meshnum[g])
            fm.geom_link[i] = l
            fm.geom_contype0[i] = int(m.geom_contype[g])
            fm.geom_conaffinity0[i] = int(m.geom_conaffinity[g])
            tag = 0
            if name == "FLOOR":
                tag |= TAG_FLOOR
            if name in lf:
                tag |= TAG_LFINGER
            if name in rf:
                tag |= TAG_RFINGER
            if name in lf2:
                tag |= TAG_LFINGER2
            if name in rf2:
                tag |= TAG_RFINGER2
            if name in robot_geoms:
                tag |= TAG_ROBOT
            bname = m.names["body"][b]
            if bname in part_names:
                pidx = part_names.index(bname)
                tag |= (pidx + 1) << TAG_PART_SHIFT
                if "collision" in name:
                    tag |= 1 << 30  # reset sets contype = conaffinity = 1 (furniture.py:1456-1461)
            fm.geom_tag[i] = tag
            if l >= 0:
                lb = order[l]
                Rl, xl = kin["xmat"][lb], kin["xpos"][lb]
                gx = kin["xpos"][b] + kin["xmat"][b] @ m.geom_pos[g]
                gR = mjcf.q_to_mat(mjcf.q_mul(kin["xquat"][b], m.geom_quat[g]))
                pos, mat = Rl.T @ (gx - xl), Rl.T @ gR
            else:
                pos = kin["xpos"][b] + kin["xmat"][b] @ m.geom_pos[g]
                mat = mjcf.q_to_mat(mjcf.q_mul(kin["xquat"][b], m.geom_quat[g]))
            fm.geom_pos[i][:] = list(pos)
            fm.geom_mat[i][:] = list(mat.ravel())
            fm.geom_size[i][:] = list(m.geom_size[g])
            fm.geom_rbound[i] = m.geom_rbound[g]
            fm.geom_friction[i] = m.geom_friction[g][0]
            fm.geom_solref[i][:] = list(m.geom_solref[g])
            assert abs(m.geom_solimp[g][3] - 0.5) < 1e-12 and abs(m.geom_solimp[g][4] - 2) < 1e-12
            fm.geom_solimp[i][:] = list(m.geom_solimp[g][:3])
            fm.geom_invweight[i] = m.body_invweight0[b][0]
            assert m.geom_condim[g] == 3, "engine assumes condim=3"
            fm.geom_margin[i], fm.geom_gap[i] = m.geom_margin[g], m.geom_gap[g]
            if name in movable:
                assert l < 0, "movable geoms sit on static bodies"
                fm.geom_mov[i] = 1 + movable.index(name)
        assert len(m.mesh_vert) <= MAXMESHVERT, "too many mesh-collider hull vertices: %d" % len(m.mesh_vert)
        for k, v in enumerate(m.mesh_vert):
            fm.mesh_vert[k][:] = list(v)
        fm.has_margin = int(any(fm.geom_margin[i] > 0 for i in range(fm.ngeom)))
        fm.has_gap = int(any(fm.geom_gap[i] > 0 for i in range(fm.ngeom)))
        pairs = []
        for g1, g2 in m.collision_pairs:
            if int(g1) in self.geom_map and int(g2) in self.geom_map:
                a, b = self.geom_map[int(g1)], self.geom_map[int(g2)]
                if fm.geom_type[a] > fm.geom_type[b]:
                    a, b = b, a
                pairs.append((a, b))
        assert len(pairs) <= MAXPAIR
        fm.npair = len(pairs)
        for k, (a, b) in enumerate(pairs):
            fm.pair_g1[k], fm.pair_g2[k] = a, b
        self.pairs = pairs
        assert m.nsite <= MAXSITE
        fm.nsite = m.nsite
        for s in range(m.nsite):
            b = int(m.site_bodyid[s])
            l = weld_link(b)
            sx = kin["xpos"][b] + kin["xmat"][b] @ m.site_pos[s]
            sq = mjcf.q_mul(kin["xquat"][b], m.site_quat[s])
            if l >= 0:
                lb = order[l]
                sx = kin["xmat"][lb].T @ (sx - kin["xpos"][lb])
                sq = mjcf.q_mul(mjcf.q_conj(kin["xquat"][lb]), sq)
            fm.site_link[s] = l
            fm.site_pos[s][:] = list(sx)
            fm.site_quat[s][:] = list(mjcf.q_norm(sq))
        assert m.neq <= MAXEQ
        fm.neq = m.neq
        for e in range(m.neq):
            b1, b2 = int(m.eq_obj1id[e]), int(m.eq_obj2id[e])
            assert b1 in self.body2link and b2 in self.body2link, "welds must join two free parts"
            fm.eq_link1[e], fm.eq_link2[e] = self.body2link[b1], self.body2link[b2]
            fm.eq_active0[e] = int(m.eq_active[e])
            fm.eq_solref[e][:] = list(m.eq_solref[e])
            fm.eq_solimp[e][:] = list(m.eq_solimp[e][:3])
            fm.eq_invw_t[e] = m.body_invweight0[b1][0] + m.body_invweight0[b2][0]
            fm.eq_invw_r[e] = m.body_invweight0[b1][1] + m.body_invweight0[b2][1]
            fm.eq_data0[e][:] = list(m.eq_data[e])
        self.fm = fm
        self.nlink, self.nrlink, self.npart = nl, nrl, npart
        self.part_names = part_names

    def blob(self):
        return C.string_at(C.addressof(self.fm), C.sizeof(self.fm))
