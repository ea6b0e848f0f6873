"""CPU ORACLE for the assembly logic (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

numpy restatement of the reference's connect path; pinned bit-exactly (decisions) against golden vectors produced by
the reference's own Python (tools/make_golden_assembly.py -> tests/golden/is_aligned.npz, connect_geom.npz):

  is_aligned               <- FurnitureEnv._is_aligned            furniture/env/furniture.py:1057-1153
  unit_vector_f32          <- transform_utils.unit_vector (2nd def, float32)  transform_utils.py:559-600
  rotate_vector            <- transform_utils.rotate_vector       transform_utils.py:739-745
  rotate_vector_cos_siml   <- transform_utils.rotate_vector_cos_siml :748-754
  lookat_to_quat           <- transform_utils.lookat_to_quat      :457-516
  transform_to_target_quat <- transform_utils.transform_to_target_quat :641-664   (pyquaternion semantics restated)
  rel_pose                 <- transform_utils.rel_pose            :633-638
  connect_masks            <- FurnitureEnv._connect collision-group rewrite  furniture.py:869-878
"""
import math

import numpy as np


def unit_vector_f32(v):
    d = np.array(v, dtype=np.float32, copy=True)
    d /= math.sqrt(np.dot(d, d))
    return d


def cos_siml(a, b):
    return np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b)


def rotate_vector(v, axis, angle_deg):
    v = np.asarray(v)
    k = unit_vector_f32(axis)
    a = angle_deg / 180 * np.pi
    return np.cos(a) * v + np.sin(a) * np.cross(k, v)


def rotate_vector_cos_siml(v, axis, cos, direction):
    v = np.asarray(v)
    k = unit_vector_f32(axis)
    return cos * v + direction * np.sqrt(1 - cos**2) * np.cross(k, v)


def _norm(x):
    return x / np.linalg.norm(x)


def lookat_to_quat(forward, up):
    """returns xyzw"""
    vector = _norm(forward)
    vector2 = _norm(np.cross(_norm(up), vector))
    vector3 = np.cross(vector, vector2)
    m00, m01, m02 = vector2
    m10, m11, m12 = vector3
    m20, m21, m22 = vector
    num8 = (m00 + m11) + m22
    q = np.zeros(4)
    if num8 > 0:
        num = np.sqrt(num8 + 1)
        q[3] = num * 0.5
        num = 0.5 / num
        q[0] = (m12 - m21) * num
        q[1] = (m20 - m02) * num
        q[2] = (m01 - m10) * num
        return q
    if (m00 >= m11) and (m00 >= m22):
        num7 = np.sqrt(((1 + m00) - m11) - m22)
        num4 = 0.5 / num7
        q[0] = 0.5 * num7
        q[1] = (m01 + m10) * num4
        q[2] = (m02 + m20) * num4
        q[3] = (m12 - m21) * num4
        return q
    if m11 > m22:
        num6 = np.sqrt(((1 + m11) - m00) - m22)
        num3 = 0.5 / num6
        q[0] = (m10 + m01) * num3
        q[1] = 0.5 * num6
        q[2] = (m21 + m12) * num3
        q[3] = (m20 - m02) * num3
        return q
    num5 = np.sqrt(((1 + m22) - m00) - m11)
    num2 = 0.5 / num5
    q[0] = (m20 + m02) * num2
    q[1] = (m21 + m12) * num2
    q[2] = 0.5 * num5
    q[3] = (m01 - m10) * num2
    return q


def is_aligned(p1, m1, p2, m2, angles, thr):
    """p*: site world position (3,), m*: site world rotation (3,3); angles: allowed angles in degrees (listed order);
    thr = (pos_dist, rot_dist_up, rot_dist_forward, project_dist). Returns (aligned, target_quat_wxyz or None)."""
    p1, p2 = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64)
    m1, m2 = np.asarray(m1, dtype=np.float64).reshape(3, 3), np.asarray(m2, dtype=np.float64).reshape(3, 3)
    up1, up2 = m1[:, 2].copy(), m2[:, 2].copy()
    f1, f2 = m1[:, 1].copy(), m2[:, 1].copy()
    pos_dist = np.linalg.norm(p1 - p2)
    rot_up = cos_siml(up1, up2)
    proj12 = np.dot(up1, unit_vector_f32(p2 - p1))
    proj21 = np.dot(up2, unit_vector_f32(p1 - p2))
    tq = None
    if len(angles) == 0:
        fwd_ok = True
        c = cos_siml(f1, f2)
        fp = rotate_vector_cos_siml(f1, up1, c, 1)
        fn = rotate_vector_cos_siml(f1, up1, c, -1)
        fr = fp if cos_siml(fp, f2) > cos_siml(fn, f2) else fn
        tq = lookat_to_quat(up1, fr)[[3, 0, 1, 2]]
    else:
        fwd_ok = False
        for a in angles:
            fr = rotate_vector(f1, up1, float(a))
            if cos_siml(fr, f2) > thr[2]:
                fwd_ok = True
                tq = lookat_to_quat(up1, fr)[[3, 0, 1, 2]]
                break
    if pos_dist < thr[0] and rot_up > thr[1] and fwd_ok and abs(proj12) > thr[3] and abs(proj21) > thr[3]:
        return True, tq
    if pos_dist < thr[0] / 2 and rot_up > thr[1] and fwd_ok:
        return True, tq
    return False, tq


# ---- pyquaternion semantics (w,x,y,z)
def _qmul(a, b):
    w, x, y, z = a
    M = np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])
    return M @ np.asarray(b, dtype=np.float64)


def _qinv(q):
    q = np.asarray(q, dtype=np.float64)
    return np.hstack((q[0], -q[1:4])) / np.dot(q, q)


def _qrotate(q, v):
    q = np.asarray(q, dtype=np.float64)
    n = np.sqrt(np.dot(q, q))
    if abs(1.0 - n) >= 1e-14 and n > 0:
        q = q / n
    qv = np.concatenate([[0.0], v])
    return _qmul(_qmul(q, qv), np.hstack((q[0], -q[1:4])))[1:4]


def transform_to_target_quat(qpos_base, qpos, target_quat):
    cur_pos, cur_rot = qpos_base[:3], qpos_base[3:]
    rel = _qmul(target_quat, _qinv(cur_rot))
    new_pos = _qrotate(rel, qpos[:3] - cur_pos) + cur_pos
    return new_pos, _qmul(rel, qpos[3:])


def rel_pose(qpos1, qpos2):
    inv = _qinv(qpos1[3:])
    return np.concatenate([_qrotate(inv, qpos2[:3] - qpos1[:3]), _qmul(inv, qpos2[3:])])


def euler_to_quat(rotation_deg, quat=None):
    def ax(a, d):
        a = np.asarray(a, dtype=np.float64)
        r = np.deg2rad(d)
        return np.concatenate([[np.cos(r / 2)], a * np.sin(r / 2)])

    q = _qmul(_qmul(ax([0, 0, 1], rotation_deg[2]), ax([0, 1, 0], rotation_deg[1])), ax([1, 0, 0], rotation_deg[0]))
    return q if quat is None else _qmul(quat, q)


def connect_masks(group1):
    """furniture.py:875-878: (contype, conaffinity) written to every colliding geom of both merged groups."""
    return (1 << 30) - 1 - (1 << (group1 + 1)), 1 << (group1 + 1)
