"""Generate golden vectors for the assembly logic by running the REFERENCE'S OWN Python, unmodified.

Runs only in the build container (needs /root/reference); writes small fixtures under tests/golden/ that travel to
the GPU box.  What is executed from the reference:
  * furniture/env/furniture.py  FurnitureEnv._is_aligned            (:1057-1153)  -> is_aligned.npz
  * furniture/env/transform_utils.py  transform_to_target_quat (:641-664), rel_pose (:633-638),
    euler_to_quat (:617-630)                                                     -> connect_geom.npz
Third-party modules that are absent here (gym, mujoco_py, pyquaternion, ...) are replaced by stubs in sys.modules;
none of them is on the _is_aligned path (pure numpy through transform_utils).  transform_to_target_quat / rel_pose
need `pyquaternion.Quaternion`: a small stand-in with pyquaternion's documented semantics is injected
(Hamilton product, inverse = conjugate / |q|^2, rotate() normalises the quaternion first) -- so connect_geom.npz
is pinned to the reference code *modulo that stand-in*, and says so in its metadata.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class Quaternion:
    """Stand-in for pyquaternion.Quaternion (only what transform_utils uses)."""

    def __init__(self, *args, **kw):
        if "axis" in kw:
            ax = np.asarray(kw["axis"], dtype=float)
            ax = ax / np.linalg.norm(ax)
            ang = np.deg2rad(kw["degrees"]) if "degrees" in kw else kw.get("radians", kw.get("angle"))
            self.q = np.concatenate([[np.cos(ang / 2.0)], ax * np.sin(ang / 2.0)])
        elif "array" in kw:
            self.q = np.asarray(kw["array"], dtype=float).copy()
        elif "vector" in kw:
            self.q = np.concatenate([[0.0], np.asarray(kw["vector"], dtype=float)])
        elif len(args) == 0:
            self.q = np.array([1.0, 0, 0, 0])
        elif len(args) == 1:
            a = args[0]
            self.q = a.q.copy() if isinstance(a, Quaternion) else np.asarray(a, dtype=float).copy()
        else:
            self.q = np.asarray(args, dtype=float)

    def _q_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])

    def __mul__(self, other):
        return Quaternion(array=np.dot(self._q_matrix(), other.q))

    @property
    def conjugate(self):
        return Quaternion(array=np.hstack((self.q[0], -self.q[1:4])))

    @property
    def inverse(self):
        ss = np.dot(self.q, self.q)
        return Quaternion(array=np.hstack((self.q[0], -self.q[1:4])) / ss)

    def rotate(self, vector):
        n = np.sqrt(np.dot(self.q, self.q))
        if abs(1.0 - n) >= 1e-14 and n > 0:
            self.q = self.q / n
        v = Quaternion(vector=vector)
        return (self * v * self.conjugate).q[1:4]

    def __iter__(self):
        return iter(self.q)

    def __getitem__(self, i):
        return self.q[i]


def import_reference():
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    class _Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any

    for name in ["gym", "gym.spaces", "gym.envs", "gym.envs.registration", "mujoco_py", "mujoco_py.generated", "colorlog", "imageio", "moviepy",
                 "moviepy.editor", "glfw", "gdown", "pybullet", "cloudpickle", "mpi4py", "hjson", "cv2", "PIL", "PIL.Image", "tqdm", "pyquaternion",
                 "scipy.interpolate", "matplotlib", "matplotlib.pyplot", "torch", "torchvision", "torchvision.utils", "h5py", "yaml"]:
        if name not in sys.modules:
            try:
                if name in ("yaml", "hjson", "scipy.interpolate", "cv2"):
                    __import__(name)
                    continue
            except Exception:
                pass
            sys.modules[name] = _Stub(name)
    sys.modules["pyquaternion"].Quaternion = Quaternion
    sys.path.insert(0, REF)
    import furniture.env.transform_utils as T  # noqa
    from furniture.env.furniture import FurnitureEnv  # noqa

    return T, FurnitureEnv


def rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def small_rot(rng, max_deg):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    a = np.deg2rad(rng.uniform(0, max_deg))
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def main():
    T, FurnitureEnv = import_reference()
    rng = np.random.RandomState(20260923)
    N = 6000
    angle_sets = [[0.0, 90.0, 180.0, 270.0], [0.0, 180.0], [], [0.0], [45.0, 135.0, 225.0, 315.0]]
    thresholds = [(0.1, 0.9, 0.9, 0.3), (0.02, 0.99, 0.99, 0.0)]  # config/furniture.py:203-226, furniture_sawyer_dense.py:11-14
    rec = dict(p1=[], m1=[], p2=[], m2=[], angles=[], nangles=[], thr=[], aligned=[], tq=[], tq_set=[])
    for n in range(N):
        angs = angle_sets[rng.randint(len(angle_sets))]
        thr = thresholds[rng.randint(2)]
        R1 = rand_rot(rng)
        p1 = rng.uniform(-0.5, 0.5, size=3)
        mode = rng.randint(4)
        if mode == 0:  # unrelated pose
            R2 = rand_rot(rng)
            p2 = p1 + rng.normal(size=3) * 0.1
        else:  # near-aligned: same up, forward rotated by an allowed angle (+ noise), offset along/around up
            base = angs[rng.randint(len(angs))] if angs else rng.uniform(0, 360)
            a = np.deg2rad(base + rng.normal() * (20.0 if mode == 1 else 4.0))
            Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
            R2 = small_rot(rng, 30.0 if mode == 1 else 6.0) @ R1 @ Rz
            scale = thr[0] * (1.5 if mode == 1 else 0.7)
            p2 = p1 + R1[:, 2] * rng.uniform(-1, 1) * scale + rng.normal(size=3) * scale * 0.25
        name1 = "a-b," + "".join("%g," % x for x in angs) + "conn_site1"
        name2 = "b-a," + "".join("%g," % x for x in angs) + "conn_site1"
        poses = {name1: (p1, R1), name2: (p2, R2)}

        class Fake:
            pass

        fake = Fake()
        fake._config = types.SimpleNamespace(alignment_pos_dist=thr[0], alignment_rot_dist_up=thr[1], alignment_rot_dist_forward=thr[2], alignment_project_dist=thr[3])
        fake._site_xpos_xquat = lambda nm: np.hstack([poses[nm][0], [1, 0, 0, 0]])
        fake._get_up_vector = lambda nm: poses[nm][1][:, 2].copy()
        fake._get_forward_vector = lambda nm: poses[nm][1][:, 1].copy()
        fake._target_connector_xquat = None
        ok = FurnitureEnv._is_aligned(fake, name1, name2)
        rec["p1"].append(p1); rec["m1"].append(R1.ravel()); rec["p2"].append(p2); rec["m2"].append(R2.ravel())
        pad = np.zeros(4); pad[: len(angs)] = angs
        rec["angles"].append(pad); rec["nangles"].append(len(angs)); rec["thr"].append(thr)
        rec["aligned"].append(bool(ok))
        tq = fake._target_connector_xquat
        rec["tq_set"].append(tq is not None)
        rec["tq"].append(np.asarray(tq, dtype=np.float64) if tq is not None else np.full(4, np.nan))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "is_aligned.npz"), **{k: np.array(v) for k, v in rec.items()},
                        source="reference furniture/env/furniture.py:_is_aligned run unmodified (tools/make_golden_assembly.py)")
    print("is_aligned: %d cases, %d aligned, %d with target quat" % (N, sum(rec["aligned"]), sum(rec["tq_set"])))

    # ---- connect geometry (transform_utils with the Quaternion stand-in)
    M = 2000
    g = dict(qb=[], q=[], tq=[], new_pos=[], new_quat=[], rel=[], eul=[], eq=[])
    for n in range(M):
        def rq():
            q = rng.normal(size=4)
            return q / np.linalg.norm(q)
        qb = np.concatenate([rng.uniform(-1, 1, 3), rq()])
        q = np.concatenate([rng.uniform(-1, 1, 3), rq()])
        tq = rq()
        npos, nquat = T.transform_to_target_quat(qb, q, tq)
        rel = T.rel_pose(qb, q)
        e = rng.uniform(-180, 180, size=3)
        eq = T.euler_to_quat(e, tq)
        for k, v in zip(("qb", "q", "tq", "new_pos", "new_quat", "rel", "eul", "eq"), (qb, q, tq, npos, nquat, rel, e, eq)):
            g[k].append(np.asarray(v, dtype=np.float64))
    np.savez_compressed(os.path.join(OUT, "connect_geom.npz"), **{k: np.array(v) for k, v in g.items()},
                        source="reference transform_utils.{transform_to_target_quat,rel_pose,euler_to_quat} run with a pyquaternion stand-in (tools/make_golden_assembly.py)")
    print("connect_geom: %d cases" % M)
    # the single known-answer doctest the reference holds on this path (transform_utils.py:35-36)
    assert np.allclose(T.quat_multiply([1, -2, 3, 4], [-5, 6, 7, 8]), [-44, -14, 48, 28])


if __name__ == "__main__":
    main()
