"""Golden vectors for what FurnitureEnv._do_ik_step does *around* the inverse-kinematics call (control_type="ik", Sawyer), made by running
the REFERENCE'S OWN Python, unmodified (furniture/env/furniture.py:2899-2996, _bounded_d_pos :1252-1258, _make_input :1332-1343;
transform_utils.euler_to_quat / quat_multiply / quat_inverse / quat2mat / mat2quat).  Runs only in the build container.

The simulator and the pybullet controller are replaced by stand-ins that show a hand pose and record what the controller is asked for:
`dpos` and `rotation` of the first get_control call, the accumulated `_initial_right_hand_quat`, and how often _do_simulation and the
closed-loop get_control() run.  pyquaternion is absent: the stand-in of tools/make_golden_assembly.py (documented semantics) is used by
euler_to_quat, so this golden is pinned to the reference code modulo that stand-in, like connect_geom.npz.
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_assembly import import_reference, rand_rot  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    T, FurnitureEnv = import_reference()
    rng = np.random.RandomState(20260925)
    rec = {k: [] for k in ("action", "hand_pos", "hand_R", "s_in", "dpos", "rotation", "s_out", "low_grip", "n_sim", "n_closed_loop")}
    for n in range(600):
        hand_R = rand_rot(rng)
        hand_pos = rng.uniform(-1.6, 1.6, size=3) if n % 3 == 0 else rng.uniform(-0.5, 0.5, size=3) + [0, 0, 0.6]
        calls = dict(sim=0, closed=0, first=None)

        class Ctl:
            def get_control(self, dpos=None, rotation=None):
                if dpos is None:
                    calls["closed"] += 1
                else:
                    calls["first"] = (np.array(dpos, dtype=np.float64), np.array(rotation, dtype=np.float64))
                return np.zeros(7)

        class Fake:
            _control_type, _agent_type, _record_demo, _action_repeat = "ik", "Sawyer", False, 3
            _move_speed, _rotate_speed = 0.1, 22.5
            _min_gripper_pos, _max_gripper_pos = np.array([-1.5, -1.5, 0.0]), np.array([1.5, 1.5, 1.5])
            _controller = Ctl()
            sim = types.SimpleNamespace(data=types.SimpleNamespace(get_body_xpos=lambda name: hand_pos.copy()))
            _bounded_d_pos = FurnitureEnv._bounded_d_pos
            _make_input = FurnitureEnv._make_input

            @property
            def _right_hand_quat(self):
                return T.mat2quat(np.ascontiguousarray(hand_R, dtype=np.float32))  # numpy 2: the reference asks for copy=False, so hand it float32

            def _setup_action(self, low):
                self.low = np.array(low, dtype=np.float64)
                return low

            def _do_simulation(self, ctrl):
                calls["sim"] += 1

        fake = Fake()
        s_in = T.mat2quat(rand_rot(rng).astype(np.float32)) if n % 2 else T.mat2quat(hand_R.astype(np.float32))
        if n % 5 == 4:  # after some steps the accumulated target is a python list of float64, not a unit float32 quaternion
            s_in = list(np.asarray(s_in, dtype=np.float64) * (1 + 1e-3 * rng.normal()))
        fake._initial_right_hand_quat = s_in
        a = rng.uniform(-1, 1, size=8)
        if n % 7 == 0:
            a[3:6] = 0
        FurnitureEnv._do_ik_step(fake, a.copy())
        rec["action"].append(a); rec["hand_pos"].append(hand_pos); rec["hand_R"].append(hand_R.ravel()); rec["s_in"].append(np.asarray(s_in, dtype=np.float64))
        rec["dpos"].append(calls["first"][0]); rec["rotation"].append(calls["first"][1].ravel())
        rec["s_out"].append(np.asarray(fake._initial_right_hand_quat, dtype=np.float64)); rec["low_grip"].append(fake.low[7])
        rec["n_sim"].append(calls["sim"]); rec["n_closed_loop"].append(calls["closed"])
    # ---- control_type="ik_quaternion" (furniture.py:2998-3058): displacement as above, the rotation given as a quaternion (w, x, y, z)
    # relative to the hand's current orientation; nothing is accumulated
    rq = {k: [] for k in ("action", "hand_pos", "hand_R", "dpos", "rotation", "low_grip", "n_sim", "n_closed_loop")}
    for n in range(300):
        hand_R = rand_rot(rng)
        hand_pos = rng.uniform(-1.6, 1.6, size=3) if n % 3 == 0 else rng.uniform(-0.5, 0.5, size=3) + [0, 0, 0.6]
        calls = dict(sim=0, closed=0, first=None)

        class CtlQ:
            def get_control(self, dpos=None, rotation=None):
                if dpos is None:
                    calls["closed"] += 1
                else:
                    calls["first"] = (np.array(dpos, dtype=np.float64), np.array(rotation, dtype=np.float64))
                return np.zeros(7)

        class FakeQ:
            _control_type, _agent_type, _record_demo, _action_repeat, _arms = "ik_quaternion", "Sawyer", False, 3, ["right"]
            _move_speed, _rotate_speed = 0.1, 22.5
            _min_gripper_pos, _max_gripper_pos = np.array([-1.5, -1.5, 0.0]), np.array([1.5, 1.5, 1.5])
            _controller = CtlQ()
            sim = types.SimpleNamespace(data=types.SimpleNamespace(get_body_xpos=lambda name: hand_pos.copy()))
            _bounded_d_pos = FurnitureEnv._bounded_d_pos
            _make_input = FurnitureEnv._make_input

            @property
            def _right_hand_quat(self):
                return T.mat2quat(np.ascontiguousarray(hand_R, dtype=np.float32))

            def _setup_action(self, low):
                self.low = np.array(low, dtype=np.float64)
                return low

            def _do_simulation(self, ctrl):
                calls["sim"] += 1

        fq = FakeQ()
        a = rng.uniform(-1, 1, size=9)
        q = rng.normal(size=4)
        a[3:7] = q / np.linalg.norm(q) * (1.0 if n % 4 else 1.3)  # now and then not a unit quaternion: quat2mat normalises
        FurnitureEnv._do_ik_step(fq, a.copy())
        rq["action"].append(a); rq["hand_pos"].append(hand_pos); rq["hand_R"].append(hand_R.ravel())
        rq["dpos"].append(calls["first"][0]); rq["rotation"].append(calls["first"][1].ravel()); rq["low_grip"].append(fq.low[7])
        rq["n_sim"].append(calls["sim"]); rq["n_closed_loop"].append(calls["closed"])
    print("ik_quaternion: %d cases" % len(rq["action"]))
    # ---- Baxter, control_type="ik" (furniture.py:2933-2970): the same per arm, right then left; both grippers pass through
    rb = {k: [] for k in ("action", "hand_pos", "hand_R", "s_in", "dpos", "rotation", "s_out", "low_grips")}
    for n in range(200):
        hand_R = [rand_rot(rng), rand_rot(rng)]
        hand_pos = [rng.uniform(-0.5, 0.5, size=3) + [0, 0, 0.6], rng.uniform(-1.6, 1.6, size=3)]
        got = {}

        class CtlB:
            def get_control(self, right=None, left=None):
                if right is not None:
                    got["right"], got["left"] = right, left
                return np.zeros(14)

        class FakeB:
            _control_type, _agent_type, _record_demo, _action_repeat = "ik", "Baxter", False, 3
            _move_speed, _rotate_speed = 0.1, 22.5
            _min_gripper_pos, _max_gripper_pos = np.array([-1.5, -1.5, 0.0]), np.array([1.5, 1.5, 1.5])
            _controller = CtlB()
            sim = types.SimpleNamespace(data=types.SimpleNamespace(get_body_xpos=lambda name: hand_pos[0 if name == "right_hand" else 1].copy()))
            _bounded_d_pos = FurnitureEnv._bounded_d_pos
            _make_input = FurnitureEnv._make_input
            _right_hand_quat = property(lambda self: T.mat2quat(np.ascontiguousarray(hand_R[0], dtype=np.float32)))
            _left_hand_quat = property(lambda self: T.mat2quat(np.ascontiguousarray(hand_R[1], dtype=np.float32)))

            def _setup_action(self, low):
                self.low = np.array(low, dtype=np.float64)
                return low

            def _do_simulation(self, ctrl):
                pass

        fb = FakeB()
        s_in = [T.mat2quat(rand_rot(rng).astype(np.float32)), T.mat2quat(hand_R[1].astype(np.float32))]
        fb._initial_right_hand_quat, fb._initial_left_hand_quat = s_in
        a = rng.uniform(-1, 1, size=15)
        FurnitureEnv._do_ik_step(fb, a.copy())
        rb["action"].append(a); rb["hand_pos"].append(np.array(hand_pos)); rb["hand_R"].append(np.array([r.ravel() for r in hand_R]))
        rb["s_in"].append(np.array(s_in, dtype=np.float64))
        rb["dpos"].append(np.array([got["right"]["dpos"], got["left"]["dpos"]], dtype=np.float64))
        rb["rotation"].append(np.array([np.asarray(got["right"]["rotation"]).ravel(), np.asarray(got["left"]["rotation"]).ravel()], dtype=np.float64))
        rb["s_out"].append(np.array([fb._initial_right_hand_quat, fb._initial_left_hand_quat], dtype=np.float64)); rb["low_grips"].append(fb.low[14:16])
    print("baxter ik: %d cases" % len(rb["action"]))
    # ---- Baxter, control_type="ik_quaternion" (furniture.py:2998-3058 with two arms): 17 numbers
    rbq = {k: [] for k in ("action", "hand_pos", "hand_R", "dpos", "rotation", "low_grips")}
    for n in range(100):
        hand_R = [rand_rot(rng), rand_rot(rng)]
        hand_pos = [rng.uniform(-0.5, 0.5, size=3) + [0, 0, 0.6], rng.uniform(-1.6, 1.6, size=3)]
        got = {}

        class CtlBQ:
            def get_control(self, right=None, left=None):
                if right is not None:
                    got["right"], got["left"] = right, left
                return np.zeros(14)

        class FakeBQ(FakeB):
            _control_type, _arms = "ik_quaternion", ["right", "left"]
            _controller = CtlBQ()
            sim = types.SimpleNamespace(data=types.SimpleNamespace(get_body_xpos=lambda name: hand_pos[0 if name == "right_hand" else 1].copy()))
            _right_hand_quat = property(lambda self: T.mat2quat(np.ascontiguousarray(hand_R[0], dtype=np.float32)))
            _left_hand_quat = property(lambda self: T.mat2quat(np.ascontiguousarray(hand_R[1], dtype=np.float32)))

        fbq = FakeBQ()
        a = rng.uniform(-1, 1, size=17)
        for off in (3, 10):
            q = rng.normal(size=4)
            a[off : off + 4] = q / np.linalg.norm(q)
        FurnitureEnv._do_ik_step(fbq, a.copy())
        rbq["action"].append(a); rbq["hand_pos"].append(np.array(hand_pos)); rbq["hand_R"].append(np.array([r.ravel() for r in hand_R]))
        rbq["dpos"].append(np.array([got["right"]["dpos"], got["left"]["dpos"]], dtype=np.float64))
        rbq["rotation"].append(np.array([np.asarray(got["right"]["rotation"]).ravel(), np.asarray(got["left"]["rotation"]).ravel()], dtype=np.float64))
        rbq["low_grips"].append(fbq.low[14:16])
    print("baxter ik_quaternion: %d cases" % len(rbq["action"]))
    np.savez_compressed(os.path.join(OUT, "ik_pre.npz"), **{"q_" + k: np.array(v) for k, v in rq.items()}, **{"b_" + k: np.array(v) for k, v in rb.items()}, **{"bq_" + k: np.array(v) for k, v in rbq.items()}, **{k: np.array(v) for k, v in rec.items()},
                        source="reference FurnitureEnv._do_ik_step run unmodified around stand-ins for the simulator and the pybullet controller (tools/make_golden_ik.py)")
    print("ik_pre: %d cases; _do_simulation calls %s, closed-loop get_control calls %s" % (len(rec["action"]), set(rec["n_sim"]), set(rec["n_closed_loop"])))


if __name__ == "__main__":
    main()
