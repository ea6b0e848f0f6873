"""control_type="ik" (SURVEY 8 f4): what _do_ik_step does around the solver is pinned to the reference's own code
(tests/golden/ik_pre.npz, tools/make_golden_ik.py); the solver itself (damped least squares on the arm's chain, in place of the
pybullet call) is checked for reaching its target and device-against-oracle."""
import os

import numpy as np
import pytest

from furniture_b200 import ik as IK
from furniture_b200 import mjcf
from oracle import ik_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
BACKENDS = [pytest.param(False, id="emu"), pytest.param(True, id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module")
def sawyer():
    return mjcf.load_scene("Sawyer", "table_lack_0825")


def test_ik_step_preamble_equals_the_reference_code():
    g = np.load(os.path.join(HERE, "golden", "ik_pre.npz"))
    p = dict(IK.IK_DEFAULTS)
    assert set(g["n_sim"]) == {p["action_repeat"]} and set(g["n_closed_loop"]) == {p["action_repeat"] - 1}
    for i in range(len(g["action"])):
        hq = O.mat2quat(g["hand_R"][i].reshape(3, 3))
        d_pos, rot, s_new, grip = O.ik_pre(g["action"][i], g["hand_pos"][i], hq, g["s_in"][i], p)
        assert np.array_equal(d_pos, g["dpos"][i]), i
        assert np.allclose(s_new, g["s_out"][i], rtol=0, atol=1e-15), i
        assert np.abs(rot - g["rotation"][i].reshape(3, 3)).max() < 2e-6, i  # float32 quaternion products; the hand quaternion's sign is free
        assert grip == g["low_grip"][i]


def test_chain_and_solver_reach_the_commanded_hand_pose(sawyer):
    m = sawyer
    p = IK.ik_params(m)
    q = np.array(m.meta["robot_init_qpos"], dtype=float)
    full = np.array(m.qpos0, dtype=float)
    full[p["chain"]["qadr"]] = q
    kin = mjcf.kinematics_np(m, full)
    hb = m.names["body"].index("right_hand")
    hp, hq, _, _ = IK.chain_fk(p["chain"], q)
    assert np.abs(hp - kin["xpos"][hb]).max() < 1e-12 and np.abs(hq - kin["xquat"][hb]).max() < 1e-12  # the chain is the model's arm
    rng = np.random.RandomState(0)
    for _ in range(20):
        tp = hp + rng.uniform(-0.06, 0.06, 3)
        tq = mjcf.q_norm(mjcf.q_mul(mjcf.q_axis_angle(rng.normal(size=3), rng.uniform(0, 0.3)), hq))
        qs, it = IK.solve_ik(p, q, tp, tq)
        a, b, _, _ = IK.chain_fk(p["chain"], qs)
        assert it < p["max_iters"] - 1 and np.linalg.norm(a - tp) < p["tol_pos"] and np.linalg.norm(IK.rot_error(tq, b)) < p["tol_rot"]
        assert (qs >= np.array(p["lower"]) - 1e-12).all() and (qs <= np.array(p["upper"]) + 1e-12).all()
