"""The five torque controllers of arm_controller.py (NEW_CONTROLLERS).

Golden: tests/golden/controllers.npz = the reference's own controller classes with the parameters of controller_config.hjson, run on a
stand-in simulator (tools/make_golden_controllers.py), 600 mj_steps per controller (3 episodes x 4 env steps x 50).
  * the numpy oracle (oracle/controller_oracle.py) reproduces the torques to 1e-13
  * the device controllers (csrc/fe_ctl.h through fe_ctl_eval) reproduce them on the same records ([emu] here)
"""
import os

import numpy as np
import pytest

from oracle.controller_oracle import ArmController

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = ["joint_torque", "joint_velocity", "joint_impedance", "position_orientation", "position"]


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(HERE, "golden", "controllers.npz"))
    return {m: {k.split("/")[1]: g[k] for k in g.files if k.startswith(m + "/")} for m in MODES}


@pytest.mark.parametrize("mode", MODES)
def test_oracle_reproduces_the_reference_controllers(golden, mode):
    G = golden[mode]
    n = len(G["torques"])
    fast = 0
    for t in range(n):
        if G["reset"][t]:
            c = ArmController(mode)
        fast += np.linalg.norm(G["qvel"][t]) > 7.0
        tau = c.torques(G["action"][t], bool(G["policy_step"][t]), G["pos"][t], G["R"][t].reshape(3, 3), G["velp"][t], G["velr"][t], G["q"][t], G["qvel"][t],
                        G["Jx"][t].reshape(3, 7), G["Jr"][t].reshape(3, 7), G["M"][t].reshape(7, 7))
        assert np.abs(tau - G["torques"][t]).max() <= 1e-13 * max(1.0, np.abs(G["torques"][t]).max()), (mode, t)
    assert n == 600 and int(np.sum(G["policy_step"])) == 12 and int(np.sum(G["reset"])) == 3
    assert fast >= 3  # the joint impedance controller's velocity-norm branch is in the data
    assert np.abs(G["action"]).max() > 1.0  # so is the clipping of transform_action


@pytest.mark.parametrize("mode", MODES)
def test_device_controllers_reproduce_the_reference_emu(golden, mode):
    """csrc/fe_ctl.h (float64, Cholesky + Jacobi in place of scipy.linalg.inv / numpy.linalg.svd) on the golden records, lane-emulated build"""
    from furniture_b200 import mjcf
    from furniture_b200.controllers import ctl_config
    from parity_util import make_engine

    eng = make_engine(mjcf.load_scene("None", "table_lack_0825"), 1, False)
    G = golden[mode]
    n = len(G["torques"])
    readings = np.concatenate([G["pos"], G["R"], G["velp"], G["velr"], G["q"], G["qvel"], G["Jx"], G["Jr"], G["M"]], axis=1)
    act = np.zeros((n, 7))
    act[:, : G["action"].shape[1]] = G["action"]
    starts = np.flatnonzero(G["reset"])
    counts = np.diff(np.append(starts, n))
    tau = eng.ctl_eval(ctl_config(mode), starts, counts, G["reset"], G["policy_step"], act, readings)
    scale = np.maximum(1.0, np.abs(G["torques"]).max(axis=1, keepdims=True))
    assert np.abs(tau - G["torques"]).max() / scale.max() < 1e-9 and (np.abs(tau - G["torques"]) / scale).max() < 1e-9, (mode, np.abs(tau - G["torques"]).max())


# ------------------------------------------------------------------ the env under a torque controller: device step against the CPU env
@pytest.fixture(scope="module")
def torque_sawyer():
    from furniture_b200 import mjcf

    return mjcf.load_scene("SawyerTorque", "table_lack_0825")


@pytest.mark.parametrize("mode", MODES)
def test_controller_env_steps_match_the_cpu_env(torque_sawyer, mode):
    """reset + 3 env steps (3 x 10 mj_steps, a controller evaluation before each; short, because the torque-actuated arm is barely held
    in the first steps of the reference's 2000-step goal ramp and a falling arm separates the float32 and float64 trajectories quickly)
    on the torque-actuated Sawyer: torques reach the actuators
    as ctrl = qfrc_bias + torques, observation and reward of the device equal the CPU env (oracle physics + oracle controller) --
    lane-emulated build; the CUDA build of this path has not been run (GPU budget spent before it was written)"""
    from furniture_b200.controllers import ctl_config
    from furniture_b200.engine import Engine, default_config
    from oracle.ref_env import OracleControllerEnv
    from parity_util import build_emu
    from test_env_parity import _sync_oracle_from_engine

    m = torque_sawyer
    cc = ctl_config(mode, model=m)
    eng = Engine(m, 1, config=default_config(nsub=10), lib_path=build_emu(), controller=cc)
    assert eng.act_dim == cc.control_dim + 2
    eng.env_reset()
    eng.set("qvel", np.zeros(m.nv))  # the reset leaves the motor-driven arm collapsing at several rad/s: start the comparison from rest
    eng.set("qacc_warmstart", np.zeros(m.nv))
    eng.forward()
    e = OracleControllerEnv(m, mode)
    e.nsub = 10
    e.reset()
    _sync_oracle_from_engine(e, eng, 0)
    e.sim.qfrc_applied[: e.nr] = eng.get("qfrc_applied")[0]  # the gravity compensation written at the end of the reset stays in force: _pre_action never renews it
    rng = np.random.RandomState(8)
    for k in range(3):
        a = rng.uniform(-1, 1, (1, eng.act_dim)).astype(np.float32)
        a[0, -1] = -0.5
        obs, rew, done, info = eng.env_step_host(a)
        ob, r, d, inf = e.step(a[0].astype(np.float64))
        tau = np.array(e.torques[-1])
        ctrl = eng.get("ctrl")[0]
        want = e.sim.ctrl
        assert np.abs(ctrl - want).max() < 2e-3 * max(1.0, np.abs(want).max()), (mode, k, ctrl, want)
        assert np.abs(obs[0] - ob).max() < 2e-3, (mode, k, np.abs(obs[0] - ob).max())
        assert abs(rew[0] - r) < 1e-5 and bool(done[0]) == d and info[0][3] == k + 1
        assert np.isfinite(tau).all()
