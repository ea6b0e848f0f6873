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
