"""Dense (phase-based) reward, furniture_sawyer_dense.py:222-1045.

Golden: tests/golden/dense_reward.npz = the reference's own _compute_reward run on a scripted world (tools/make_golden_dense.py).
  * the numpy oracle (oracle/dense_oracle.py) must reproduce it: phases / subtasks / flags bit-exactly, rewards to 1e-12
  * the device state machine (fe_dense.h, through fe_dense_eval of the C-ABI) must reproduce it on the same records:
    [emu] build on CPU, [cuda] build on the GPU
"""
import json
import os

import numpy as np
import pytest

from oracle import dense_oracle as D

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "dense_reward.npz"))


class GoldenWorld:
    """feeds one golden record to the oracle by name"""

    def __init__(self, g, recipe):
        self.g, self.t = g, 0
        self.names = {}
        n = len(recipe["recipe"])
        for s in range(n):
            leg = recipe["recipe"][s][0]
            self.names[leg] = ("leg_pos", s)
            self.names[recipe["site_recipe"][s][0]] = ("leg_site", s)
            self.names[recipe["site_recipe"][s][1]] = ("table_site", s)
            for i in range(n):
                self.names["%s_ltgt_site%d" % (leg, i)] = ("gl", s)
                self.names["%s_rtgt_site%d" % (leg, i)] = ("gr", s)

    def pos(self, name):
        g, t = self.g, self.t
        if name == "griptip_site":
            return g["eef"][t].copy()
        kind, s = self.names[name]
        key = {"leg_site": "leg_site_pos", "table_site": "table_site_pos"}.get(kind, kind)
        return g[key][t, s].copy()

    def _mat(self, name):
        g, t = self.g, self.t
        if name == "grip_site":
            return g["grip_mat"][t].reshape(3, 3)
        kind, s = self.names[name]
        return g["leg_site_mat" if kind == "leg_site" else "table_site_mat"][t, s].reshape(3, 3)

    def up(self, name):
        return self._mat(name)[:, 2].copy()

    def forward(self, name):
        return self._mat(name)[:, 1].copy()

    def finger_contact(self, leg):
        s = self.names[leg][1]
        return bool(self.g["touchL"][self.t, s]), bool(self.g["touchR"][self.t, s])


def golden_cfg(g, ep):
    cfg = {k: float(v) for k, v in zip(g["coef_names"], g["coefs"]) if k in D.DEFAULTS}
    for k in ("diff_rew", "early_termination", "phase_ob", "reset_robot_after_attach"):
        cfg[k] = bool(g["ep_" + k][ep])
    return cfg


def test_oracle_reproduces_the_reference_reward_machine(golden):
    g = {k: golden[k] for k in golden.files}
    recipes = [json.loads(str(r)) for r in g["recipe_json"]]
    n = len(g["reward"])
    worst = 0.0
    orc = None
    for t in range(n):
        ep = int(g["episode"][t])
        if g["is_reset"][t]:
            recipe = recipes[int(g["ep_recipe"][ep])]
            world = GoldenWorld(g, recipe)
            world.t = t
            orc = D.DenseOracle(world, recipe, golden_cfg(g, ep))
            orc.begin_episode()
        else:
            world.t = t
            r, done, info = orc.step(g["ac"][t], bool(g["connected"][t]))
            assert done == bool(g["done"][t]), t
            assert orc.success == bool(g["success"][t]), t
            ref = g["reward"][t]
            if np.isnan(ref):
                assert np.isnan(r), t
            else:
                err = abs(r - ref) / max(1.0, abs(ref))
                worst = max(worst, err)
                assert err < 1e-12, (t, r, ref)
            for k, want in zip(D.INFO_KEYS, g["info"][t]):
                got = float(info[k])
                assert (np.isnan(want) and np.isnan(got)) or abs(got - want) <= 1e-12 * max(1.0, abs(want)), (t, k, got, want)
        assert orc.phase == int(g["phase"][t]), t
        assert orc.subtask == int(g["subtask"][t]), t
    assert int(np.sum(g["success"])) > 20 and len(set(g["phase"].tolist())) == 8  # the script did reach every phase and finish episodes
