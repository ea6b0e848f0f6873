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


# ------------------------------------------------------------------ the device machine on the same records
def _device_inputs(g, recipes):
    """golden records -> the arrays of fe_dense_eval; site ids: leg_site s, table_site 8+s, g_l 16+s, g_r 24+s, griptip 32, grip_site 33"""
    n = len(g["reward"])
    spos = np.zeros((n, 34, 3))
    smat = np.zeros((n, 34, 9))
    for k, (pk, mk) in enumerate((("leg_site_pos", "leg_site_mat"), ("table_site_pos", "table_site_mat"), ("gl", None), ("gr", None))):
        spos[:, 8 * k : 8 * k + 4] = g[pk]
        if mk:
            smat[:, 8 * k : 8 * k + 4] = g[mk]
    spos[:, 32] = g["eef"]
    smat[:, 33] = g["grip_mat"]
    touch = (g["touchL"] & g["touchR"]).astype(np.uint8)
    return spos, smat, g["leg_pos"], touch


def _packed_recipe(recipe):
    from furniture_b200.dense import pack_dense_recipe

    ids = {}
    for s, (leg, _) in enumerate(recipe["recipe"]):
        ids[recipe["site_recipe"][s][0]] = s
        ids[recipe["site_recipe"][s][1]] = 8 + s
        for k in range(len(recipe["recipe"])):
            ids["%s_ltgt_site%d" % (leg, k)] = 16 + s
            ids["%s_rtgt_site%d" % (leg, k)] = 24 + s
    legs = [r[0] for r in recipe["recipe"]]
    return pack_dense_recipe(recipe, ids.get, legs.index, 32, 33)


def _check_device_machine(eng, golden):
    from furniture_b200.dense import dense_config

    g = {k: golden[k] for k in golden.files}
    recipes = [json.loads(str(r)) for r in g["recipe_json"]]
    spos, smat, ppos, touch = _device_inputs(g, recipes)
    starts = np.flatnonzero(g["is_reset"])
    ends = np.append(starts[1:], len(g["reward"]))
    thr = [0.02, 0.99, 0.99, 0.0]
    groups = {}
    for e, (a, b) in enumerate(zip(starts, ends)):
        key = (int(g["ep_recipe"][e]),) + tuple(bool(g["ep_" + k][e]) for k in ("diff_rew", "early_termination", "phase_ob", "reset_robot_after_attach"))
        groups.setdefault(key, []).append((a, b - a))
    checked = 0
    for key, eps in groups.items():
        recipe = recipes[key[0]]
        cfg = golden_cfg(g, int(g["episode"][eps[0][0]]))
        dc = dense_config(**{k: v for k, v in cfg.items() if not k.startswith("alignment")})
        rc = _packed_recipe(recipe)
        first, count = [a for a, _ in eps], [c for _, c in eps]
        rew, done, info = eng.dense_eval(dc, rc, thr, len(recipe["recipe"]), first, count, spos, smat, ppos, touch, g["is_reset"], g["connected"], g["ac"])
        for a, c in eps:
            sl = slice(a, a + c)
            assert np.array_equal(info[sl, 0].astype(int), g["phase"][sl]), key
            assert np.array_equal(info[sl, 1].astype(int), g["subtask"][sl]), key
            assert np.array_equal(done[sl] & 1, g["done"][sl].astype(np.uint8)), key
            assert np.array_equal((done[sl] >> 1) & 1, g["success"][sl].astype(np.uint8)), key
            ref = g["reward"][sl]
            assert np.array_equal(np.isnan(ref), np.isnan(rew[sl])), key
            ok = ~np.isnan(ref)
            assert np.all(np.abs(rew[sl][ok] - ref[ok]) <= 1e-11 * np.maximum(1.0, np.abs(ref[ok]))), (key, np.max(np.abs(rew[sl][ok] - ref[ok])))
            # info columns shared with the golden: phase_bonus ... stable_grip_succ, then the two skip flags as bits
            gi = g["info"][sl]
            step = ~g["is_reset"][sl]
            for col, name in enumerate(D.INFO_KEYS[:9]):
                want, got = gi[step, col], info[sl][step, 2 + col]
                assert np.allclose(got, want, rtol=1e-11, atol=1e-11, equal_nan=True), (key, name)
            assert np.array_equal(info[sl][step, 11].astype(int), (gi[step, 9] + 2 * gi[step, 10]).astype(int)), key
            checked += c
    assert checked == len(g["reward"])


def test_device_reward_machine_matches_the_reference_emu(golden):
    from furniture_b200 import mjcf
    from parity_util import make_engine

    eng = make_engine(mjcf.load_scene("None", "table_lack_0825"), 1, False)
    _check_device_machine(eng, golden)


@pytest.mark.gpu
def test_device_reward_machine_matches_the_reference_cuda(golden):
    from furniture_b200 import mjcf
    from parity_util import make_engine

    eng = make_engine(mjcf.load_scene("None", "table_lack_0825"), 1, True)
    _check_device_machine(eng, golden)
