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


# ------------------------------------------------------------------ the dense env: device step against the CPU env
BACKENDS = [pytest.param(False, id="emu"), pytest.param(True, id="cuda", marks=pytest.mark.gpu)]
DENSE_BASE = dict(max_episode_steps=150, auto_align=0, alignment_pos_dist=0.02, alignment_rot_dist_up=0.99, alignment_rot_dist_forward=0.99, alignment_project_dist=0.0)


def _dense_engine(m, n, gpu, dense_kw=None, **cfg):
    from furniture_b200.dense import dense_config
    from furniture_b200.engine import Engine, default_config
    from parity_util import build_emu

    c = default_config(**dict(DENSE_BASE, **cfg))
    dc = dense_config(**(dense_kw or {}))
    return Engine(m, n, device=0, config=c, dense=dc) if gpu else Engine(m, n, config=c, lib_path=build_emu(), dense=dc)


def _dense_state(eng, i):
    """per-env FeDenseState as a dict (8 int32, then 12 + 11 doubles)"""
    raw = eng.get("dense_state")[i].tobytes()
    ints = np.frombuffer(raw[:32], np.int32)
    dbl = np.frombuffer(raw[32:], np.float64)
    return dict(phase=int(ints[0]), subtask=int(ints[1]), dropped=int(ints[2]), table_moved=int(ints[3]), lifted=int(ints[4]), fine_aligned=int(ints[5]),
                success=int(ints[6]), table_site0=dbl[0:3], leg0=dbl[3:6], lift_target=dbl[6:9], init_eef=dbl[9:12], prev=dbl[12:23])


PREV_KEYS = ["init_eef", "above_leg", "eef_leg", "grasp", "lift_z", "lift_xy", "move_pos", "move_up", "move_fwd", "proj_t", "proj_l"]


def _adopt_device_anchors(orc, ds):
    """after a reset the device took its anchors from the kinematics of the last mj_step's forward pass (as the reference does), the
    synced oracle from a fresh forward pass one integration later; continue from the device's numbers so that the steps compare tightly"""
    orc.table_site0, orc.leg0, orc.lift_target, orc.init_eef = ds["table_site0"].copy(), ds["leg0"].copy(), ds["lift_target"].copy(), ds["init_eef"].copy()
    for k, name in enumerate(PREV_KEYS):
        if name in orc.prev:
            orc.prev[name] = float(ds["prev"][k])


def _assert_same_machine(ds, orc, where, tol=1e-4):
    assert ds["phase"] == orc.phase and ds["subtask"] == orc.subtask, (where, ds["phase"], orc.phase, ds["subtask"], orc.subtask)
    assert (ds["dropped"], ds["table_moved"], ds["lifted"], ds["fine_aligned"]) == (int(orc.dropped), int(orc.table_moved), int(orc.lifted), orc.fine_aligned), where
    if orc.subtask < len(orc.sub):
        assert np.abs(ds["table_site0"] - orc.table_site0).max() < tol and np.abs(ds["leg0"] - orc.leg0).max() < tol, where
        assert np.abs(ds["lift_target"] - orc.lift_target).max() < tol, where
    for k, name in enumerate(PREV_KEYS):
        if name in orc.prev:
            assert abs(ds["prev"][k] - orc.prev[name]) < tol, (where, name, ds["prev"][k], orc.prev[name])


@pytest.mark.parametrize("gpu", BACKENDS)
@pytest.mark.parametrize("dense_kw", [dict(), dict(diff_rew=False, phase_ob=True)], ids=["diff", "plain-phase_ob"])
def test_dense_env_steps_match_the_cpu_env(gpu, dense_kw):
    """reset + 3 env steps of the dense-reward Sawyer env: reward, done, phase machine state and the phase observation of the device
    equal the CPU env (oracle physics + oracle/dense_oracle.py) started from the same post-reset state"""
    from furniture_b200 import mjcf
    from oracle.ref_env import DenseCfg, OracleDenseEnv
    from test_env_parity import _sync_oracle_from_engine

    m = mjcf.load_scene("Sawyer", "table_lack_0825")
    n = 2
    eng = _dense_engine(m, n, gpu, dense_kw)
    assert eng.obs_dim == 35 + 29 + (8 if dense_kw.get("phase_ob") else 0)
    eng.env_reset()
    envs = [OracleDenseEnv(m, DenseCfg(), dense_kw) for _ in range(n)]
    for i, e in enumerate(envs):
        e.reset()
        _sync_oracle_from_engine(e, eng, i)
        e.sim.qfrc_bias[: e.nr] = eng.get("qfrc_bias")[i]
        e.dense.begin_episode()  # _reset_reward_variables on the state the device reset produced
        _assert_same_machine(_dense_state(eng, i), e.dense, ("reset", i), tol=5e-4)
        _adopt_device_anchors(e.dense, _dense_state(eng, i))
        assert e.dense.phase == 1  # table_lack's first subtask has no grip_init_pos: it starts at move_eef_above_leg
    rng = np.random.RandomState(11)
    for k in range(3):
        a = rng.uniform(-1, 1, (n, eng.act_dim)).astype(np.float32)
        a[:, -1] = -0.5
        obs, rew, done, info = eng.env_step_host(a)
        dinfo = eng.get("dense_info")
        for i, e in enumerate(envs):
            ob, r, d, inf = e.step(a[i].astype(np.float64))
            assert np.abs(obs[i] - ob).max() < 2e-4, (k, i)
            assert abs(rew[i] - r) < 2e-2 + 1e-5 * abs(r), (k, i, rew[i], r)
            assert bool(done[i]) == d and info[i][1] == int(e.dense.success)
            _assert_same_machine(_dense_state(eng, i), e.dense, (k, i))
            assert int(dinfo[i][0]) == e.dense.phase and abs(dinfo[i][3] - e.dense_info["ctrl_penalty"]) < 1e-6
            if dense_kw.get("phase_ob"):
                assert np.array_equal(obs[i][-8:], np.eye(8, dtype=np.float32)[e.dense.phase])


@pytest.mark.parametrize("gpu", BACKENDS)
def test_dense_env_grasp_connect_and_next_subtask(gpu):
    """the recipe's first leg held between the fingers at its connector: the machine jumps to lift_leg (safe grasp), a connect action
    attaches it (correct connection outside move_leg_fine: 2 x phase_bonus), and the next subtask starts from the new world --
    device and CPU env agree on every integer and on the reward"""
    from furniture_b200 import mjcf
    from oracle.ref_env import DenseCfg, OracleDenseEnv
    from test_env_parity import _grasp_and_align_state

    m = mjcf.load_scene("Sawyer", "table_lack_0825")
    cfg = DenseCfg()
    cfg.auto_align = True  # snap on connect, so that the connection is a correct one for the reward machine as well
    env = OracleDenseEnv(m, cfg)
    env.reset()
    q = _grasp_and_align_state(m, env, leg=1, leg_site="leg-table,0,90,180,270,conn_site2", table_site="table-leg,0,90,180,270,conn_site2",
                               arm_qpos=[-0.28, -0.9, 0.0, 1.86, 0.0, 0.6, 1.57])  # arm raised: the table top hangs clear of the floor
    env.nsub = 1
    env.sim.qvel[:] = 0; env.sim.qacc_warmstart[:] = 0; env.sim.ctrl[:] = 0
    env.sim.forward()
    env.dense.begin_episode()
    eng = _dense_engine(m, 2, gpu, nsub=1, auto_align=1)
    eng.env_reset()
    eng.set("qpos", q); eng.set("qvel", np.zeros(m.nv)); eng.set("qacc_warmstart", np.zeros(m.nv))
    eng.forward()
    # the device machine starts its episode inside fe_env_reset; restart it on the hand-made state through the state field
    st = eng.get("dense_state").copy()
    raw = bytearray(st[0].tobytes())
    ints = np.frombuffer(raw[:32], np.int32).copy()
    dbl = np.frombuffer(raw[32:], np.float64).copy()
    o = env.dense
    ints[:7] = [o.phase, o.subtask, 0, 0, 0, 0, 0]
    dbl[0:3], dbl[3:6], dbl[6:9] = o.table_site0, o.leg0, o.lift_target
    for k, name in enumerate(PREV_KEYS):
        dbl[12 + k] = o.prev.get(name, 0.0)
    row = np.frombuffer(ints.tobytes() + dbl.tobytes(), np.uint8)
    eng.set("dense_state", np.stack([row, row]).view(st.dtype).reshape(st.shape))
    a = np.zeros((2, eng.act_dim), np.float32)
    a[:, -2] = 1.0
    a[0, -1], a[1, -1] = -1.0, -1.0
    obs, rew, done, info = eng.env_step_host(a)          # step 1: no connect asked; the held leg is noticed
    ob, r, d, inf = env.step(a[0].astype(np.float64))
    assert env.dense_info["skip_to_lift_leg"] == 1 and env.dense.phase >= 4, "the CPU env did not see a safe grasp: test state is wrong"
    assert abs(rew[0] - r) < 0.5 + 1e-4 * abs(r), (rew[0], r)
    _assert_same_machine(_dense_state(eng, 0), env.dense, "grasp")
    a[0, -1] = 1.0                                       # step 2: env 0 connects, env 1 does not
    obs, rew, done, info = eng.env_step_host(a)
    ob, r, d, inf = env.step(a[0].astype(np.float64))
    assert inf["num_connected"] == 1 and env.dense.subtask == 1, "the CPU env did not attach the leg: test state is wrong"
    assert info[0][0] == 1 and info[1][0] == 0
    assert abs(rew[0] - r) < 0.5 + 1e-4 * abs(r) and rew[0] > 9000 and rew[1] < 5000, (rew, r)
    assert bool(done[0]) == d and not d
    _assert_same_machine(_dense_state(eng, 0), env.dense, "attached")
    s0, s1 = _dense_state(eng, 0), _dense_state(eng, 1)
    assert (s0["subtask"], s0["phase"]) == (1, 0) and (s1["subtask"], s1["phase"]) == (0, 4)  # next leg from init_eef; env 1 still lifting
    assert np.abs(s0["init_eef"] - env.dense.init_eef).max() < 1e-4
