"""FurnitureCursorEnv (BASELINE.json config 1: Cursor + toy_table, one env; furniture/env/furniture_cursor.py + the Cursor
branches of furniture.py).  The host logic of furniture_b200/cursor_env.py is run twice from the same seed and actions: over the
engine (lane-emulated build here, the sm_100a library when marked gpu) and over the fp64 CPU oracle through the same backend
interface.  Decisions (selection, rollbacks, connect steps, welds) must be identical, poses equal to fp32 tolerance."""
import numpy as np
import pytest

from furniture_b200 import mjcf
from furniture_b200.cursor_env import EngineBackend, FurnitureCursorEnvB200
from oracle import assembly_oracle as A
from oracle.oracle import OracleSim
from parity_util import build_emu

BACKENDS = [pytest.param(False, id="emu"), pytest.param(True, id="cuda", marks=pytest.mark.gpu)]


class OracleBackend:
    """the backend interface of cursor_env.EngineBackend over the CPU physics oracle (test infrastructure)"""

    def __init__(self, model):
        self.model, self.sim = model, OracleSim(model)
        self.part_body = [model.names["body"].index(n) for n in model.meta["part_names"]]
        self.cursor_body = [model.names["body"].index("cursor0"), model.names["body"].index("cursor1")]
        self.body_pos = np.array(model.body_pos, dtype=np.float64).copy()

    def reset_data(self): self.sim.reset()
    def forward(self): self.sim.forward()
    def step(self, n=1): self.sim.step(n)
    def qpos(self): return self.sim.qpos.copy()
    def set_qpos(self, q): self.sim.qpos[:] = q
    def qvel(self): return self.sim.qvel.copy()
    def set_qvel(self, v): self.sim.qvel[:] = v
    def zero_warmstart(self): self.sim.qacc_warmstart[:] = 0

    def set_gravcomp(self, f):
        for p, b in enumerate(self.part_body):
            self.sim.xfrc_applied[6 * b : 6 * b + 6] = [0, 0, -f[p] * self.model.opt_gravity[2] * self.model.body_mass[b], 0, 0, 0]

    def part_poses(self):
        return np.array([self.sim.xpos[3 * b : 3 * b + 3] for b in self.part_body]), np.array([self.sim.xquat[4 * b : 4 * b + 4] for b in self.part_body])

    def cursor_pos(self, i): return self.sim.xpos[3 * self.cursor_body[i] : 3 * self.cursor_body[i] + 3].copy()

    def set_cursor_pos(self, i, pos):
        self.body_pos[self.cursor_body[i]] = pos
        self.sim.set_model("body_pos", self.body_pos)

    def touch_bits(self):  # on_collision(cursor_i, part): a contact between the cursor geom and a geom of the part's body
        bits = np.zeros(len(self.part_body), dtype=np.int32)
        names = self.model.names["geom"]
        for c in self.sim.contacts():
            for ga, gb in ((c.geom1, c.geom2), (c.geom2, c.geom1)):
                b = int(self.model.geom_bodyid[gb])
                if names[ga] in ("cursor0", "cursor1") and b in self.part_body:
                    bits[self.part_body.index(b)] |= 1 << int(names[ga][-1])
        return bits

    def geom_masks(self): return self.sim.geom_contype.copy(), self.sim.geom_conaffinity.copy()
    def set_geom_masks(self, ct, ca): self.sim.geom_contype[:] = ct; self.sim.geom_conaffinity[:] = ca
    def engine_geom(self, g): return g
    def eq(self): return self.sim.eq_active.copy(), self.sim.eq_data.reshape(-1, 7).copy()
    def set_eq(self, act, data): self.sim.eq_active[:] = act; self.sim.eq_data[:] = np.ravel(data)

    def is_aligned(self, p1, m1, p2, m2, angles, thr):
        ok, tq = A.is_aligned(p1, np.ravel(m1), p2, np.ravel(m2), angles, thr)
        return bool(ok), tq


def _pair(gpu, seed=11):
    m = mjcf.load_scene("Cursor", "toy_table")
    dev = FurnitureCursorEnvB200(backend=EngineBackend(m, lib_path=None if gpu else build_emu()), seed=seed)
    ref = FurnitureCursorEnvB200(backend=OracleBackend(m), seed=seed)
    return m, dev, ref


def test_cursor_scene_dimensions():
    m = mjcf.load_scene("Cursor", "toy_table")
    assert (m.nq, m.nv, m.nu, m.neq) == (35, 30, 0, 4)
    g = [m.names["geom"].index(n) for n in ("cursor0", "cursor1")]
    assert np.allclose(m.geom_margin[g], 0.05) and np.allclose(m.geom_gap[g], 10) and np.allclose(m.geom_size[g], 0.05)


@pytest.mark.parametrize("gpu", BACKENDS)
def test_cursor_reset_and_random_steps_agree_with_the_cpu_env(gpu):
    m, dev, ref = _pair(gpu)
    ob_d, ob_r = dev.reset(), ref.reset()
    assert ob_d["object_ob"].shape == (35,) and ob_d["robot_ob"].shape == (8,)  # furniture_cursor.py:40-43
    assert np.abs(ob_d["object_ob"] - ob_r["object_ob"]).max() < 1e-4 and np.array_equal(ob_d["robot_ob"], ob_r["robot_ob"])
    # script: cursor 0 walks to the first leg and selects it, lifts it clear of the others, carries and turns it; cursor 1 idles.
    # (Carrying a part *through* others teleports it into deep penetration, 0.1 m per action: the response is chaotic and not
    # comparable between fp32 and fp64, so the script keeps the carried part in free space.)
    leg = 0
    script, selected_any = [], False
    for k in range(14):
        a = np.zeros(15)
        a[6] = 1.0
        if dev.cursor_selected[0] is None:
            d = ob_r["object_ob"][7 * leg : 7 * leg + 3] - ob_r["robot_ob"][0:3]
            a[0:2] = np.clip(np.round(d[:2] / 0.1), -1, 1)
        else:
            n = sum(1 for s in script if s)  # steps since the selection
            a[0:6] = [[0, 0, 1, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 1, 0, 0, 0], [-1, 0, 0, 0, 0, 0], [-1, 0, 0, 0, 0, 1], [0, 1, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0]][min(n, 6)]
        od, rd, dd, _ = dev.step(a)
        orf, rr, dr, _ = ref.step(a)
        script.append(ref.cursor_selected[0] is not None)
        assert dev.cursor_selected == ref.cursor_selected, (k, dev.cursor_selected, ref.cursor_selected)
        assert np.array_equal(od["robot_ob"], orf["robot_ob"]) and rd == rr and dd == dr
        assert np.abs(od["object_ob"] - orf["object_ob"]).max() < 1e-3, (k, np.abs(od["object_ob"] - orf["object_ob"]).max())
        selected_any |= ref.cursor_selected[0] is not None
        ob_r = orf
    assert ob_r["object_ob"][7 * dev.cursor_selected[0] + 2] > 0.25  # the selected part hangs in the air under gravity compensation
    assert selected_any  # the sensor contacts of the cursors select parts in both envs
    dev.close()


@pytest.mark.parametrize("gpu", BACKENDS)
def test_cursor_connect_takes_ten_approach_steps_then_welds(gpu):
    """two selected parts held close to alignment: ten connect actions interpolate part 2 towards the target (furniture.py:993-1034),
    the eleventh welds them (_connect): num_connected, weld activation and the released cursor are identical on both backends"""
    m, dev, ref = _pair(gpu, seed=5)
    dev.reset(); ref.reset()
    s_leg = m.names["site"].index([n for n in m.names["site"] if n.startswith("leg-top") and "conn_site" in n][0])
    partner = m.names["site"][s_leg].split(",")[0].split("-")[::-1]
    s_top = [s for s, n in enumerate(m.names["site"]) if "conn_site" in n and n.split(",")[0].split("-") == partner][0]
    for env in (dev, ref):
        leg, top = env.part_body.index(int(m.site_bodyid[s_leg])), env.part_body.index(int(m.site_bodyid[s_top]))
        # lift the table top, hold the leg under its connector a few cm away and a few degrees off
        env._set_q(top, [0.0, 0.0, 0.4], [1, 0, 0, 0])
        env.sim.forward()
        xpos, xquat = env.sim.part_poses()
        p_top, _, q_top = env._site_pose(s_top, xpos, xquat)
        tq = mjcf.q_mul(mjcf.q_axis_angle([0, 1, 0], 0.05), q_top)  # _is_aligned wants the two connectors' up vectors parallel (furniture.py:1077-1081)
        leg_q = mjcf.q_mul(tq, mjcf.q_conj(m.site_quat[s_leg]))
        leg_p = p_top + np.array([0.01, -0.01, -0.03]) - mjcf.q_to_mat(leg_q) @ m.site_pos[s_leg]
        env._set_q(leg, leg_p, leg_q)
        env._stop(range(env.npart), 1)
        env.sim.set_cursor_pos(0, env._q(leg)[:3]); env.sim.set_cursor_pos(1, env._q(top)[:3])
        env.sim.forward()
        env.cursor_selected = [leg, top]
    a = np.zeros(15); a[6] = a[13] = 1.0; a[14] = 1.0
    counts = []
    for k in range(12):
        od, rd, dd, _ = dev.step(a)
        orf, rr, dr, _ = ref.step(a)
        assert dev.connect_step == ref.connect_step and dev.num_connected == ref.num_connected and dev.cursor_selected == ref.cursor_selected, k
        assert rd == rr and np.abs(od["object_ob"] - orf["object_ob"]).max() < 5e-2  # the approach drives the two parts into contact: fp32 / fp64 responses drift (2e-2 seen on CUDA); the decisions above are what is pinned
        counts.append(ref.num_connected)
        if ref.num_connected:
            break
    assert counts[-1] == 1 and len(counts) == 11 and ref.cursor_selected[1] is None  # ten approach steps, then the weld; cursor 1 released
    assert list(dev.sim.eq()[0]) == list(ref.sim.eq()[0]) and sum(ref.sim.eq()[0]) == 1
    assert np.abs(dev.sim.eq()[1] - ref.sim.eq()[1]).max() < 5e-2  # the welded relative pose inherits the drift of the approach
    dev.close()
