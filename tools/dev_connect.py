"""Developer script: run the connect scenario on emu and (if present) cuda and print per-part errors vs the CPU env."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from furniture_b200 import mjcf
from oracle.ref_env import OracleFurnitureEnv
from parity_util import *
from test_env_parity import _grasp_and_align_state
m = mjcf.load_scene("Sawyer", "table_lack_0825")
for gpu in ([False, True] if have_gpu() else [False]):
    env = OracleFurnitureEnv(m); env.reset()
    q = _grasp_and_align_state(m, env)
    env.nsub = 1
    env.sim.qvel[:] = 0; env.sim.qacc_warmstart[:] = 0; env.sim.ctrl[:] = 0; env.sim.forward()
    eng = make_engine(m, 2, gpu, nsub=1)
    eng.env_reset()
    eng.set("qpos", q); eng.set("qvel", np.zeros(m.nv)); eng.set("qacc_warmstart", np.zeros(m.nv)); eng.forward()
    a = np.zeros((2, eng.act_dim), np.float32); a[:, -2] = 1.0; a[0, -1] = 1.0; a[1, -1] = -1.0
    obs, rew, done, info = eng.env_step_host(a)
    ob, r, d, inf = env.step(a[0].astype(np.float64))
    qe = eng.get("qpos")[0]
    print("gpu" if gpu else "emu", "info", info[0], "oracle", inf, "rew", rew[0], r)
    print("  robot err", np.abs(qe[:9] - env.sim.qpos[:9]).max(), "parts err", [float(np.abs(qe[9+7*p:16+7*p] - env.sim.qpos[9+7*p:16+7*p]).max()) for p in range(5)])
    print("  qvel err", np.abs(eng.get("qvel")[0] - env.sim.qvel).max(), "|qvel|", np.abs(env.sim.qvel).max(), "ncon", eng.get("ncon")[0], env.sim.ncon, "flags", eng.get("flags")[0])
    print("  eq_data err", np.abs(eng.get("eq_data")[0] - env.sim.eq_data).max())
