"""Developer script: step engine and oracle side by side, report divergence."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, time
from furniture_b200 import mjcf
from oracle.oracle import OracleSim
from parity_util import *

gpu = have_gpu()
m = mjcf.load_scene("Sawyer", "table_lack_0825")
eng = make_engine(m, 1, gpu)
sim = OracleSim(m)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
q = settled_state(m, seed, robot_noise=0.0)
rng = np.random.RandomState(seed)
sim.qpos[:] = q; eng.set("qpos", q)
sim.forward(); eng.forward()
t0 = time.time()
for es in range(nsteps // 50):
    ctrl = rng.uniform(-1, 1, m.nu) * (m.actuator_ctrlrange[:, 1])
    sim.ctrl[:] = ctrl; eng.set("ctrl", ctrl)
    sim.qfrc_applied[:9] = sim.qfrc_bias[:9]
    eng.set("qfrc_applied", eng.get("qfrc_bias"))
    for k in range(5):
        sim.step(10); eng.step(10)
        qe = eng.get("qpos")[0]; ve = eng.get("qvel")[0]
        print("step %4d  qpos err %.2e  qvel err %.2e  |qvel| %.2f  ncon %d/%d niter %d/%d flags %d" % (
            es * 50 + (k + 1) * 10, np.abs(qe - sim.qpos).max(), np.abs(ve - sim.qvel).max(), np.abs(sim.qvel).max(),
            eng.get("ncon")[0][0], sim.ncon, eng.get("niter")[0][0], sim.scalar("solver_niter"), eng.get("flags")[0][0]))
print("time", time.time() - t0)
