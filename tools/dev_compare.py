"""Developer script: one forward pass, engine (emulated or CUDA) vs oracle, printed stage by stage."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from furniture_b200 import mjcf
from oracle.oracle import OracleSim
from parity_util import *

gpu = have_gpu()
m = mjcf.load_scene("Sawyer", "table_lack_0825")
eng = make_engine(m, 2, gpu)
em = eng.em
sim = OracleSim(m)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dz = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
q = settled_state(m, seed, dz=dz)
rng = np.random.RandomState(seed + 1)
v = rng.normal(size=m.nv) * 0.2
ctrl = rng.uniform(-1, 1, m.nu)
sim.qpos[:] = q; sim.qvel[:] = v; sim.ctrl[:] = ctrl
sim.forward()
eng.set("qpos", q); eng.set("qvel", v); eng.set("ctrl", ctrl)
eng.forward()
xp, xq, xm = oracle_link_poses(sim, em)
lp = eng.get("link_xpos")[0].reshape(-1, 3); lq = eng.get("link_xquat")[0].reshape(-1, 4)
print("link pos err", np.abs(lp - xp).max(), "quat err", max(quat_err(a, b) for a, b in zip(lq, xq)))
print("bias err", np.abs(eng.get("qfrc_bias")[0] - sim.qfrc_bias[:9]).max(), "|bias|", np.abs(sim.qfrc_bias[:9]).max())
Mr = eng.get("dbg_Mr")[0].reshape(9, 9); M = sim.qM.reshape(m.nv, m.nv)
print("Mr err", np.abs(Mr - M[:9, :9]).max(), "|M|", np.abs(M[:9, :9]).max())
fs = eng.get("dbg_fs")[0]; as_ = eng.get("dbg_as")[0]
print("fs robot err", np.abs(fs[:9] - sim.qfrc_smooth[:9]).max(), "as robot err", np.abs(as_[:9] - sim.qacc_smooth[:9]).max(), "|as|", np.abs(sim.qacc_smooth[:9]).max())
zs = to_z(m, em, xm, sim.qacc_smooth)
print("as parts err", np.abs(as_[9:] - zs[9:]).max(), "|as parts|", np.abs(zs[9:]).max())
print("ncon", eng.get("ncon")[0], sim.ncon, "niter", eng.get("niter")[0], sim.scalar("solver_niter"), "flags", eng.get("flags")[0])
nc = int(eng.get("ncon")[0][0])
cd = eng.get("con_dist")[0][:nc]; cp = eng.get("con_pos")[0].reshape(-1, 3)[:nc]; cf = eng.get("con_frame")[0].reshape(-1, 9)[:nc]
oc = sim.contacts()
if nc == len(oc):
    print("contact dist err", max(abs(cd[i] - oc[i].dist) for i in range(nc)) if nc else 0,
          "pos err", max(np.abs(cp[i] - np.array(list(oc[i].pos))).max() for i in range(nc)) if nc else 0,
          "frame err", max(np.abs(cf[i] - np.array(list(oc[i].frame))).max() for i in range(nc)) if nc else 0)
    ar = eng.get("con_aref")[0].reshape(-1, 3)[:nc]
    oar = np.array([sim.efc_aref[c.efc_address:c.efc_address + 3] for c in oc])
    print("aref err", np.abs(ar - oar).max() if nc else 0, "|aref|", np.abs(oar).max() if nc else 0)
    D = eng.get("con_D")[0].reshape(-1, 2)[:nc]
    oD = np.array([sim.efc_D[c.efc_address:c.efc_address + 2] for c in oc])
    print("D rel err", np.abs(D / oD - 1).max() if nc else 0)
    fo = np.array([sim.efc_force[c.efc_address:c.efc_address + 3] for c in oc]); fe = eng.get("con_force")[0].reshape(-1, 3)[:nc]
    print("force err", np.abs(fe - fo).max() if nc else 0, "|f|", np.abs(fo).max() if nc else 0)
x = eng.get("dbg_x")[0]
zo = to_z(m, em, xm, sim.qacc)
print("qacc err", np.abs(x - zo).max(), "rel", np.abs(x - zo).max() / np.abs(zo).max(), "|qacc|", np.abs(zo).max())
