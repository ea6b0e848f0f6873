"""Golden vectors for the reset placement, produced by running the REFERENCE'S OWN sampler, unmodified:
furniture/env/models/tasks/placement_sampler.py UniformRandomSampler.setup / sample (:68-190) with numpy RandomState(seed),
interleaved with the robot-noise draws of FurnitureEnv._reset (`_init_random(shape, "agent")` = rng.uniform(-r, r, size=7),
furniture.py:336-349, called 1 + 100 times per reset, :1581 and :1609).  Needs /root/reference (build container); writes
tests/golden/placement.npz.  The MujocoObject the sampler queries is replaced by a two-method fake fed from the same XML
numbers (`*_initpos`, `*_horizontal_radius_site`); pyquaternion by the stand-in of make_golden_assembly.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import make_golden_assembly as G  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def main():
    G.import_reference()
    from furniture.env.models.tasks.placement_sampler import UniformRandomSampler
    from furniture.util import Qpos

    from furniture_b200 import mjcf

    rec = {}
    for furn in ("table_lack_0825", "swivel_chair_0700"):
        xml, meta = mjcf.compose_scene("Sawyer", furn)
        names = list(meta["part_names"])

        class FakeObject:
            def get_horizontal_radius(self, name):
                return meta["part_radius"][name]

        objs = __import__("collections").OrderedDict((n, FakeObject()) for n in names)
        seeds = [123, 124, 500, 2026]
        nreset = 3
        pos = np.zeros((len(seeds), nreset, len(names), 3)); quat = np.zeros((len(seeds), nreset, len(names), 4))
        noise = np.zeros((len(seeds), nreset, 7)); mt = np.zeros((len(seeds), 624), np.uint32); mtpos = np.zeros(len(seeds), np.int64)
        for si, seed in enumerate(seeds):
            rng = np.random.RandomState(seed)  # furniture.py:72
            init = {n: Qpos(q[0], q[1], q[2], G.Quaternion(q[3], q[4], q[5], q[6])) for n, q in meta["part_init_qpos"].items()}
            s = UniformRandomSampler(rng, r_xyz=0.02, r_rot=3, init_qpos=init)  # floor_task.py:33-34, config/furniture.py:177-194
            s.setup(objs, (0, 0, 0), (0.7, 0.7, 0))                              # floor_task.py:37
            for r in range(nreset):
                p, q = s.sample(placed_objects_orig=[])                          # _place_objects -> place_objects, furniture.py:1404
                for k, n in enumerate(names):
                    pos[si, r, k] = p[n]; quat[si, r, k] = np.asarray(list(q[n]), dtype=np.float64)
                for _ in range(101):                                             # _initialize_robot_pos, furniture.py:1581, :1609
                    noise[si, r] = rng.uniform(low=-0.001, high=0.001, size=(7,))
            st = rng.get_state()
            mt[si], mtpos[si] = st[1], st[2]
        rec[furn] = dict(seeds=np.array(seeds), pos=pos, quat=quat, noise=noise, mt=mt, mtpos=mtpos)
        print(furn, "parts", names, "first placement", pos[0, 0, 0], quat[0, 0, 0])
    os.makedirs(OUT, exist_ok=True)
    flat = {"%s/%s" % (f, k): v for f, d in rec.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "placement.npz"), **flat,
                        source="reference UniformRandomSampler.sample run unmodified with numpy RandomState (tools/make_golden_placement.py)")


if __name__ == "__main__":
    main()
