"""Physical facts read off the reference's own demo recording (demos/Sawyer_7.pkl: states logged from a real MuJoCo run of
FurnitureSawyerEnv + swivel_chair_0700; legacy format, qpos only): the height at which the chair base rests on the floor and
the small tilt it rests with (Cursor_7.pkl, same furniture).
That number is an equilibrium of the soft-contact model (geom masses from density, gravity, cylinder-plane contacts, solref /
solimp impedance), so it pins those parts of the physics oracle against MuJoCo itself.  The assembled relative poses at the
end of the recording are NOT usable: they match neither the current XML sites nor its weld data (older model version).
Needs /root/reference; writes tests/golden/demo_facts.json."""
import io
import json
import os
import pickle

import numpy as np


class _DataOnlyUnpickler(pickle.Unpickler):
    """the demo files come from the untrusted reference tree: only numpy array reconstruction and builtin containers may be
    instantiated; any other global in the stream (the hook arbitrary-code pickles rely on) is refused"""

    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
                ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("collections", "OrderedDict"),
                ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "frozenset")}

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED:
            raise pickle.UnpicklingError("refusing to load %s.%s from a demo file" % (module, name))
        return super().find_class(module, name)


def load_data_only(path):
    with open(path, "rb") as f:
        return _DataOnlyUnpickler(io.BytesIO(f.read())).load()

REF = "/root/reference/demos/Sawyer_7.pkl"
REF2 = "/root/reference/demos/Cursor_7.pkl"  # same furniture driven by the cursor agent; the base starts with no yaw
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "demo_facts.json")


def main():
    q = load_data_only(REF)["qpos"]
    z = np.array([s["1_chair_base"][2] for s in q])
    rest = z[:10]  # the base has not been touched yet in the first steps
    assert rest.std() < 1e-7
    facts = {"source": "demos/Sawyer_7.pkl of the reference, key 1_chair_base, first 10 recorded states (tools/make_golden_demo_facts.py)",
             "swivel_chair_base_rest_z": float(rest.mean()), "swivel_chair_base_rest_z_std": float(rest.std()), "n_states": int(len(q))}
    q2 = load_data_only(REF2)["qpos"]
    base = np.array([s["1_chair_base"] for s in q2[:10]])
    assert base.std(0).max() < 1e-7 and abs(base[0, 2] - rest.mean()) < 1e-7  # same rest height in both recordings
    # at rest the base leans by 0.028 degrees about its x axis (its five cylinders are not arranged symmetrically about the
    # centre of mass): quaternion x component; the z component (yaw) is where it happened to be put
    facts["swivel_chair_base_rest_quat"] = [float(v) for v in base[:, 3:].mean(0)]
    # the whole first state of the Cursor recording: all three parts standing untouched (base flat, column and seat upright)
    facts["cursor7_rest_state"] = {k: [float(v) for v in np.mean([s[k] for s in q2[:10]], axis=0)] for k in ("1_chair_base", "2_chair_column", "3_chair_seat")}
    # demos/Baxter_0.pkl: the two boxes of the `block` furniture at rest (solref 0.001 < 2 h: the refsafe clamp is in play)
    q3 = load_data_only("/root/reference/demos/Baxter_0.pkl")["qpos"]
    blocks = {k: np.array([s[k] for s in q3[:10]]) for k in ("1_block_l", "2_block_r")}
    assert all(v[:, 2:].std(0).max() < 1e-7 for v in blocks.values())  # height and orientation are still; x / y creep by 2e-7 per recorded step
    facts["baxter0_rest_state"] = {k: [float(x) for x in v[0]] for k, v in blocks.items()}
    json.dump(facts, open(OUT, "w"), indent=1)
    print(facts)


if __name__ == "__main__":
    main()
