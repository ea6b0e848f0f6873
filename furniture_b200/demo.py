"""Demonstration recording in the reference's file format (furniture/util/demo_recorder.py:58-87) for a batch of envs.

The reference's DemoRecorder collects, for ONE env, `states` (get_env_state dicts: qpos, qvel), `obs`, `actions`, `rewards`,
`low_level_obs`, `low_level_actions` (with the connect action appended at save time) and `connect_actions`, and pickles them
under `prefix + "%04d.pkl"`.  `BatchDemoRecorder` keeps one such record per env of a `BatchedFurnitureEnv` and writes one file
per env with the same keys, so that the reference's loaders (`--load_demo`, `--load_init_states`: furniture.py:121-130) and
its demo tools read them unchanged.  Observations are stored as the reference stores them: an OrderedDict of float64 arrays.
"""
from __future__ import annotations

import glob
import os
import pickle
from collections import OrderedDict

import numpy as np


class BatchDemoRecorder:
    def __init__(self, env, demo_dir="./", metadata=None):
        self.env, self.demo_dir, self.metadata = env, demo_dir, metadata
        os.makedirs(demo_dir, exist_ok=True)
        self.reset()

    def reset(self):
        n = self.env.num_envs
        self._rec = [dict(states=[], obs=[], actions=[], rewards=[], low_level_obs=[], low_level_actions=[], connect_actions=[]) for _ in range(n)]

    def _ob(self, obs_dict, i):
        return OrderedDict((k, np.asarray(v[i].detach().cpu().numpy() if hasattr(v, "detach") else v[i], dtype=np.float64)) for k, v in obs_dict.items())

    def add_reset(self, obs_dict):
        """after env.reset(): first observation, first low-level observation and the state (furniture.py:1613-1614, :324-330)"""
        st = self.env.get_env_state()
        for i, r in enumerate(self._rec):
            ob = self._ob(obs_dict, i)
            r["obs"].append(ob)
            r["low_level_obs"].append(ob)
            r["states"].append({"qpos": st["qpos"][i].astype(np.float64), "qvel": st["qvel"][i].astype(np.float64)})

    def add_step(self, actions, obs_dict, rewards):
        """after env.step(actions): with control_type="impedance" every env step is also one low-level step (furniture.py:1283-1289)"""
        a = np.asarray(actions.detach().cpu().numpy() if hasattr(actions, "detach") else actions, dtype=np.float64)
        rw = np.asarray(rewards.detach().cpu().numpy() if hasattr(rewards, "detach") else rewards, dtype=np.float64)
        st = self.env.get_env_state()
        for i, r in enumerate(self._rec):
            ob = self._ob(obs_dict, i)
            r["actions"].append(a[i].copy())
            r["rewards"].append(float(rw[i]))
            r["obs"].append(ob)
            r["low_level_obs"].append(ob)
            r["low_level_actions"].append(a[i, :-1].copy())
            r["connect_actions"].append(float(a[i, -1]))
            r["states"].append({"qpos": st["qpos"][i].astype(np.float64), "qvel": st["qvel"][i].astype(np.float64)})

    def save(self, prefix, envs=None):
        """one file per env, numbered like the reference (count = files already carrying the prefix); returns the paths"""
        paths = []
        for i in (range(len(self._rec)) if envs is None else envs):
            r = self._rec[i]
            lla = [np.concatenate([x, [c]]) for x, c in zip(r["low_level_actions"], r["connect_actions"])]  # demo_recorder.py:73-77
            assert len(r["low_level_obs"]) == len(lla) + 1 and len(r["obs"]) == len(r["actions"]) + 1
            demo = {"states": r["states"], "obs": r["obs"], "actions": r["actions"], "rewards": r["rewards"], "low_level_obs": r["low_level_obs"],
                    "low_level_actions": lla, "connect_actions": r["connect_actions"], "metadata": self.metadata}
            count = min(9999, len(glob.glob(os.path.join(self.demo_dir, prefix) + "*")))
            path = os.path.join(self.demo_dir, prefix + "{:04d}.pkl".format(count))
            with open(path, "wb") as f:
                pickle.dump(demo, f)
            paths.append(path)
        self.reset()
        return paths


def load_init_states(path):
    """states usable with `BatchedFurnitureEnv.set_env_state` from a demo file or an init-states pickle of the reference
    (a list of get_env_state dicts, furniture.py:127-130): returns the list of {"qpos", "qvel"} dicts"""
    import io

    class _DataOnly(pickle.Unpickler):
        _OK = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
               ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("collections", "OrderedDict")}

        def find_class(self, module, name):
            if (module, name) not in self._OK:
                raise pickle.UnpicklingError("refusing to load %s.%s" % (module, name))
            return super().find_class(module, name)

    with open(path, "rb") as f:
        d = _DataOnly(io.BytesIO(f.read())).load()
    states = d["states"] if isinstance(d, dict) else d
    return [{"qpos": np.asarray(s["qpos"]), "qvel": np.asarray(s["qvel"])} for s in states]
