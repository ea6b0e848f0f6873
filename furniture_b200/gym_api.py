"""Single-environment gym-shaped surface (numpy in / numpy out) over the engine: the call surface of the reference's
`FurnitureGym` (furniture/env/furniture_gym.py:11-80) and of `FurnitureEnv.step()/reset()` behind it.

  env = FurnitureGymB200(name="FurnitureSawyerEnv", furniture_name="table_lack_0825", seed=123)
  ob = env.reset()                                  # OrderedDict(object_ob (7 n_objects,), robot_ob (29,))
  ob, reward, done, info = env.step(action)         # action: (dof,) array or {"default": array}

It is the batched engine with one env (the batch entry point `furniture_b200.env.make_vec_env` is the one to use for
throughput).  Differences from the reference, all consequences of the step kernel resetting a finished env on the
device like a `SubprocVecEnv` worker does (subproc_vec_env.py:16-20):
  * the observation returned together with `done=True` is the first observation of the next episode, not the terminal one;
  * the `reset()` the caller issues after `done` returns that same observation and does not reset a second time, so the
    env consumes exactly one reset's worth of numpy random draws per episode, like the reference (furniture.py:72).
`info` carries the reference's episode keys at the end of an episode (`_after_step`, furniture.py:451-480).
"""
from __future__ import annotations

import time
from collections import OrderedDict

import numpy as np

from . import mjcf
from .engine import Engine, default_config

AGENTS = {"FurnitureSawyerDenseRewardEnv": "Sawyer", "IKEASawyerDense-v0": "Sawyer", "furniture-sawyer-densereward-v0": "Sawyer", "FurnitureSawyerEnv": "Sawyer", "IKEASawyer-v0": "Sawyer", "Sawyer": "Sawyer", "FurnitureBaxterEnv": "Baxter", "IKEABaxter-v0": "Baxter", "Baxter": "Baxter"}


class FurnitureGymB200:
    metadata = {"render.modes": []}

    def __init__(self, name="FurnitureSawyerEnv", furniture_name=None, device=0, lib_path=None, id=None, **config):
        """`id`, `name` and the remaining keywords are what gym passes from the registration (env/__init__.py:19-114: id, name,
        furniture_name / furniture_id, background, port ...); options of the renderer are ignored, see env.split_config"""
        from .dense import dense_config
        from .env import DENSE_IDS, NEW_CONTROLLERS, control_options, split_config, split_dense_config

        if name not in AGENTS:
            raise ValueError("unknown env %s (this build accelerates %s)" % (name, sorted(AGENTS)))
        if furniture_name is not None:
            config["furniture_name"] = furniture_name
        self.dense_cfg = None
        if name in DENSE_IDS:  # FurnitureSawyerDenseRewardEnv: the phase-based reward, computed inside the step kernel
            furniture_name, over, dense, self.ignored_config = split_dense_config(config)
            self.dense_cfg = dense_config(**dense)
        else:
            furniture_name, over, self.ignored_config = split_config(config)
        ctl = control_options(config)
        self.control_type, self.ik_cfg, self.ctl_cfg = ctl["control_type"], None, None
        agent = AGENTS[name]
        if self.control_type in NEW_CONTROLLERS:  # the torque controllers run on the torque-actuated robot (furniture.py:1893-1899)
            if agent != "Sawyer":
                raise NotImplementedError("the torque controllers are built for the Sawyer env")
            agent = "SawyerTorque"
        self.model = mjcf.load_scene(agent, furniture_name)
        self.cfg = default_config(**over)
        if self.control_type in NEW_CONTROLLERS:
            from .controllers import ctl_config

            self.ctl_cfg = ctl_config(self.control_type, model=self.model, move_speed=ctl.get("ik", {}).get("move_speed", 0.1))
        if self.control_type in ("ik", "ik_quaternion"):  # "ik" is the reference's default control type (config/furniture.py:57)
            if AGENTS[name] not in ("Sawyer", "Baxter"):
                raise NotImplementedError("control_type='%s' is built for the Sawyer and Baxter envs" % self.control_type)
            from .ik import ik_config

            self.ik_cfg = ik_config(self.model, **dict(ctl["ik"], quaternion_mode=int(self.control_type == "ik_quaternion")))
        self.engine = Engine(self.model, 1, device=device, config=self.cfg, lib_path=lib_path, dense=self.dense_cfg, ik=self.ik_cfg, controller=self.ctl_cfg)
        self.n_objects = self.engine.scene.npart
        self.object_ob_dim = 7 * self.n_objects
        self.robot_ob_dim = self.engine.scene.robot_ob_dim
        self.phase_ob_dim = 8 if self.engine.scene.phase_ob else 0
        self._robot_skip = 0 if self.control_type == "impedance" else 14  # ik: robot_ob has no joint positions / velocities (furniture_sawyer.py:110-125)
        self._narms = max(1, int(self.engine.scene.narms))
        per = self.robot_ob_dim // self._narms
        self._robot_cols = np.array([self.object_ob_dim + a * per + k for a in range(self._narms) for k in range(self._robot_skip, per)])
        self.robot_ob_dim -= self._robot_skip * self._narms
        self.dof = self.engine.act_dim
        self._max_episode_steps = self.cfg.max_episode_steps
        self._pending_ob = None  # observation of an episode the device has already started
        self._episode_reward = 0.0
        self._episode_time = time.time()

    # ---- spaces (furniture.py:215-252, :293-310): plain shape / bound records, gym.spaces objects when gym is installed
    @property
    def observation_space(self):
        shapes = OrderedDict(object_ob=(self.object_ob_dim,), robot_ob=(self.robot_ob_dim,))
        if self.phase_ob_dim:
            shapes["phase_ob"] = (8,)
        try:
            import gym.spaces as sp

            return sp.Dict(OrderedDict((k, sp.Box(low=-np.inf, high=np.inf, shape=v)) for k, v in shapes.items()))
        except Exception:
            return shapes

    @property
    def action_space(self):
        try:
            import gym.spaces as sp

            return sp.Dict([("default", sp.Box(shape=(self.dof,), low=-1, high=1, dtype=np.float32))])
        except Exception:
            return OrderedDict(default=dict(shape=(self.dof,), low=-1.0, high=1.0))

    def _ob(self, obs_row):
        a, b = self.object_ob_dim, self.object_ob_dim + self._robot_skip * self._narms + self.robot_ob_dim
        ob = OrderedDict(object_ob=obs_row[:a].astype(np.float64), robot_ob=obs_row[self._robot_cols].astype(np.float64))
        if self.phase_ob_dim:
            ob["phase_ob"] = obs_row[b:].astype(np.float64)
        return ob

    def reset(self):
        if self._pending_ob is not None:
            ob, self._pending_ob = self._pending_ob, None
        else:
            self.engine.env_reset()
            ob = self._ob(self.engine.get("obs")[0])
        self._episode_reward = 0.0
        self._episode_time = time.time()
        return ob

    def step(self, action):
        if isinstance(action, dict):
            action = np.concatenate([np.asarray(v, dtype=np.float32).ravel() for v in action.values()])
        a = np.asarray(action, dtype=np.float32).reshape(1, self.dof)
        if self._pending_ob is not None:  # stepping on after `done` without reset(): the device is already in the next episode
            self._pending_ob, self._episode_reward, self._episode_time = None, 0.0, time.time()
        obs, rew, done, info = self.engine.env_step_host(a)
        reward, done = float(rew[0]), bool(done[0])
        self._episode_reward += reward
        out = OrderedDict()
        ob = self._ob(obs[0])
        if self.dense_cfg is not None:  # the step's reward terms (_compute_reward's info, furniture_sawyer_dense.py:574-586)
            from .dense import INFO_KEYS

            row = self.engine.get("dense_info")[0]
            out.update((k, float(v)) for k, v in zip(INFO_KEYS, row))
        if done:
            unstable = int(info[0][2])
            out["episode_success"] = int(info[0][1])
            out["episode_reward"] = self._episode_reward
            out["episode_length"] = int(info[0][3])
            out["episode_time"] = time.time() - self._episode_time
            out["episode_unstable"] = -float(self.cfg.unstable_penalty_coef) if unstable else 0
            out["episode_num_connected"] = int(info[0][0])
            self._pending_ob = ob
        return ob, reward, done, out

    # ---- pass-throughs of FurnitureGym (furniture_gym.py:35-50)
    def set_max_episode_steps(self, max_episode_steps):
        self._max_episode_steps = int(max_episode_steps)
        self.engine.set_max_episode_steps(self._max_episode_steps)

    def get_env_state(self):
        q, v = self.engine.get_state()
        return {"qpos": q[0].astype(np.float64), "qvel": v[0].astype(np.float64)}

    def set_env_state(self, state):
        self.engine.set_state(np.asarray(state["qpos"])[None], np.asarray(state["qvel"])[None])
        self.engine.set("ctrl", np.zeros((1, self.model.nu), np.float32))
        self.engine.forward()

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the accelerated path")

    def close(self):
        self.engine.close()


def register_gym_envs():
    """Registers the accelerated counterpart of the reference's gym ids (furniture/env/__init__.py:19-114) when gym or
    gymnasium is importable: `gym.make("IKEASawyer-v0")` then builds a FurnitureGymB200 with the reference's kwargs
    (furniture_name="swivel_chair_0700").  Returns the ids registered (empty without gym)."""
    try:
        from gym.envs.registration import register
    except Exception:
        try:
            from gymnasium.envs.registration import register
        except Exception:
            return []
    specs = {"IKEASawyerDense-v0": {"id": "IKEASawyerDense-v0", "name": "FurnitureSawyerDenseRewardEnv", "unity": False},
             "furniture-sawyer-densereward-v0": {"id": "IKEASawyerDense-v0", "name": "FurnitureSawyerDenseRewardEnv", "unity": False},
             "IKEASawyer-v0": {"id": "IKEASawyer-v0", "name": "FurnitureSawyerEnv", "furniture_name": "swivel_chair_0700", "background": "Industrial", "port": 1050},
             "IKEABaxter-v0": {"id": "IKEABaxter-v0", "name": "FurnitureBaxterEnv", "furniture_id": 1, "background": "Interior", "port": 1050}}
    done = []
    try:  # the Cursor agent is host logic over the simulator surface (furniture_b200/cursor_env.py)
        register(id="IKEACursor-v0", entry_point="furniture_b200.cursor_env:FurnitureCursorEnvB200",
                 kwargs={"id": "IKEACursor-v0", "name": "FurnitureCursorEnv", "furniture_id": 0, "background": "Lab", "port": 1050})
        done.append("IKEACursor-v0")
    except Exception:
        pass
    for env_id, kwargs in specs.items():
        try:
            register(id=env_id, entry_point="furniture_b200.gym_api:FurnitureGymB200", kwargs=kwargs)
            done.append(env_id)
        except Exception:  # already registered (e.g. by the reference package itself)
            pass
    return done
