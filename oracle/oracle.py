"""ctypes binding of the CPU ORACLE (oracle/fe_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The physics restated here is "parity unpinned" against MuJoCo (see fe_oracle.h); it is the checker the CUDA
path is compared with, never the thing shipped or measured as the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INT_SCALARS = "nq nv nu nbody njnt ngeom nsite neq npair opt_cone_elliptic opt_iterations".split()
DBL_SCALARS = "opt_timestep opt_impratio opt_tolerance stat_meaninertia".split()
INT_ARRAYS = (
    "body_parentid body_weldid body_rootid body_jntadr body_jntnum body_dofadr body_dofnum jnt_type jnt_bodyid "
    "jnt_qposadr jnt_dofadr jnt_limited dof_bodyid dof_jntid dof_parentid geom_type geom_bodyid geom_contype "
    "geom_conaffinity geom_condim site_bodyid actuator_type actuator_jntid actuator_ctrllimited actuator_forcelimited "
    "eq_obj1id eq_obj2id eq_active collision_pairs geom_meshadr geom_meshnum"
).split()
DBL_ARRAYS = (
    "opt_gravity body_pos body_quat body_ipos body_iquat body_mass body_inertia body_invweight0 jnt_pos jnt_axis "
    "jnt_range jnt_solref jnt_solimp dof_damping dof_armature dof_invweight0 qpos0 geom_size geom_pos geom_quat "
    "geom_friction geom_solref geom_solimp geom_margin geom_gap geom_rbound site_pos site_quat actuator_gear "
    "actuator_gainprm actuator_biasprm actuator_ctrlrange actuator_forcerange eq_solref eq_solimp eq_data mesh_vert"
).split()


class Contact(C.Structure):
    _fields_ = [
        ("dist", C.c_double), ("pos", C.c_double * 3), ("frame", C.c_double * 9), ("friction", C.c_double * 5),
        ("solref", C.c_double * 2), ("solimp", C.c_double * 5), ("mu", C.c_double), ("margin", C.c_double), ("gap", C.c_double), ("dim", C.c_int),
        ("geom1", C.c_int), ("geom2", C.c_int), ("efc_address", C.c_int),
    ]


def build(force=False):
    so = os.path.join(_HERE, "libfe_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fe_oracle.c", "fe_oracle_collide.c", "fe_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.om_model_new.restype = C.c_void_p
        L.om_data_new.restype = C.c_void_p
        L.om_data_new.argtypes = [C.c_void_p]
        L.om_data_dbl.restype = C.POINTER(C.c_double)
        L.om_data_dbl.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.om_data_int.restype = C.POINTER(C.c_int)
        L.om_data_int.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.om_model_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.om_model_set_dbl.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        for f in ("om_forward", "om_step", "om_kinematics", "om_smooth", "om_collision", "om_make_constraint", "om_solve", "om_reset_data"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, f).restype = None
        L.om_model_free.argtypes = [C.c_void_p]
        L.om_data_free.argtypes = [C.c_void_p]
        L.om_site_velocity.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.om_collide_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.om_collide_pair.restype = C.c_int
        L.om_data_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.om_data_scalar.restype = C.c_int
        L.om_data_contacts.argtypes = [C.c_void_p]
        L.om_data_contacts.restype = C.c_void_p
        L.om_clear_warning.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class OracleSim:
    """Single-env CPU sim over a compiled furniture_b200.mjcf.Model; mirrors the mujoco_py.MjSim surface the
    reference uses (forward/step/reset, data fields as numpy views)."""

    def __init__(self, model):
        self.L = lib()
        self.model = model
        self.m = self.L.om_model_new()
        a = dict(model.a)
        a["npair"] = len(a["collision_pairs"])
        for k in INT_SCALARS:
            v = np.array([int(a[k])], dtype=np.int32)
            assert self.L.om_model_set_int(self.m, k.encode(), v.ctypes.data, 1) == 0, k
        for k in DBL_SCALARS:
            v = np.array([float(a[k])], dtype=np.float64)
            assert self.L.om_model_set_dbl(self.m, k.encode(), v.ctypes.data, 1) == 0, k
        for k in INT_ARRAYS:
            v = np.ascontiguousarray(a[k], dtype=np.int32).ravel()
            assert self.L.om_model_set_int(self.m, k.encode(), v.ctypes.data, v.size) == 0, k
        for k in DBL_ARRAYS:
            v = np.ascontiguousarray(a[k], dtype=np.float64).ravel()
            assert self.L.om_model_set_dbl(self.m, k.encode(), v.ctypes.data, v.size) == 0, k
        self.d = self.L.om_data_new(self.m)
        self._views = {}

    def __del__(self):
        try:
            self.L.om_data_free(self.d)
            self.L.om_model_free(self.m)
        except Exception:
            pass

    # ---- data views
    def f(self, name):
        """numpy view of a double data array (efc_* arrays are re-fetched: their length changes)."""
        n = C.c_int()
        p = self.L.om_data_dbl(self.d, self.m, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)) if n.value else np.zeros(0)

    def i(self, name):
        n = C.c_int()
        p = self.L.om_data_int(self.d, self.m, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)) if n.value else np.zeros(0, dtype=np.int32)

    def __getattr__(self, name):
        if name.startswith("_") or name in ("L", "m", "d", "model", "ncon", "nefc"):
            raise AttributeError(name)
        try:
            return self.f(name)
        except KeyError:
            try:
                return self.i(name)
            except KeyError:
                raise AttributeError(name)

    def scalar(self, name):
        return self.L.om_data_scalar(self.d, name.encode())

    def set_model(self, name, value):
        """overwrite a float array of the model (sim.model.body_pos[...] = ... of the reference's _set_pos, furniture.py:3133-3145)"""
        v = np.ascontiguousarray(value, dtype=np.float64).ravel()
        assert self.L.om_model_set_dbl(self.m, name.encode(), v.ctypes.data, v.size) == 0, name

    @property
    def ncon(self):
        return self.scalar("ncon")

    @property
    def nefc(self):
        return self.scalar("nefc")

    def contacts(self):
        p = self.L.om_data_contacts(self.d)
        return list((Contact * self.ncon).from_address(p)) if self.ncon else []

    def solver_info(self):
        return {k: self.scalar(k) for k in ("nefc", "ne", "nl", "nc", "solver_niter", "warning")}

    # ---- sim surface
    def reset(self):
        self.L.om_reset_data(self.m, self.d)

    def forward(self):
        self.L.om_forward(self.m, self.d)

    def step(self, n=1):
        for _ in range(n):
            self.L.om_step(self.m, self.d)

    def stage(self, name):
        getattr(self.L, "om_" + name)(self.m, self.d)

    def site_velocity(self, site):
        out = np.zeros(6)
        self.L.om_site_velocity(self.m, self.d, site, out.ctypes.data)
        return out

    def collide_pair(self, g1, g2):
        buf = (Contact * 8)()
        n = self.L.om_collide_pair(self.m, self.d, g1, g2, C.byref(buf))
        return [buf[i] for i in range(n)]
