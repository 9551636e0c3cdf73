"""ctypes wrapper around oracle/_build/liboracle.so (fp64 CPU restatement of mj_step).

TEST INFRASTRUCTURE ONLY.  Mirrors the slice of the `mujoco` Python API the reference touches
(MuJoCo_Gym/mujoco_parent.py:126-127,325-336,349-355,375,390-391): zero-copy numpy views of
qpos / qvel / ctrl / sensordata / xipos ..., `step`, `forward`, `reset`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def oracle_lib_path():
    return os.path.join(_HERE, "_build", "liboracle.so")


def build_oracle(force=False):
    """Compile the oracle with g++ (a few seconds). Building the checker is not using it."""
    path = oracle_lib_path()
    srcs = [os.path.join(_HERE, f) for f in ("mj_oracle.cpp", "orc_base.h", "orc_collide.h") if os.path.exists(os.path.join(_HERE, f))]
    if force or not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return path


def _lib():
    global _LIB
    if _LIB is None:
        path = oracle_lib_path()
        if not os.path.exists(path):
            build_oracle()
        lib = ctypes.CDLL(path)
        lib.orc_create.restype = ctypes.c_void_p
        lib.orc_create.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.orc_destroy.argtypes = [ctypes.c_void_p]
        lib.orc_set_solver.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        for fn in ("orc_reset", "orc_forward", "orc_step"):
            getattr(lib, fn).argtypes = [ctypes.c_void_p]
        lib.orc_array.restype = ctypes.POINTER(ctypes.c_double)
        lib.orc_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        lib.orc_time.restype = ctypes.c_double
        lib.orc_time.argtypes = [ctypes.c_void_p]
        lib.orc_energy.restype = ctypes.c_double
        lib.orc_energy.argtypes = [ctypes.c_void_p]
        for fn in ("orc_ncon", "orc_nefc", "orc_solver_iters"):
            getattr(lib, fn).argtypes = [ctypes.c_void_p]
        lib.orc_contact.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        lib.orc_render.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _LIB = lib
    return _LIB


class OracleSim:
    """One (model, data) pair backed by the fp64 oracle.  `blob` is the packed model blob
    (bytes) produced by the product's MJCF compiler (`mjb_model_blob`)."""

    SOLVER_EXACT, SOLVER_NEWTON_FIXED, SOLVER_PGS_FIXED, SOLVER_PGS_CONVERGED = 0, 1, 2, 3

    def __init__(self, blob: bytes):
        self._lib = _lib()
        self._blob = bytes(blob)
        self._h = self._lib.orc_create(self._blob, len(self._blob))
        if not self._h:
            raise RuntimeError("oracle: could not create simulation from blob")
        # fixed-size state arrays never move: cache the zero-copy views (as `MjData.qpos` etc. are in mujoco)
        self._fixed = {n: self.array(n) for n in ("qpos", "qvel", "ctrl", "qacc", "qacc_warmstart", "sensordata", "xpos",
                                                  "xmat", "xipos", "geom_xpos", "geom_xmat", "site_xpos", "site_xmat")}
        for n in ("xpos", "xipos", "geom_xpos", "site_xpos"):
            self._fixed[n] = self._fixed[n].reshape(-1, 3)
        for n in ("xmat", "geom_xmat", "site_xmat"):
            self._fixed[n] = self._fixed[n].reshape(-1, 9)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def array(self, name):
        n = ctypes.c_int()
        p = self._lib.orc_array(self._h, name.encode(), ctypes.byref(n))
        if n.value < 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    # state views (valid until the next call that resizes: efc_* arrays change size every forward)
    qpos = property(lambda s: s._fixed["qpos"])
    qvel = property(lambda s: s._fixed["qvel"])
    ctrl = property(lambda s: s._fixed["ctrl"])
    qacc = property(lambda s: s._fixed["qacc"])
    qacc_warmstart = property(lambda s: s._fixed["qacc_warmstart"])
    sensordata = property(lambda s: s._fixed["sensordata"])
    xpos = property(lambda s: s._fixed["xpos"])
    xmat = property(lambda s: s._fixed["xmat"])
    xipos = property(lambda s: s._fixed["xipos"])
    geom_xpos = property(lambda s: s._fixed["geom_xpos"])
    geom_xmat = property(lambda s: s._fixed["geom_xmat"])
    site_xpos = property(lambda s: s._fixed["site_xpos"])
    site_xmat = property(lambda s: s._fixed["site_xmat"])
    time = property(lambda s: s._lib.orc_time(s._h))
    ncon = property(lambda s: s._lib.orc_ncon(s._h))
    nefc = property(lambda s: s._lib.orc_nefc(s._h))
    solver_iters = property(lambda s: s._lib.orc_solver_iters(s._h))

    def set_solver(self, mode=0, iters=100, warmstart=True):
        self._lib.orc_set_solver(self._h, mode, iters, int(warmstart))

    def reset(self):
        self._lib.orc_reset(self._h)

    def forward(self):
        self._lib.orc_forward(self._h)

    def step(self):
        self._lib.orc_step(self._h)

    def energy(self):
        return self._lib.orc_energy(self._h)

    def render(self, cam, width=64, height=64):
        """u8 [height, width, 3] image of fixed camera `cam` at the current qpos (rows bottom-up)"""
        out = np.zeros((height, width, 3), np.uint8)
        if self._lib.orc_render(self._h, cam, width, height, out.ctypes.data) != 0:
            raise ValueError(f"camera {cam} cannot be rendered")
        return out

    def contact(self, i):
        g = (ctypes.c_int * 2)()
        d = ctypes.c_double()
        pos = (ctypes.c_double * 3)()
        fr = (ctypes.c_double * 9)()
        if self._lib.orc_contact(self._h, i, g, ctypes.byref(d), pos, fr) != 0:
            raise IndexError(i)
        return {"geom1": g[0], "geom2": g[1], "dist": d.value, "pos": np.array(pos[:]),
                "frame": np.array(fr[:]).reshape(3, 3)}

    def contact_pairs(self):
        return [(c["geom1"], c["geom2"]) for c in (self.contact(i) for i in range(self.ncon))]
