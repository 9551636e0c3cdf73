"""Host SIMT-emulator harness (test tooling): runs the device step code on the CPU over numpy buffers."""
import ctypes
import os
import subprocess

import numpy as np

from mujoco_rl_environment_wrapper_b200 import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = None


def emu_lib():
    global _EMU
    if _EMU is None:
        so = os.path.join(_HERE, "emu", "libmjb_emu.so")
        src = os.path.join(_HERE, "emu", "emu_lib.cpp")
        csrc = os.path.join(_HERE, "..", "mujoco_rl_environment_wrapper_b200", "csrc")
        deps = [src, os.path.join(_HERE, "emu", "simt_emu.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                   "-o", so, src])
        lib = ctypes.CDLL(so)
        lib.emu_last_error.restype = ctypes.c_char_p
        lib.emu_layout.argtypes = [ctypes.c_char_p, ctypes.POINTER(L.EnvSpec), ctypes.c_int, ctypes.POINTER(L.Layout)]
        lib.emu_run.argtypes = [ctypes.c_char_p, ctypes.POINTER(L.EnvSpec), ctypes.c_int, ctypes.POINTER(L.Buffers),
                                ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _EMU = lib
    return _EMU


MODE_STEP, MODE_PHYSICS, MODE_FORWARD, MODE_RESET = 0, 1, 2, 3


class EmuBatch:
    """numpy-backed stand-in for the CUDA batch, same buffers and semantics."""

    def __init__(self, blob, spec, num_envs, keepalive=()):
        self.lib = emu_lib()
        self.blob, self.spec, self.n = bytes(blob), spec, num_envs
        self._keep = keepalive
        lay = L.Layout()
        if self.lib.emu_layout(self.blob, ctypes.byref(spec), num_envs, ctypes.byref(lay)) != 0:
            raise Exception(self.lib.emu_last_error().decode())
        self.layout = lay
        A, N = spec.n_agents, num_envs
        f, i, u = np.float32, np.int32, np.uint8
        self.buf = {
            "qpos": np.zeros((N, lay.qpos_stride), f), "qvel": np.zeros((N, lay.qvel_stride), f),
            "ctrl": np.zeros((N, lay.ctrl_stride), f), "warmstart": np.zeros((N, lay.qvel_stride), f),
            "sensordata": np.zeros((N, lay.sensor_stride), f), "probe": np.zeros((N, max(1, lay.probe_count), 4), f),
            "actions": np.zeros((N, max(1, A), lay.act_stride), f), "obs": np.zeros((N, max(1, A), lay.obs_stride), f),
            "reward": np.zeros((N, max(1, A)), f), "term": np.zeros((N, A + 1), u), "trunc": np.zeros((N, A + 1), u),
            "timestep": np.zeros((N,), i), "store_i": np.zeros((N, max(1, A), lay.store_i32), i),
            "store_f": np.zeros((N, max(1, A), lay.store_f32), f), "ncon": np.zeros((N,), i),
            "contact_geom": np.zeros((N, lay.maxcon, 2), i), "contact_dist": np.zeros((N, lay.maxcon), f), "niter": np.zeros((N,), i), "nreset": np.zeros((N,), i), "ncon_dropped": np.zeros((N,), i),
        }
        self.B = L.Buffers()
        for k, v in self.buf.items():
            setattr(self.B, k, v.ctypes.data)

    def run(self, mode, skip_frames=1, mask=None, reverse=False):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data
        rc = self.lib.emu_run(self.blob, ctypes.byref(self.spec), self.n, ctypes.byref(self.B), mode, skip_frames, m,
                              1 if reverse else 0)
        if rc != 0:
            raise Exception(self.lib.emu_last_error().decode())

    def __getattr__(self, k):
        if k in self.__dict__.get("buf", {}):
            return self.buf[k]
        raise AttributeError(k)


def simple_spec(model, agent_bodies, act_index, n_phys, free_joint=False, skip_frames=1, max_steps=1024,
                obs_sensors=None):
    """EnvSpec with every agent observing (its sensors,) all qpos and all qvel; no plugins."""
    A = len(agent_bodies)
    spec = L.EnvSpec()
    spec.n_agents, spec.free_joint, spec.skip_frames, spec.max_steps = A, int(free_joint), skip_frames, max_steps
    spec.n_phys_act, spec.act_dim = n_phys, n_phys
    ai = (ctypes.c_int32 * max(1, len(act_index)))(*act_index)
    obs = []
    adr = [0]
    for a in range(A):
        sens = obs_sensors[a] if obs_sensors else []
        ent = [(0 << 24) | s for s in sens] + [(1 << 24) | q for q in range(model.nq)] + [(2 << 24) | v for v in range(model.nv)]
        obs += ent
        adr.append(len(obs))
        spec.obs_dim[a] = len(ent)
        spec.agent_body[a] = agent_bodies[a]
    oi = (ctypes.c_int32 * max(1, len(obs)))(*obs)
    for a in range(A + 1):
        spec.obs_adr[a] = adr[a]
    spec.act_index = ctypes.cast(ai, ctypes.POINTER(ctypes.c_int32))
    spec.obs_index = ctypes.cast(oi, ctypes.POINTER(ctypes.c_int32))
    spec.seed = 1234
    return spec, (ai, oi)
