"""Generate golden fixtures by running the REAL reference host code (MuJoCo_Gym/mujoco_rl.py,
mujoco_parent.py, sensor.py, helper.py — imported unmodified from /root/reference) on top of import
stubs: `mujoco` is backed by the fp64 physics oracle, `xmltodict` / `gymnasium` / `pettingzoo` / `glfw`
by a few lines each.  The reference cannot run any other way here: `mujoco==2.3.3` is not installable
(no wheel, no network).  What the goldens pin is therefore the reference's HOST LOOP (index tables,
action routing, observation assembly, dynamics / reward / done ordering, truncation timing, reset
semantics) — physics numbers in them come from the oracle and are labelled as such.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Writes tests/golden/host_loop_2A.json, tables.json.
"""
import json
import os
import random
import sys
import types
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

from mujoco_rl_environment_wrapper_b200 import _lib as L  # noqa: E402  (MJCF compiler only)
from oracle import OracleSim  # noqa: E402


# ---- xmltodict stand-in: attributes -> "@k", repeated children -> list, empty element -> None
def _elem_to_dict(e):
    d = {}
    for k, v in e.attrib.items():
        d["@" + k] = v
    for ch in list(e):
        val = _elem_to_dict(ch)
        if ch.tag in d:
            if not isinstance(d[ch.tag], list):
                d[ch.tag] = [d[ch.tag]]
            d[ch.tag].append(val)
        else:
            d[ch.tag] = val
    return d if d else None


def xmltodict_parse(text):
    root = ET.fromstring(text)
    return {root.tag: _elem_to_dict(root)}


# ---- mujoco stand-in over the oracle
class _Named:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class MjModel:
    def __init__(self, path):
        self._m = L.Model.from_xml_path(path)
        f = self._m.fields
        self.jnt_qposadr, self.jnt_dofadr, self.jnt_type = f["jnt_qposadr"], f["jnt_dofadr"], f["jnt_type"]
        self.opt = _Named(timestep=self._m.timestep)
        self.nq, self.nv, self.nu = self._m.nq, self._m.nv, self._m.nu

    @staticmethod
    def from_xml_path(path):
        return MjModel(path)

    def joint(self, name):
        i = self._m.name2id(L.OBJ_JOINT, name)
        if i < 0:
            raise KeyError(name)
        return _Named(id=i, dofadr=np.array([self.jnt_dofadr[i]]), name=name)

    def body(self, name):
        i = self._m.name2id(L.OBJ_BODY, name)
        if i < 0:
            raise KeyError(name)
        return _Named(id=i, mass=np.array([self._m.fields["body_mass"][i]]), name=name)

    def geom(self, name):
        i = self._m.name2id(L.OBJ_GEOM, name)
        if i < 0:
            raise KeyError(name)
        f = self._m.fields
        return _Named(id=i, rgba=f["geom_rgba"][4 * i:4 * i + 4], type=np.array([f["geom_type"][i]]), name=name)

    def camera(self, name):
        raise KeyError(name)


class MjData:
    def __init__(self, model):
        self._model = model
        self._s = OracleSim(model._m.blob)
        self.cfrc_ext = np.zeros((model._m.nbody, 6))

    qpos = property(lambda s: s._s.qpos)
    qvel = property(lambda s: s._s.qvel)
    ctrl = property(lambda s: s._s.ctrl)
    time = property(lambda s: s._s.time)
    ncon = property(lambda s: s._s.ncon)

    @property
    def sensordata(self):
        return self._s.sensordata[:self._model._m.nsensordata]

    @property
    def contact(self):
        return [_Named(geom1=c["geom1"], geom2=c["geom2"]) for c in (self._s.contact(i) for i in range(self._s.ncon))]

    def body(self, name):
        i = self._model._m.name2id(L.OBJ_BODY, name)
        if i < 0:
            raise KeyError(name)
        return _Named(id=i, name=name, xipos=self._s.xipos[i], xpos=self._s.xpos[i], xmat=self._s.xmat[i])

    def geom(self, name):
        i = self._model._m.name2id(L.OBJ_GEOM, name)
        if i < 0:
            raise KeyError(name)
        return _Named(id=i, name=name, xpos=self._s.geom_xpos[i], xmat=self._s.geom_xmat[i])

    def sensor(self, name):
        m = self._model._m
        i = m.name2id(L.OBJ_SENSOR, name)
        if i < 0:
            raise KeyError(name)
        adr, dim = int(m.fields["sensor_adr"][i]), int(m.fields["sensor_dim"][i])
        return _Named(id=i, name=name, data=self._s.sensordata[adr:adr + dim])


def install_stubs():
    mj = types.ModuleType("mujoco")
    mj.MjModel, mj.MjData = MjModel, MjData
    mj.mj_step = lambda m, d: d._s.step()
    mj.mj_forward = lambda m, d: d._s.forward()
    mj.mj_resetData = lambda m, d: d._s.reset()
    mj.mjtJoint = _Named(mjJNT_FREE=0)
    mj.MjvCamera = mj.MjvOption = lambda *a, **k: _Named()
    glfw_mod = types.ModuleType("mujoco.glfw")
    glfw_mod.glfw = _Named()
    mj.glfw = glfw_mod
    sys.modules["mujoco"], sys.modules["mujoco.glfw"] = mj, glfw_mod
    xd = types.ModuleType("xmltodict")
    xd.parse = xmltodict_parse
    sys.modules["xmltodict"] = xd

    class Box:
        def __init__(self, low, high, **kw):
            self.low, self.high = np.asarray(low, np.float32), np.asarray(high, np.float32)
            self.shape = self.low.shape
            self._rng = np.random.default_rng(99)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1)
            hi = np.where(np.isfinite(self.high), self.high, 1)
            return self._rng.uniform(lo, hi).astype(np.float32)
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box, spaces.Space = Box, object
    gym.spaces = spaces
    sys.modules["gymnasium"], sys.modules["gymnasium.spaces"] = gym, spaces
    pz = types.ModuleType("pettingzoo")
    pz.ParallelEnv = type("ParallelEnv", (), {})
    sys.modules["pettingzoo"] = pz


# ---- the README plugins, in reference form (4-tuple Language, Q3; reward with the evident intent, Q2)
def make_plugins(draw_log):
    class Language:
        def __init__(self, mujoco_gym):
            self.mujoco_gym = mujoco_gym
            self.observation_space = {"low": [0], "high": [3]}
            self.action_space = {"low": [0], "high": [3]}

        def dynamic(self, agent, actions):
            if "utterance" not in self.mujoco_gym.data_store[agent].keys():
                self.mujoco_gym.data_store[agent]["utterance"] = 0
            utterance = int(actions[0])
            self.mujoco_gym.data_store[agent]["utterance"] = utterance
            otherAgent = [other for other in self.mujoco_gym.agents if other != agent][0]
            if "utterance" in self.mujoco_gym.data_store[otherAgent]:
                return 0, np.array([self.mujoco_gym.data_store[otherAgent]["utterance"]]), False, {}
            return 0, np.array([0]), False, {}

    def reward_function(mujoco_gym, agent):
        st = mujoco_gym.data_store[agent]
        if "current_target" not in st.keys():
            targets = mujoco_gym.filter_by_tag("target")
            k = random.randint(0, len(targets) - 1)
            draw_log.append(k)
            st["current_target"] = targets[k]["name"]
            st["distance"] = mujoco_gym.distance(agent, st["current_target"])
            new_reward = 0
        else:
            distance = mujoco_gym.distance(agent, st["current_target"])
            new_reward = st["distance"] - distance
            st["distance"] = distance
        return new_reward * 10

    def done_function(mujoco_gym, agent):
        return bool(mujoco_gym.data_store[agent].get("distance", 0.0) <= 1)

    return Language, reward_function, done_function


def main():
    install_stubs()
    sys.path.insert(0, REF)
    from MuJoCo_Gym.mujoco_rl import MuJoCoRL  # the real reference class
    levels = REF   # the reference's own level files, fed to the reference's own host code
    out_tables = {}
    for name, xml, agents, fj in [("2A", "benchmarking/levels/MultiAgentModel.xml", ["sender", "receiver"], False),
                                   ("2A_free", "benchmarking/levels/MultiAgentModel.xml", ["sender", "receiver"], True),
                                   ("3S_free", "benchmarking/levels/MultiAgentModel3Sensors.xml", ["sender", "receiver"], True),
                                   ("1A", "benchmarking/levels/Ant.xml", ["torso"], False),
                                   ("C1", "benchmarking/levels/SingleAgentModel.xml", ["sender"], False),
                                   ("S1_free", "Testing/sensor_levels/Model1.xml", ["receiver"], True),
                                   ("S3_free", "Testing/sensor_levels/Model3.xml", ["receiver"], True)]:
        cfg = {"xmlPath": os.path.join(levels, xml), "agents": agents, "freeJoint": fj, "skipFrames": 1}
        if name in ("1A", "S1_free", "S3_free") and not fj:
            pass
        try:
            env = MuJoCoRL(cfg)
        except Exception as e:  # e.g. reference needs >= 2 motors; sensor levels have no actuator in ctrl mode
            out_tables[name] = {"error": repr(e)}
            continue
        out_tables[name] = {
            "agents_action_index": {a: [int(i) for i in env.agents_action_index[a]] for a in agents},
            "agents_observation_index": {a: {k: [int(i) for i in v] for k, v in env.agents_observation_index[a].items()} for a in agents},
            "action_routing": env.action_routing,
            "obs_low": {a: [float(x) for x in env.observation_space(a).low] for a in agents},
            "obs_high": {a: [float(x) for x in env.observation_space(a).high] for a in agents},
            "act_low": {a: [float(x) for x in env.action_space(a).low] for a in agents},
            "act_high": {a: [float(x) for x in env.action_space(a).high] for a in agents},
        }
    json.dump(out_tables, open(os.path.join(HERE, "tables.json"), "w"), indent=1)

    # ---- host loop trace on the two-agent level (config C2)
    draw_log = []
    Language, reward_function, done_function = make_plugins(draw_log)
    random.seed(7)
    cfg = {"xmlPath": os.path.join(levels, "benchmarking/levels/MultiAgentModel.xml"), "infoJson": os.path.join(ROOT, "tests", "levels", "info_2A.json"),
           "agents": ["sender", "receiver"], "freeJoint": False, "skipFrames": 1, "maxSteps": 5,
           "environmentDynamics": [Language], "rewardFunctions": [reward_function], "doneFunctions": [done_function]}
    env = MuJoCoRL(cfg)
    init_draws = len(draw_log)  # the constructor's validators run the plugins once for agents[0]
    rng = np.random.default_rng(5)
    trace = {"max_steps": 5, "steps": [], "init_draws": init_draws}
    obs, infos = env.reset()
    trace["reset_obs"] = {a: [float(x) for x in obs[a]] for a in env.agents}
    trace["store_after_reset"] = {a: dict(env.data_store[a]) for a in env.agents}
    for t in range(8):
        act = {a: np.concatenate([rng.uniform(-1, 1, 8), rng.uniform(0, 3, 1)]).astype(np.float32) for a in env.agents}
        o, r, term, trunc, info = env.step(act)
        trace["steps"].append({
            "action": {a: [float(x) for x in act[a]] for a in env.agents},
            "obs": {a: [float(x) for x in o[a]] for a in env.agents},
            "reward": {a: float(r[a]) for a in env.agents},
            "reward_is_int": {a: isinstance(r[a], int) for a in env.agents},
            "term": {k: bool(v) for k, v in term.items()},
            "trunc": {k: bool(v) for k, v in trunc.items()},
            "info_keys": {a: sorted(info[a].keys()) for a in env.agents},
        })
    trace["draws"] = [int(d) for d in draw_log]
    trace["targets"] = [d["name"] for d in env.filter_by_tag("target")]
    json.dump(trace, open(os.path.join(HERE, "host_loop_2A.json"), "w"), indent=1)
    print("wrote goldens:", sorted(out_tables), "steps", len(trace["steps"]), "draws", trace["draws"])


if __name__ == "__main__":
    main()
