"""CPU restatement of the reference's host loop on top of the fp64 physics oracle.

TEST INFRASTRUCTURE ONLY (see oracle/mj_oracle.cpp header).  One instance = one environment, as in
the reference.  Restates, in the reference's order of operations:
  MuJoCoRL.step              MuJoCo_Gym/mujoco_rl.py:243-289
  MuJoCoRL.__apply_dynamics  MuJoCo_Gym/mujoco_rl.py:215-241
  MuJoCoRL.reset             MuJoCo_Gym/mujoco_rl.py:291-331
  __check_truncations        MuJoCo_Gym/mujoco_rl.py:406-417
  apply_action               MuJoCo_Gym/mujoco_parent.py:316-336
  get_observations           MuJoCo_Gym/mujoco_parent.py:380-392
  distance                   MuJoCo_Gym/mujoco_parent.py:428-449
and the example plugins (README.md:108-173, Testing/Pick_Up_Dynamic.py, fps_custom_env.py:4-27) in
their reference form (Python objects in a per-agent dict store).  Random draws (`random.randint`,
README.md:154) are taken from an injected `draw(agent_index, counter) -> int` callable so that the
CUDA path can be checked on identical draws.  Pinned against the REAL reference host code run over
import stubs by tests/golden/make_golden.py.
"""
import math

import numpy as np

from .sim import OracleSim


class Language:
    """README.md:108-137, 4-tuple return (mujoco_rl.py:124,236)."""

    def __init__(self, env):
        self.env = env
        self.observation_space = {"low": [0], "high": [3]}
        self.action_space = {"low": [0], "high": [3]}

    def dynamic(self, agent, actions):
        store = self.env.data_store
        if "utterance" not in store[agent].keys():
            store[agent]["utterance"] = 0
        utterance = int(actions[0])
        store[agent]["utterance"] = utterance
        other = [o for o in self.env.agents if o != agent][0]
        if "utterance" in store[other]:
            return 0, np.array([store[other]["utterance"]]), False, {}
        return 0, np.array([0]), False, {}


class PickUp:
    """Testing/Pick_Up_Dynamic.py:4-41 per agent."""

    def __init__(self, env):
        self.env = env
        self.observation_space = {"low": [-70, -70, -70, 0], "high": [70, 70, 70, 1]}
        self.action_space = {"low": [], "high": []}

    def dynamic(self, agent, actions):
        env, st = self.env, self.env.data_store[agent]
        reward = 0
        if "inventory" not in st:
            st["inventory"] = [0]
        if "current_target" not in st:
            st["current_target"] = env.draw_target(agent)
        d = env.distance(agent, st["current_target"])
        if d < 2:
            st["inventory"][0] = 1 - st["inventory"][0]
            reward = 1
            st["current_target"] = env.draw_target(agent)
            st["distance"] = env.distance(agent, st["current_target"])
        pos = env.position(st["current_target"])
        return reward, np.concatenate((pos, st["inventory"])), False, {}


def tag_distance_reward(env, agent):
    """README.md:149-163, evident intent."""
    st = env.data_store[agent]
    if "current_target" not in st:
        st["current_target"] = env.draw_target(agent)
        st["distance"] = env.distance(agent, st["current_target"])
        new_reward = 0
    else:
        distance = env.distance(agent, st["current_target"])
        new_reward = st["distance"] - distance
        st["distance"] = distance
    return new_reward * 10


def distance_done(env, agent):
    """README.md:168-173"""
    return bool(env.data_store[agent].get("distance", 0.0) <= 1)


def ant_reward(env, agent):
    """benchmarking/fps_gym/fps_custom_env.py:4-27 (cfrc_ext is zero on these models)."""
    st = env.data_store[agent]
    xpos_before = st.get("xpos_before", None)
    xpos_after = env.position(agent)[0]
    if xpos_before is None:
        st["xpos_before"] = xpos_after
        return 0
    dt = env.timestep_len
    reward = (xpos_after - xpos_before) / dt - 0.5 * np.square(env.sim.ctrl).sum()
    st["xpos_before"] = xpos_after
    return reward


class OracleEnv:
    """Single-environment mirror of MuJoCoRL on the oracle.  `tables` is the product's Tables object
    (index lists derived with the reference's rules), `targets` the filter_by_tag("target") names."""

    def __init__(self, model, tables, agents, free_joint=False, skip_frames=1, max_steps=1024, dynamics=(),
                 reward_functions=(), done_functions=(), targets=(), draw=None, resolve=None):
        self.model, self.tables, self.agents = model, tables, list(agents)
        self.sim = OracleSim(model.blob)
        self.free_joint, self.skip_frames, self.max_steps = free_joint, skip_frames, max_steps
        self.timestep_len = model.timestep
        self.targets = list(targets)
        self._draw = draw
        self._draw_count = {a: 0 for a in self.agents}
        self._resolve = resolve
        self.data_store = {a: {} for a in self.agents}
        self.environment_dynamics = [d(self) for d in dynamics]
        self.reward_functions, self.done_functions = list(reward_functions), list(done_functions)
        n_phys = len(tables.act_space[self.agents[0]]["low"])
        self.action_routing = {"physical": [0, n_phys], "dynamic": {}}
        pos = n_phys
        for d in self.environment_dynamics:
            self.action_routing["dynamic"][d.__class__.__name__] = [pos, pos + len(d.action_space["low"])]
            pos += len(d.action_space["low"])
        self.act_dim = pos
        self.timestep = 0

    # ---- helpers the plugins use
    def draw_target(self, agent):
        a = self.agents.index(agent)
        k = self._draw(a, self._draw_count[agent]) % len(self.targets)
        self._draw_count[agent] += 1
        return self.targets[k]

    def position(self, name):
        ot, oid = self._resolve(name)
        return (self.sim.xipos[oid] if ot == 1 else self.sim.geom_xpos[oid]).copy()

    def distance(self, a, b):
        return math.dist(self.position(a), self.position(b))

    def get_observations(self, agent):
        oi = self.tables.agents_observation_index[agent]
        s = self.sim
        return np.array([s.sensordata[i] for i in oi["sensors"]] + [s.qpos[i] for i in oi["qpos"]] + [s.qvel[i] for i in oi["qvel"]])

    def apply_action(self, actions):
        for agent, act in actions.items():
            idx = self.tables.agents_action_index[agent]
            if self.free_joint:
                self.sim.qvel[idx] = act
            else:
                self.sim.ctrl[idx] = act[:len(idx)]
        for _ in range(self.skip_frames):
            self.sim.step()

    def _apply_dynamics(self, action, observations, rewards, terminations, infos):
        for dyn in self.environment_dynamics:
            for agent in self.agents:
                lo, hi = self.action_routing["dynamic"][dyn.__class__.__name__]
                reward, obs, done, info = dyn.dynamic(agent, action[agent][lo:hi])
                observations[agent] = np.concatenate((observations[agent], obs))
                rewards[agent] += reward
                terminations[agent] = any([terminations[agent], done])
                infos[agent][dyn.__class__.__name__] = info

    def step(self, action):
        lo, hi = self.action_routing["physical"]
        self.apply_action({k: action[k][lo:hi] for k in action})
        observations = {a: self.get_observations(a) for a in self.agents}
        rewards = {a: 0 for a in self.agents}
        terminations = {a: False for a in self.agents}
        infos = {a: {} for a in self.agents}
        self._apply_dynamics(action, observations, rewards, terminations, infos)
        for fn in self.reward_functions:
            rewards = {a: rewards[a] + fn(self, a) for a in self.agents}
        trunc = self.timestep >= self.max_steps
        truncations = {a: trunc for a in self.agents}
        truncations["__all__"] = all(truncations.values())
        if len(self.done_functions) != 0:
            for fn in self.done_functions:
                terminations = {a: any([terminations[a], fn(self, a)]) for a in self.agents}
                terminations["__all__"] = any(terminations.values())
                if terminations["__all__"]:
                    break
        self.timestep += 1
        return observations, rewards, terminations, truncations, infos

    def reset(self, action):
        """`action`: the sampled action the reference applies the dynamics with (mujoco_rl.py:315)."""
        self.sim.reset()
        self.sim.forward()
        self.data_store = {a: {} for a in self.agents}
        observations = {a: self.get_observations(a) for a in self.agents}
        rewards = {a: 0 for a in self.agents}
        terminations = {a: False for a in self.agents}
        infos = {a: {} for a in self.agents}
        self._apply_dynamics(action, observations, rewards, terminations, infos)
        self.data_store = {a: {} for a in self.agents}
        self.timestep = 0
        return observations, infos
