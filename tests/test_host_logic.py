"""Host-loop restatement (oracle/host_loop.py) and the product's table builder (tables.py) against
goldens produced by the REAL reference host code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from common import LEVELS, load_scene
from mujoco_rl_environment_wrapper_b200 import _lib as L
from mujoco_rl_environment_wrapper_b200.tables import Tables
from oracle import host_loop as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tables_for(name):
    xml, agents, fj = {
        "2A": ("two_ants.xml", ["sender", "receiver"], False),
        "2A_free": ("two_ants.xml", ["sender", "receiver"], True),
        "3S_free": ("two_ants_touch_acc.xml", ["sender", "receiver"], True),
        "1A": ("ant_rk4.xml", ["torso"], False), "C1": ("one_ant_arena.xml", ["sender"], False),
        "S1_free": ("box_touch.xml", ["receiver"], True),
        "S3_free": ("box_rangefinder.xml", ["receiver"], True)}[name]
    text = open(os.path.join(LEVELS, xml)).read()
    model = L.Model(text)
    return model, Tables(text, model, agents, fj), agents


@pytest.mark.parametrize("name", ["2A", "2A_free", "3S_free", "1A", "C1", "S1_free", "S3_free"])
def test_tables_equal_reference(name):
    gold = json.load(open(os.path.join(GOLD, "tables.json")))[name]
    model, t, agents = _tables_for(name)
    assert t.agents_action_index == gold["agents_action_index"]
    assert t.agents_observation_index == gold["agents_observation_index"]
    for a in agents:
        assert [float(x) for x in t.obs_space[a]["low"]] == gold["obs_low"][a]
        assert [float(x) for x in t.obs_space[a]["high"]] == gold["obs_high"][a]
        assert [float(x) for x in t.act_space[a]["low"]] == gold["act_low"][a]
        assert [float(x) for x in t.act_space[a]["high"]] == gold["act_high"][a]


def _resolver(model):
    def resolve(name):
        b = model.name2id(L.OBJ_BODY, name)
        if b >= 0:
            return 1, b
        return 5, model.name2id(L.OBJ_GEOM, name)
    return resolve


def test_host_loop_equals_reference_trace():
    g = json.load(open(os.path.join(GOLD, "host_loop_2A.json")))
    model, tables, agents = _tables_for("2A")
    draws = g["draws"][g["init_draws"]:]           # the constructor's validator consumed the first ones
    per_agent = {0: [draws[0]], 1: [draws[1]]}     # first step: sender then receiver (fn outer, agent inner)
    env = H.OracleEnv(model, tables, agents, max_steps=g["max_steps"], dynamics=[H.Language],
                      reward_functions=[H.tag_distance_reward], done_functions=[H.distance_done],
                      targets=g["targets"], draw=lambda a, k: per_agent[a][k], resolve=_resolver(model))
    obs, infos = env.reset({a: np.zeros(9) for a in agents})
    for a in agents:
        assert np.allclose(obs[a][:59], g["reset_obs"][a][:59], rtol=0, atol=1e-12)
    assert env.data_store == {a: {} for a in agents} and g["store_after_reset"] == {a: {} for a in agents}
    for t, st in enumerate(g["steps"]):
        act = {a: np.array(st["action"][a]) for a in agents}
        o, r, term, trunc, info = env.step(act)
        for a in agents:
            assert np.allclose(o[a], st["obs"][a], rtol=0, atol=1e-12), (t, a)
            assert r[a] == pytest.approx(st["reward"][a], abs=1e-12)
            assert isinstance(r[a], int) == st["reward_is_int"][a]
            assert sorted(info[a].keys()) == st["info_keys"][a]
        assert {k: bool(v) for k, v in term.items()} == st["term"]
        assert {k: bool(v) for k, v in trunc.items()} == st["trunc"]


def test_reference_semantics_documented_in_survey():
    """SURVEY.md 3.3 [RUN] facts: Gauss-Seidel agent order in Language, truncation on the (max_steps+1)-th
    call, `__all__` keys, int rewards without plugins."""
    model, tables, agents = _tables_for("2A")
    env = H.OracleEnv(model, tables, agents, max_steps=3, dynamics=[H.Language], resolve=_resolver(model))
    env.reset({a: np.zeros(9) for a in agents})
    mk = lambda u1, u2: {"sender": np.array([0.0] * 8 + [u1]), "receiver": np.array([0.0] * 8 + [u2])}
    o, r, term, trunc, _ = env.step(mk(1, 2))
    assert (o["sender"][-1], o["receiver"][-1]) == (0, 1)
    o, *_ = env.step(mk(3, 0))
    assert (o["sender"][-1], o["receiver"][-1]) == (2, 3)
    assert "__all__" not in term and trunc["__all__"] is False and r == {"sender": 0, "receiver": 0}
    env.step(mk(0, 0))
    *_, trunc, _ = env.step(mk(0, 0))
    assert trunc["__all__"] is True   # 4th call with maxSteps = 3
