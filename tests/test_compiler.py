"""MJCF compiler (csrc/mjcf_compile.cpp) against the sizes and tables SURVEY.md A.1 derived by running
the reference's own index helpers, plus closed-form mass properties."""
import math
import re

import numpy as np
import pytest

from common import SCENES, load_scene
from mujoco_rl_environment_wrapper_b200 import _lib as L

SIZES = {  # nq, nv, nu, nbody, ngeom, nsensordata  
    "1A": (15, 14, 8, 14, 14, 0), "C1": (15, 14, 8, 21, 20, 1), "2A": (30, 28, 16, 36, 35, 2),
    "3S": (30, 28, 16, 36, 35, 10), "S1": (7, 6, 0, 11, 10, 1), "S2": (7, 6, 0, 11, 10, 3),
    "S3": (7, 6, 0, 11, 10, 1), "S4": (7, 6, 0, 11, 10, 3),
}
PAIRS = {"1A": 13, "C1": 135, "2A": 491, "S3": 9}


@pytest.mark.parametrize("name", sorted(SIZES))
def test_sizes(name):
    model, tables, agents, fj = load_scene(name)
    assert (model.nq, model.nv, model.nu, model.nbody, model.ngeom, model.nsensordata) == SIZES[name]


@pytest.mark.parametrize("name", sorted(PAIRS))
def test_collision_pair_table(name):
    model, *_ = load_scene(name)
    assert model.npair == PAIRS[name]


def test_pair_types_2A():
    model, *_ = load_scene("2A")
    f = model.fields
    from collections import Counter
    cnt = Counter((int(f["geom_type"][a]), int(f["geom_type"][b])) for a, b in zip(f["pair_geom1"], f["pair_geom2"]))
    assert cnt == {(0, 2): 2, (0, 3): 24, (2, 6): 16, (3, 6): 192, (2, 3): 32, (2, 2): 1, (3, 3): 224}


def test_index_tables_match_reference_run():
    """[RUN] values in SURVEY.md 8a rows a2 / a5 (produced by the reference's own helpers)."""
    model, tables, agents, fj = load_scene("2A")
    assert tables.agents_action_index == {"sender": [2, 3, 4, 5, 6, 7, 0, 1], "receiver": [10, 11, 12, 13, 14, 15, 8, 9]}
    oi = tables.agents_observation_index
    assert oi["sender"]["sensors"] == [0] and oi["receiver"]["sensors"] == [1]
    assert oi["sender"]["qpos"] == list(range(30)) and oi["sender"]["qvel"] == list(range(28))
    assert len(tables.obs_space["sender"]["low"]) == 59
    model, tables, agents, fj = load_scene("3S")
    oi = tables.agents_observation_index
    assert oi["sender"]["sensors"] == [0, 2, 4, 5, 6] and oi["receiver"]["sensors"] == [1, 3, 7, 8, 9]
    assert tables.agents_action_index == {"sender": [0, 1, 5], "receiver": [14, 15, 19]}
    model, tables, agents, fj = load_scene("1A")
    assert tables.agents_action_index == {"torso": [2, 3, 4, 5, 6, 7, 0, 1]}
    assert len(tables.obs_space["torso"]["low"]) == 29


def test_sensor_bounds_match_reference_sensor_test():
    """Testing/sensor_test.py:25-83 expects these bounds for the sensor slice."""
    exp = {"S1": ([0], [20.0]), "S2": ([-5.0] * 3, [5.0] * 3), "S3": ([-1], [10.0]), "S4": ([-1] * 3, [1] * 3)}
    for name, (lo, hi) in exp.items():
        model, tables, agents, fj = load_scene(name)
        n = len(lo)
        assert tables.obs_space["receiver"]["low"][:n] == lo and tables.obs_space["receiver"]["high"][:n] == hi
        assert len(tables.obs_space["receiver"]["low"]) == n + 13


def test_mass_properties_closed_form():
    model, *_ = load_scene("1A")
    f = model.fields
    torso = model.name2id(L.OBJ_BODY, "torso")
    m = 5.0 * 4.0 / 3.0 * math.pi * 0.25 ** 3
    assert f["body_mass"][torso] == pytest.approx(m, rel=1e-12)
    assert f["body_inertia"][3 * torso] == pytest.approx(0.4 * m * 0.25 ** 2, rel=1e-12)
    # capsule from fromto 0 0 0 -> 0.2 0.2 0, radius 0.08
    leg = model.name2id(L.OBJ_BODY, "front_left_leg")
    r, h = 0.08, math.sqrt(0.08)
    vol = math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3
    assert f["body_mass"][leg] == pytest.approx(5.0 * vol, rel=1e-12)
    assert np.allclose(f["body_ipos"][3 * leg:3 * leg + 3], [0.1, 0.1, 0.0])
    # hinge range in radians, ankle axis normalised
    j = model.name2id(L.OBJ_JOINT, "ankle_1")
    assert np.allclose(f["jnt_range"][2 * j:2 * j + 2], np.deg2rad([30, 70]))
    assert np.allclose(f["jnt_axis"][3 * j:3 * j + 3], np.array([-1, 1, 0]) / math.sqrt(2))
    # free joint: the translational inverse weight is at least 1 / total mass (the hinges add mobility)
    assert 1.0 / f["body_subtreemass"][torso] <= f["dof_invweight0"][0] < 1.02 / f["body_subtreemass"][torso]
    assert f["dof_invweight0"][0] == f["dof_invweight0"][1] == f["dof_invweight0"][2]
    assert f["opt_integrator"][0] == 1 and f["opt_timestep"][0] == 0.01


def test_qpos0_and_euler():
    model, *_ = load_scene("S3")
    q = model.fields["qpos0"]
    assert np.allclose(q[:3], [4.595446, 1.222577, 0.4743838])
    assert np.allclose(q[3:], [math.cos(math.pi / 2), 1, 0, 0], atol=1e-12)  # euler="180 0 0"


@pytest.mark.parametrize("xml,msg", [
    ("<mujoco><worldbody><body><geom type='mesh'/></body></worldbody></mujoco>", "unsupported geom type"),
    ("<mujoco><worldbody><body><joint type='ball'/><geom size='1'/></body></worldbody></mujoco>", "unsupported joint type"),
    ("<mujoco><worldbody><body>", "XML parse error"),
    ("<mujoco><worldbody><body><joint/></body></worldbody></mujoco>", "has no mass"),
    ("<notmujoco/>", "root element"),
])
def test_errors_are_reported_not_thrown(xml, msg):
    with pytest.raises(Exception, match=msg):
        L.Model(xml)


def test_name_lookup_fallback():
    model, *_ = load_scene("2A")
    assert model.name2id(L.OBJ_BODY, "choice_1") > 0
    assert model.name2id(L.OBJ_BODY, "choice_1_geom") == -1      # lets callers fall back to geoms
    assert model.name2id(L.OBJ_GEOM, "choice_1_geom") >= 0
    assert model.id2name(L.OBJ_BODY, model.name2id(L.OBJ_BODY, "sender")) == "sender"
    assert model.name2id(L.OBJ_BODY, "") == -1
