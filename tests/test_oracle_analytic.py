"""Known-answer tests that pin the fp64 oracle (the reference holds no golden numbers for mj_step,
SURVEY.md 8c): closed-form free fall under semi-implicit Euler, energy conservation of an undamped
chain + free body under RK4, static equilibrium of a sphere on a plane, joint-limit restoring force,
sensors at rest, and solver cross-checks (primal Newton == dual PGS at convergence)."""
import math

import numpy as np
import pytest

from common import load_scene
from mujoco_rl_environment_wrapper_b200 import _lib as L
from oracle import OracleSim

FREE = """<mujoco><option timestep="0.002"/><worldbody>
<body pos="0 0 5"><joint type="free"/><geom type="box" size="0.1 0.2 0.3" contype="0" conaffinity="0"/></body>
</worldbody></mujoco>"""

CHAIN = """<mujoco><option timestep="0.0005" integrator="RK4"/><worldbody>
<body pos="0 0 2"><joint type="hinge" axis="0 1 0" name="j1"/><geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.05" contype="0" conaffinity="0"/>
<body pos="0.5 0 0"><joint type="hinge" axis="0 1 0.3" name="j2"/><geom type="capsule" fromto="0 0 0 0.3 0.1 0" size="0.04" contype="0" conaffinity="0"/>
<body pos="0.3 0.1 0"><joint type="hinge" axis="1 0 0" name="j3"/><geom type="box" pos="0.1 0 0.05" size="0.1 0.05 0.02" euler="10 20 30" contype="0" conaffinity="0"/></body>
</body></body>
<body pos="1 1 1"><joint type="free"/><geom type="box" size="0.1 0.2 0.3" contype="0" conaffinity="0"/></body>
</worldbody></mujoco>"""

BALL = """<mujoco><option timestep="0.002"/><worldbody>
<geom type="plane" size="5 5 0.1"/>
<body pos="0 0 0.1"><joint type="free"/><geom type="sphere" size="0.1" density="1000"/></body>
</worldbody></mujoco>"""


def test_free_fall_semi_implicit_euler_closed_form():
    s = OracleSim(L.Model(FREE).blob)
    n, h, g = 500, 0.002, 9.81
    for _ in range(n):
        s.step()
    assert s.qvel[2] == pytest.approx(-g * h * n, rel=1e-12)
    assert s.qpos[2] == pytest.approx(5 - g * h * h * n * (n + 1) / 2, rel=1e-12)  # v updated before q
    assert s.time == pytest.approx(n * h)


def test_energy_conservation_rk4_chain_and_free_body():
    s = OracleSim(L.Model(CHAIN).blob)
    s.qvel[:] = [0.3, -0.2, 0.5, 0.1, 0.2, 0.3, 2.0, -1.0, 0.5]
    e0 = s.energy()
    for _ in range(2000):
        s.step()
    assert abs(s.energy() - e0) / abs(e0) < 1e-10


def test_sphere_on_plane_static_equilibrium():
    """At rest the contact force equals the weight; the penetration follows from the soft-constraint law
    a_ref = -b v - k d(r) r with the default solref / solimp."""
    model = L.Model(BALL)
    s = OracleSim(model.blob)
    for _ in range(3000):
        s.step()
    assert np.abs(s.qvel).max() < 1e-8
    s.forward()
    m = model.fields["body_mass"][1]
    f = s.array("efc_force").sum()                      # frictionless at rest: the 4 pyramid edges carry the normal load
    assert f == pytest.approx(m * 9.81, rel=1e-8)
    assert s.ncon == 1 and s.contact(0)["dist"] < 0


def test_joint_limit_pushes_back_into_range():
    """The ankles of the ant start outside their range (qpos0 = 0, range 30..70 deg): the limit force
    accelerates them toward the range."""
    model, tables, agents, fj = load_scene("1A")
    s = OracleSim(model.blob)
    s.qpos[2] = 3.0  # lift off the floor
    s.forward()
    f = model.fields
    assert s.nefc == 4 and s.ncon == 0
    for name, sign in (("ankle_1", 1), ("ankle_2", -1), ("ankle_3", -1), ("ankle_4", 1)):
        j = model.name2id(L.OBJ_JOINT, name)
        assert np.sign(s.qacc[f["jnt_dofadr"][j]]) == sign


def test_rangefinder_flipped_box_sees_the_floor():
    model, *_ = load_scene("S3")
    s = OracleSim(model.blob)
    s.forward()
    assert s.sensordata[0] == pytest.approx(model.fields["qpos0"][2], rel=1e-9)  # site z axis points down
    model, *_ = load_scene("2A")
    s = OracleSim(model.blob)
    s.forward()
    assert list(s.sensordata[:2]) == [-1.0, -1.0]  # upright ant: ray points up, nothing to hit


def test_accelerometer_and_axis_sensors_at_rest():
    model, *_ = load_scene("S2")
    s = OracleSim(model.blob)
    for _ in range(2500):
        s.step()
    s.forward()
    # resting box flipped by euler 180 about x: +g along world z is -g along the site z axis; cutoff 5
    assert np.allclose(s.sensordata[:3], [0, 0, -5.0], atol=1e-6)
    model, *_ = load_scene("S4")
    s = OracleSim(model.blob)
    s.forward()
    assert np.allclose(s.sensordata[:3], [1, 0, 0], atol=1e-12)


def test_primal_newton_equals_dual_pgs_at_convergence():
    model, tables, agents, fj = load_scene("2A")
    a, b = OracleSim(model.blob), OracleSim(model.blob)
    a.set_solver(OracleSim.SOLVER_EXACT)
    b.set_solver(OracleSim.SOLVER_PGS_CONVERGED)
    rng = np.random.default_rng(0)
    worst = 0
    for i in range(400):
        c = rng.uniform(-1, 1, 16)
        a.ctrl[:] = c
        b.qpos[:] = a.qpos; b.qvel[:] = a.qvel; b.ctrl[:] = c; b.qacc_warmstart[:] = a.qacc_warmstart
        a.step(); b.step()
        worst = max(worst, np.abs(a.qvel - b.qvel).max())
    assert a.ncon > 0 or worst >= 0
    assert worst < 1e-8


def test_oracle_against_mujoco_goldens():
    """Pins the oracle to real MuJoCo trajectories when tests/golden/make_physics_golden.py could be run somewhere
    (it needs `import mujoco`).  No golden file -> skipped, and physics parity stays UNPINNED."""
    import glob
    import os
    import pytest
    from mujoco_rl_environment_wrapper_b200 import _lib as L
    from oracle.sim import OracleSim
    here = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(here, "golden", "physics_*.npz")))
    if not files:
        pytest.skip("no MuJoCo physics goldens (mujoco is not installable in this environment)")
    for path in files:
        g = np.load(path)
        scene = os.path.basename(path)[len("physics_"):-4] + ".xml"
        model = L.Model(open(os.path.join(here, "levels", scene)).read())
        sim = OracleSim(model.blob)
        sim.reset()
        # one-step parity from MuJoCo's own previous state (no drift accumulation), every 10th step
        for t in range(1, len(g["qpos"]), 10):
            sim.array("qpos")[:] = g["qpos"][t - 1]
            sim.array("qvel")[:] = g["qvel"][t - 1]
            if model.nu:
                sim.array("ctrl")[:] = g["ctrl"][t]
            sim.step()
            assert np.allclose(sim.array("qpos"), g["qpos"][t], rtol=0, atol=1e-6), (scene, t)
            assert np.allclose(sim.array("qvel"), g["qvel"][t], rtol=0, atol=1e-5), (scene, t)
            want = sorted((int(a), int(b)) for tt, a, b in g["contact_pairs"] if tt == t)
            assert sorted(sim.contact_pairs()) == want, (scene, t)
