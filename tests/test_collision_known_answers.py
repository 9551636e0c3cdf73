"""Closed-form narrow-phase answers (tests/collision_cases.py) for the fp64 oracle AND the kernel source (host SIMT
emulator here; the same cases run through the C-ABI on the GPU in test_gpu_parity.py::test_collision_known_answers).
Box - box (SAT + face clipping + edge - edge) and capsule - box follow the specification in DESIGN.md 3c; the two
implementations share the specification, not the algorithm (exact breakpoint minimisation + Sutherland - Hodgman in
the oracle, golden-section search + lane-parallel vertex enumeration in the kernel)."""
import numpy as np
import pytest

import emu_harness as E
from collision_cases import CASES
from mujoco_rl_environment_wrapper_b200 import _lib as L
from oracle import OracleSim


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_known_answer(case):
    name, xml, qpos, ncon, dists, normal = case
    model = L.Model(xml)
    sim = OracleSim(model.blob)
    sim.qpos[:] = qpos
    sim.forward()
    cons = [sim.contact(i) for i in range(sim.ncon)]
    assert len(cons) == ncon, [c["dist"] for c in cons]
    assert np.allclose(sorted(c["dist"] for c in cons), sorted(dists), atol=1e-9)
    for c in cons:
        assert np.allclose(c["frame"][0], normal, atol=1e-9), (name, c["frame"][0])
        # position = midway between the two surfaces along the normal
        assert (c["geom1"], c["geom2"]) in ((0, 1), (1, 0))


def _emu_forward(xml, qpos):
    model = L.Model(xml)
    spec, keep = E.simple_spec(model, [1], [0, 1, 5], 3, free_joint=True)
    eb = E.EmuBatch(model.blob, spec, 1, keep)
    eb.qpos[0, :7] = qpos
    eb.run(E.MODE_FORWARD)
    return eb


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_source_known_answer(case):
    name, xml, qpos, ncon, dists, normal = case
    eb = _emu_forward(xml, qpos)
    n = int(eb.ncon[0])
    assert n == ncon, (name, eb.contact_dist[0, :n])
    assert np.allclose(sorted(eb.contact_dist[0, :n]), sorted(dists), atol=2e-6)
    assert int(eb.ncon_dropped[0]) == 0


def test_contact_positions_box_overhanging():
    """the clipped polygon itself: corners (0.8, +-0.3) of the small box and the crossings (1.0, +-0.3) with the edge"""
    name, xml, qpos, *_ = [c for c in CASES if c[0] == "box_overhanging"][0]
    sim = OracleSim(L.Model(xml).blob)
    sim.qpos[:] = qpos
    sim.forward()
    pts = sorted((round(float(sim.contact(i)["pos"][0]), 9), round(float(sim.contact(i)["pos"][1]), 9)) for i in range(sim.ncon))
    assert pts == [(0.8, -0.3), (0.8, 0.3), (1.0, -0.3), (1.0, 0.3)]
    assert all(abs(sim.contact(i)["pos"][2] - (0.5 + 0.002)) < 1e-12 for i in range(4))


def test_static_box_on_box_force_balance():
    """a box resting on a box: the four contact forces carry its weight (oracle and kernel source agree on the
    settled penetration to 1e-5)"""
    from collision_cases import BIGBOX, SMALL, scene
    xml = scene(BIGBOX, SMALL)
    model = L.Model(xml)
    sim = OracleSim(model.blob)
    sim.qpos[:] = [0.1, 0.2, 0.5 + 0.1 + 0.001, 1, 0, 0, 0]
    for _ in range(1500):
        sim.step()
    assert sim.ncon == 4 and abs(sim.qvel).max() < 1e-6
    z_ref = sim.qpos[2]
    assert 0.5 + 0.1 - 0.01 < z_ref < 0.5 + 0.1 + 0.01
    spec, keep = E.simple_spec(model, [1], [0, 1, 5], 3, free_joint=False)
    spec.n_phys_act = 0
    eb = E.EmuBatch(model.blob, spec, 1, keep)
    eb.qpos[0, :7] = [0.1, 0.2, 0.5 + 0.1 + 0.001, 1, 0, 0, 0]
    for _ in range(60):
        eb.run(E.MODE_PHYSICS, 25)
    assert abs(eb.qpos[0, 2] - z_ref) < 1e-5 and int(eb.ncon[0]) == 4


def random_poses(static, free, n, seed):
    """poses of the free geom scattered around the surface of the static box (touching, penetrating, just clear)"""
    from collision_cases import scene
    rng = np.random.default_rng(seed)
    model = L.Model(scene(static, free))
    sim = OracleSim(model.blob)
    qs = []
    while len(qs) < n:
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        p = rng.uniform(-1, 1, 3) * np.array([0.7, 0.6, 0.5])
        ax = rng.integers(0, 3)
        p[ax] = rng.choice([-1, 1]) * (np.array([0.5, 0.4, 0.3])[ax] + rng.uniform(0.0, 0.45))
        sim.qpos[:] = list(p) + list(q)
        sim.forward()
        if sim.ncon == 0 and rng.uniform() < 0.9:
            continue
        qs.append(np.array(list(p) + list(q), dtype=np.float32))
    return model, sim, qs


RANDOM_PAIRS = [("capsule-box", '<geom name="S" type="box" size="0.5 0.4 0.3"/>', '<geom name="G" type="capsule" size="0.08 0.25"/>'),
                ("box-box", '<geom name="S" type="box" size="0.5 0.4 0.3"/>', '<geom name="G" type="box" size="0.15 0.2 0.1"/>')]


@pytest.mark.parametrize("kind,static,free", RANDOM_PAIRS, ids=[p[0] for p in RANDOM_PAIRS])
def test_random_poses_kernel_source_agrees_with_oracle(kind, static, free):
    """two different algorithms, one specification: identical contact counts and distances on random poses"""
    n = 160
    model, sim, qs = random_poses(static, free, n, seed=11)
    spec, keep = E.simple_spec(model, [1], [0, 1, 5], 3, free_joint=True)
    eb = E.EmuBatch(model.blob, spec, n, keep)
    eb.qpos[:, :7] = np.array(qs)
    eb.run(E.MODE_FORWARD)
    multi = 0
    for e, q in enumerate(qs):
        sim.qpos[:] = q.astype(np.float64)
        sim.forward()
        want = sorted(sim.contact(i)["dist"] for i in range(sim.ncon))
        got = sorted(eb.contact_dist[e, :eb.ncon[e]])
        assert len(want) == len(got), (kind, e, want, got)
        if want:
            assert np.abs(np.array(want) - np.array(got)).max() < 2e-5, (kind, e, want, got)
        multi += len(want) >= 2
    assert multi > 20
