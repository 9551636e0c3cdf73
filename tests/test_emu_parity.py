"""The device step code (csrc/step_kernel.cuh + env_kernel.cuh), compiled for the host SIMT emulator,
against the fp64 oracle from identical states.  Runs without a GPU; the same comparisons run through
the C-ABI on the B200 in test_gpu_parity.py.  Tolerance: fp32 state after one step within
rtol = 1e-4 of the fp64 result (|x_gpu - x_ref| <= 1e-4 * max(1, |x_ref|)), contact-pair sets exact."""
import numpy as np
import pytest

import emu_harness as E
from common import load_scene, make_spec, oracle_states
from mujoco_rl_environment_wrapper_b200 import _lib as L
from oracle import host_loop as H

RTOL = 1e-4


def rel_err(x, ref):
    return float((np.abs(x - ref) / np.maximum(1.0, np.abs(ref))).max()) if len(ref) else 0.0


@pytest.mark.parametrize("scene,n,stride,settle", [("2A", 24, 25, 150), ("1A", 10, 20, 40), ("S3", 12, 20, 100), ("3S", 8, 30, 150)])
def test_one_step_parity(scene, n, stride, settle):
    model, tables, agents, fj = load_scene(scene)
    spec, keep = make_spec(model, tables, agents, fj)
    states = oracle_states(model, tables, agents, fj, n, stride=stride, settle=settle)
    eb = E.EmuBatch(model.blob, spec, len(states), keep)
    nq, nv, nu, ns = model.nq, model.nv, model.nu, model.nsensordata
    n_phys = spec.n_phys_act
    for e, (pre, post) in enumerate(states):
        eb.qpos[e, :nq], eb.qvel[e, :nv], eb.warmstart[e, :nv] = pre[0], pre[1], pre[2]
        if nu:
            eb.ctrl[e, :nu] = pre[3]
        eb.actions[e, :, :n_phys] = pre[4]
    eb.run(E.MODE_PHYSICS, 1)
    seen_contacts = 0
    for e, (pre, post) in enumerate(states):
        assert rel_err(eb.qpos[e, :nq], post["qpos"]) < RTOL
        assert rel_err(eb.qvel[e, :nv], post["qvel"]) < RTOL
        if ns:
            assert rel_err(eb.sensordata[e, :ns], post["sensordata"][:ns]) < RTOL
        nc = eb.ncon[e]
        assert sorted((int(a), int(b)) for a, b in eb.contact_geom[e, :nc]) == post["pairs"]
        seen_contacts += nc
    assert seen_contacts > 0, "the sample must exercise contacts"


def test_lane_order_independence():
    """Running the lanes in reverse order must not change a single bit (missing-barrier detector)."""
    model, tables, agents, fj = load_scene("2A")
    spec, keep = make_spec(model, tables, agents, fj)
    states = oracle_states(model, tables, agents, fj, 4, stride=60, settle=200)
    outs = []
    for rev in (False, True):
        eb = E.EmuBatch(model.blob, spec, len(states), keep)
        for e, (pre, post) in enumerate(states):
            eb.qpos[e, :model.nq], eb.qvel[e, :model.nv], eb.warmstart[e, :model.nv] = pre[0], pre[1], pre[2]
            eb.actions[e, :, :8] = pre[4]
        eb.run(E.MODE_PHYSICS, 2, reverse=rev)
        outs.append((eb.qpos.copy(), eb.qvel.copy(), eb.sensordata.copy()))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def _resolver(model):
    def resolve(name):
        b = model.name2id(L.OBJ_BODY, name)
        return (1, b) if b >= 0 else (5, model.name2id(L.OBJ_GEOM, name))
    return resolve


def test_full_step_with_plugins_against_host_loop():
    """C2: Language + tag-distance reward + done, state re-synchronised every step so that the epilogue
    (ordering, draws, flags) is compared exactly."""
    model, tables, agents, fj = load_scene("2A")
    lib = L.load()
    targets = ["choice_1", "choice_2"]
    tspec = [(L.OBJ_BODY, model.name2id(L.OBJ_BODY, t)) for t in targets]
    seed, max_steps = 99, 6
    spec, keep = make_spec(model, tables, agents, fj, max_steps=max_steps, dynamics=[(L.DYN_LANGUAGE, 1, 1, 0.0)],
                           rewards=[(L.REW_TAG_DISTANCE, 10.0)], dones=[(L.DONE_DISTANCE_LE, 1.0)], targets=tspec, seed=seed)
    env = H.OracleEnv(model, tables, agents, max_steps=max_steps, dynamics=[H.Language],
                      reward_functions=[H.tag_distance_reward], done_functions=[H.distance_done], targets=targets,
                      draw=lambda a, k: lib.mjb_draw_u32(seed, 0, a, k), resolve=_resolver(model))
    eb = E.EmuBatch(model.blob, spec, 1, keep)
    rng = np.random.default_rng(11)
    act0 = {a: np.concatenate([rng.uniform(-1, 1, 8), rng.uniform(0, 3, 1)]) for a in agents}
    obs, _ = env.reset(act0)
    eb.actions[0, :, :9] = np.stack([act0[a] for a in agents])
    eb.run(E.MODE_RESET)
    for i, a in enumerate(agents):
        assert rel_err(eb.obs[0, i, :60], obs[a]) < RTOL
        assert eb.obs[0, i, 59] == obs[a][59]
    assert not eb.store_i[0, :, :5].any() and eb.timestep[0] == 0
    nq, nv = model.nq, model.nv
    for t in range(9):
        act = {a: np.concatenate([rng.uniform(-1, 1, 8), rng.uniform(0, 3, 1)]) for a in agents}
        eb.qpos[0, :nq], eb.qvel[0, :nv], eb.warmstart[0, :nv] = env.sim.qpos, env.sim.qvel, env.sim.qacc_warmstart
        eb.actions[0, :, :9] = np.stack([act[a] for a in agents])
        o, r, term, trunc, _ = env.step(act)
        eb.run(E.MODE_STEP, 1)
        for i, a in enumerate(agents):
            assert rel_err(eb.obs[0, i, :59], o[a][:59]) < RTOL
            assert eb.obs[0, i, 59] == o[a][59]                       # Language: integer, exact
            assert abs(eb.reward[0, i] - r[a]) < 1e-3                 # 10 * delta of fp32 distances
            assert bool(eb.term[0, i]) == term[a] and bool(eb.trunc[0, i]) == trunc[a]
        assert bool(eb.term[0, 2]) == term["__all__"] and bool(eb.trunc[0, 2]) == trunc["__all__"]
        assert eb.timestep[0] == env.timestep
        tgt = [targets.index(env.data_store[a]["current_target"]) + 1 for a in agents]
        assert list(eb.store_i[0, :, L.STORE_I["current_target"]]) == tgt


def test_packed_envs_ragged_count_and_masked_reset():
    """Env packing (csrc/replicate.h): an Ant has 14 dofs, so two real envs share a warp.  Five envs (the last
    virtual env holds a single copy), ant reward in the epilogue, then a masked reset of envs 1 and 4 only."""
    model, tables, agents, fj = load_scene("1A")
    spec, keep = make_spec(model, tables, agents, fj, skip_frames=2, rewards=[(L.REW_ANT, 0.0)], max_steps=4)
    N = 5
    eb = E.EmuBatch(model.blob, spec, N, keep)

    def resolve(name):
        b = model.name2id(L.OBJ_BODY, name)
        return (1, b) if b >= 0 else (5, model.name2id(L.OBJ_GEOM, name))
    envs = [H.OracleEnv(model, tables, agents, skip_frames=2, max_steps=4, reward_functions=[H.ant_reward], resolve=resolve)
            for _ in range(N)]
    eb.run(E.MODE_RESET)
    for e in envs:
        e.reset({"torso": np.zeros(8)})
    rng = np.random.default_rng(17)
    for t in range(6):
        act = rng.uniform(-1, 1, (N, 1, 8))
        for e, env in enumerate(envs):
            eb.qpos[e, :15], eb.qvel[e, :14], eb.ctrl[e, :8] = env.sim.qpos, env.sim.qvel, env.sim.ctrl
        eb.actions[:, :, :8] = act
        eb.run(E.MODE_STEP, 2)
        for e, env in enumerate(envs):
            o, r, term, trunc, _ = env.step({"torso": act[e, 0]})
            assert rel_err(eb.obs[e, 0, :29], o["torso"]) < 3 * RTOL, (t, e)
            assert abs(eb.reward[e, 0] - r["torso"]) < 2e-2 * max(1.0, abs(r["torso"])), (t, e)
            assert bool(eb.trunc[e, 0]) == trunc["torso"] and eb.timestep[e] == env.timestep
    q_before = eb.qpos.copy()
    mask = np.array([0, 1, 0, 0, 1], np.uint8)
    eb.run(E.MODE_RESET, mask=mask)
    for e in range(N):
        if mask[e]:
            assert np.allclose(eb.qpos[e, :15], model.fields["qpos0"], atol=1e-6) and eb.timestep[e] == 0
            assert not eb.store_i[e, :, :5].any()
        else:
            assert np.array_equal(eb.qpos[e], q_before[e]) and eb.timestep[e] == 6


def test_non_finite_state_resets_only_that_env():
    """mj_checkPos / mj_checkVel behaviour: an env whose state became NaN / absurd is reset to qpos0 (counted in
    `nreset`); its warp neighbour in a packed pair is untouched."""
    model, tables, agents, fj = load_scene("1A")
    spec, keep = make_spec(model, tables, agents, fj)
    eb = E.EmuBatch(model.blob, spec, 3, keep)
    eb.run(E.MODE_RESET)
    eb.actions[:, :, :8] = 0.3
    eb.run(E.MODE_PHYSICS, 3)
    healthy = eb.qpos.copy()
    eb.qvel[1, 3] = np.nan
    eb.qpos[2, 0] = 3e12
    eb.run(E.MODE_PHYSICS, 1)
    assert list(eb.nreset) == [0, 1, 1]
    for e in (1, 2):
        assert np.allclose(eb.qpos[e, :15], model.fields["qpos0"], atol=1e-6) and not eb.qvel[e].any()
    assert np.isfinite(eb.qpos).all() and np.isfinite(eb.qvel).all()
    assert not np.array_equal(eb.qpos[0], healthy[0]) and abs(eb.qpos[0, 2] - healthy[0, 2]) < 0.1


@pytest.mark.parametrize("scene,n", [("2A", 6), ("1A", 5)])
def test_reset_noise_is_bounded_seeded_and_off_by_default(scene, n):
    """`reset_noise` (config "resetNoise", new: the reference always restarts at qpos0): hinge / slide qpos within
    +-noise of qpos0, free joints untouched, qvel within +-noise; different per env and per reset, identical for the
    same seed; also with two envs per warp ("1A") and a masked reset."""
    model, tables, agents, fj = load_scene(scene)
    f = model.fields
    qpos0 = f["qpos0"]
    hinge = np.array([int(f["jnt_qposadr"][j]) for j in range(model.njnt) if int(f["jnt_type"][j]) != L.JNT_FREE])
    other = np.setdiff1d(np.arange(model.nq), hinge)

    def fresh(noise, seed=3):
        spec, keep = make_spec(model, tables, agents, fj)
        spec.reset_noise, spec.seed = noise, seed
        eb = E.EmuBatch(model.blob, spec, n, keep)
        eb.run(E.MODE_RESET)
        return eb
    plain = fresh(0.0)
    assert np.allclose(plain.qpos[:, :model.nq], qpos0, atol=1e-6) and not plain.qvel.any()
    a, b, c = fresh(0.1), fresh(0.1), fresh(0.1, seed=4)
    d = a.qpos[:, :model.nq] - qpos0
    assert np.abs(d[:, hinge]).max() <= 0.1 + 1e-6 and np.abs(d[:, hinge]).min() > 0
    assert np.abs(d[:, other]).max() < 1e-6
    assert np.abs(a.qvel[:, :model.nv]).max() <= 0.1 + 1e-6 and np.abs(a.qvel[:, :model.nv]).min() > 0
    assert len({a.qpos[e].tobytes() for e in range(n)}) == n
    assert np.array_equal(a.qpos, b.qpos) and np.array_equal(a.qvel, b.qvel)
    assert not np.array_equal(a.qpos, c.qpos)
    # the observation returned by reset is that of the perturbed state
    od = a.spec.obs_dim[0]
    assert rel_err(a.obs[0, 0, od - model.nv:od], a.qvel[0, :model.nv]) < 1e-6
    # a later (masked) reset draws again and leaves the other envs alone
    first = a.qpos.copy()
    n_phys = a.spec.n_phys_act
    a.actions[:, :, :n_phys] = 0.3
    for _ in range(3):
        a.run(E.MODE_STEP)
    stepped = a.qpos.copy()
    mask = np.zeros(n, np.uint8)
    mask[1] = mask[n - 1] = 1
    a.run(E.MODE_RESET, mask=mask)
    for e in range(n):
        if mask[e]:
            assert np.abs(a.qpos[e, hinge] - qpos0[hinge]).max() <= 0.1 + 1e-6
            assert not np.array_equal(a.qpos[e], first[e])
        else:
            assert np.array_equal(a.qpos[e], stepped[e])


def test_two_level_broad_phase_never_drops_a_contact():
    """The block-level cull (tree bounding sphere vs tree / static geom) is conservative: for ants teleported next to
    walls, target boxes and each other with random joint angles, the contact set equals the oracle's, which tests
    every pair."""
    model, tables, agents, fj = load_scene("2A")
    spec, keep = make_spec(model, tables, agents, fj)
    f = model.fields
    rng = np.random.default_rng(42)
    N = 48
    eb = E.EmuBatch(model.blob, spec, N, keep)
    eb.run(E.MODE_RESET)
    hinge = [int(f["jnt_qposadr"][j]) for j in range(model.njnt) if int(f["jnt_type"][j]) != L.JNT_FREE]
    free = [int(f["jnt_qposadr"][j]) for j in range(model.njnt) if int(f["jnt_type"][j]) == L.JNT_FREE]
    # places of interest: arena walls at about +-10 m, the three target boxes, mid-field
    spots = [(9.3, 0.0), (-9.4, 1.0), (0.5, 4.3), (0.2, -4.4), (7.0, -2.0), (1.4, -2.1), (-5.5, -2.4), (0.0, 0.0), (3.0, 3.0)]
    qpos = np.tile(f["qpos0"], (N, 1))
    for e in range(N):
        sx, sy = spots[rng.integers(len(spots))]
        for k, qa in enumerate(free):
            if k == 0 or rng.random() < 0.5:
                qpos[e, qa:qa + 2] = [sx + rng.uniform(-1.2, 1.2), sy + rng.uniform(-1.2, 1.2)]
            else:   # the second ant right next to the first
                qpos[e, qa:qa + 2] = qpos[e, free[0]:free[0] + 2] + rng.uniform(-0.9, 0.9, 2)
            qpos[e, qa + 2] = rng.uniform(0.3, 0.9)
            q = rng.normal(size=4)
            qpos[e, qa + 3:qa + 7] = q / np.linalg.norm(q)
        qpos[e, hinge] += rng.uniform(-0.6, 0.6, len(hinge))
    eb.qpos[:, :model.nq] = qpos
    eb.qvel[:] = 0
    eb.run(E.MODE_FORWARD)
    sim = OracleSimLazy(model)
    total = 0
    kinds = set()
    for e in range(N):
        sim.array("qpos")[:] = qpos[e]
        sim.array("qvel")[:] = 0
        sim.forward()
        want = sorted(sim.contact_pairs())
        nc = eb.ncon[e]
        got = sorted((int(a), int(b)) for a, b in eb.contact_geom[e, :nc])
        if len(want) <= eb.layout.maxcon:
            assert got == want, (e, got, want)
        else:
            assert set(got) <= set(want) and len(got) == eb.layout.maxcon
        total += len(want)
        for a, b in want:
            kinds.add((int(f["geom_type"][a]), int(f["geom_type"][b])))
    assert total > 60 and len(kinds) >= 4, (total, kinds)   # floor, walls / boxes, ant-ant contacts all occur


def OracleSimLazy(model):
    from oracle.sim import OracleSim
    return OracleSim(model.blob)


def test_inter_ant_contacts_single_block_factorisation():
    """Ants dropped onto each other: contacts between the two kinematic trees couple them in the Newton Hessian, which is
    then ONE 28 x 28 block (register Cholesky over all dofs instead of two 14 x 14 L D L' blocks).  One step from identical
    states vs the oracle, kernel source on the emulator."""
    from oracle import OracleSim
    model, tables, agents, fj = load_scene("2A")
    spec, keep = make_spec(model, tables, agents, fj)
    sim = OracleSim(model.blob)
    rng = np.random.default_rng(21)
    idx = np.array(tables.agents_action_index["sender"] + tables.agents_action_index["receiver"])
    geom_body, root = model.fields["geom_bodyid"], model.fields["body_rootid"]
    states, cross = [], 0
    for trial in range(3):
        sim.reset()
        sim.qpos[15:18] = sim.qpos[0:3] + np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 0.45 + 0.2 * rng.random()])
        yaw = rng.uniform(0, np.pi)
        sim.qpos[18:22] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
        for t in range(200):
            c = rng.uniform(-1, 1, 16)
            sim.ctrl[idx] = c
            if t >= 60 and t % 20 == 0:
                pre = (sim.qpos.copy(), sim.qvel.copy(), sim.qacc_warmstart.copy(), sim.ctrl.copy(), c.reshape(2, 8).copy())
                sim.step()
                pairs = sim.contact_pairs()
                cross += any(root[geom_body[a]] != root[geom_body[b]] and root[geom_body[a]] and root[geom_body[b]] for a, b in pairs)
                states.append((pre, {"qpos": sim.qpos.copy(), "qvel": sim.qvel.copy(), "pairs": sorted(pairs)}))
            else:
                sim.step()
    assert cross > 5, cross
    eb = E.EmuBatch(model.blob, spec, len(states), keep)
    for e, (pre, post) in enumerate(states):
        eb.qpos[e, :30], eb.qvel[e, :28], eb.warmstart[e, :28] = pre[0], pre[1], pre[2]
        eb.ctrl[e, :16] = pre[3]
        eb.actions[e, :, :8] = pre[4]
    eb.run(E.MODE_PHYSICS, 1)
    for e, (pre, post) in enumerate(states):
        assert rel_err(eb.qpos[e, :30], post["qpos"]) < RTOL, e
        assert rel_err(eb.qvel[e, :28], post["qvel"]) < RTOL, e
        assert sorted((int(a), int(b)) for a, b in eb.contact_geom[e, :eb.ncon[e]]) == post["pairs"], e
