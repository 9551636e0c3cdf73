"""Upper face of the drop-in boundary: `env.data` / `env.model` views, plugin recognition by source, orientation and
position queries (reference call sites mujoco_parent.py:394-449, fps_custom_env.py:4-27, README.md:108-173).
CPU part: the pure-host helpers.  GPU part (marked): the reference's plugin source text, held verbatim under
tests/golden/ref_plugins/, fed to MuJoCoRL unmodified."""
import os

import numpy as np
import pytest
import torch

from common import LEVELS
from mujoco_rl_environment_wrapper_b200 import _lib as L
from mujoco_rl_environment_wrapper_b200 import plugins as P
from mujoco_rl_environment_wrapper_b200.dataview import quat_to_euler_zyx_deg, static_pose

REFP = os.path.join(os.path.dirname(__file__), "golden", "ref_plugins")


def load_ref(name):
    """the fixture as a real module (inspect.getsource must find the text, as it does for a user's own file)"""
    import importlib.util
    import sys
    path = os.path.join(REFP, name)
    spec = importlib.util.spec_from_file_location("ref_plugin_" + name[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return vars(mod)


def test_euler_matches_scipy():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    q = rng.normal(size=(200, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    want = Rotation.from_quat(q[:, [1, 2, 3, 0]]).as_euler("zyx", degrees=True)   # helper.py:6-18
    got = quat_to_euler_zyx_deg(q)
    assert np.abs(got - want).max() < 1e-9
    got_t = quat_to_euler_zyx_deg(torch.tensor(q)).numpy()
    assert np.abs(got_t - want).max() < 1e-9


def test_static_pose_matches_oracle_kinematics():
    from oracle import OracleSim
    text = open(os.path.join(LEVELS, "two_ants.xml")).read()
    model = L.Model(text)
    sim = OracleSim(model.blob)
    sim.reset(); sim.forward()
    f = model.fields
    n_static = 0
    for b in range(1, model.nbody):
        sp = static_pose(f, L.OBJ_BODY, b)
        if sp is not None:
            assert np.abs(sp[0] - sim.xipos[b]).max() < 1e-12
            n_static += 1
    for g in range(model.ngeom):
        sp = static_pose(f, L.OBJ_GEOM, g)
        if sp is not None:
            assert np.abs(sp[0] - sim.geom_xpos[g]).max() < 1e-12
    assert n_static >= 8 and static_pose(f, L.OBJ_BODY, model.name2id(L.OBJ_BODY, "sender")) is None


def test_torch_draw_stream_equals_the_kernels():
    lib = L.load()
    env_ids = torch.arange(0, 5000, 7, dtype=torch.int64)
    for seed, agent, counter in ((1234, 0, 0), (99, 1, 3), (2 ** 63 + 12345, 7, 250)):
        cnt = torch.full_like(env_ids, counter)
        got = P.draw_u32(seed, env_ids, agent, cnt).tolist()
        want = [lib.mjb_draw_u32(seed, int(e), agent, counter) for e in env_ids.tolist()]
        assert got == want


def test_reference_plugin_sources_are_recognised():
    ant = load_ref("ant_reward_function.py")["ant_reward_function"]
    lang = load_ref("readme_language.py")["Language"]
    rd = load_ref("readme_reward_done.py")
    assert P.recognise(ant)[:2] == ("reward", L.REW_ANT)
    assert P.recognise(lang)[:2] == ("dynamic", L.DYN_LANGUAGE)
    assert P.recognise(rd["reward_function"]) == ("reward", L.REW_TAG_DISTANCE, {"scale": 10.0})
    assert P.recognise(rd["done_function"]) == ("done", L.DONE_DISTANCE_LE, {"threshold": 1.0})
    # comments, blank lines and the function's name do not matter; a changed statement does
    src = open(os.path.join(REFP, "readme_reward_done.py")).read()
    h0 = P.normalised_source_hash(src[src.index("def done_function"):])
    assert h0 == P.normalised_source_hash("def my_done(mujoco_gym, agent):\n\n  # end?\n  if mujoco_gym.data_store[agent][\"distance\"] <= 1:\n"
                                          "    return True\n  else:\n    return False\n")
    assert h0 != P.normalised_source_hash(src[src.index("def done_function"):].replace("<= 1", "<= 2"))
    assert P.recognise(lambda env, agent: 0.0) is None and P.recognise(P.tag_distance_reward)[:2] == ("reward", L.REW_TAG_DISTANCE)


# ------------------------------------------------------------------------------------------------------------ GPU
def _env(**kw):
    from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
    cfg = {"xmlPath": os.path.join(LEVELS, "two_ants.xml"), "infoJson": os.path.join(LEVELS, "info_2A.json"), "agents": ["sender", "receiver"]}
    cfg.update(kw)
    return MuJoCoRL(cfg)


@pytest.mark.gpu
def test_verbatim_reference_plugins_take_the_fused_path():
    """the README's Language / reward / done and fps_custom_env's ant reward, source text unmodified, give exactly what
    the package's own (marked) plugins give, for a batch"""
    lang = load_ref("readme_language.py")["Language"]
    rd = load_ref("readme_reward_done.py")
    N = 64
    a = _env(num_envs=N, seed=5, environmentDynamics=[lang], rewardFunctions=[rd["reward_function"]], doneFunctions=[rd["done_function"]])
    b = _env(num_envs=N, seed=5, environmentDynamics=[P.Language], rewardFunctions=[P.tag_distance_reward], doneFunctions=[P.distance_done])
    assert a._spec.n_dynamics == a._spec.n_rewards == a._spec.n_dones == 1 and not a._host_rew and not a._host_dyn
    a.reset(); b.reset()
    for t in range(30):
        act = b.sample_actions()
        assert torch.equal(act, a.sample_actions())
        ra, rb = a.step(act), b.step(act)
        for x, y in zip(ra[:4], rb[:4]):
            for k in y:
                assert torch.equal(x[k], y[k]), (t, k)
    ant = load_ref("ant_reward_function.py")["ant_reward_function"]
    cfg = dict(xmlPath=os.path.join(LEVELS, "ant_rk4.xml"), infoJson=None, agents=["torso"], num_envs=N, seed=3)
    a, b = _env(**cfg, rewardFunctions=[ant]), _env(**cfg, rewardFunctions=[P.ant_reward_function])
    a.reset(); b.reset()
    for t in range(10):
        act = b.sample_actions(); a.sample_actions()
        ra, rb = a.step(act), b.step(act)
        assert torch.equal(ra[1]["torso"], rb[1]["torso"]) and float(rb[1]["torso"].abs().max()) > 0 or t == 0


@pytest.mark.gpu
def test_host_plugins_through_the_data_shim_single_env():
    """num_envs = 1: reference-style Python plugins that are NOT recognised (here: the reference's ant reward with one
    extra statement) run on the host against env.data / env.model / env.get_data with the reference's shapes, and give
    the fused path's reward"""
    ns = load_ref("ant_reward_function.py")
    src = open(os.path.join(REFP, "ant_reward_function.py")).read().replace("    dt = env.model.opt.timestep", "    dt = float(env.model.opt.timestep)")
    g = {}
    exec(compile(src, "<modified>", "exec"), g)
    user_fn = g["ant_reward_function"]
    assert P.recognise(user_fn) is None
    cfg = dict(xmlPath=os.path.join(LEVELS, "ant_rk4.xml"), infoJson=None, agents=["torso"], seed=3)
    a, b = _env(**cfg, rewardFunctions=[user_fn]), _env(**cfg, rewardFunctions=[ns["ant_reward_function"]])
    assert a._host_rew and not b._host_rew
    a.reset(); b.reset()
    rng = np.random.default_rng(1)
    for t in range(12):
        act = {"torso": rng.uniform(-1, 1, 8).astype(np.float32)}
        ra, rb = a.step(act), b.step(act)
        assert isinstance(ra[1]["torso"], float) or isinstance(ra[1]["torso"], int)
        assert abs(ra[1]["torso"] - rb[1]["torso"]) <= 1e-6 * max(1.0, abs(rb[1]["torso"])), (t, ra[1], rb[1])
    d = a.data
    assert d.qpos.shape == (15,) and d.ctrl.shape == (8,) and d.cfrc_ext.shape == (a.model.nbody, 6) and a.model.opt.timestep == 0.01
    assert d.body("torso").xipos.shape == (3,) and abs(d.body("torso").xipos[2] - a.get_data("torso")["position"][2]) == 0
    assert a.model.body("torso").mass.shape == (1,) and isinstance(d.ncon, int) and len(d.contact) == d.ncon


@pytest.mark.gpu
def test_torch_plugins_equal_the_fused_kernel():
    """the package's plugins as batched torch code (forced onto the host path by a leading user plugin) against the fused
    epilogue: identical rewards, flags, store columns and Language / Pick_Up observations"""
    class Nop:
        def __init__(self, env):
            self.env, self.observation_space, self.action_space = env, {"low": [], "high": []}, {"low": [], "high": []}

        def dynamic(self, agent, actions):
            n = self.env.num_envs
            return 0, torch.zeros(n, 0, device=self.env.device), torch.zeros(n, dtype=torch.bool, device=self.env.device), {}

    nop_r = lambda env, agent: torch.zeros(env.num_envs, device=env.device)
    nop_d = lambda env, agent: torch.zeros(env.num_envs, dtype=torch.bool, device=env.device)
    N = 128
    for dyn in (P.Language, P.PickUpDynamic):
        host = _env(num_envs=N, seed=9, environmentDynamics=[Nop, dyn], rewardFunctions=[nop_r, P.tag_distance_reward], doneFunctions=[nop_d, P.distance_done])
        fused = _env(num_envs=N, seed=9, environmentDynamics=[dyn], rewardFunctions=[P.tag_distance_reward], doneFunctions=[P.distance_done])
        assert host._host_dyn and host._host_rew and host._host_done and fused._spec.n_dynamics == 1
        host.reset(); fused.reset()
        events = 0
        for t in range(40):
            act = fused.sample_actions()
            host.sample_actions()
            if t == 10:   # bring some agents next to a target so that done / pick-up events happen
                for e in (host, fused):
                    e.batch.qpos[::4, 0:3] = torch.tensor([6.3, -2.0, 0.8], device=e.device)
            rh, rf = host.step(act), fused.step(act)
            for a in fused.agents:
                assert torch.equal(rh[0][a], rf[0][a]), (dyn.__name__, t, a, "obs")
                assert torch.equal(rh[1][a].float(), rf[1][a]), (dyn.__name__, t, a, "reward")
                assert torch.equal(rh[2][a], rf[2][a]) and torch.equal(rh[3][a], rf[3][a])
            assert torch.equal(rh[2]["__all__"], rf[2]["__all__"])
            assert host.batch.store_f.cpu().numpy().tobytes() == fused.batch.store_f.cpu().numpy().tobytes()
            assert torch.equal(host.batch.store_i, fused.batch.store_i)
            events += int(rf[2]["__all__"].sum()) + int(fused.batch.store_i[:, :, L.STORE_I["inventory"]].sum())
        assert events > 0, "the run must contain done / pick-up events"


@pytest.mark.gpu
def test_get_data_orientation_and_arbitrary_names():
    """get_data()["orientation"] (zyx Euler of xmat, mujoco_parent.py:407,419) and distance() / data.body(n).xipos for objects
    that are neither agents nor targets: static ones are constants, moving ones are exported on request"""
    from oracle import OracleSim
    env = _env(num_envs=3, exportPositions=["aux_1", "left_ankle_geom"], exportOrientation=True)
    env.reset()
    for _ in range(5):
        env.step(env.sample_actions())
    # the exported values belong to the LAST forward pass of the step, i.e. to the state before the last integration:
    # reproduce that with the oracle: mj_forward at the current state equals what the NEXT step exports
    sim = OracleSim(env.model.blob)
    e = 1
    sim.qpos[:] = env.batch.qpos[e, :30].double().cpu().numpy()
    sim.qvel[:] = env.batch.qvel[e, :28].double().cpu().numpy()
    sim.forward()
    env.step(env.sample_actions())
    from scipy.spatial.transform import Rotation

    def euler(mat):
        return Rotation.from_matrix(np.asarray(mat).reshape(3, 3)).as_euler("zyx", degrees=True)   # helper.py:6-18

    m = env.model
    for name in ("sender", "receiver", "choice_1", "reference", "aux_1"):
        d = env.get_data(name)
        bid = m.name2id(L.OBJ_BODY, name)
        assert d["type"] == "body" and d["id"] == bid
        assert np.abs(d["position"][e].cpu().numpy() - sim.xipos[bid]).max() < 2e-5, name
        assert np.abs(d["orientation"][e].cpu().numpy() - euler(sim.xmat[bid])).max() < 2e-3, name
    g = env.get_data("left_ankle_geom")
    gid = m.name2id(L.OBJ_GEOM, "left_ankle_geom")
    assert g["type"] == "geom" and np.abs(g["position"][e].cpu().numpy() - sim.geom_xpos[gid]).max() < 2e-5
    assert np.abs(g["orientation"][e].cpu().numpy() - euler(sim.geom_xmat[gid])).max() < 2e-3
    assert env.distance("reference", "aux_1").shape == (3,)
    assert abs(float(env.distance("reference", "aux_1")[e]) - np.linalg.norm(sim.xipos[m.name2id(L.OBJ_BODY, "reference")] - sim.xipos[m.name2id(L.OBJ_BODY, "aux_1")])) < 1e-4
    xm = env.data.body("aux_1").xmat
    assert xm.shape == (3, 9) and np.abs(xm[e].cpu().numpy() - sim.xmat[m.name2id(L.OBJ_BODY, "aux_1")]).max() < 1e-4
    with pytest.raises(Exception, match="exportPositions"):
        env.distance("sender", "aux_2")
    # everything at once
    env_all = _env(num_envs=2, exportPositions="all")
    env_all.reset()
    env_all.step(env_all.sample_actions())
    assert env_all.distance("aux_4_2", "back_leg").shape == (2,)


@pytest.mark.gpu
def test_done_threshold_and_reward_are_the_references_fp64_arithmetic():
    """Agents placed within +-1e-6 of the done threshold (distance 1 to their target): the done flag and the reward are
    exactly what the reference's arithmetic (math.dist on float64 views, README.md:149-173) gives on the fp32 positions
    the step exported — d <= 1 decided in fp64, reward = float32(10 * (d_prev - d))."""
    N, A = 512, 2
    env = _env(num_envs=N, seed=17, environmentDynamics=[P.Language], rewardFunctions=[P.tag_distance_reward], doneFunctions=[P.distance_done])
    env.reset()
    b = env.batch
    env.step(env.sample_actions())          # first call per agent: target drawn, distance stored, reward 0
    torch.cuda.synchronize()
    names = env._probe_names
    tgt_probe = torch.tensor([names.index(t) for t in env._target_names], device=env.device)
    rows = torch.arange(N, device=env.device)
    deltas = torch.linspace(-1e-6, 1e-6, N, device=env.device, dtype=torch.float64)
    d_prev, tgt_idx = {}, {}
    for a in range(A):
        tgt = b.store_i[:, a, L.STORE_I["current_target"]].long() - 1
        assert bool((tgt >= 0).all())
        tgt_idx[a] = tgt_probe[tgt]
        tp = b.probe[rows, tgt_idx[a], :3].double()
        pos = b.qpos[:, 15 * a:15 * a + 3].double()
        u = pos - tp
        u = u / u.norm(dim=1, keepdim=True)
        b.qpos[:, 15 * a:15 * a + 3] = (tp + u * (1.0 + deltas)[:, None]).float()
        d_prev[a] = b.store_f.view(torch.float64)[:, a, 1].clone()
    obs, rew, term, trunc, _ = env.step(env.sample_actions())
    torch.cuda.synchronize()
    probe = b.probe.cpu().numpy()
    both = set()
    for a, name in enumerate(env.agents):
        pa = probe[:, a, :3].astype(np.float64)
        pt = probe[np.arange(N), tgt_idx[a].cpu().numpy(), :3].astype(np.float64)
        dx, dy, dz = (pa - pt).T
        d = np.sqrt((dx * dx + dy * dy) + dz * dz)
        assert np.abs(d - 1.0).max() < 5e-6, "the agents must sit at the threshold"
        want_done = d <= 1.0
        want_rew = (10.0 * (d_prev[a].cpu().numpy() - d)).astype(np.float32)
        assert np.array_equal(term[name].cpu().numpy(), want_done), name
        assert np.array_equal(rew[name].cpu().numpy(), want_rew), name
        both |= set(want_done.tolist())
    assert both == {True, False}, "both sides of the threshold must occur"
