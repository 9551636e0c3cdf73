"""The reference's example plugins for the batched environment.

Signatures are the reference's: dynamics are classes with `__init__(self, mujoco_gym)`, `observation_space`,
`action_space` and `dynamic(self, agent, actions)`; reward / done functions are `f(mujoco_gym, agent)`.

Every plugin here exists twice, with identical results:
  * as a FUSED kind (`mjb_kind`): `MuJoCoRL` hands it to the CUDA epilogue (csrc/env_kernel.cuh::run_plugins), which
    runs the whole plugin list per env in the reference's order inside the step kernel;
  * as a real batched torch implementation (the function / method body below) over the same device store columns.
    It runs whenever the plugin cannot be fused — it follows a user plugin in its list (order matters), or the
    validators call it once at construction (mujoco_rl.py:114-169).

`recognise()` maps the reference's own source text (benchmarking/fps_gym/fps_custom_env.py:4-27, README.md:109-173,
held verbatim as fixtures under tests/golden/ref_plugins/) to the fused kind, so that a user who passes the
reference's functions unmodified gets the fused path for `num_envs > 1` too.
"""
import hashlib
import inspect
import io
import tokenize

import torch

from . import _lib as L

_M64 = (1 << 64) - 1


def _i64(c):
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(z, k):
    """logical right shift of int64 tensors (torch's >> is arithmetic)"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def draw_u32(seed, env_ids, agent, counter):
    """torch twin of csrc/env_kernel.cuh::draw_u32 (the counter-based stream replacing random.randint, README.md:154):
    int64 arithmetic wraps like the kernel's uint64"""
    z = (torch.full_like(env_ids, _i64(seed)) + _i64(0x9E3779B97F4A7C15) * (1 + env_ids) + _i64(0xBF58476D1CE4E5B9 * agent)
         + _i64(0x94D049BB133111EB) * counter)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    return _lsr(z, 32) & 0xFFFFFFFF


def _store(env):
    b = env._batch
    return b.store_i, b.store_f, b.store_f.view(torch.float64)   # fp64 view: column 1 = MJB_STORE_F_DISTANCE64


def _draw_target(env, a, mask):
    """for the envs in `mask`: a new random "target" index (1-based) from the per-(env, agent) stream"""
    si = env._batch.store_i
    ids = torch.arange(env.num_envs, device=env.device, dtype=torch.int64)
    cnt = si[:, a, L.STORE_I["draws"]].to(torch.int64)
    drawn = 1 + (draw_u32(env.seed, ids, a, cnt) % len(env._target_names)).to(torch.int32)
    si[:, a, L.STORE_I["draws"]] += mask.to(torch.int32)
    return drawn


def _dist_to_target(env, a, tgt):
    """fp64 distance agent a -> its target (1-based index tensor) from the exported fp32 positions"""
    pr = env._batch.probe.double()
    A = len(env.agents)
    tp = pr[torch.arange(env.num_envs, device=env.device), A + (tgt.long() - 1).clamp(min=0), :3]
    d = pr[:, a, :3] - tp
    # (dx*dx + dy*dy) + dz*dz as separate IEEE operations: the kernel evaluates the same expression without fused multiply-add
    return ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).sqrt(), tp


class Language:
    """README.md:108-137 (4-tuple form required by mujoco_rl.py:124,236)."""
    mjb_kind = ("dynamic", L.DYN_LANGUAGE)

    def __init__(self, mujoco_gym):
        self.mujoco_gym = mujoco_gym
        self.observation_space = {"low": [0], "high": [3]}
        self.action_space = {"low": [0], "high": [3]}

    def dynamic(self, agent, actions):
        env = self.mujoco_gym
        a = env.agents.index(agent)
        actions = torch.as_tensor(actions, dtype=torch.float32, device=env.device).reshape(-1, 1)
        utt = actions[:, 0].to(torch.int32).expand(env.num_envs)  # int(): truncation toward zero
        si = env._batch.store_i
        si[:, a, L.STORE_I["utterance"]] = utt
        si[:, a, L.STORE_I["has_utterance"]] = 1
        others = [i for i in range(len(env.agents)) if i != a]
        if not others:
            raise IndexError("Language needs a second agent")
        o = others[0]
        val = torch.where(si[:, o, L.STORE_I["has_utterance"]] != 0, si[:, o, L.STORE_I["utterance"]],
                          torch.zeros_like(utt)).to(torch.float32)
        return 0, val.reshape(-1, 1), torch.zeros(env.num_envs, dtype=torch.bool, device=env.device), {}


class PickUpDynamic:
    """Testing/Pick_Up_Dynamic.py:4-41 re-expressed per agent with `dynamic(agent, actions)`."""
    mjb_kind = ("dynamic", L.DYN_PICKUP)
    threshold = 2.0

    def __init__(self, mujoco_gym):
        self.mujoco_gym = mujoco_gym
        self.observation_space = {"low": [-70, -70, -70, 0], "high": [70, 70, 70, 1]}
        self.action_space = {"low": [], "high": []}

    def dynamic(self, agent, actions):
        env = self.mujoco_gym
        a = env.agents.index(agent)
        si, sf, sf64 = _store(env)
        N = env.num_envs
        reward = torch.zeros(N, dtype=torch.float64, device=env.device)
        obs = torch.zeros(N, 4, device=env.device)
        if not env._target_names:
            return reward, obs, torch.zeros(N, dtype=torch.bool, device=env.device), {}
        tgt = si[:, a, L.STORE_I["current_target"]]
        new = tgt == 0
        tgt = torch.where(new, _draw_target(env, a, new), tgt)
        d, _ = _dist_to_target(env, a, tgt)
        hit = d < self.threshold
        si[:, a, L.STORE_I["inventory"]] ^= hit.to(torch.int32)
        reward += hit.double()
        tgt = torch.where(hit, _draw_target(env, a, hit), tgt)
        d2, tp = _dist_to_target(env, a, tgt)
        sf64[:, a, 1] = torch.where(hit, d2, sf64[:, a, 1])
        sf[:, a, L.STORE_F["distance"]] = sf64[:, a, 1].float()
        si[:, a, L.STORE_I["current_target"]] = tgt
        obs[:, :3] = tp.float()
        obs[:, 3] = si[:, a, L.STORE_I["inventory"]].float()
        return reward, obs, torch.zeros(N, dtype=torch.bool, device=env.device), {}


Pick_Up_Dynamic = PickUpDynamic


def tag_distance_reward(mujoco_gym, agent):
    """README.md:149-163 with the evident intent: draw a "target" once per agent and episode, then
    reward 10 * (previous distance - distance); fp64 arithmetic like the reference's math.dist."""
    env = mujoco_gym
    a = env.agents.index(agent)
    si, sf, sf64 = _store(env)
    if not env._target_names:
        return torch.zeros(env.num_envs, dtype=torch.float64, device=env.device)
    tgt = si[:, a, L.STORE_I["current_target"]]
    new = tgt == 0
    tgt = torch.where(new, _draw_target(env, a, new), tgt)
    d, _ = _dist_to_target(env, a, tgt)
    reward = torch.where(new, torch.zeros_like(d), (sf64[:, a, 1] - d) * tag_distance_reward.scale)
    sf64[:, a, 1] = d
    sf[:, a, L.STORE_F["distance"]] = d.float()
    si[:, a, L.STORE_I["current_target"]] = tgt
    return reward


tag_distance_reward.mjb_kind = ("reward", L.REW_TAG_DISTANCE)
tag_distance_reward.scale = 10.0
reward_function = tag_distance_reward


def distance_done(mujoco_gym, agent):
    """README.md:168-173: data_store[agent]["distance"] <= 1."""
    env = mujoco_gym
    return _store(env)[2][:, env.agents.index(agent), 1] <= distance_done.threshold


distance_done.mjb_kind = ("done", L.DONE_DISTANCE_LE)
distance_done.threshold = 1.0
done_function = distance_done


def ant_reward_function(env, agent):
    """benchmarking/fps_gym/fps_custom_env.py:4-27: forward progress / dt - 0.5 * sum(ctrl^2) - contact cost; the
    contact cost is identically zero for models without force / torque / accelerometer sensors (cfrc_ext = 0)."""
    a = env.agents.index(agent)
    si, sf, _ = _store(env)
    if any(int(t) == 1 for t in env.model.fields["sensor_type"]):
        raise NotImplementedError("ant_reward_function: models with accelerometers have a non-zero cfrc_ext contact cost")
    x_after = env._batch.probe[:, a, 0]
    has = si[:, a, L.STORE_I["has_xpos"]] != 0
    ctrl = env._batch.ctrl[:, :env.model.nu].double()
    r = (x_after.double() - sf[:, a, L.STORE_F["xpos_before"]].double()) / float(env.model.timestep) - 0.5 * ctrl.pow(2).sum(dim=1)
    si[:, a, L.STORE_I["has_xpos"]] = 1
    sf[:, a, L.STORE_F["xpos_before"]] = x_after
    return torch.where(has, r, torch.zeros_like(r))


ant_reward_function.mjb_kind = ("reward", L.REW_ANT)


# ---- recognition of the reference's own source text -------------------------------------------------------------------
def normalised_source_hash(src: str) -> str:
    """hash of a function / class source that ignores comments, blank lines, indentation style, docstrings and the
    name after `def` / `class` (users rename)"""
    toks, prev = [], None
    try:
        stream = list(tokenize.generate_tokens(io.StringIO(src).readline))
    except (tokenize.TokenError, IndentationError):
        import textwrap
        stream = list(tokenize.generate_tokens(io.StringIO(textwrap.dedent(src)).readline))
    for t in stream:
        if t.type in (tokenize.COMMENT, tokenize.NL, tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT, tokenize.ENDMARKER):
            continue
        if t.type == tokenize.STRING and prev in (None, ":", ")") and t.string[:3] in ('"""', "'''"):
            continue   # docstring
        s = t.string
        if prev in ("def", "class") and t.type == tokenize.NAME:
            s = "_"
        toks.append(s)
        prev = t.string
    return hashlib.sha256(" ".join(toks).encode()).hexdigest()[:16]


# normalised hashes of the reference's plugins (tests/golden/make_ref_plugin_fixtures.py prints them from /root/reference)
REFERENCE_SOURCES = {
    "21ca1977f26ac013": ("reward", L.REW_ANT, {}),                         # fps_custom_env.py:4-27 ant_reward_function
    "f8b5b57f47f29996": ("dynamic", L.DYN_LANGUAGE, {}),                  # README.md:109-136 class Language
    "f223b6778582385f": ("reward", L.REW_TAG_DISTANCE, {"scale": 10.0}),   # README.md:149-163 reward_function
    "f53352e96e4500c8": ("done", L.DONE_DISTANCE_LE, {"threshold": 1.0}),    # README.md:168-172 done_function
}


def recognise(obj):
    """(list, kind, attrs) when `obj` (function or dynamics class) is one of the reference's example plugins, either by
    its `mjb_kind` marker or by its normalised source text; None otherwise"""
    kind = getattr(obj, "mjb_kind", None)
    if kind:
        return kind[0], kind[1], {}
    try:
        src = inspect.getsource(obj)
    except (OSError, TypeError):
        return None
    import textwrap
    return REFERENCE_SOURCES.get(normalised_source_hash(textwrap.dedent(src)))
