"""The reference's example plugins, re-expressed so that the environment can recognise them and
run them inside the fused CUDA epilogue (attribute `mjb_kind`).  Signatures are the reference's:
dynamics are classes with `__init__(self, mujoco_gym)`, `observation_space`, `action_space` and
`dynamic(self, agent, actions)`; reward / done functions are `f(mujoco_gym, agent)`.

Calling them directly (what `MuJoCoRL.__check_*` does once at construction, mujoco_rl.py:114-169) runs
a small torch implementation over the same store tensors; the per-step path never calls them.
"""
import torch

from . import _lib as L


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
        obs = torch.zeros(env.num_envs, 4, device=env.device)
        return 0, obs, torch.zeros(env.num_envs, dtype=torch.bool, device=env.device), {}


Pick_Up_Dynamic = PickUpDynamic


def tag_distance_reward(mujoco_gym, agent):
    """README.md:149-163 with the evident intent: draw a "target" once per agent and
    episode, then reward 10 * (previous distance - distance)."""
    return 0.0


tag_distance_reward.mjb_kind = ("reward", L.REW_TAG_DISTANCE)
tag_distance_reward.scale = 10.0
reward_function = tag_distance_reward


def distance_done(mujoco_gym, agent):
    """README.md:168-173: data_store[agent]["distance"] <= 1."""
    return False


distance_done.mjb_kind = ("done", L.DONE_DISTANCE_LE)
distance_done.threshold = 1.0
done_function = distance_done


def ant_reward_function(env, agent):
    """benchmarking/fps_gym/fps_custom_env.py:4-27."""
    return 0.0


ant_reward_function.mjb_kind = ("reward", L.REW_ANT)
