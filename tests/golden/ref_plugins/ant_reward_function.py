import random

import numpy as np


def ant_reward_function(env, agent):
    """
    Custom reward function that mimics the reward function of Gym's Ant environment.
    Args:
        env: The environment instance.
        agent: The agent ID.
    Returns:
        float: The reward for the agent.
    """
    xpos_before = env.data_store[agent].get('xpos_before', None)
    xpos_after = env.get_data(agent)['position'][0]

    if xpos_before is None:
        env.data_store[agent]['xpos_before'] = xpos_after
        return 0

    dt = env.model.opt.timestep
    forward_reward = (xpos_after - xpos_before) / dt
    control_cost = 0.5 * np.square(env.data.ctrl).sum()
    contact_cost = 0.5 * 1e-3 * np.sum(np.square(np.clip(env.data.cfrc_ext, -1, 1)))
    reward = forward_reward - control_cost - contact_cost
    env.data_store[agent]['xpos_before'] = xpos_after

    return reward
