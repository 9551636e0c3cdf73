import random

import numpy as np


def reward_function(mujoco_gym, agent):
    # Creates all the necessary fields to store the needed data within the dataStore at timestep 0 
    if "targets" not in mujoco_gym.data_store[agent].keys():
        mujoco_gym.data_store["targets"] = mujoco_gym.filter_by_tag("target")
        mujoco_gym.data_store[agent]["current_target"] = mujoco_gym.data_store["targets"][random.randint(0, len(mujoco_gym.data_store["targets"]) - 1)]["name"]
        distance = mujoco_gym.distance(agent, mujoco_gym.data_store[agent]["current_target"])
        mujoco_gym.data_store[agent]["distance"] = distance
        new_reward = 0
    else:  # Calculates the distance between the agent and the current target
        distance = mujoco_gym.distance(agent, mujoco_gym.data_store[agent]["current_target"])
        new_reward = mujoco_gym.data_store[agent]["distance"] - distance
        mujoco_gym.data_store[agent]["distance"] = distance
    reward = new_reward * 10
    return reward

def done_function(mujoco_gym, agent):
    if mujoco_gym.data_store[agent]["distance"] <= 1:
        return True
    else:
        return False
