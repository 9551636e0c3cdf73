"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel mode on small batches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mujoco_rl_environment_wrapper_b200 import plugins as P
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
LV = os.path.join(ROOT, "tests", "levels")
cases = [
    dict(xmlPath="two_ants_cams.xml", infoJson="info_2A.json", agents=["sender", "receiver"], environmentDynamics=[P.Language],
         rewardFunctions=[P.tag_distance_reward], doneFunctions=[P.distance_done], num_envs=37, agentCameras=True, sensorResolution=(32, 24)),
    dict(xmlPath="ant_rk4.xml", agents=["torso"], rewardFunctions=[P.ant_reward_function], num_envs=9, skipFrames=2),
    dict(xmlPath="ant_rk4.xml", agents=["torso"], rewardFunctions=[P.ant_reward_function], num_envs=9, skipFrames=0, freeJoint=True),
    dict(xmlPath="two_ants_touch_acc.xml", agents=["sender", "receiver"], num_envs=5, freeJoint=True, skipFrames=3, resetNoise=0.05),
    dict(xmlPath="box_rangefinder.xml", agents=["receiver"], num_envs=11, freeJoint=True),
]
for cfg in cases:
    cfg = dict(cfg)
    cfg["xmlPath"] = os.path.join(LV, cfg["xmlPath"])
    if "infoJson" in cfg:
        cfg["infoJson"] = os.path.join(LV, cfg["infoJson"])
    env = MuJoCoRL(cfg)
    env.reset()
    for t in range(12):
        env.step(env.sample_actions())
    m = torch.zeros(env.num_envs, dtype=torch.bool); m[::2] = True
    env.reset(mask=m)
    env.step(env.sample_actions())
    if cfg.get("agentCameras"):
        env.get_camera_data(env.agents[0])
    b = env.batch
    arrs = b.host_arrays()
    arrs[0][:] = env.sample_actions().cpu().numpy().reshape(arrs[0].shape[0], arrs[0].shape[1], -1)[:, :, :arrs[0].shape[2]] if False else 0
    b.step_host(*arrs)
    torch.cuda.synchronize()
    print("ok", os.path.basename(cfg["xmlPath"]), float(b.qpos.abs().sum()))
