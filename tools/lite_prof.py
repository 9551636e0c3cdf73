import os, sys, torch
sys.path.insert(0, os.getcwd())
from mujoco_rl_environment_wrapper_b200 import plugins as P
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
LV = os.path.join(os.getcwd(), "tests", "levels")
for n in (4096, 65536):
    env = MuJoCoRL(dict(xmlPath=os.path.join(LV, "ant_rk4.xml"), agents=["torso"], freeJoint=True, skipFrames=0,
                        rewardFunctions=[P.ant_reward_function], num_envs=n))
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for k in range(4):
        flush.zero_(); env.batch.step(); torch.cuda.synchronize()
    print("envs", n, flush=True)
