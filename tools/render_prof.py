"""Small driver for profiling k_render: 4096 envs, 2 cameras, 64x64, settled state."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
LV = os.path.join(ROOT, "tests", "levels")
env = MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants_cams.xml"), "agents": ["sender", "receiver"], "agentCameras": True, "num_envs": 4096})
env.reset()
for _ in range(150):
    env.step(env.sample_actions())
out = env.batch.render([0, 1], 64, 64)
for _ in range(3):
    env.batch.render([0, 1], 64, 64, out=out)
torch.cuda.synchronize()
print("ok", int(out.sum()))
