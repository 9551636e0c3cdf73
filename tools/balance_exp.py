"""Experiment: kernel time of the settled C2 step with / without cost-ordered scheduling (kernel events only)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mujoco_rl_environment_wrapper_b200 import plugins as P
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
lv = os.path.join(ROOT, "tests", "levels")
n = int(os.environ.get("N", "4096"))
env = MuJoCoRL({"xmlPath": os.path.join(lv, "two_ants.xml"), "infoJson": os.path.join(lv, "info_2A.json"), "agents": ["sender", "receiver"],
                "num_envs": n, "seed": 1234, "environmentDynamics": [P.Language], "rewardFunctions": [P.tag_distance_reward], "doneFunctions": [P.distance_done]})
b, ad = env.batch, env._act_dim
pool = torch.stack([env.sample_actions() for _ in range(16)])
env.reset()
for k in range(300):
    b.actions[:, :, :ad].copy_(pool[k % 16]); b.step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in ("none", "sort", "snake"):
    os.environ["MJB_BALANCE_SNAKE"] = "1" if mode == "snake" else "0"
    ts = []
    for k in range(60):
        b.actions[:, :, :ad].copy_(pool[k % 16])
        if mode != "none":
            b.rebalance()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.step(); e1.record()
        torch.cuda.synchronize()
        if k >= 10:
            ts.append(e0.elapsed_time(e1))
    print(mode, "kernel ms mean", sum(ts) / len(ts), "niter hist", torch.bincount(b.niter.flatten().clamp(0, 12)).tolist())
