"""Throughput through the Python drop-in API (MuJoCoRL.step with a packed action tensor / a dict of tensors), C2."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mujoco_rl_environment_wrapper_b200 import plugins as P
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
LV = os.path.join(ROOT, "tests", "levels")
N = int(os.environ.get("MJB_BENCH_ENVS", "4096"))
env = MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants.xml"), "infoJson": os.path.join(LV, "info_2A.json"), "agents": ["sender", "receiver"],
                "environmentDynamics": [P.Language], "rewardFunctions": [P.tag_distance_reward], "doneFunctions": [P.distance_done], "num_envs": N})
env.reset()
pool = [env.sample_actions() for _ in range(8)]
dpool = [{a: p[:, i].contiguous() for i, a in enumerate(env.agents)} for p in pool]
for k in range(300):
    env.step(pool[k % 8])
res = {}
for name, acts in (("packed_tensor", pool), ("dict_of_tensors", dpool)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(1000):
        out = env.step(acts[k % 8])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 1000
    res[name] = {"ms_per_step": dt * 1e3, "agent_steps_per_s": N * 2 / dt}
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(1000):
    env.batch.actions[:, :, :env._act_dim] = pool[k % 8]
    env.batch.step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 1000
res["batch_step_only"] = {"ms_per_step": dt * 1e3, "agent_steps_per_s": N * 2 / dt}
print(json.dumps(res))
