"""Fixed cost of a launch of the step kernel: C2 at 1, 2, 4, 8, 16 full lock-step rounds (148 SMs x 14 env-warps = 2072
envs per round), with and without the L2 flush between steps.  T = a + b * rounds: `a` is launch + image staging + cold
instruction / data fetch + grid tail, `b` the steady round.  -> gpurun_out/r02_rounds.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MJB_WARPS", "14")
import torch  # noqa: E402

from mujoco_rl_environment_wrapper_b200 import plugins as P  # noqa: E402
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL  # noqa: E402

LV = os.path.join(ROOT, "tests", "levels")


def run(n, flush_l2, steps=40):
    env = MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants.xml"), "infoJson": os.path.join(LV, "info_2A.json"), "agents": ["sender", "receiver"],
                    "num_envs": n, "seed": 99, "environmentDynamics": [P.Language], "rewardFunctions": [P.tag_distance_reward],
                    "doneFunctions": [P.distance_done], "maxSteps": 100000})
    b, ad = env.batch, env._act_dim
    env.reset()
    pool = torch.stack([env.sample_actions() for _ in range(8)])
    for k in range(300):
        b.actions[:, :, :ad] = pool[k % 8]; b.step()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for k in range(steps + 3):
        b.actions[:, :, :ad] = pool[k % 8]
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.step(); e1.record()
        torch.cuda.synchronize()
        if k >= 3:
            ts.append(e0.elapsed_time(e1))
    del env
    return sum(ts) / len(ts)


if __name__ == "__main__":
    out = []
    for rounds in (1, 2, 4, 8, 16):
        n = 2072 * rounds
        r = {"rounds": rounds, "envs": n, "ms_flushed": run(n, True), "ms_unflushed": run(n, False)}
        print(json.dumps(r), flush=True)
        out.append(r)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_rounds.json"), "w"), indent=1)
