"""Throughput of every BASELINE.json config on one GPU -> profiles/r01_configs.json.

bench.py measures the headline config (C2) under the driver's contract; this script records the others with
the same timing hygiene (warm-up, CUDA events on the launching stream, settled state) so that each row of the
scope table has a number:
  C1  SingleAgentModel.xml, 1 agent, ctrl mode, no plugins
  C2  MultiAgentModel.xml, Language + tag-distance reward + done            (4096 and 65536 envs)
  C3-literal  Ant.xml, freeJoint, skipFrames=0, ant reward  (no physics: gather/scatter, HBM bound)
  C3-physics  Ant.xml, ctrl mode, RK4, skipFrames=1 and 5, ant reward
  C4  MultiAgentModel3Sensors.xml, freeJoint, skipFrames=5 (rangefinder + touch + accelerometer)
  C5  MultiAgentModel.xml, Pick_Up dynamic
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mujoco_rl_environment_wrapper_b200 import plugins as P  # noqa: E402
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL  # noqa: E402

LV = os.path.join(ROOT, "tests", "levels")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

CONFIGS = [
    ("C1", dict(xmlPath="one_ant_arena.xml", agents=["sender"]), 16384),
    ("C2", dict(xmlPath="two_ants.xml", infoJson="info_2A.json", agents=["sender", "receiver"], environmentDynamics=[P.Language],
                rewardFunctions=[P.tag_distance_reward], doneFunctions=[P.distance_done]), 4096),
    ("C2", dict(xmlPath="two_ants.xml", infoJson="info_2A.json", agents=["sender", "receiver"], environmentDynamics=[P.Language],
                rewardFunctions=[P.tag_distance_reward], doneFunctions=[P.distance_done]), 65536),
    ("C3-literal", dict(xmlPath="ant_rk4.xml", agents=["torso"], freeJoint=True, skipFrames=0, rewardFunctions=[P.ant_reward_function]), 65536),
    ("C3-physics-skip1", dict(xmlPath="ant_rk4.xml", agents=["torso"], skipFrames=1, rewardFunctions=[P.ant_reward_function]), 65536),
    ("C3-physics-skip5", dict(xmlPath="ant_rk4.xml", agents=["torso"], skipFrames=5, rewardFunctions=[P.ant_reward_function]), 65536),
    ("C4", dict(xmlPath="two_ants_touch_acc.xml", agents=["sender", "receiver"], freeJoint=True, skipFrames=5), 16384),
    ("C5", dict(xmlPath="two_ants.xml", infoJson="info_2A.json", agents=["sender", "receiver"], environmentDynamics=[P.PickUpDynamic]), 32768),
]


def run(name, cfg, n, steps=60, settle=150):
    cfg = dict(cfg)
    cfg["xmlPath"] = os.path.join(LV, cfg["xmlPath"])
    if "infoJson" in cfg:
        cfg["infoJson"] = os.path.join(LV, cfg["infoJson"])
    cfg["num_envs"] = n
    env = MuJoCoRL(cfg)
    b = env.batch
    A = len(cfg["agents"])
    pool = torch.stack([env.sample_actions() for _ in range(8)])
    env.reset()
    for k in range(settle):
        b.actions[:, :, :env._act_dim] = pool[k % 8]
        b.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        b.actions[:, :, :env._act_dim] = pool[k % 8]
        b.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    m, lay = env.model, b.layout
    obs_dim = max(env._spec.obs_dim[a] for a in range(A))
    bytes_env = 4 * (2 * (m.nq + 2 * m.nv + m.nu) + A * env._act_dim + m.nsensordata + A * obs_dim + A + 4 * lay.probe_count) + 2 * (A + 1) + 48
    out = {"config": name, "envs": n, "agents": A, "skip_frames": int(cfg.get("skipFrames", 1)), "ms_per_step": ms,
           "agent_steps_per_s": n * A / (ms * 1e-3), "env_steps_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_env_step": bytes_env,
           "achieved_GBps": bytes_env * n / (ms * 1e-3) / 1e9, "hbm_frac_of_measured_peak": bytes_env * n / (ms * 1e-3) / 1e9 / PEAK,
           "geometry": b.geometry(), "ncon_mean": float(b.ncon.float().mean()), "newton_iters_mean": float(b.niter.float().mean())}
    del env
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    res = []
    for name, cfg, n in CONFIGS:
        r = run(name, cfg, n)
        print(json.dumps(r), flush=True)
        res.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r01_configs.json"), "w"), indent=1)
