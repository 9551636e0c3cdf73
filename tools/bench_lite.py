"""Literal C3 step (Ant.xml, freeJoint, skipFrames = 0, ant reward; fps_custom_env.py:39-48) at 65536 envs: kernel
time by CUDA events around each launch with L2 flushed between steps -> achieved HBM bandwidth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mujoco_rl_environment_wrapper_b200 import plugins as P  # noqa: E402
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL  # noqa: E402

LV = os.path.join(ROOT, "tests", "levels")


def lite_bytes(env):
    """bytes one env-step of the tile kernel has to move (lite_kernel.cuh): rows read + rows written"""
    lay, A, m = env.batch.layout, len(env.agents), env.model
    rd = 4 * (lay.qpos_stride + lay.qvel_stride + (lay.ctrl_stride if m.nu else 0) + (lay.sensor_stride if m.nsensordata else 0) +
              A * lay.act_stride + 4 * lay.probe_count + A * lay.store_i32 + A * lay.store_f32 + 1)
    wr = 4 * (A * lay.obs_stride + (lay.qvel_stride if env.free_joint else lay.ctrl_stride) + A * lay.store_i32 + A * lay.store_f32 + A + 1) + 2 * (A + 1)
    return rd + wr


def run(n, steps=200, flush_l2=True):
    env = MuJoCoRL(dict(xmlPath=os.path.join(LV, "ant_rk4.xml"), agents=["torso"], freeJoint=True, skipFrames=0,
                        rewardFunctions=[P.ant_reward_function], num_envs=n))
    b = env.batch
    env.reset()
    pool = torch.stack([env.sample_actions() for _ in range(4)])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for k in range(10):
        b.actions[:, :, :3] = pool[k % 4]; b.step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        if flush_l2:
            flush.zero_()
        ev[k][0].record(); b.step(); ev[k][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(c) for a, c in ev)
    ms = sum(ts) / len(ts)
    by = lite_bytes(env)
    return {"envs": n, "l2_flushed": flush_l2, "kernel_ms_mean": ms, "kernel_ms_median": ts[len(ts) // 2], "kernel_ms_min": ts[0],
            "bytes_per_env_step": by, "GBps": by * n / (ms * 1e-3) / 1e9, "agent_steps_per_s": n / (ms * 1e-3)}


if __name__ == "__main__":
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = []
    for n, fl in ((65536, True), (65536, False), (262144, True), (4096, True)):
        r = run(n, flush_l2=fl)
        r["frac_of_peak"] = r["GBps"] / peak
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_lite.json"), "w"), indent=1)
