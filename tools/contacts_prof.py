"""Throughput of the C2 step against the contact load: the same 4096-env batch is put into states
with different numbers of contacts per env — airborne (0), the bench's random-action state (~2), standing still on
the floor (all feet down), ants driven into the floor with a constant push (legs folded: many contacts), and the two
ants of every env stacked on each other (contacts BETWEEN the kinematic trees: 28 x 28 shared-memory factorisation
instead of two register-resident 14 x 14 blocks) — and 20 L2-flushed steps of each are timed.  Reports ms/step,
agent-steps/s, contacts per env (mean / max), contacts dropped, and the Newton-iteration histogram.
-> gpurun_out/r02_contacts.json (copied to profiles/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mujoco_rl_environment_wrapper_b200 import plugins as P  # noqa: E402
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL  # noqa: E402

LV = os.path.join(ROOT, "tests", "levels")
N = int(os.environ.get("N", "4096"))


def make():
    return MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants.xml"), "infoJson": os.path.join(LV, "info_2A.json"), "agents": ["sender", "receiver"],
                     "num_envs": N, "seed": 99, "environmentDynamics": [P.Language], "rewardFunctions": [P.tag_distance_reward],
                     "doneFunctions": [P.distance_done], "maxSteps": 100000})


def timed(env, actions_fn, steps=20):
    b, ad = env.batch, env._act_dim
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts, ncon_sum, ncon_max, hist = [], 0.0, 0, torch.zeros(16, dtype=torch.long, device="cuda")
    for k in range(steps + 3):
        b.actions[:, :, :ad] = actions_fn(k)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.step(); e1.record()
        torch.cuda.synchronize()
        if k >= 3:
            ts.append(e0.elapsed_time(e1))
            ncon_sum += float(b.ncon.float().mean()); ncon_max = max(ncon_max, int(b.ncon.max()))
            hist += torch.bincount(b.niter.flatten().clamp(0, 15), minlength=16)
    ms = sum(ts) / len(ts)
    h = hist.cpu().tolist()
    while h and h[-1] == 0:
        h.pop()
    return {"ms_per_step": ms, "agent_steps_per_s": 2 * N / (ms * 1e-3), "ncon_mean": ncon_sum / steps, "ncon_max": ncon_max,
            "ncon_dropped_total": int(b.ncon_dropped.sum()), "nreset_total": int(b.nreset.sum()) if hasattr(b, "nreset") else None,
            "newton_iterations_hist": h, "newton_iterations_mean": sum(i * c for i, c in enumerate(h)) / max(1, sum(h))}


def main():
    out = []
    env = make()
    b, ad = env.batch, env._act_dim
    pool = torch.stack([env.sample_actions() for _ in range(16)])
    zero = torch.zeros_like(pool[0])
    zero[:, :, 8] = 1.0   # Language action stays valid
    only_stacked = os.environ.get("ONLY") == "stacked"
    # 1. airborne: right after reset the ants are still falling
    env.reset()
    if not only_stacked:
        out.append(dict(state="airborne (first steps after reset, random actions)", **timed(env, lambda k: pool[k % 16], steps=10)))
        states_1_to_4(env, b, ad, pool, zero, out)
    stacked(env, b, ad, pool, out)


def states_1_to_4(env, b, ad, pool, zero, out):
    # 2. the bench's state: 300 random-action steps
    env.reset()
    for k in range(300):
        b.actions[:, :, :ad] = pool[k % 16]; b.step()
    out.append(dict(state="random actions, settled 300 steps (bench state)", **timed(env, lambda k: pool[k % 16])))
    # 3. standing still: zero torques for 400 steps
    env.reset()
    for k in range(400):
        b.actions[:, :, :ad] = zero; b.step()
    out.append(dict(state="standing still (zero torques, settled 400 steps)", **timed(env, lambda k: zero)))
    # 4. legs driven outwards / down with constant full torque: folded legs, many leg-floor contacts
    push = zero.clone(); push[:, :, :8] = 1.0
    for k in range(300):
        b.actions[:, :, :ad] = push; b.step()
    out.append(dict(state="constant full torque on every motor (settled 300 steps)", **timed(env, lambda k: push)))
    push2 = zero.clone(); push2[:, :, :8] = -1.0
    for k in range(300):
        b.actions[:, :, :ad] = push2; b.step()
    out.append(dict(state="constant full negative torque on every motor (settled 300 steps)", **timed(env, lambda k: push2)))


def stacked(env, b, ad, pool, out):
    # 5. the two ants of every env stacked: contacts couple the two kinematic trees
    env.reset()
    torch.cuda.synchronize()
    g = torch.Generator(device="cuda").manual_seed(5)
    off = torch.rand((N, 3), device="cuda", generator=g)
    b.qpos[:, 15] = b.qpos[:, 0] + 0.6 * (off[:, 0] - 0.5)
    b.qpos[:, 16] = b.qpos[:, 1] + 0.6 * (off[:, 1] - 0.5)
    b.qpos[:, 17] = b.qpos[:, 2] + 0.45 + 0.2 * off[:, 2]
    for k in range(150):
        b.actions[:, :, :ad] = pool[k % 16]; b.step()
    out.append(dict(state="receiver dropped onto the sender (inter-ant contacts: one 28x28 Hessian block), random actions", **timed(env, lambda k: pool[k % 16])))
    for r in out:
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"envs": N, "workload": "C2 (two ants, Language + tag reward + done)", "l2": "flushed between steps", "rows": out},
              open(os.path.join(ROOT, "gpurun_out", "r02_contacts_stacked.json" if os.environ.get("ONLY") == "stacked" else "r02_contacts.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
