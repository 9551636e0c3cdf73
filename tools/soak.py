"""Stability soak: C2 at 65536 envs, 3000 random-action steps with per-env auto-reset on term | trunc.
Reports non-finite states, divergence resets (mj_checkPos / mj_checkVel semantics), contact-slot overflow."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mujoco_rl_environment_wrapper_b200 import plugins as P
from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
LV = os.path.join(ROOT, "tests", "levels")
N, STEPS = 65536, 3000
env = MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants.xml"), "infoJson": os.path.join(LV, "info_2A.json"), "agents": ["sender", "receiver"],
                "environmentDynamics": [P.Language], "rewardFunctions": [P.tag_distance_reward], "doneFunctions": [P.distance_done],
                "num_envs": N, "maxSteps": 1024})
env.reset()
b = env.batch
max_ncon = max_iter = 0
episodes = 0
for t in range(STEPS):
    obs, rew, term, trunc, _ = env.step(env.sample_actions())
    if t % 50 == 49:
        max_ncon = max(max_ncon, int(b.ncon.max())); max_iter = max(max_iter, int(b.niter.max()))
        done = term["__all__"] | trunc["__all__"]
        if bool(done.any()):
            episodes += int(done.sum())
            env.reset(mask=done)
res = {"envs": N, "steps": STEPS, "finite": bool(torch.isfinite(b.qpos).all() and torch.isfinite(b.qvel).all() and torch.isfinite(b.obs).all()),
       "divergence_resets": int(b.nreset.sum()), "max_contacts_seen": max_ncon, "contact_slots": b.layout.maxcon,
       "contacts_dropped_total": int(b.ncon_dropped.sum()),
       "max_newton_iters_seen": max_iter, "episodes_reset": episodes,
       "max_abs_qvel": float(b.qvel.abs().max()), "max_height": float(b.qpos[:, 2].max())}
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r02_soak.json"), "w"), indent=1)
