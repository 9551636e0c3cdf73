"""Per-phase clock cycles of the step kernel (debug build -DMJB_PHASE_PROF of the same sources, built on the fly into
libmjb_prof.so and selected through MJB_LIB).  Prints, for a settled C2 batch, the share of warp-cycles spent in each
phase of one env-step (lane 0 of every env-warp accumulates clock64 deltas into a global counter per phase)."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "mujoco_rl_environment_wrapper_b200")
PROF_LIB = os.path.join(PKG, "libmjb_prof.so")
PHASES = ["load", "fk", "crb", "rne", "collide", "sensors", "constraints", "newton_init", "newton_grad", "newton_hess", "newton_factor",
          "newton_linesearch", "integrate", "store", "epilogue", "round_barrier", "align", "ls_rowsmul", "ls_Mv", "ls_loop", "epi_obs", "epi_stage", "epi_plugins"]


def build():
    from mujoco_rl_environment_wrapper_b200 import build as b
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-DMJB_PHASE_PROF", "-o", PROF_LIB] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
    subprocess.run(cmd, check=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
        return
    os.environ["MJB_LIB"] = PROF_LIB
    import torch
    from mujoco_rl_environment_wrapper_b200 import _lib as L
    from mujoco_rl_environment_wrapper_b200 import plugins as P
    from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
    lv = os.path.join(ROOT, "tests", "levels")
    lib = L.load()
    out = []
    for n in [int(x) for x in os.environ.get("PROF_ENVS", "4096,65536").split(",")]:
        env = MuJoCoRL({"xmlPath": os.path.join(lv, "two_ants.xml"), "infoJson": os.path.join(lv, "info_2A.json"), "agents": ["sender", "receiver"],
                        "num_envs": n, "seed": 1234, "environmentDynamics": [P.Language], "rewardFunctions": [P.tag_distance_reward],
                        "doneFunctions": [P.distance_done]})
        b, ad = env.batch, env._act_dim
        pool = torch.stack([env.sample_actions() for _ in range(16)])
        env.reset()
        for k in range(300):
            b.actions[:, :, :ad].copy_(pool[k % 16]); b.step()
        torch.cuda.synchronize()
        buf = (ctypes.c_uint64 * 32)()
        lib.mjb_phase_cycles(buf, 32)   # clear
        steps = 20
        for k in range(steps):
            b.actions[:, :, :ad].copy_(pool[k % 16]); b.step()
        torch.cuda.synchronize()
        nph = lib.mjb_phase_cycles(buf, 32)
        tot = sum(buf[i] for i in range(nph))
        rec = {"envs": n, "cycles_per_env_step": tot / (n * steps), "niter_mean": float(b.niter.float().mean()), "ncon_mean": float(b.ncon.float().mean()),
               "phases": {PHASES[i]: round(buf[i] / tot, 4) for i in range(nph)},
               "linesearch_evals_per_newton_iter": buf[24] / max(1, buf[25]), "newton_iters_per_env_step": buf[25] / (n * steps),
               "rows_switched_after_iteration": {str(i): buf[25 + i] / (n * steps) for i in range(1, 6)},
               "steps_with_alpha_1": buf[31] / max(1, buf[25])}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del env, b
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "phase_prof.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
