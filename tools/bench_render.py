"""Agent-camera raycaster throughput -> gpurun_out/r01_render.json.
C2 scene with the reference's two agent cameras, 64x64 (the reference's default sensorResolution), settled
state; the output tensor (N x 2 x 64 x 64 x 3 bytes) is larger than L2 from N = 5200 up."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL  # noqa: E402

LV = os.path.join(ROOT, "tests", "levels")
res = []
for n, wh in ((4096, 64), (16384, 64), (4096, 128)):
    env = MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants_cams.xml"), "agents": ["sender", "receiver"], "agentCameras": True,
                    "num_envs": n, "sensorResolution": (wh, wh)})
    env.reset()
    for _ in range(150):
        env.step(env.sample_actions())
    b = env.batch
    out = b.render([0, 1], wh, wh)
    for _ in range(3):
        b.render([0, 1], wh, wh, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        b.render([0, 1], wh, wh, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rays = n * 2 * wh * wh
    r = {"envs": n, "cameras": 2, "resolution": [wh, wh], "ms_per_render": ms, "frames_per_s": n * 2 / (ms * 1e-3),
         "rays_per_s": rays / (ms * 1e-3), "ray_geom_tests_per_s": rays * env.model.ngeom / (ms * 1e-3),
         "output_GBps": out.numel() / (ms * 1e-3) / 1e9, "lit_fraction": float((out.view(-1, 3).sum(1) > 0).float().mean())}
    print(json.dumps(r), flush=True)
    res.append(r)
    del env, b, out
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r01_render.json"), "w"), indent=1)
