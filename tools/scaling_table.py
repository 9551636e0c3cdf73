"""gpurun_out/scale_n{1,2,4,8}.json (tools/gpu_scale.sh) -> profiles/r02_scaling.json + the table of DESIGN 10."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for n in (1, 2, 4, 8):
    d = json.loads(open(os.path.join(ROOT, "gpurun_out", f"scale_n{n}.json")).read().strip().splitlines()[-1])
    cfgs = {c["name"]: {k: c.get(k) for k in ("envs_per_gpu", "ms_per_step", "agent_steps_per_s", "scaling") if k in c} for c in d.get("configs", [])}
    out[str(n)] = {"value": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"], "clocks": d.get("clocks"), "configs": cfgs}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_scaling.json"), "w"), indent=1)
cols = ("C2-65536", "C3-physics", "C5", "C3-literal")
M = lambda x: round(x / 1e6, 1)
print("| GPUs | C2 weak | e2e | " + " | ".join(cols) + " |")
for n in ("1", "2", "4", "8"):
    o = out[n]
    print(f"| {n} | {M(o['value'])} | {M(o['e2e'])} | " + " | ".join(str(M(o['configs'][k]['agent_steps_per_s'])) for k in cols) + " |")
o1, o8 = out["1"], out["8"]
print("| efficiency at 8 | %.2f | %.2f | " % (o8["value"] / 8 / o1["value"], o8["e2e"] / 8 / o1["e2e"]) +
      " | ".join("%.2f" % (o8["configs"][k]["agent_steps_per_s"] / 8 / o1["configs"][k]["agent_steps_per_s"]) for k in cols) + " |")
