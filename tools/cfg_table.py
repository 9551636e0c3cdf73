"""Print the per-config table of one bench.py JSON line (stdin or file)."""
import json, sys
d = json.loads((open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read())
print("headline", round(d["value"] / 1e6, 2), "M/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"] / 1e6, 2))
for c in d["configs"]:
    print(f'{c["name"]:15s} envs/gpu {c["envs_per_gpu"]:6d} ms {c["ms_per_step"]:.4f} M/s {c["agent_steps_per_s"] / 1e6:7.1f} ncon {c["ncon_mean"]:.2f} nit {c["niter_mean"]} '
          f'drop {c["ncon_dropped"]} hbm {c["hbm"]["frac"]:.4f} warps {c["geometry"]["warps_per_cta"]}')
