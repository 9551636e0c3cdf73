"""profiles/r02_counts.json from ncu captures: executed warp instructions per env and DRAM bytes per launch of the step
kernel, stamped with the fingerprint of the kernel sources they were measured on (bench.py only uses them while the
sources still match).  usage: make_counts.py NAME=report.ncu-rep:ENVS [...]   e.g.  C2=gpurun_out/prof.ncu-rep:4096"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_sha16  # noqa: E402


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, v = rows[0], rows[2]
    return {n: v[i] for i, n in enumerate(h)}, {n: rows[1][i] for i, n in enumerate(h)}


def num(x):
    return float(str(x).replace(",", ""))


def main():
    cfgs = {}
    for arg in sys.argv[1:]:
        name, rest = arg.split("=")
        rep, envs = rest.rsplit(":", 1)
        envs = int(envs)
        r, units = raw(rep)
        inst = num(r["smsp__inst_executed.sum"])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = num(r["dram__bytes_read.sum"]) * scale[units["dram__bytes_read.sum"]] + num(r["dram__bytes_write.sum"]) * scale[units["dram__bytes_write.sum"]]
        cfgs[name] = {"envs": envs, "warp_instructions_per_env": inst / envs, "warp_instructions": inst, "dram_bytes": dram,
                      "kernel_us_under_ncu": num(r["gpu__time_duration.sum"]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units["gpu__time_duration.sum"], 1.0),
                      "report": os.path.basename(rep)}
    out = {"csrc_sha16": csrc_sha16(), "configs": cfgs,
           "note": "executed warp instructions (smsp__inst_executed.sum) and DRAM bytes of one k_env launch, from ncu --set full captures"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_counts.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
