"""Summarise an ncu report: headline metrics, stall reasons, per-function shares (run where ncu is installed)."""
import collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout
det = list(csv.reader(io.StringIO(ncu("--page", "details", "--csv"))))
idx = {h: i for i, h in enumerate(det[0])}
want = ["Duration", "Elapsed Cycles", "SM Frequency", "Executed Ipc Active", "Issue Slots Busy", "Issued Warp Per Scheduler",
        "Active Warps Per Scheduler", "Eligible Warps Per Scheduler", "No Eligible", "Warp Cycles Per Issued Instruction",
        "Avg. Active Threads Per Warp", "Avg. Not Predicated Off Threads Per Warp", "Executed Instructions", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Block Size", "Grid Size", "Theoretical Occupancy", "Achieved Occupancy",
        "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Compute (SM) Throughput", "Branch Efficiency"]
print("== headline (kernel id 0)")
for r in det[1:]:
    if r[idx["ID"]] == "0" and r[idx["Metric Name"]] in want:
        print(f'{r[idx["Metric Name"]]:45s} {r[idx["Metric Value"]]:>16s} {r[idx["Metric Unit"]]}')
raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
if len(raw) > 2:
    h = raw[0]
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
                 "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"):
        if name in h:
            print(f"{name:60s} {raw[2][h.index(name)]} {raw[1][h.index(name)]}")
rows = csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass")))
cur_file, cur_line, hdr = None, None, None
per = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
stall = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0].isdigit():
        cur_line = int(r[0])
        try:
            per[(cur_file, cur_line)][0] += float(r[7]); per[(cur_file, cur_line)][1] += float(r[6]); per[(cur_file, cur_line)][2] += float(r[8])
        except Exception:
            pass
        for i, name in enumerate(hdr):
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    stall[name] += float(r[i])
                except Exception:
                    pass
fn_starts = []
for fname in ("step_kernel.cuh", "env_kernel.cuh", "warp_prims.cuh", "mjb_batch.cu"):
    for i, l in enumerate(open(os.path.join(root, "mujoco_rl_environment_wrapper_b200", "csrc", fname)), 1):
        m = re.match(r"^(?:MJB_DEV_NOINLINE|MJB_DEV|MJB_HD|__global__|__device__ __forceinline__)\s+[\w:<>\*& ]+?\s+\**(\w+)\s*\(", l)
        if m:
            fn_starts.append((fname, i, m.group(1)))
def fn_of(f, ln):
    best = None
    for ff, i, n in fn_starts:
        if ff == f and i <= ln:
            best = n
    return best
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (f, ln), v in per.items():
    k = (f, fn_of(f, ln))
    for i in range(3):
        agg[k][i] += v[i]
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"== warp stall samples (all): total {sum(stall.values()):.0f}")
for k, v in stall.most_common(9):
    print(f"{k:28s} {100 * v / max(1, sum(stall.values())):5.1f}%")
print(f"== per source function: executed warp instructions {ti:.0f}, samples {ts:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{k[0]:18s} {str(k[1]):18s} instr {100 * v[0] / ti:5.1f}%  samples {100 * v[1] / ts:5.1f}%  lane-util {v[2] / max(1, v[0]) / 32:4.2f}")
