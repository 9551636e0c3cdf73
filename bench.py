#!/usr/bin/env python
"""bench.py — agent-steps/s of the batched MuJoCoRL.step hot path on B200 (BASELINE.json metric).

Workload (config C2, BASELINE.json configs[1]): benchmarking/levels/MultiAgentModel.xml,
agents sender + receiver, Language dynamic + tag-distance reward + done, freeJoint = False,
skipFrames = 1, 4096 envs PER GPU (weak scaling: envs shard by index, no collective on the step path),
synthetic random actions (pre-generated on the device, one device-to-device copy per step).

One "step" = one `MuJoCoRL.step` for every env of the rank = ONE fused kernel launch.
  value    device-resident throughput, per-step CUDA events on the launching stream, L2 flushed between steps
  e2e      the same through the C-ABI host-buffer call (mjb_step_host): pinned H2D of the actions and D2H of
           obs / rewards / flags inside the timed region
  roofline algorithmic HBM bytes per launch / kernel time vs MEASURED_PEAKS.json (physics is latency /
           issue bound, so the fraction is tiny by construction — see DESIGN.md)
  cpu_baseline / --impl reference: the fp64 oracle + reference-order host loop on the host cores
           (kind "port": the reference's MuJoCo dependency is not installable here).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
LEVELS = os.path.join(ROOT, "tests", "levels")
ENVS_PER_GPU = int(os.environ.get("MJB_BENCH_ENVS", "4096"))
SETTLE = int(os.environ.get("MJB_BENCH_SETTLE", "300"))
WORKLOAD = "C2: MultiAgentModel.xml, 2 agents, Language + tag-distance reward + done, ctrl mode, skipFrames=1"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def algorithmic_bytes_per_env_step(nq, nv, nu, nsens, n_agents, act_dim, obs_dim, nprobe):
    """fp32 state read + written once, actions read, obs / reward / flags written, plugin store
    read + written, qacc warm start read + written."""
    read = 4 * (nq + nv + nu + nv + n_agents * act_dim) + 28
    write = 4 * (nq + nv + nu + nv + nsens + n_agents * obs_dim + n_agents + 4 * nprobe) + 2 * (n_agents + 1) + 4 + 16
    return read + write


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle physics + reference-order host loop, one env per process
_CPU_ENV = None


def _cpu_env(seed):
    global _CPU_ENV
    if _CPU_ENV is None:
        sys.path.insert(0, ROOT)
        import numpy as np
        from mujoco_rl_environment_wrapper_b200 import _lib as L
        from mujoco_rl_environment_wrapper_b200.tables import Tables
        from oracle import host_loop as H
        text = open(os.path.join(LEVELS, "two_ants.xml")).read()
        model = L.Model(text)
        agents = ["sender", "receiver"]
        tables = Tables(text, model, agents, False)

        def resolve(name):
            b = model.name2id(L.OBJ_BODY, name)
            return (1, b) if b >= 0 else (5, model.name2id(L.OBJ_GEOM, name))
        rng = np.random.default_rng(seed)
        env = H.OracleEnv(model, tables, agents, dynamics=[H.Language], reward_functions=[H.tag_distance_reward],
                          done_functions=[H.distance_done], targets=["choice_1", "choice_2"],
                          draw=lambda a, k: int(rng.integers(0, 1 << 30)), resolve=resolve)
        env.reset({a: np.zeros(9) for a in agents})
        _CPU_ENV = (env, rng, agents, np)
    return _CPU_ENV


def _cpu_worker(args):
    seed, seconds = args
    env, rng, agents, np = _cpu_env(seed)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        act = {a: np.concatenate([rng.uniform(-1, 1, 8), rng.uniform(0, 3, 1)]) for a in agents}
        env.step(act)
        steps += 1
        if env.timestep >= 1024:
            env.reset({a: np.zeros(9) for a in agents})
    return steps, time.perf_counter() - t0


def cpu_arm(seconds, pool, cores):
    res = pool.map(_cpu_worker, [(1000 + i, seconds) for i in range(cores)], chunksize=1)
    env_steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return 2.0 * env_steps / wall, env_steps


def make_pool():
    cores = os.cpu_count() or 1
    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_cpu_worker, [(1000 + i, 0.05) for i in range(cores)], chunksize=1)  # build the envs
    return pool, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build_oracle
    build_oracle()
    pool, cores = make_pool()
    per_step_s = min(10.0, max(0.05, 100.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        v, n = cpu_arm(per_step_s, pool, cores)
        if i >= args.warmup:
            vals.append(v)
    pool.close()
    value = sum(vals) / len(vals)
    sample = f"{cores} processes x 1 env, {per_step_s:.2f} s of random-action steps per timed step, {WORKLOAD}"
    print(json.dumps({
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_process": 1, "processes": cores},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated-CPU (fp64 oracle + reference-order Python host loop), NOT MuJoCo: mujoco==2.3.3 is not installable here",
    }))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 10 ms on a thread (NVML), DURING the timed region."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_sm, self._stop, self._thr = index, [], set(), None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _run(self):
        nv = self._nv
        while not self._stop:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop = True
        if self._thr:
            self._thr.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(sm)}


def run_ours(args):
    import numpy as np
    import torch
    from mujoco_rl_environment_wrapper_b200 import dist as D
    from mujoco_rl_environment_wrapper_b200 import plugins as P
    from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
    rank, local, world = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = ENVS_PER_GPU
    env = MuJoCoRL({"xmlPath": os.path.join(LEVELS, "two_ants.xml"), "infoJson": os.path.join(LEVELS, "info_2A.json"),
                    "agents": ["sender", "receiver"], "skipFrames": 1, "maxSteps": 1024, "num_envs": N,
                    "seed": 1234 + 1000 * rank, "device": dev, "environmentDynamics": [P.Language],
                    "rewardFunctions": [P.tag_distance_reward], "doneFunctions": [P.distance_done]})
    b = env.batch
    A, act_dim = 2, env._act_dim
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + 1000 * rank)
    POOL = 32
    lo = torch.tensor([-1.0] * 8 + [0.0], device=dev)
    hi = torch.tensor([1.0] * 8 + [3.0], device=dev)
    pool = torch.zeros((POOL,) + tuple(b.actions.shape), device=dev)
    pool[..., :act_dim] = lo + torch.rand((POOL, N, A, act_dim), generator=gen, device=dev) * (hi - lo)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    env.reset()
    # settle onto the floor so that the timed steps carry contacts (the reference loop times whole episodes)
    for k in range(SETTLE):
        b.actions.copy_(pool[k % POOL])
        b.step()
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm
    for k in range(max(3, args.warmup)):
        b.actions.copy_(pool[k % POOL]); b.step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = b.launch_count
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush, outside the per-step events
        ev[k][0].record()
        b.actions.copy_(pool[k % POOL])    # this step's synthetic actions (device to device)
        kev[k][0].record()
        b.step()                           # the fused kernel
        kev[k][1].record()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = sum(a.elapsed_time(c) for a, c in ev) / args.steps
    kern_ms = sum(a.elapsed_time(c) for a, c in kev) / args.steps
    launches = b.launch_count - launches0
    step_ms = D.max_over_ranks(step_ms, dev)
    kern_ms_max = D.max_over_ranks(kern_ms, dev)

    # ---- end-to-end arm: C-ABI host-buffer call, pinned H2D / D2H inside the timed region
    h_act, h_obs, h_rew, h_term, h_trunc = b.host_arrays()
    # each step's actions arrive in page-locked host memory (four different buffers in rotation, as a policy's output
    # ring would be); results land in page-locked host arrays
    pinned_keep = [torch.zeros(h_act.shape, dtype=torch.float32, pin_memory=True) for _ in range(4)]
    host_acts = [t.numpy() for t in pinned_keep]
    for k in range(4):
        np.copyto(host_acts[k], pool[k].cpu().numpy())
    for k in range(4):
        b.step_host(host_acts[k % 4], h_obs, h_rew, h_term, h_trunc)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        b.step_host(host_acts[k % 4], h_obs, h_rew, h_term, h_trunc)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = D.max_over_ranks(e2e_ms, dev)
    clk = clocks.stop()

    # optional end-of-rollout collective (outside every timed region)
    stats = torch.tensor([float(b.reward.sum()), float(b.term[:, A].sum()), float(b.ncon.float().mean())], device=dev)
    allstats = D.allgather_episode_stats(stats)

    if rank == 0:
        m = env.model
        bytes_env = algorithmic_bytes_per_env_step(m.nq, m.nv, m.nu, m.nsensordata, A, act_dim, 60, b.layout.probe_count)
        peak, which = peaks()
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("envs_per_gpu") == N:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        achieved = bytes_env * N / (kern_ms * 1e-3) / 1e9
        # the bound that actually binds: warp-instruction issue.  Executed warp instructions per launch come from the
        # committed ncu capture of this workload; the rate is live (this run's kernel time and SM clock).
        issue = None
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("envs_per_gpu") == N and tj.get("warp_instructions"):
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = (clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz") or 1965
                peak_ips = sms * 4 * mhz * 1e6
                ips = tj["warp_instructions"] / (kern_ms * 1e-3)
                issue = {"warp_instructions_per_launch": tj["warp_instructions"], "achieved_ginst_s": ips / 1e9,
                         "peak_ginst_s": peak_ips / 1e9, "frac": ips / peak_ips,
                         "note": "4 issue slots per SM per cycle; instruction count from profiles/ (ncu), time and clock live"}
        agent_steps = N * A * world
        out = {
            "metric": "agent-steps/sec", "value": agent_steps / (step_ms * 1e-3), "unit": "agent-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": N, "agents": A, "l2": "flushed between steps (256 MiB write, outside the per-step events)",
                       "state": f"settled for {SETTLE} steps before timing", "geometry": b.geometry(), "wall_ms_per_step_incl_flush": t_wall * 1e3 / args.steps},
            "e2e": {"value": agent_steps / (e2e_ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h_act.nbytes), "d2h_bytes_per_step": int(h_obs.nbytes + h_rew.nbytes + h_term.nbytes + h_trunc.nbytes)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": which, "algorithmic_bytes_per_env_step": bytes_env, "kernel_ms": kern_ms_max,
                         "note": "physics-on step is issue/latency bound (SURVEY 8d): see profiles/ for issue-slot and stall evidence",
                         "issue": issue},
            "clocks": clk,
            "episode_stats_allgather": allstats.cpu().tolist(),
        }
        if world == 1 and not args.no_cpu:
            from oracle import build_oracle
            build_oracle()
            cpool, cores = make_pool()
            v, n = cpu_arm(12.0, cpool, cores)
            cpool.close()
            out["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                   "sample": f"{cores} processes x 1 env x 12 s of the same workload ({n} env-steps), fp64 oracle + reference-order host loop"}
        print(json.dumps(out))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
