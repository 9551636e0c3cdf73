"""Multi-GPU plumbing on CPU: env-index sharding and the end-of-rollout stats all-gather (gloo, world 2)."""
import os
import subprocess
import sys

import pytest

from mujoco_rl_environment_wrapper_b200.dist import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.environ["MJB_ROOT"])
from mujoco_rl_environment_wrapper_b200 import dist as D
rank, local, world = D.init_from_env("gloo")
lo, hi = D.shard_range(65537, rank, world)
stats = torch.tensor([float(rank), float(hi - lo), 1.5 * (rank + 1)])
allst = D.allgather_episode_stats(stats)
assert allst.shape == (world, 3)
assert int(allst[:, 1].sum().item()) == 65537
assert allst[:, 0].tolist() == [float(r) for r in range(world)]
m = D.max_over_ranks(10.0 + rank, device="cpu")
assert m == 10.0 + world - 1
open(os.path.join(os.environ["MJB_OUT"], f"rank{rank}.ok"), "w").write("ok")
"""


@pytest.mark.parametrize("total,world", [(4096, 1), (65536, 8), (10, 3), (7, 8)])
def test_shard_range_partitions(total, world):
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MJB_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MJB_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists(), res.stdout + res.stderr
