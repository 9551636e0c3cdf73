"""fp64 CPU oracle of the MuJoCoRL.step hot path.

TEST INFRASTRUCTURE ONLY (parity unpinned, see mj_oracle.cpp header).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the shipped package never does.
"""
from .sim import OracleSim, build_oracle, oracle_lib_path  # noqa: F401
