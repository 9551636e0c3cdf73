"""First GPU contact: physics parity vs the fp64 oracle from identical states + a raw timing loop."""
import sys, time, ctypes, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from mujoco_rl_environment_wrapper_b200 import _lib as L
from mujoco_rl_environment_wrapper_b200.batch import Batch
from oracle import OracleSim
import emu_harness as E

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
xml = os.path.join(root, "tests", "levels", "two_ants.xml")
model = L.Model.from_xml_path(xml)
spec, keep = E.simple_spec(model, [model.name2id(L.OBJ_BODY, 'sender'), model.name2id(L.OBJ_BODY, 'receiver')],
                           [2, 3, 4, 5, 6, 7, 0, 1, 10, 11, 12, 13, 14, 15, 8, 9], 8, obs_sensors=[[0], [1]])
nq, nv, nu = model.nq, model.nv, model.nu
# ---- parity: K states along an oracle rollout
K = 256
ref = OracleSim(model.blob)
rng = np.random.default_rng(0)
states = []
for i in range(K * 4):
    c = rng.uniform(-1, 1, nu)
    if i % 4 == 0:
        states.append((ref.qpos.copy(), ref.qvel.copy(), ref.qacc_warmstart.copy(), c.copy()))
    ref.ctrl[:] = c
    ref.step()
b = Batch(model, spec, K, keepalive=keep)
print("geometry", b.geometry(), flush=True)
idx = [spec.act_index[k] for k in range(16)]
for e, (q, v, w, c) in enumerate(states):
    b.qpos[e, :nq] = torch.tensor(q, dtype=torch.float32); b.qvel[e, :nv] = torch.tensor(v, dtype=torch.float32)
    b.warmstart[e, :nv] = torch.tensor(w, dtype=torch.float32)
    b.actions[e, :, :8] = torch.tensor(c[idx], dtype=torch.float32).reshape(2, 8)
b.physics(1); b.sync()
gq, gv = b.qpos.cpu().numpy(), b.qvel.cpu().numpy()
ncon = b.ncon.cpu().numpy(); cg = b.contact_geom.cpu().numpy()
worst = 0; mism = 0
for e, (q, v, w, c) in enumerate(states):
    ref.qpos[:] = q; ref.qvel[:] = v; ref.qacc_warmstart[:] = w; ref.ctrl[:] = c
    ref.step()
    rel = (np.abs(gv[e, :nv] - ref.qvel) / np.maximum(1, np.abs(ref.qvel))).max()
    worst = max(worst, rel, np.abs(gq[e, :nq] - ref.qpos).max())
    if sorted(ref.contact_pairs()) != sorted((int(a), int(b_)) for a, b_ in cg[e, :ncon[e]]):
        mism += 1
print(f"PARITY K={K}: worst rel err {worst:.3e}, contact-set mismatches {mism}", flush=True)
# ---- timing
for N in (4096, 16384, 65536):
    bb = Batch(model, spec, N, keepalive=keep)
    bb.reset(); bb.sync()
    bb.actions.uniform_(-1, 1)
    for _ in range(20): bb.physics(1)
    bb.sync()
    steps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        bb.physics(1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"N={N}: {ms:.3f} ms/step, {N * 2 / ms * 1e3:.3e} agent-steps/s, geometry {bb.geometry()}, ncon mean {bb.ncon.float().mean().item():.2f}", flush=True)
    del bb
