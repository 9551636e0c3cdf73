"""Physics goldens from REAL MuJoCo — runs only where `import mujoco` succeeds (it does not in the build container
nor on the GPU boxes of this round: `mujoco==2.3.3` is not installable offline).  For every scene fixture
it steps MuJoCo with seeded ctrl sequences and stores qpos / qvel / sensordata / contact pairs per step in
tests/golden/physics_<scene>.npz; tests/test_oracle_analytic.py::test_oracle_against_mujoco_goldens then pins the
fp64 oracle to them (and through the oracle, the CUDA path).  Until such a file exists the oracle is pinned by
closed-form known answers only and physics parity stays UNPINNED."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LEVELS = os.path.join(HERE, "..", "levels")
SCENES = ["two_ants.xml", "one_ant_arena.xml", "ant_rk4.xml", "two_ants_touch_acc.xml", "box_rangefinder.xml", "box_touch.xml",
          "box_accelerometer.xml", "box_framexaxis.xml"]
STEPS, HOLD = 400, 5


def main():
    try:
        import mujoco
    except ImportError:
        sys.exit("mujoco is not importable here: no physics goldens written (parity stays unpinned)")
    for scene in SCENES:
        model = mujoco.MjModel.from_xml_path(os.path.join(LEVELS, scene))
        data = mujoco.MjData(model)
        rng = np.random.default_rng(2024)
        mujoco.mj_resetData(model, data)
        mujoco.mj_forward(model, data)
        rec = {k: [] for k in ("ctrl", "qpos", "qvel", "sensordata", "ncon")}
        pairs = []
        ctrl = np.zeros(model.nu)
        for t in range(STEPS):
            if t % HOLD == 0 and model.nu:
                ctrl = rng.uniform(-1, 1, model.nu)
            data.ctrl[:] = ctrl
            mujoco.mj_step(model, data)
            rec["ctrl"].append(ctrl.copy()); rec["qpos"].append(data.qpos.copy()); rec["qvel"].append(data.qvel.copy())
            rec["sensordata"].append(data.sensordata.copy()); rec["ncon"].append(data.ncon)
            pairs.append(sorted((int(c.geom1), int(c.geom2)) for c in data.contact[:data.ncon]))
        out = os.path.join(HERE, "physics_" + scene.replace(".xml", ".npz"))
        flat = np.array([(t, a, b) for t, ps in enumerate(pairs) for a, b in ps], dtype=np.int32).reshape(-1, 3)
        np.savez_compressed(out, mujoco_version=mujoco.__version__, contact_pairs=flat, **{k: np.array(v) for k, v in rec.items()})
        print("wrote", out)


if __name__ == "__main__":
    main()
