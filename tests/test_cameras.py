"""Agent cameras (reference get_camera_data, mujoco_parent.py:496-575).

The reference renders with MuJoCo's OpenGL pipeline, which cannot run here: image parity with the reference is
UNPINNED.  What is pinned: the camera frames the compiler derives from the MJCF (against hand-computed
values), the oracle's image formation (against closed-form silhouettes / shading), and — on the GPU — the CUDA
raycaster against the oracle."""
import os

import numpy as np
import pytest

from mujoco_rl_environment_wrapper_b200 import _lib as L
from oracle.sim import OracleSim

LV = os.path.join(os.path.dirname(__file__), "levels")

SCENE = """<mujoco><compiler angle="degree"/><option timestep="0.002"/>
 <worldbody>
  <geom name="floor" type="plane" size="0 0 1" pos="0 0 0" rgba="0.2 0.6 0.4 1"/>
  <camera name="front" pos="0 0 1" xyaxes="0 -1 0 0 0 1" fovy="60"/>
  <camera name="down" pos="0 0 5" quat="1 0 0 0" fovy="40"/>
  <camera name="follow" mode="trackcom" pos="0 -3 0.3" xyaxes="1 0 0 0 0 1"/>
  <body name="ball" pos="4 0 1"><freejoint name="root"/><geom name="ball_geom" type="sphere" size="0.5" rgba="1 0 0.5 1"/>
   <camera name="onboard" pos="0 0 0.7" euler="90 0 0"/></body>
  <body name="crate" pos="0 0 0.25"><geom name="crate_geom" type="box" size="0.5 0.3 0.25" rgba="0 0 255 1"/></body>
  <body name="ghost" pos="2 0 1"><geom name="ghost_geom" type="sphere" size="0.4" rgba="1 1 1 0"/></body>
 </worldbody></mujoco>"""


def test_compiler_reads_cameras():
    m = L.Model(SCENE)
    assert m.ncam == 4
    assert [m.name2id(L.OBJ_CAMERA, n) for n in ("front", "down", "follow", "onboard", "nope")] == [0, 1, 2, 3, -1]
    f = m.fields
    assert list(f["cam_bodyid"]) == [0, 0, 0, m.name2id(L.OBJ_BODY, "ball")]
    assert list(f["cam_mode"]) == [0, 0, 1, 0]
    assert np.allclose(f["cam_fovy"], [60, 40, 45, 45])
    q = f["cam_quat"].reshape(4, 4)
    # xyaxes "0 -1 0  0 0 1": x = -Y, y = +Z, z = x cross y = -X  -> the camera looks along -z = +X
    w, x, y, z = q[0]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    assert np.allclose(R, np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0]]), atol=1e-12)
    assert np.allclose(q[1], [1, 0, 0, 0])
    assert np.allclose(q[3], [np.sqrt(0.5), np.sqrt(0.5), 0, 0])   # euler 90 0 0


def test_reference_levels_keep_their_agent_cameras():
    from mujoco_rl_environment_wrapper_b200.tables import Tables
    text = open(os.path.join(LV, "two_ants_cams.xml")).read()
    m = L.Model(text)
    t = Tables(text, m, ["sender", "receiver"], False)
    assert t.rgb_sensors == {"sender": ["sender_camera"], "receiver": ["receiver_camera"]}
    assert m.ncam == 2 and np.allclose(m.fields["cam_pos"], [0, -0.6, 0, 0, -0.6, 0])
    # cameras do not change the physics model
    m0 = L.Model(open(os.path.join(LV, "two_ants.xml")).read())
    for k in ("body_pos", "geom_size", "qpos0", "body_mass"):
        assert np.array_equal(m.fields[k], m0.fields[k])


def test_oracle_image_formation_closed_forms():
    m = L.Model(SCENE)
    s = OracleSim(m.blob)
    s.reset()
    s.forward()
    W = H = 128
    front = s.render(0, W, H).astype(int)
    # (1) the ball: centre at distance 4 on the optical axis, radius 0.5 -> silhouette is a disc of image-plane
    # radius tan(asin(r / D)) / tan(fovy / 2) * H / 2 pixels; the invisible ghost sphere in front of it is skipped
    ball = (front[:, :, 0] > 100) & (front[:, :, 1] == 0)
    rho = np.tan(np.arcsin(0.5 / 4.0)) / np.tan(np.radians(30)) * H / 2
    assert abs(ball.sum() - np.pi * rho * rho) / (np.pi * rho * rho) < 0.04
    ys, xs = np.nonzero(ball)
    assert abs(xs.mean() - (W - 1) / 2) < 0.51 and abs(ys.mean() - (H - 1) / 2) < 0.51
    # head-on at the centre: |n.d| = 1 -> full intensity of rgba (1, 0, 0.5)
    assert abs(front[H // 2, W // 2, 0] - 255) <= 1 and abs(front[H // 2, W // 2, 2] - 128) <= 1
    # limb darkening towards the ambient term
    edge = front[H // 2, int((W - 1) / 2 + rho - 1), 0]
    assert 0.4 * 255 - 2 <= edge < 0.75 * 255
    # (2) horizon: the camera is 1 m above an infinite floor and looks horizontally -> rows below the middle see
    # the floor (green dominant), rows above see nothing except the ball
    left = front[:, 5]
    assert (left[: H // 2, 1] > 0).all() and (left[H // 2:].sum(axis=1) == 0).all()
    # floor intensity at pixel row iy: |n.d| = |d_z|
    iy = 10
    dl = np.array([((5 + 0.5) / W * 2 - 1) * np.tan(np.radians(30)), ((iy + 0.5) / H * 2 - 1) * np.tan(np.radians(30)), -1.0])
    nd = abs(dl[1]) / np.linalg.norm(dl)
    assert abs(front[iy, 5, 1] - round(0.6 * (0.4 + 0.6 * nd) * 255)) <= 1
    # (3) looking straight down onto the crate's top face: uniform full-intensity blue (rgba 255 clamps to 1), and
    # its footprint 1.0 x 0.6 at distance 4.5
    down = s.render(1, W, H).astype(int)
    blue = (down[:, :, 2] > 250) & (down[:, :, 0] == 0) & (down[:, :, 1] == 0)
    px_per_m = H / 2 / (np.tan(np.radians(20)) * 4.5)
    assert abs(blue.sum() - 1.0 * 0.6 * px_per_m ** 2) / (0.6 * px_per_m ** 2) < 0.08
    assert down[H // 2, W // 2].tolist() == [0, 0, 255]
    # camera +y (image up) is world +Y for the identity orientation: the 1.0-long side (x) is horizontal
    ys, xs = np.nonzero(blue)
    assert (xs.max() - xs.min()) > 1.4 * (ys.max() - ys.min())
    # (4) a camera on a moving body follows it
    before = s.render(3, 32, 32)
    s.array("qpos")[:3] = [0.0, -3.0, 1.0]   # the onboard camera looks along body +y: now at the crate
    s.forward()
    after = s.render(3, 32, 32)
    assert not np.array_equal(before, after) and (after[:, :, 2] > 120).any() and not (before[:, :, 2] > 120).any()
    # tracking cameras are refused, not silently rendered as fixed ones
    with pytest.raises(ValueError):
        s.render(2, 8, 8)


@pytest.mark.gpu
def test_cuda_raycaster_matches_oracle():
    import torch
    from common import load_scene, make_spec
    from mujoco_rl_environment_wrapper_b200.batch import Batch
    from mujoco_rl_environment_wrapper_b200.tables import Tables
    text = open(os.path.join(LV, "two_ants_cams.xml")).read()
    model = L.Model(text)
    tables = Tables(text, model, ["sender", "receiver"], False)
    spec, keep = make_spec(model, tables, ["sender", "receiver"], False)
    N, W, H = 24, 64, 64
    b = Batch(model, spec, N, keepalive=keep)
    b.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(120):   # let the ants fall, flail and turn so that the cameras see different things
        b.actions[:, :, :spec.n_phys_act] = torch.rand((N, 2, spec.n_phys_act), generator=g, device="cuda") * 2 - 1
        b.step()
    img = b.render([0, 1], W, H).cpu().numpy()
    assert img.shape == (N, 2, H, W, 3) and img.dtype == np.uint8
    qpos = b.qpos.cpu().numpy()
    sim = OracleSim(model.blob)
    bad = total = 0
    for e in range(N):
        sim.array("qpos")[:] = qpos[e, :model.nq]
        sim.forward()
        for k in range(2):
            ref = sim.render(k, W, H).astype(int)
            diff = np.abs(img[e, k].astype(int) - ref).max(axis=2)
            bad += int((diff > 1).sum())
            total += diff.size
            # away from silhouette edges / grazing hits the two agree to one 8-bit level
    assert bad / total < 0.005, f"{bad} of {total} pixels differ by more than one level"
    assert len({img[e].tobytes() for e in range(N)}) == N, "every env renders its own state"
    assert (img.reshape(N, -1).max(axis=1) > 0).all()
    # non multiple-of-4 width takes the byte-wise store path
    odd = b.render([1], 30, 20).cpu().numpy()
    sim.array("qpos")[:] = qpos[3, :model.nq]
    sim.forward()
    ref = sim.render(1, 30, 20).astype(int)
    assert (np.abs(odd[3, 0].astype(int) - ref).max(axis=2) > 1).mean() < 0.01
    with pytest.raises(Exception, match="camera id"):
        b.render([5], W, H)


@pytest.mark.gpu
def test_get_camera_data_api():
    import torch
    from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
    cfg = {"xmlPath": os.path.join(LV, "two_ants_cams.xml"), "agents": ["sender", "receiver"], "agentCameras": True}
    env = MuJoCoRL(dict(cfg, num_envs=6, sensorResolution=(32, 32)))
    env.reset()
    a = env.get_camera_data("sender")
    assert torch.is_tensor(a) and tuple(a.shape) == (6, 1, 32, 32, 3) and a.dtype == torch.uint8
    c = env.get_camera_data("receiver_camera")
    assert tuple(c.shape) == (6, 32, 32, 3)
    assert torch.equal(c, env.get_camera_data("receiver")[:, 0])
    with pytest.raises(KeyError):
        env.get_camera_data("no_such_camera")
    one = MuJoCoRL(cfg)   # num_envs = 1: numpy, the reference's [n_cams, 64, 64, 3]
    one.reset()
    img = one.get_camera_data("sender")
    assert isinstance(img, np.ndarray) and img.shape == (1, 64, 64, 3) and img.dtype == np.uint8
    assert one.get_camera_data("sender_camera").shape == (64, 64, 3)
    with pytest.raises(Exception, match="agentCameras"):
        MuJoCoRL({"xmlPath": cfg["xmlPath"], "agents": cfg["agents"]}).get_camera_data("sender")
    nocam = MuJoCoRL({"xmlPath": os.path.join(LV, "two_ants.xml"), "agents": cfg["agents"], "agentCameras": True, "num_envs": 2})
    assert tuple(nocam.get_camera_data("sender").shape) == (2, 0, 64, 64, 3)


@pytest.mark.gpu
def test_cameras_follow_each_envs_level(tmp_path):
    """xmlPath list + agentCameras: every env is rendered with the geometry / colours of its own level"""
    import torch
    from mujoco_rl_environment_wrapper_b200.mujoco_rl import MuJoCoRL
    a = open(os.path.join(LV, "two_ants_cams.xml")).read()
    b = a.replace('size="0.5 0.5 0.5" rgba="0 0 255 1"', 'size="0.9 0.9 0.9" rgba="0 255 0 1"')
    assert a != b
    pa, pb = str(tmp_path / "A.xml"), str(tmp_path / "B.xml")
    open(pa, "w").write(a)
    open(pb, "w").write(b)
    common = dict(agents=["sender", "receiver"], agentCameras=True, sensorResolution=(32, 32), num_envs=16, seed=5)
    multi = MuJoCoRL(dict(common, xmlPath=[pa, pb]))
    singles = [MuJoCoRL(dict(common, xmlPath=pa)), MuJoCoRL(dict(common, xmlPath=pb))]
    for e in [multi] + singles:
        e.reset()
    lid = multi.level_id.cpu().numpy()
    assert set(lid.tolist()) == {0, 1}
    for t in range(3):
        acts = multi.sample_actions()
        for e in [multi] + singles:
            e.step(acts)
    got = multi.get_camera_data("sender").cpu().numpy()
    want = [s.get_camera_data("sender").cpu().numpy() for s in singles]
    for e in range(16):
        assert np.array_equal(got[e], want[lid[e]][e]), e
    assert not np.array_equal(want[0], want[1]), "the changed box must be visible to some camera"


def test_oracle_reproduces_committed_camera_fixture():
    """tests/golden/camera_two_ants.npz: both agent cameras of the two-ant level at a stored state, written by the
    oracle (the script is this test's body run with np.savez).  Guards the image formation against silent changes."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_two_ants.npz"))
    m = L.Model(open(os.path.join(LV, "two_ants_cams.xml")).read())
    s = OracleSim(m.blob)
    s.array("qpos")[:] = g["qpos"]
    s.forward()
    imgs = np.stack([s.render(k, 64, 64) for k in range(2)])
    assert np.array_equal(imgs, g["images"])
    assert 0.3 < (imgs.sum(axis=-1) > 0).mean() < 0.9   # floor below the horizon, sky above
