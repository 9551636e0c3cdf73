"""Every field of the compiled model blob (csrc/mjcf_compile.cpp) against an INDEPENDENT derivation from the XML
(oracle/model_ref.py: numpy + ElementTree, rotation matrices, parallel-axis inertias, dense J M^-1 J^T).

The oracle consumes the product compiler's blob, so CUDA-vs-oracle parity cannot see a wrong model constant;
this test is what pins them.  Closed forms that anchor the derivation itself are at the bottom (free sphere:
invweight0 = 1/m, 1/I; capsule inertia by numerical quadrature; pair counts of SURVEY A.1b)."""
import glob
import os

import numpy as np
import pytest

from common import LEVELS
from mujoco_rl_environment_wrapper_b200 import _lib as L
from oracle.model_ref import RefModel, quat2mat

ALL = sorted(glob.glob(os.path.join(LEVELS, "*.xml")))
NOT_DERIVED = {"opt_impratio", "opt_tolerance", "opt_iterations", "sensor_objtype", "ncam", "cam_bodyid", "cam_mode", "cam_pos",
               "cam_quat", "cam_fovy"}


@pytest.mark.parametrize("path", ALL, ids=[os.path.basename(p) for p in ALL])
def test_blob_equals_independent_derivation(path):
    text = open(path).read()
    model = L.Model(text)
    ref = RefModel(text)
    checked = 0
    for name, got in model.fields.items():
        if name in NOT_DERIVED:
            continue
        assert name in ref.f, f"blob field '{name}' has no independent derivation"
        want = ref.f[name]
        got = np.asarray(got)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if name in RefModel.QUAT_FIELDS:
            # q and -q are the same rotation: compare the rotation matrices
            for k in range(len(want) // 4):
                a, b = got[4 * k:4 * k + 4], want[4 * k:4 * k + 4]
                if np.linalg.norm(a) == 0 and np.linalg.norm(b) == 0:
                    continue
                assert np.abs(quat2mat(a) - quat2mat(b)).max() < 1e-12, (name, k)
        elif want.dtype.kind == "i":
            assert np.array_equal(got.astype(np.int64), want), name
        else:
            scale = np.maximum(1.0, np.abs(want))
            assert (np.abs(got - want) / scale).max() < 1e-12 if len(want) else True, (name, np.abs(got - want).max())
        checked += 1
    assert checked >= 70, checked
    # nothing the derivation knows about is missing from the blob
    assert set(ref.f) <= set(model.fields), set(ref.f) - set(model.fields)


def test_free_sphere_invweight_is_one_over_mass_and_inertia():
    xml = """<mujoco><worldbody><body name="s" pos="0.3 -0.2 1.5"><freejoint/><geom type="sphere" size="0.25" density="5"/></body>
             </worldbody></mujoco>"""
    for fields in (L.Model(xml).fields, RefModel(xml).f):
        m = 5 * 4 / 3 * np.pi * 0.25 ** 3
        I = 0.4 * m * 0.25 ** 2
        assert abs(fields["body_mass"][1] - m) < 1e-15
        assert np.allclose(fields["body_inertia"][3:6], I, rtol=1e-14)
        assert np.allclose(fields["dof_invweight0"], [1 / m] * 3 + [1 / I] * 3, rtol=1e-12)
        assert np.allclose(fields["body_invweight0"][2:4], [1 / m, 1 / I], rtol=1e-12)


def test_hinge_pendulum_invweight_closed_form():
    """point-like sphere of mass m at distance l from a hinge with armature a: M = m l^2 + I + a; dof_invweight0 = 1 / M;
    body_invweight0 (translation) = l^2 / M * (1/3) * |direction|^2 summed over xyz = l^2 / (3 M)"""
    xml = """<mujoco><compiler angle="radian"/><worldbody><body name="p" pos="0 0 2"><joint type="hinge" axis="0 1 0" armature="0.3"/>
             <geom type="sphere" size="0.1" pos="0.7 0 0" density="1000"/></body></worldbody></mujoco>"""
    m = 1000 * 4 / 3 * np.pi * 1e-3
    I = 0.4 * m * 0.01
    Mj = m * 0.49 + I + 0.3
    for fields in (L.Model(xml).fields, RefModel(xml).f):
        assert abs(fields["dof_invweight0"][0] - 1 / Mj) < 1e-12
        assert abs(fields["body_invweight0"][2] - 0.49 / (3 * Mj)) < 1e-12
        assert abs(fields["body_invweight0"][3] - 1 / (3 * Mj)) < 1e-12


def test_capsule_inertia_by_quadrature():
    """the capsule formula (both derivations) against a direct numerical integration over the solid"""
    r, half = 0.08, 0.2
    m, I = RefModel.geom_inertia(3, [r, half, 0], 5.0, None)
    z = np.linspace(-(half + r), half + r, 400001)
    rad2 = np.where(np.abs(z) <= half, r * r, np.maximum(0.0, r * r - (np.abs(z) - half) ** 2))
    dz = z[1] - z[0]
    vol = np.trapezoid(np.pi * rad2, dx=dz)
    izz = np.trapezoid(5.0 * np.pi * rad2 ** 2 / 2, dx=dz)
    ixx = np.trapezoid(5.0 * (np.pi * rad2 ** 2 / 4 + np.pi * rad2 * z * z), dx=dz)
    assert abs(m - 5.0 * vol) / m < 1e-9
    assert abs(I[2] - izz) / izz < 1e-8 and abs(I[0] - ixx) / ixx < 1e-8
    xml = f'<mujoco><worldbody><body><freejoint/><geom type="capsule" size="{r} {half}" density="5"/></body></worldbody></mujoco>'
    f = L.Model(xml).fields
    assert abs(f["body_mass"][1] - m) < 1e-15 and np.allclose(f["body_inertia"][3:6], I, rtol=1e-13)


def test_pair_table_counts_of_the_reference_levels():
    """candidate pairs by type for the reference's levels, from the independent filter"""
    def by_type(path):
        f = RefModel(open(os.path.join(LEVELS, path)).read()).f
        t = f["geom_type"]
        out = {}
        for a, b in zip(f["pair_geom1"], f["pair_geom2"]):
            out[(int(t[a]), int(t[b]))] = out.get((int(t[a]), int(t[b])), 0) + 1
        return out
    assert by_type("ant_rk4.xml") == {(0, 2): 1, (0, 3): 12}
    assert by_type("two_ants.xml") == {(0, 2): 2, (0, 3): 24, (2, 6): 16, (3, 6): 192, (2, 3): 32, (2, 2): 1, (3, 3): 224}
    assert by_type("box_rangefinder.xml") == {(0, 6): 1, (6, 6): 8}
