"""The C-ABI library loads on a machine without a GPU and exports every symbol include/mjb.h declares."""
import ctypes
import os
import re

import pytest

from mujoco_rl_environment_wrapper_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mjb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mjb_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol():
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libmjb.so does not export {n}"


def test_struct_sizes_match_header():
    # the header's structs are POD; sizes computed by hand from the declarations
    assert ctypes.sizeof(L.Plugin) == 32
    assert ctypes.sizeof(L.Layout) == 44
    assert ctypes.sizeof(L.Buffers) == 21 * 8
    assert ctypes.sizeof(L.Dims) == 56


def test_version_and_error_channel():
    lib = L.load()
    assert b"sm_100a" in lib.mjb_version()
    h = ctypes.c_void_p()
    assert lib.mjb_model_create(b"<mujoco>", ctypes.byref(h)) != 0
    assert b"XML" in lib.mjb_last_error()


def test_draw_stream_is_exported_and_deterministic():
    lib = L.load()
    a = [lib.mjb_draw_u32(1234, e, 0, 0) for e in range(4)]
    assert a == [lib.mjb_draw_u32(1234, e, 0, 0) for e in range(4)]
    assert len(set(a)) == 4


def test_batch_layout_without_gpu():
    from common import load_scene, make_spec
    model, tables, agents, fj = load_scene("2A")
    spec, keep = make_spec(model, tables, agents, fj, dynamics=[(L.DYN_LANGUAGE, 1, 1, 0.0)])
    lay = L.Layout()
    L.check(L.load().mjb_batch_layout(model._h, ctypes.byref(spec), 4096, ctypes.byref(lay)))
    assert (lay.qpos_stride, lay.qvel_stride, lay.obs_stride, lay.act_stride) == (32, 28, 60, 12)


def test_product_does_not_import_oracle():
    """The shipped package must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "mujoco_rl_environment_wrapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "mj_oracle" not in text, f


def test_oracle_includes_no_product_source():
    """the checker must stay independent of the thing it checks: oracle/ may include the public blob format
    (include/mjb_blob.h) and nothing from the package (csrc/)"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "oracle", "*.cpp")) + glob.glob(os.path.join(root, "oracle", "*.h")):
        for inc in re.findall(r'#include\s+"([^"]+)"', open(path).read()):
            assert "csrc" not in inc and "mujoco_rl_environment_wrapper_b200" not in inc, (path, inc)
