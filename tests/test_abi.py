"""CPU: the C-ABI library loads and exports every symbol include/kgmt_c.h declares; without a GPU the
product refuses to compute (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kgmt_c.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kgmt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    import cudasbmp_b200 as k
    if not os.path.exists(k.LIB_PATH):
        g.build()
    return k.load()


def test_header_and_binding_agree():
    import cudasbmp_b200 as k
    assert _declared() == sorted(k.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), name


def test_abi_version_and_defaults(lib):
    import cudasbmp_b200 as k
    assert lib.kgmt_abi_version() == 2
    p = k.kgmt.default_params()
    # demos/main.cu:19-28
    assert (p.width, p.height, p.N, p.n, p.num_iterations, p.max_tree_size, p.num_disc) == (20.0, 20.0, 16, 8, 100, 30000, 10)
    assert (p.agent_length, p.goal_threshold) == (1.0, 0.5)
    # statePropagator.cu:17-19: a in (-5, 5], steering in (-pi, pi], duration in (0.05, 1.05]
    import math
    import numpy as np
    assert (p.accel_min, p.accel_max, p.steer_min, p.steer_max) == (-5.0, 5.0, -math.pi, math.pi)
    assert p.duration_min == float(np.float32(0.05)) and p.duration_max == float(np.float32(0.05)) + 1.0


def test_car_yaml_loader(lib, tmp_path):
    """systems/car.yaml of the reference is an EMPTY file: loading it must leave the defaults; a populated file
    overrides exactly its keys; unknown keys are reported with their line."""
    import cudasbmp_b200 as k
    empty = tmp_path / "car.yaml"
    empty.write_text("")
    p = k.kgmt.default_params()
    bad = C.c_int(-1)
    assert lib.kgmt_params_from_yaml(os.fsencode(str(empty)), C.byref(p), C.byref(bad)) == 0
    d = k.kgmt.default_params()
    assert all(getattr(p, f) == getattr(d, f) for f, _ in p._fields_ if f != "reserved")
    full = tmp_path / "car2.yaml"
    full.write_text("# kinematic bicycle\nwheelbase: 2.5\ncontrols:\n  accel_min: -3\n  accel_max: 2.5  # m/s^2\n"
                    "  steer_min: -0.25pi\n  steer_max: 0.5\n  duration_min: 0.1\n  duration_max: 0.6\nnum_disc: 20\n")
    assert lib.kgmt_params_from_yaml(os.fsencode(str(full)), C.byref(p), C.byref(bad)) == 0
    import math
    assert (p.agent_length, p.num_disc, p.accel_min, p.accel_max) == (2.5, 20, -3.0, 2.5)
    assert (p.steer_min, p.steer_max, p.duration_min, p.duration_max) == (-0.25 * math.pi, 0.5, 0.1, 0.6)
    wrong = tmp_path / "car3.yaml"
    wrong.write_text("wheelbase: 1\nturbo: 9\n")
    assert lib.kgmt_params_from_yaml(os.fsencode(str(wrong)), C.byref(p), C.byref(bad)) == -1 and bad.value == 2


def test_shipped_car_yaml_is_the_reference_model(lib):
    """systems/car.yaml of this repo spells out the literals of statePropagator.cu:17-19: loading it must give the same
    EFFECTIVE control transform as the defaults (accelerations and durations are narrowed to float, steering stays double)."""
    import numpy as np
    import cudasbmp_b200 as k
    p, d = k.kgmt.default_params(), k.kgmt.default_params()
    bad = C.c_int(0)
    assert lib.kgmt_params_from_yaml(os.fsencode(os.path.join(ROOT, "systems", "car.yaml")), C.byref(p), C.byref(bad)) == 0, bad.value
    f = np.float32
    assert (p.agent_length, p.num_disc) == (d.agent_length, d.num_disc)
    assert (f(p.accel_max - p.accel_min), f(p.accel_min)) == (f(d.accel_max - d.accel_min), f(d.accel_min))
    assert (p.steer_max - p.steer_min, p.steer_min) == (d.steer_max - d.steer_min, d.steer_min)
    assert (f(p.duration_max - p.duration_min), f(p.duration_min)) == (f(d.duration_max - d.duration_min), f(d.duration_min))


def test_no_cpu_fallback(lib):
    import torch
    import cudasbmp_b200 as k
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(k.KgmtError, match="no CPU fallback"):
        k.KGMT()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cudasbmp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "kgmt_oracle" not in text and "libref_" not in text, f
