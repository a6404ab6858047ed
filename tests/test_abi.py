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
    assert lib.kgmt_abi_version() == 1
    p = k.kgmt.default_params()
    # demos/main.cu:19-28
    assert (p.width, p.height, p.N, p.n, p.num_iterations, p.max_tree_size, p.num_disc) == (20.0, 20.0, 16, 8, 100, 30000, 10)
    assert (p.agent_length, p.goal_threshold) == (1.0, 0.5)


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
