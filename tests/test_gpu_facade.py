"""GPU: the reference-compatible C++ face.  cudasbmp_b200/bin/ref_main_on_b200 is the REFERENCE's demos/main.cu
compiled unchanged against cudasbmp_b200/include (built by __graft_entry__.build() where /root/reference exists);
kgmt_demo is our data-driven demo over the same KGMT class."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cudasbmp_b200", "bin")
CSV13 = ["samples.csv", "unexploredSamples.csv", "parentRelations.csv", "uParentIdx.csv", "G.csv", "R2Avail.csv",
         "R1Avail.csv", "R1Valid.csv", "R2Valid.csv", "R1Invalid.csv", "R2Invalid.csv", "R1Score.csv", "R1.csv"]


def _layout(tmp_path):
    from cudasbmp_b200 import workloads as w
    cfg = tmp_path / "configurations"
    (cfg / "obstacles").mkdir(parents=True)
    with open(cfg / "obstacles" / "obstacles.csv", "w") as f:
        for o in w.C1_OBSTACLES:
            f.write(",".join("%g" % v for v in o) + "\n")
        f.write("\n")                                    # the shipped file ends with a blank line
    build = tmp_path / "build"
    build.mkdir()
    return cfg, build


def test_reference_main_runs_unchanged_on_the_b200_library(tmp_path):
    exe = os.path.join(BIN, "ref_main_on_b200")
    if not os.path.exists(exe):
        pytest.skip("ref_main_on_b200 not built (needs /root/reference at build time)")
    _, build = _layout(tmp_path)
    out = subprocess.run([exe], cwd=build, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "numObstacles: 5" in out.stdout and "Goal: 2.000000, 18.000000" in out.stdout
    assert "time inside KGMT is" in out.stdout and "Tree size" in out.stdout
    for nm in CSV13:
        assert (build / nm).exists(), nm
    par = np.loadtxt(build / "parentRelations.csv", dtype=np.int64)
    assert len(par) == 30000 and par[0] == -1
    T = int((par >= 0).sum()) + 1
    s = np.loadtxt(build / "samples.csv", delimiter=",")
    assert s.shape == (30000, 7) and tuple(s[0, :2]) == (5.0, 5.0)
    assert (par[1:T] < np.arange(1, T)).all() and (s[T:] == 0).all()


def test_data_driven_demo(tmp_path):
    exe = os.path.join(BIN, "kgmt_demo")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    cfg, build = _layout(tmp_path)
    for sub, txt in (("init", "5,5,0,0,0,0,0"), ("goal", "2,18,0,0,0,0,0"), ("numR1", "16"), ("R2", "8")):
        (cfg / sub).mkdir()
        name = {"init": "init.csv", "goal": "goal.csv", "numR1": "numR1.csv", "R2": "numR2.csv"}[sub]
        (cfg / sub / name).write_text(txt)
    out = subprocess.run([exe, str(cfg), "5"], cwd=build, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "obstacles 5 seed 5" in out.stdout
    if "stop 1" in out.stdout:
        lines = [l for l in out.stdout.splitlines() if l.startswith("  ")]
        first, last = lines[0].split(), lines[-1].split()
        assert float(first[2]) == 5.0 and float(first[4]) == 5.0
        assert np.hypot(float(last[2]) - 2.0, float(last[4]) - 18.0) < 0.5
    # the same seed through the Python face gives the same tree size
    from cudasbmp_b200 import kgmt as K, workloads as w
    p = K.KGMT(**w.C1, seed=5, record_candidates=True)
    r = p.plan(w.C1_INIT, w.C1_GOAL, w.C1_OBSTACLES)
    assert ("Tree size %d" % r["tree_size"]) in out.stdout


def _multi_demo(args, tmp_path):
    exe = os.path.join(BIN, "kgmt_multi_demo")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    return subprocess.run([exe] + [str(a) for a in args], cwd=tmp_path, capture_output=True, text=True, timeout=300)


def test_cpp_multi_gpu_demo_one_rank(tmp_path):
    """The C++ multi-GPU host program (one process per GPU, NCCL unique id through a shared page, kgmt_comm_init,
    kgmt_plan_sharded / kgmt_plan_batch_sharded / kgmt_plan_portfolio / kgmt_expand_sharded) with a world of ONE: every
    collective and the fused kernel run against themselves; the sharded plan must equal kgmt_plan."""
    out = _multi_demo([1], tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "multi-GPU demo ok (world 1)" in out.stdout and "identical 1" in out.stdout
    assert "plan_portfolio world 1: winner rank 0" in out.stdout
    for ex in (0, 1, 2):
        assert ("expand_sharded world 1 exchange %d: iteration 4" % ex) in out.stdout, out.stdout
    # the three exchanges expand the same four iterations into the same tree
    sizes = set(l.split("tree ")[1].split()[0] for l in out.stdout.splitlines() if l.startswith("expand_sharded"))
    assert len(sizes) == 1, out.stdout


def test_cpp_multi_gpu_demo_all_gpus(tmp_path):
    """The same program over every GPU of the box (>= 2, else skipped), reference demo map and the config-2 map."""
    import torch
    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip("needs at least 2 GPUs")
    from cudasbmp_b200 import workloads as w
    out = _multi_demo([G], tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    assert ("multi-GPU demo ok (world %d)" % G) in out.stdout and "identical 1" in out.stdout
    csv = tmp_path / "c2.csv"
    with open(csv, "w") as f:
        for o in w.c2_obstacles(1000):
            f.write(",".join("%.9g" % v for v in o) + "\n")
    out = _multi_demo([G, csv, 16, 32, 1 << 20], tmp_path)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "identical 1" in out.stdout
