"""Generates tests/golden/*.npz from the REFERENCE'S OWN code built in this container
(oracle/_ref, see oracle/Makefile).  Run here (needs /root/reference); the vectors travel, the
reference does not.

    python tests/golden/make_golden.py

propagate_*.npz : inputs (parents, parentOf, key, obstacles, numDisc, L, W, H) and the outputs of the
                  reference's unmodified propagateAndCheck + isMotionValid (host build, libref_host.so):
                  x1[M,7], valid[M], u3[M].
regions.npz     : (x, y) -> getR1 / getR2 of the reference (host side of its __host__ __device__ functions,
                  libref_gpu.so), for the C1 and C2 grids, including points on cell borders and outside.
philox.npz      : Random123 / cuRAND known answers, cross-checked against the glue's generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from cudasbmp_b200 import workloads as w  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def propagate_case(name, obstacles, P, children, key, num_disc, seed):
    parents = w.random_parents(P, obstacles, seed=seed)
    parent_of = np.repeat(np.arange(P, dtype=np.int32), children)
    _, x1, valid, u3 = po.ref_host_batch(parents, parent_of, key, 0, num_disc, 1.0, obstacles, 20.0, 20.0)
    np.savez_compressed(os.path.join(OUT, name), parents=parents, parent_of=parent_of, key=np.uint32(key),
                        obstacles=np.asarray(obstacles, np.float32), num_disc=np.int32(num_disc), L=np.float32(1.0),
                        W=np.float32(20.0), H=np.float32(20.0), x1=x1, valid=valid, u3=u3)
    print(name, "M=%d valid=%.3f" % (len(valid), valid.mean()))


def main():
    po.build()
    assert po.ref_host() is not None and po.ref_gpu() is not None, "reference builds missing (make -C oracle ref)"
    propagate_case("propagate_c1.npz", w.C1_OBSTACLES, 96, 32, 1234, 10, 3)
    propagate_case("propagate_c2.npz", w.c2_obstacles(), 64, 32, 77, 10, 5)
    propagate_case("propagate_c3.npz", w.c3_obstacles()[:2000], 32, 16, 99, 40, 9)
    # root at rest (v = 0): the first iteration of the demo
    root = w.C1_INIT[None, :]
    po_of = np.zeros(32, dtype=np.int32)
    _, x1, valid, u3 = po.ref_host_batch(root, po_of, 2, 0, 10, 1.0, w.C1_OBSTACLES, 20.0, 20.0)
    np.savez_compressed(os.path.join(OUT, "propagate_root.npz"), parents=root, parent_of=po_of, key=np.uint32(2),
                        obstacles=w.C1_OBSTACLES, num_disc=np.int32(10), L=np.float32(1.0), W=np.float32(20.0),
                        H=np.float32(20.0), x1=x1, valid=valid, u3=u3)

    R = po.ref_gpu()
    rng = np.random.default_rng(11)
    cases = {}
    for tag, (N, n) in {"c1": (16, 8), "c2": (16, 32), "n64": (64, 8)}.items():
        R1 = np.float32(20.0) / np.float32(N)
        R2 = np.float32(20.0) / np.float32(n * N)
        x = rng.uniform(-1.0, 21.0, 4000).astype(np.float32)
        y = rng.uniform(-1.0, 21.0, 4000).astype(np.float32)
        # exact cell borders and their float neighbours
        b = (np.arange(0, N * n + 1, dtype=np.float32) * R2).astype(np.float32)
        bx = np.concatenate([b, np.nextafter(b, np.float32(-1e9)), np.nextafter(b, np.float32(1e9))])
        x = np.concatenate([x, bx, rng.uniform(0, 20, len(bx)).astype(np.float32)])
        y = np.concatenate([y, rng.uniform(0, 20, len(bx)).astype(np.float32), bx])
        r1 = np.array([R.ref_getR1(float(a), float(c), float(R1), N) for a, c in zip(x, y)], dtype=np.int32)
        r2 = np.array([R.ref_getR2(float(a), float(c), int(r), float(R1), N, float(R2), n)
                       for a, c, r in zip(x, y, r1)], dtype=np.int32)
        cases.update({tag + "_x": x, tag + "_y": y, tag + "_r1": r1, tag + "_r2": r2,
                      tag + "_Nn": np.array([N, n], dtype=np.int32)})
        print("regions", tag, len(x), "outside:", int((r1 < 0).sum()))
    np.savez_compressed(os.path.join(OUT, "regions.npz"), **cases)

    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
           ((0, 0, 7, 0), (1234, 0), (0x56e604f4, 0x2107acfd, 0xe9ac28d3, 0x1debf147))]
    H = po.ref_host()
    import ctypes as C
    for ctr, key, out in kat:
        c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
        H.ref_philox4x32_10(c.ctypes.data_as(po.u32p), k.ctypes.data_as(po.u32p), o.ctypes.data_as(po.u32p))
        assert tuple(int(v) for v in o) == out, (ctr, key, o)
    np.savez_compressed(os.path.join(OUT, "philox.npz"), ctr=np.array([k[0] for k in kat], dtype=np.uint32),
                        key=np.array([k[1] for k in kat], dtype=np.uint32),
                        out=np.array([k[2] for k in kat], dtype=np.uint32))
    print("philox KATs ok")


if __name__ == "__main__":
    main()
