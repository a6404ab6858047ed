"""ctypes bindings for the CPU oracle and the reference builds — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (cudasbmp_b200) never does.

  liboracle  = oracle/libkgmt_oracle.so        our C restatement (oracle/kgmt_oracle.c)
  ref_host   = oracle/_ref/libref_host.so      reference propagate+collision, unmodified, host build
  ref_gpu    = oracle/_ref/libref_gpu.so       reference CUDA kernels, unmodified, sm_100a + Philox
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libkgmt_oracle.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libref_host.so")
REF_GPU_SO = os.path.join(HERE, "_ref", "libref_gpu.so")
REF_MAIN = os.path.join(HERE, "_ref", "ref_main")

MATH_HOST, MATH_FMA = 0, 1
STATUS = {0: "running", 1: "solved", 2: "tree_full", 3: "iter_limit", 4: "frontier_empty"}

(ARR_TREE_SAMPLES, ARR_UNEXPLORED, ARR_TREE_PARENT, ARR_U_PARENT, ARR_G, ARR_R2AVAIL, ARR_R1AVAIL,
 ARR_R1VALID, ARR_R2VALID, ARR_R1INVALID, ARR_R2INVALID, ARR_R1SCORE, ARR_R1, ARR_R2, ARR_COSTS,
 ARR_U_VALID, ARR_U_R1, ARR_U_R2, ARR_U_U3, ARR_U_MARGIN) = range(20)

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build(ref=True):
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_uniform.restype = C.c_float
        L.orc_uniform.argtypes = [C.c_uint32]
        L.orc_getR1.restype = C.c_int
        L.orc_getR1.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_getR2.restype = C.c_int
        L.orc_getR2.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_float, C.c_uint32, C.c_int]
        for name in ("orc_destroy", "orc_reset"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_set_obstacles.restype = None
        L.orc_set_obstacles.argtypes = [C.c_void_p, f32p, C.c_int]
        L.orc_begin.restype = None
        L.orc_begin.argtypes = [C.c_void_p, f32p, f32p]
        L.orc_iterate.restype = C.c_int
        L.orc_iterate.argtypes = [C.c_void_p]
        L.orc_plan.restype = C.c_int
        L.orc_plan.argtypes = [C.c_void_p, f32p, f32p]
        for name in ("orc_tree_size", "orc_iterations", "orc_goal_index", "orc_status", "orc_frontier_start",
                     "orc_frontier_count", "orc_last_M", "orc_last_accepted", "orc_last_mode", "orc_last_children"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_cost_to_goal.restype = C.c_float
        L.orc_cost_to_goal.argtypes = [C.c_void_p]
        L.orc_R1Threshold.restype = C.c_float
        L.orc_R1Threshold.argtypes = [C.c_void_p]
        L.orc_expansions.restype = C.c_longlong
        L.orc_expansions.argtypes = [C.c_void_p]
        L.orc_array.restype = C.c_void_p
        L.orc_array.argtypes = [C.c_void_p, C.c_int]
        L.orc_propagate_slot.restype = C.c_int
        L.orc_propagate_slot.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_int, C.c_float, f32p, C.c_int,
                                         C.c_float, C.c_float, C.c_int, f32p, f32p, f32p, i32p]
        L.orc_propagate_batch.restype = None
        L.orc_propagate_batch.argtypes = [f32p, i32p, C.c_long, f32p, u8p, f32p, f32p, C.c_int, C.c_float,
                                          C.c_uint32, C.c_uint32, f32p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_getR2_fma.restype = C.c_int
        L.orc_getR2_fma.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]
        L.orc_regions_batch.restype = None
        L.orc_regions_batch.argtypes = [f32p, C.c_int, C.c_long, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int,
                                        i32p, i32p]
        L.orc_scores.restype = None
        L.orc_scores.argtypes = [i32p, i32p, i32p, i32p, i32p, C.c_int, C.c_int, C.c_float, f32p, f32p]
        L.orc_update_maps.restype = None
        L.orc_update_maps.argtypes = [C.c_int, i32p, i32p, u8p, f32p, f32p, i32p] + [i32p] * 8 + [u8p]
        L.orc_insert.restype = C.c_int
        L.orc_insert.argtypes = [C.c_int, u8p, f32p, i32p, C.c_int, f32p, i32p, f32p, u8p, f32p, C.c_float,
                                 f32p, i32p]
        L.orc_expansion_shape.restype = None
        L.orc_expansion_shape.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_slot_uniforms.restype = None
        L.orc_slot_uniforms.argtypes = [C.c_uint32, C.c_uint32, f32p]
        L.orc_set_car_ranges.restype = None
        L.orc_set_car_ranges.argtypes = [C.POINTER(C.c_double)]
        L.orc_controls.restype = None
        L.orc_controls.argtypes = [f32p, C.c_int, f32p, f32p, f32p]
        L.orc_controls_general.restype = None
        L.orc_controls_general.argtypes = [f32p, C.POINTER(C.c_double), C.c_int, f32p, f32p, f32p]
        L.orc_in_goal.restype = C.c_int
        L.orc_in_goal.argtypes = [f32p, f32p, C.c_float]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- small helpers
def set_car_ranges(r6=None):
    """Control ranges (accel_min, accel_max, steer_min, steer_max, duration_min, duration_max) for every later
    orc_* call of this process; None restores the reference's literals (statePropagator.cu:17-19)."""
    if r6 is None:
        lib().orc_set_car_ranges(None)
    else:
        lib().orc_set_car_ranges((C.c_double * 6)(*[float(v) for v in r6]))


def controls(u3, math_mode=MATH_FMA, ranges=None):
    """(a, steering, duration) from three uniforms: the literal path, or the general one when ranges is given."""
    u = np.ascontiguousarray(u3, dtype=np.float32)
    a, s, d = C.c_float(), C.c_float(), C.c_float()
    if ranges is None:
        lib().orc_controls(_p(u, f32p), math_mode, C.byref(a), C.byref(s), C.byref(d))
    else:
        lib().orc_controls_general(_p(u, f32p), (C.c_double * 6)(*[float(v) for v in ranges]), math_mode,
                                   C.byref(a), C.byref(s), C.byref(d))
    return a.value, s.value, d.value


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, u32p), _p(k, u32p), _p(o, u32p))
    return o


def slot_uniforms(key0, slot):
    u = np.zeros(4, dtype=np.float32)
    lib().orc_slot_uniforms(int(key0) & 0xFFFFFFFF, int(slot), _p(u, f32p))
    return u


def getR1(x, y, R1Size, N):
    return lib().orc_getR1(float(x), float(y), float(R1Size), int(N))


def getR2(x, y, r1, R1Size, N, R2Size, n):
    return lib().orc_getR2(float(x), float(y), int(r1), float(R1Size), int(N), float(R2Size), int(n))


def getR2_fma(x, y, r1, R1Size, N, R2Size, n):
    return lib().orc_getR2_fma(float(x), float(y), int(r1), float(R1Size), int(N), float(R2Size), int(n))


def regions_batch(rows, R1Size, N, R2Size, n, math_mode=MATH_FMA):
    """rows: float32 [M, stride>=2] with x, y in the first two columns.  Returns (r1[M], r2[M])."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    M, stride = rows.shape
    r1 = np.zeros(M, dtype=np.int32)
    r2 = np.zeros(M, dtype=np.int32)
    lib().orc_regions_batch(_p(rows, f32p), stride, M, float(R1Size), int(N), float(R2Size), int(n), math_mode,
                            _p(r1, i32p), _p(r2, i32p))
    return r1, r2


def insert(accept, cand7, cand_parent, tree_size, tree7, tree_parent, costs, G, goal7, r, cost_to_goal=0.0, goal_idx=-1):
    """orc_insert on numpy arrays (tree arrays updated in place).  Returns (accepted, costToGoal, goalIdx)."""
    accept = np.ascontiguousarray(accept, dtype=np.uint8)
    cand7 = np.ascontiguousarray(cand7, dtype=np.float32)
    cand_parent = np.ascontiguousarray(cand_parent, dtype=np.int32)
    goal7 = np.ascontiguousarray(goal7, dtype=np.float32)
    ctg = C.c_float(cost_to_goal)
    gi = C.c_int(goal_idx)
    k = lib().orc_insert(len(accept), _p(accept, u8p), _p(cand7, f32p), _p(cand_parent, i32p), int(tree_size),
                         _p(tree7, f32p), _p(tree_parent, i32p), _p(costs, f32p), _p(G, u8p), _p(goal7, f32p), float(r),
                         C.byref(ctg), C.byref(gi))
    return k, ctg.value, gi.value


def expansion_shape(active, tree_size, max_tree):
    m, c, M = C.c_int(), C.c_int(), C.c_int()
    lib().orc_expansion_shape(active, tree_size, max_tree, C.byref(m), C.byref(c), C.byref(M))
    return m.value, c.value, M.value


def propagate_batch(parents7, parent_of, key0, slot0, num_disc, L, obstacles, W, H, math_mode=MATH_FMA,
                    want_margin=True):
    """Returns (x1[M,7] f32, valid[M] u8, u3[M] f32, margin[M] f32|None)."""
    parents7 = np.ascontiguousarray(parents7, dtype=np.float32)
    parent_of = np.ascontiguousarray(parent_of, dtype=np.int32)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 4)
    M = parent_of.shape[0]
    x1 = np.zeros((M, 7), dtype=np.float32)
    valid = np.zeros(M, dtype=np.uint8)
    u3 = np.zeros(M, dtype=np.float32)
    margin = np.zeros(M, dtype=np.float32) if want_margin else None
    lib().orc_propagate_batch(_p(parents7, f32p), _p(parent_of, i32p), M, _p(x1, f32p), _p(valid, u8p),
                              _p(u3, f32p), _p(margin, f32p), num_disc, L, int(key0) & 0xFFFFFFFF, int(slot0),
                              _p(obstacles, f32p), obstacles.shape[0], W, H, math_mode)
    return x1, valid, u3, margin


def scores(R1Avail, R2Avail, R1Valid, R1Invalid, R1, N, n, eps=0.01):
    a = [np.ascontiguousarray(v, dtype=np.int32) for v in (R1Avail, R2Avail, R1Valid, R1Invalid, R1)]
    out = np.zeros(N * N, dtype=np.float32)
    thr = C.c_float()
    lib().orc_scores(*[_p(v, i32p) for v in a], N, n, eps, _p(out, f32p), C.byref(thr))
    return out, thr.value


def update_maps(r1, r2, valid, u3, R1Score, R2AvailSnap, maps):
    """maps: dict of int32 arrays R1,R2,R1Valid,R2Valid,R1Invalid,R2Invalid,R1Avail,R2Avail (updated in place).
    Returns accept[M] u8."""
    M = len(r1)
    r1 = np.ascontiguousarray(r1, dtype=np.int32)
    r2 = np.ascontiguousarray(r2, dtype=np.int32)
    valid = np.ascontiguousarray(valid, dtype=np.uint8)
    u3 = np.ascontiguousarray(u3, dtype=np.float32)
    acc = np.zeros(M, dtype=np.uint8)
    lib().orc_update_maps(M, _p(r1, i32p), _p(r2, i32p), _p(valid, u8p), _p(u3, f32p),
                          _p(np.ascontiguousarray(R1Score, dtype=np.float32), f32p),
                          _p(np.ascontiguousarray(R2AvailSnap, dtype=np.int32), i32p),
                          *[_p(maps[k], i32p) for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid",
                                                        "R1Avail", "R2Avail")],
                          _p(acc, u8p))
    return acc


class Planner:
    """The reference planner's observable state (same array names) driven by the C restatement."""

    def __init__(self, width, height, N, n, num_iterations, max_tree, num_disc, agent_length, goal_threshold,
                 seed=0, math_mode=MATH_FMA):
        self.N, self.n, self.max_tree = N, n, max_tree
        self.h = lib().orc_create(width, height, N, n, num_iterations, max_tree, num_disc, agent_length,
                                  goal_threshold, seed & 0xFFFFFFFF, math_mode)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_obstacles(self, aabb):
        a = np.ascontiguousarray(aabb, dtype=np.float32).reshape(-1, 4)
        lib().orc_set_obstacles(self.h, _p(a, f32p), a.shape[0])

    def reset(self):
        lib().orc_reset(self.h)

    def begin(self, init7, goal7):
        i = np.ascontiguousarray(init7, dtype=np.float32)
        g = np.ascontiguousarray(goal7, dtype=np.float32)
        lib().orc_begin(self.h, _p(i, f32p), _p(g, f32p))

    def iterate(self):
        return lib().orc_iterate(self.h)

    def plan(self, init7, goal7):
        i = np.ascontiguousarray(init7, dtype=np.float32)
        g = np.ascontiguousarray(goal7, dtype=np.float32)
        return lib().orc_plan(self.h, _p(i, f32p), _p(g, f32p))

    tree_size = property(lambda s: lib().orc_tree_size(s.h))
    iterations = property(lambda s: lib().orc_iterations(s.h))
    cost_to_goal = property(lambda s: lib().orc_cost_to_goal(s.h))
    goal_index = property(lambda s: lib().orc_goal_index(s.h))
    status = property(lambda s: lib().orc_status(s.h))
    expansions = property(lambda s: lib().orc_expansions(s.h))
    frontier_start = property(lambda s: lib().orc_frontier_start(s.h))
    frontier_count = property(lambda s: lib().orc_frontier_count(s.h))
    last_M = property(lambda s: lib().orc_last_M(s.h))
    last_accepted = property(lambda s: lib().orc_last_accepted(s.h))
    last_mode = property(lambda s: lib().orc_last_mode(s.h))
    last_children = property(lambda s: lib().orc_last_children(s.h))

    def array(self, aid):
        """A numpy VIEW onto the oracle's array (valid until close())."""
        T, c1 = self.max_tree, self.N * self.N
        c2 = c1 * self.n * self.n
        shapes = {
            ARR_TREE_SAMPLES: ((T, 7), np.float32), ARR_UNEXPLORED: ((T, 7), np.float32),
            ARR_TREE_PARENT: ((T,), np.int32), ARR_U_PARENT: ((T,), np.int32), ARR_G: ((T,), np.uint8),
            ARR_R2AVAIL: ((c2,), np.int32), ARR_R1AVAIL: ((c1,), np.int32), ARR_R1VALID: ((c1,), np.int32),
            ARR_R2VALID: ((c2,), np.int32), ARR_R1INVALID: ((c1,), np.int32), ARR_R2INVALID: ((c2,), np.int32),
            ARR_R1SCORE: ((c1,), np.float32), ARR_R1: ((c1,), np.int32), ARR_R2: ((c2,), np.int32),
            ARR_COSTS: ((T,), np.float32), ARR_U_VALID: ((T,), np.uint8), ARR_U_R1: ((T,), np.int32),
            ARR_U_R2: ((T,), np.int32), ARR_U_U3: ((T,), np.float32), ARR_U_MARGIN: ((T,), np.float32),
        }
        shape, dt = shapes[aid]
        ptr = lib().orc_array(self.h, aid)
        count = int(np.prod(shape))
        buf = (C.c_byte * (count * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shape)


# ----------------------------------------------------------------------------- reference builds
_ref_host = None
_ref_gpu = None


def ref_host():
    """The reference's own propagate+collision, host build.  None when not built."""
    global _ref_host
    if _ref_host is None and os.path.exists(REF_HOST_SO):
        L = C.CDLL(REF_HOST_SO)
        L.ref_host_propagate.restype = C.c_int
        L.ref_host_propagate.argtypes = [f32p, f32p, C.c_int, C.c_float, C.c_uint32, C.c_uint32, f32p, C.c_int,
                                         C.c_float, C.c_float, f32p]
        L.ref_host_propagate_batch.restype = C.c_double
        L.ref_host_propagate_batch.argtypes = [f32p, i32p, C.c_long, f32p, u8p, f32p, C.c_int, C.c_float,
                                               C.c_uint32, C.c_uint32, f32p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.ref_philox4x32_10.restype = None
        L.ref_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.ref_host_hw_threads.restype = C.c_int
        _ref_host = L
    return _ref_host


def ref_host_batch(parents7, parent_of, key0, slot0, num_disc, L, obstacles, W, H, threads=1, outputs=True):
    """Returns (seconds, x1, valid, u3) from the reference's unmodified host-built code."""
    R = ref_host()
    parents7 = np.ascontiguousarray(parents7, dtype=np.float32)
    parent_of = np.ascontiguousarray(parent_of, dtype=np.int32)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 4)
    M = parent_of.shape[0]
    x1 = np.zeros((M, 7), dtype=np.float32) if outputs else None
    valid = np.zeros(M, dtype=np.uint8) if outputs else None
    u3 = np.zeros(M, dtype=np.float32) if outputs else None
    sec = R.ref_host_propagate_batch(_p(parents7, f32p), _p(parent_of, i32p), M, _p(x1, f32p), _p(valid, u8p),
                                     _p(u3, f32p), num_disc, L, int(key0) & 0xFFFFFFFF, int(slot0),
                                     _p(obstacles, f32p), obstacles.shape[0], W, H, threads)
    return sec, x1, valid, u3


def ref_gpu():
    """The reference's own CUDA kernels (sm_100a, Philox).  None when not built."""
    global _ref_gpu
    if _ref_gpu is None and os.path.exists(REF_GPU_SO):
        L = C.CDLL(REF_GPU_SO)
        L.ref_getR1.restype = C.c_int
        L.ref_getR1.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
        L.ref_getR2.restype = C.c_int
        L.ref_getR2.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]
        L.ref_gpu_expand.restype = C.c_int
        L.ref_gpu_expand.argtypes = ([C.c_int, C.c_int, f32p, C.c_int, i32p, C.c_int] + [i32p] * 8 +
                                     [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, f32p,
                                      C.c_int, C.c_float, C.c_float, C.c_int, f32p, i32p, u8p, C.c_int, f32p])
        L.ref_gpu_scores.restype = C.c_int
        L.ref_gpu_scores.argtypes = [i32p] * 5 + [C.c_int, C.c_float, f32p, f32p]
        L.ref_gpu_insert.restype = C.c_int
        L.ref_gpu_insert.argtypes = [C.c_int, u8p, f32p, i32p, C.c_int, f32p, i32p, f32p, u8p, f32p, C.c_float,
                                     f32p]
        L.ref_gpu_plan.restype = C.c_int
        L.ref_gpu_plan.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_float, f32p, f32p, f32p, C.c_int, i32p, f32p, C.POINTER(C.c_double)]
        L.ref_gpu_last_error.restype = C.c_char_p
        _ref_gpu = L
    return _ref_gpu


def ref_gpu_expand(mode, children, tree7, frontier, maps, R1Score, N, n, R1Size, R2Size, num_disc, L, obstacles,
                   W, H, seed_key, reps=1):
    """Launch the reference's propagateG / propagateGV2.  maps updated in place.
    Returns (unexplored[M,7], uParentIdx[M], GNew[M], ms)."""
    R = ref_gpu()
    tree7 = np.ascontiguousarray(tree7, dtype=np.float32).reshape(-1, 7)
    frontier = np.ascontiguousarray(frontier, dtype=np.int32)
    obstacles = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 4)
    M = frontier.shape[0] * children
    unx = np.zeros((M, 7), dtype=np.float32)
    upar = np.zeros(M, dtype=np.int32)
    gnew = np.zeros(M, dtype=np.uint8)
    ms = C.c_float()
    rc = R.ref_gpu_expand(mode, children, _p(tree7, f32p), tree7.shape[0], _p(frontier, i32p), frontier.shape[0],
                          *[_p(maps[k], i32p) for k in ("R1", "R2", "R1Valid", "R2Valid", "R1Invalid", "R2Invalid",
                                                        "R1Avail", "R2Avail")],
                          _p(np.ascontiguousarray(R1Score, dtype=np.float32), f32p), N, n, R1Size, R2Size, num_disc,
                          L, _p(obstacles, f32p), obstacles.shape[0], W, H, int(seed_key), _p(unx, f32p),
                          _p(upar, i32p), _p(gnew, u8p), reps, C.byref(ms))
    if rc != 0:
        raise RuntimeError("ref_gpu_expand failed rc=%d: %s" % (rc, R.ref_gpu_last_error().decode()))
    return unx, upar, gnew, ms.value


def ref_gpu_plan(cfg, init7, goal7, obstacles):
    """The reference's own KGMT::plan end to end (unmodified sources, sm_100a, XORWOW, time(NULL) seed):
    dict(tree_size, cost_to_goal, plan_wall_ms = wall clock around plan() incl. its CSV dumps, inside_ms = the
    reference's own printed 'time inside KGMT' (std::clock around its loop, KGMT.cu:82,294-295)).  cfg: the nine
    constructor arguments under cudasbmp_b200.workloads' names."""
    R = ref_gpu()
    if R is None:
        raise RuntimeError("oracle/_ref/libref_gpu.so not built")
    i7 = np.ascontiguousarray(init7, dtype=np.float32)
    g7 = np.ascontiguousarray(goal7, dtype=np.float32)
    ob = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 4)
    ts, cost, ms = C.c_int(), C.c_float(), (C.c_double * 2)()
    rc = R.ref_gpu_plan(cfg["width"], cfg["height"], cfg["N"], cfg["n"], cfg["numIterations"], cfg["maxTreeSize"],
                        cfg["numDisc"], cfg["agentLength"], cfg["goalThreshold"], _p(i7, f32p), _p(g7, f32p), _p(ob, f32p),
                        ob.shape[0], C.byref(ts), C.byref(cost), ms)
    if rc != 0:
        raise RuntimeError("ref_gpu_plan failed rc=%d" % rc)
    return dict(tree_size=ts.value, cost_to_goal=cost.value, plan_wall_ms=ms[0], inside_ms=ms[1])
