"""ctypes binding of libkgmt_b200.so (include/kgmt_c.h) — the KGMT tree-expansion path on B200.

This is the Python face of the C ABI; the reference-compatible C++ face is
cudasbmp_b200/include/planners/KGMT.cuh.  Names follow the reference planner
(`KGMT(width, height, N, n, numIterations, maxTreeSize, numDisc, agentLength, goalThreshold)`,
`plan(initial, goal, obstacles)`: /root/reference include/planners/KGMT.cuh:28,31).

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but any call
that computes raises KgmtError without the shared library or without a B200.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# KGMT_LIB selects another build of the same ABI (the bounds-checked twin libkgmt_b200_check.so: scripts/bounds_check.sh)
LIB_PATH = os.environ.get("KGMT_LIB") or os.path.join(HERE, "libkgmt_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_COMM = 0, -1, -2, -3, -4, -5
STOP = {0: "running", 1: "solved", 2: "tree_full", 3: "iter_limit", 4: "frontier_empty", 5: "peer_solved"}
COLLIDE_GRID, COLLIDE_BRUTE = 0, 1

(ARR_SAMPLES, ARR_UNEXPLORED, ARR_PARENT, ARR_U_PARENT, ARR_G, ARR_R2AVAIL, ARR_R1AVAIL, ARR_R1VALID, ARR_R2VALID,
 ARR_R1INVALID, ARR_R2INVALID, ARR_R1SCORE, ARR_R1, ARR_R2, ARR_COSTS, ARR_U_VALID, ARR_U_R1, ARR_U_R2, ARR_U_U3,
 ARR_U_ACCEPT) = range(20)

_DTYPE = {ARR_SAMPLES: (np.float32, 7), ARR_UNEXPLORED: (np.float32, 7), ARR_PARENT: (np.int32, 1),
          ARR_U_PARENT: (np.int32, 1), ARR_G: (np.uint8, 1), ARR_R2AVAIL: (np.int32, 1), ARR_R1AVAIL: (np.int32, 1),
          ARR_R1VALID: (np.int32, 1), ARR_R2VALID: (np.int32, 1), ARR_R1INVALID: (np.int32, 1),
          ARR_R2INVALID: (np.int32, 1), ARR_R1SCORE: (np.float32, 1), ARR_R1: (np.int32, 1), ARR_R2: (np.int32, 1),
          ARR_COSTS: (np.float32, 1), ARR_U_VALID: (np.uint8, 1), ARR_U_R1: (np.int32, 1), ARR_U_R2: (np.int32, 1),
          ARR_U_U3: (np.float32, 1), ARR_U_ACCEPT: (np.uint8, 1)}

# every symbol include/kgmt_c.h declares
ABI_SYMBOLS = [
    "kgmt_abi_version", "kgmt_default_params", "kgmt_create", "kgmt_destroy", "kgmt_last_error", "kgmt_reset", "kgmt_set_seed",
    "kgmt_set_obstacles", "kgmt_set_obstacles_host", "kgmt_plan", "kgmt_begin", "kgmt_expand_iteration", "kgmt_expand_iterations",
    "kgmt_get_result", "kgmt_extract_path", "kgmt_plan_batch", "kgmt_stage_scores", "kgmt_stage_propagate", "kgmt_seed_frontier",
    "kgmt_set_children", "kgmt_checkpoint", "kgmt_restore", "kgmt_export", "kgmt_import", "kgmt_array_bytes",
    "kgmt_dump_csv", "kgmt_tree_size", "kgmt_cost_to_goal", "kgmt_r1_size", "kgmt_r2_size", "kgmt_stream",
    "kgmt_launch_count", "kgmt_get_config", "kgmt_iteration_log",
    "kgmt_set_stream", "kgmt_shard_delta_ints", "kgmt_shard_expand", "kgmt_shard_pack", "kgmt_shard_commit",
    "kgmt_peer_handle_bytes", "kgmt_peer_export", "kgmt_peer_attach", "kgmt_peer_attach_local", "kgmt_peer_expand_begin",
    "kgmt_peer_expand_end", "kgmt_peer_detach", "kgmt_peer_race",
    "kgmt_params_from_yaml", "kgmt_stage_update_maps", "kgmt_stage_insert", "kgmt_work_counters", "kgmt_batch_cluster_size",
    "kgmt_peer_expand_iterations", "kgmt_peer_plan",
    "kgmt_comm_unique_id", "kgmt_comm_init", "kgmt_comm_destroy", "kgmt_comm_barrier", "kgmt_comm_rank", "kgmt_comm_world",
    "kgmt_plan_batch_sharded", "kgmt_plan_portfolio", "kgmt_expand_sharded", "kgmt_plan_sharded", "kgmt_debug_checks",
]
EXCHANGE_FUSED, EXCHANGE_PEER_LAUNCHES, EXCHANGE_NCCL = 0, 1, 2


class KgmtError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("width", C.c_float), ("height", C.c_float), ("N", C.c_int), ("n", C.c_int),
                ("num_iterations", C.c_int), ("max_tree_size", C.c_int), ("num_disc", C.c_int),
                ("agent_length", C.c_float), ("goal_threshold", C.c_float), ("seed", C.c_uint32),
                ("device", C.c_int), ("max_candidates", C.c_int), ("collision_mode", C.c_int),
                ("record_candidates", C.c_int), ("cull_cells", C.c_int), ("reserved", C.c_int * 5),
                ("accel_min", C.c_double), ("accel_max", C.c_double), ("steer_min", C.c_double),
                ("steer_max", C.c_double), ("duration_min", C.c_double), ("duration_max", C.c_double)]


class IterStats(C.Structure):
    _fields_ = [("iteration", C.c_int), ("mode", C.c_int), ("children", C.c_int), ("frontier", C.c_int),
                ("candidates", C.c_int), ("accepted", C.c_int), ("tree_size", C.c_int), ("stop", C.c_int),
                ("cost_to_goal", C.c_float), ("goal_index", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ShardInfo(C.Structure):
    _fields_ = [("iteration", C.c_int), ("candidates", C.c_int), ("children", C.c_int), ("frontier", C.c_int),
                ("chunk_lo", C.c_int), ("chunk_hi", C.c_int), ("accepted_local", C.c_int), ("stop", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [("stop", C.c_int), ("iterations", C.c_int), ("tree_size", C.c_int), ("cost_to_goal", C.c_float),
                ("goal_index", C.c_int), ("expansions", C.c_longlong), ("device_ms", C.c_float),
                ("kernel_launches", C.c_int), ("done_ms", C.c_float), ("service_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    """Load libkgmt_b200.so.  Raises KgmtError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KgmtError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(make -C cudasbmp_b200/csrc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, f32p = C.c_void_p, C.POINTER(C.c_float)
    L.kgmt_abi_version.restype = C.c_int
    L.kgmt_default_params.argtypes = [C.POINTER(Params)]
    L.kgmt_default_params.restype = None
    L.kgmt_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.kgmt_destroy.argtypes = [vp]
    L.kgmt_destroy.restype = None
    L.kgmt_last_error.argtypes = [vp]
    L.kgmt_last_error.restype = C.c_char_p
    L.kgmt_reset.argtypes = [vp]
    L.kgmt_set_seed.argtypes = [vp, C.c_uint32]
    L.kgmt_set_obstacles.argtypes = [vp, vp, C.c_int]
    L.kgmt_set_obstacles_host.argtypes = [vp, f32p, C.c_int]
    L.kgmt_plan.argtypes = [vp, f32p, f32p, C.POINTER(Result)]
    L.kgmt_begin.argtypes = [vp, f32p, f32p]
    L.kgmt_expand_iteration.argtypes = [vp, C.POINTER(IterStats)]
    L.kgmt_expand_iterations.argtypes = [vp, C.c_int, C.POINTER(IterStats)]
    L.kgmt_get_result.argtypes = [vp, C.POINTER(Result)]
    L.kgmt_extract_path.argtypes = [vp, C.c_int, f32p, C.c_int]
    L.kgmt_plan_batch.argtypes = [vp, f32p, f32p, C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(Result), f32p, C.c_int,
                                  C.POINTER(C.c_int), f32p]
    L.kgmt_stage_scores.argtypes = [vp]
    L.kgmt_stage_propagate.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, f32p]
    L.kgmt_seed_frontier.argtypes = [vp, f32p, C.c_int, f32p]
    L.kgmt_set_children.argtypes = [vp, C.c_int]
    L.kgmt_checkpoint.argtypes = [vp]
    L.kgmt_restore.argtypes = [vp]
    L.kgmt_export.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.kgmt_import.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.kgmt_array_bytes.argtypes = [vp, C.c_int]
    L.kgmt_array_bytes.restype = C.c_size_t
    L.kgmt_dump_csv.argtypes = [vp, C.c_char_p]
    L.kgmt_tree_size.argtypes = [vp]
    L.kgmt_cost_to_goal.argtypes = [vp]
    L.kgmt_cost_to_goal.restype = C.c_float
    L.kgmt_r1_size.argtypes = [vp]
    L.kgmt_r1_size.restype = C.c_float
    L.kgmt_r2_size.argtypes = [vp]
    L.kgmt_r2_size.restype = C.c_float
    L.kgmt_stream.argtypes = [vp]
    L.kgmt_stream.restype = vp
    L.kgmt_launch_count.argtypes = [vp]
    L.kgmt_launch_count.restype = C.c_longlong
    L.kgmt_get_config.argtypes = [vp, C.POINTER(C.c_int)]
    L.kgmt_iteration_log.argtypes = [vp, C.c_int, C.POINTER(C.c_ulonglong), C.c_int]
    L.kgmt_set_stream.argtypes = [vp, vp]
    L.kgmt_shard_delta_ints.argtypes = [vp]
    L.kgmt_shard_delta_ints.restype = C.c_size_t
    L.kgmt_shard_expand.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(ShardInfo)]
    L.kgmt_shard_pack.argtypes = [vp, vp, C.c_int]
    L.kgmt_shard_commit.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), C.c_int, vp, C.POINTER(IterStats)]
    L.kgmt_peer_handle_bytes.restype = C.c_size_t
    L.kgmt_peer_export.argtypes = [vp, vp, C.c_size_t]
    L.kgmt_peer_attach.argtypes = [vp, C.c_int, C.c_int, vp]
    L.kgmt_peer_attach_local.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.kgmt_peer_expand_begin.argtypes = [vp]
    L.kgmt_peer_expand_end.argtypes = [vp, C.POINTER(IterStats)]
    L.kgmt_peer_detach.argtypes = [vp]
    L.kgmt_peer_race.argtypes = [vp, f32p, f32p, C.c_int, C.POINTER(Result)]
    L.kgmt_params_from_yaml.argtypes = [C.c_char_p, C.POINTER(Params), C.POINTER(C.c_int)]
    L.kgmt_stage_update_maps.argtypes = [vp, f32p, C.POINTER(C.c_ubyte), f32p, C.POINTER(C.c_int), C.c_int]
    L.kgmt_stage_insert.argtypes = [vp, C.POINTER(IterStats)]
    L.kgmt_work_counters.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.kgmt_batch_cluster_size.argtypes = [vp, C.c_int]
    L.kgmt_peer_expand_iterations.argtypes = [vp, C.c_int, C.POINTER(IterStats)]
    L.kgmt_peer_plan.argtypes = [vp, f32p, f32p, C.POINTER(Result)]
    L.kgmt_comm_unique_id.argtypes = [vp]
    L.kgmt_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.kgmt_comm_destroy.argtypes = [vp]
    L.kgmt_comm_barrier.argtypes = [vp]
    L.kgmt_comm_rank.argtypes = [vp]
    L.kgmt_comm_world.argtypes = [vp]
    L.kgmt_plan_batch_sharded.argtypes = [vp, f32p, f32p, C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(Result), f32p]
    L.kgmt_plan_portfolio.argtypes = [vp, f32p, f32p, C.c_uint32, C.c_int, C.POINTER(Result), C.POINTER(C.c_int), f32p, C.c_int,
                                      C.POINTER(C.c_int)]
    L.kgmt_expand_sharded.argtypes = [vp, C.c_int, C.POINTER(IterStats), f32p]
    L.kgmt_plan_sharded.argtypes = [vp, f32p, f32p, C.POINTER(Result)]
    L.kgmt_debug_checks.argtypes = [vp, C.POINTER(C.c_int)]
    _lib = L
    return L


def default_params():
    p = Params()
    load().kgmt_default_params(C.byref(p))
    return p


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None and a.size != n:
        raise ValueError("expected %d floats, got %d" % (n, a.size))
    return a


class KGMT:
    """The reference's planner object over the B200 library.

    KGMT(width, height, N, n, numIterations, maxTreeSize, numDisc, agentLength, goalThreshold)
    mirrors the reference constructor; keyword arguments expose what the reference hard-codes
    (seed: it uses time(NULL), KGMT.cu:111) or does not have (collision back end, recording).
    """

    def __init__(self, width=20.0, height=20.0, N=16, n=8, numIterations=100, maxTreeSize=30000, numDisc=10,
                 agentLength=1.0, goalThreshold=0.5, *, seed=1, device=-1, max_candidates=0,
                 collision_mode=COLLIDE_GRID, record_candidates=False, cull_cells=0, stage_limit_bytes=0,
                 ctas_per_sm=0, car=None, car_yaml=None):
        L = load()
        p = default_params()
        p.width, p.height, p.N, p.n = width, height, N, n
        p.num_iterations, p.max_tree_size, p.num_disc = numIterations, maxTreeSize, numDisc
        p.agent_length, p.goal_threshold = agentLength, goalThreshold
        p.seed, p.device, p.max_candidates = seed & 0xFFFFFFFF, device, max_candidates
        p.collision_mode, p.record_candidates, p.cull_cells = collision_mode, int(bool(record_candidates)), cull_cells
        p.reserved[0] = stage_limit_bytes
        p.reserved[1] = ctas_per_sm          # 0 = as many as fit
        if car_yaml is not None:             # systems/car.yaml-style model file (flat key: value; empty = defaults)
            bad = C.c_int(0)
            if L.kgmt_params_from_yaml(os.fsencode(car_yaml), C.byref(p), C.byref(bad)) != OK:
                raise KgmtError("cannot read car model %s (line %d)" % (car_yaml, bad.value))
        for key, val in (car or {}).items():  # accel_min/max, steer_min/max, duration_min/max
            if key not in ("accel_min", "accel_max", "steer_min", "steer_max", "duration_min", "duration_max"):
                raise KgmtError("unknown car model key %r" % key)
            setattr(p, key, float(val))
        self.params = p
        self.N, self.n, self.max_tree = N, n, maxTreeSize
        self.max_cand = max_candidates if max_candidates > 0 else maxTreeSize
        self._h = C.c_void_p()
        rc = L.kgmt_create(C.byref(p), C.byref(self._h))
        if rc != OK:
            msg = L.kgmt_last_error(self._h).decode() if self._h else "invalid parameters"
            if self._h:
                L.kgmt_destroy(self._h)
                self._h = C.c_void_p()
            raise KgmtError("kgmt_create failed (%d): %s" % (rc, msg))
        self.R1Size_ = L.kgmt_r1_size(self._h)
        self.R2Size_ = L.kgmt_r2_size(self._h)
        self.treeSize_ = 0
        self.costToGoal_ = 0.0

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc < 0:
            raise KgmtError("libkgmt_b200 error %d: %s" % (rc, load().kgmt_last_error(self._h).decode()))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            load().kgmt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference surface
    def set_obstacles(self, aabb):
        """aabb: host float[K][4] (minx, miny, maxx, maxy) — configurations/obstacles/obstacles.csv rows."""
        a = _f32(aabb).reshape(-1, 4)
        self._ck(load().kgmt_set_obstacles_host(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[0]))

    def set_obstacles_device(self, dptr, K):
        """dptr: device pointer (int) to float[K][4], as KGMT::plan's d_obstacles."""
        self._ck(load().kgmt_set_obstacles(self._h, C.c_void_p(int(dptr)), int(K)))

    def plan(self, initial, goal, obstacles=None):
        """KGMT::plan (KGMT.cu:80-317): returns the result dict; treeSize_/costToGoal_ as the reference members."""
        if obstacles is not None:
            self.set_obstacles(obstacles)
        i, g = _f32(initial, 7), _f32(goal, 7)
        r = Result()
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_plan(self._h, i.ctypes.data_as(f32p), g.ctypes.data_as(f32p), C.byref(r)))
        self.treeSize_, self.costToGoal_ = r.tree_size, r.cost_to_goal
        return r.as_dict()

    def batch_cluster_size(self, Q):
        """The cluster size plan_batch(cluster_size=0) uses for Q queries (kgmt_batch_cluster_size)."""
        return self._ck(load().kgmt_batch_cluster_size(self._h, int(Q)))

    def plan_batch(self, inits, goals, seeds, cluster_size=0, max_path=0):
        """Q independent queries on this planner's map in one launch (kgmt_plan_batch); cluster_size 0 = chosen from Q.
        Returns (results: list of dict incl. done_ms / service_ms per query, device_ms, paths: list of np [L,7] or None,
        workspaces)."""
        a = _f32(inits).reshape(-1, 7)
        g = _f32(goals).reshape(-1, 7)
        sd = np.ascontiguousarray(seeds, dtype=np.uint32)
        Q = a.shape[0]
        res = (Result * Q)()
        ms = C.c_float()
        f32p = C.POINTER(C.c_float)
        paths = np.zeros((Q, max_path, 7), dtype=np.float32) if max_path else None
        plen = np.zeros(Q, dtype=np.int32) if max_path else None
        ws = self._ck(load().kgmt_plan_batch(
            self._h, a.ctypes.data_as(f32p), g.ctypes.data_as(f32p), sd.ctypes.data_as(C.POINTER(C.c_uint32)), Q,
            int(cluster_size), res, paths.ctypes.data_as(f32p) if max_path else None, int(max_path),
            plen.ctypes.data_as(C.POINTER(C.c_int)) if max_path else None, C.byref(ms)))
        out = [r.as_dict() for r in res]
        pl = [paths[q, :min(plen[q], max_path)].copy() for q in range(Q)] if max_path else None
        return out, ms.value, pl, ws

    # ------------------------------------------------------------------ stepwise
    def begin(self, initial, goal):
        i, g = _f32(initial, 7), _f32(goal, 7)
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_begin(self._h, i.ctypes.data_as(f32p), g.ctypes.data_as(f32p)))

    def iterate(self):
        s = IterStats()
        self._ck(load().kgmt_expand_iteration(self._h, C.byref(s)))
        self.treeSize_, self.costToGoal_ = s.tree_size, s.cost_to_goal
        return s.as_dict()

    def iterate_many(self, count):
        """Up to `count` iterations in one launch; returns the stats of the last one executed."""
        s = IterStats()
        self._ck(load().kgmt_expand_iterations(self._h, int(count), C.byref(s)))
        self.treeSize_, self.costToGoal_ = s.tree_size, s.cost_to_goal
        return s.as_dict()

    def result(self):
        r = Result()
        self._ck(load().kgmt_get_result(self._h, C.byref(r)))
        return r.as_dict()

    def reset(self):
        self._ck(load().kgmt_reset(self._h))

    def set_seed(self, seed):
        self._ck(load().kgmt_set_seed(self._h, int(seed) & 0xFFFFFFFF))

    def seed_frontier(self, nodes7, goal):
        a = _f32(nodes7).reshape(-1, 7)
        g = _f32(goal, 7)
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_seed_frontier(self._h, a.ctypes.data_as(f32p), a.shape[0], g.ctypes.data_as(f32p)))

    def set_children(self, children):
        self._ck(load().kgmt_set_children(self._h, int(children)))

    def checkpoint(self):
        self._ck(load().kgmt_checkpoint(self._h))

    def restore(self):
        self._ck(load().kgmt_restore(self._h))

    def stage_scores(self):
        self._ck(load().kgmt_stage_scores(self._h))

    def stage_propagate(self, parents7, children, key0, slot0=0):
        """Stages 2-4 on explicit parents.  Returns device milliseconds; results via export()."""
        a = _f32(parents7).reshape(-1, 7)
        ms = C.c_float()
        self._ck(load().kgmt_stage_propagate(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[0], int(children),
                                             int(key0) & 0xFFFFFFFF, int(slot0) & 0xFFFFFFFF, C.byref(ms)))
        return ms.value

    def stage_update_maps(self, cand7, valid, u3, parent):
        """Stage 5a alone on caller-supplied candidates (needs record_candidates=True and a begun context)."""
        c = _f32(cand7).reshape(-1, 7)
        M = len(c)
        v = np.ascontiguousarray(valid, dtype=np.uint8)
        u = _f32(u3, M)
        pa = np.ascontiguousarray(parent, dtype=np.int32)
        if v.size != M or pa.size != M:
            raise ValueError("valid / parent must have one entry per candidate")
        self._ck(load().kgmt_stage_update_maps(self._h, c.ctypes.data_as(C.POINTER(C.c_float)),
                                               v.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                               u.ctypes.data_as(C.POINTER(C.c_float)),
                                               pa.ctypes.data_as(C.POINTER(C.c_int)), M))

    def stage_insert(self):
        """Stage 5b alone: ordered insertion of what stage_update_maps accepted; returns the iteration stats."""
        st = IterStats()
        self._ck(load().kgmt_stage_insert(self._h, C.byref(st)))
        self.treeSize_, self.costToGoal_ = st.tree_size, st.cost_to_goal
        return st.as_dict()

    def work_counters(self):
        """{steps, pairs, expansions} executed by the recording kernels since the plan began."""
        out = (C.c_ulonglong * 4)()
        self._ck(load().kgmt_work_counters(self._h, out))
        return {"steps": int(out[0]), "pairs": int(out[1]), "expansions": int(out[2])}

    def debug_checks(self):
        """Bounds-checked build only: {site, failures, value, limit} of the device-side index checks since the last call."""
        out = (C.c_int * 4)()
        self._ck(load().kgmt_debug_checks(self._h, out))
        return {"site": out[0], "failures": out[1], "value": out[2], "limit": out[3]}

    def extract_path(self, node=-1, max_rows=4096):
        buf = np.zeros((max_rows, 7), dtype=np.float32)
        n = self._ck(load().kgmt_extract_path(self._h, int(node), buf.ctypes.data_as(C.POINTER(C.c_float)), max_rows))
        return buf[:min(n, max_rows)].copy()

    # ------------------------------------------------------------------ sharded expansion (config 5)
    def set_stream(self, cuda_stream):
        """Launch on the caller's CUDA stream (int handle, e.g. torch.cuda.current_stream().cuda_stream); 0/None = own."""
        self._ck(load().kgmt_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))

    def shard_delta_ints(self):
        return load().kgmt_shard_delta_ints(self._h)

    def shard_expand(self, rank, world, delta_ptr):
        """delta_ptr: device pointer (int) of the zeroed int32 slab of shard_delta_ints() elements."""
        info = ShardInfo()
        self._ck(load().kgmt_shard_expand(self._h, int(rank), int(world), C.c_void_p(int(delta_ptr)), C.byref(info)))
        return info.as_dict()

    def shard_pack(self, send_ptr, cap_rows):
        self._ck(load().kgmt_shard_pack(self._h, C.c_void_p(int(send_ptr)), int(cap_rows)))

    def shard_commit(self, recv_ptr, cap_rows, counts, delta_ptr):
        cnt = (C.c_int * len(counts))(*[int(c) for c in counts])
        s = IterStats()
        self._ck(load().kgmt_shard_commit(self._h, C.c_void_p(int(recv_ptr)), int(cap_rows), cnt, len(counts),
                                          C.c_void_p(int(delta_ptr)), C.byref(s)))
        self.treeSize_, self.costToGoal_ = s.tree_size, s.cost_to_goal
        return s.as_dict()

    # ------------------------------------------------------------------ sharded expansion over peer memory
    def peer_export(self):
        """cudaIpc handles of this planner's tree, maps and exchange block (bytes) for the other ranks."""
        n = load().kgmt_peer_handle_bytes()
        buf = (C.c_ubyte * n)()
        self._ck(load().kgmt_peer_export(self._h, buf, n))
        return bytes(buf)

    def peer_attach(self, rank, world, all_handles):
        """all_handles: the peer_export() bytes of every rank, concatenated in rank order."""
        raw = bytes(all_handles)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        self._ck(load().kgmt_peer_attach(self._h, int(rank), int(world), buf))

    def peer_attach_local(self, rank, planners):
        """Wire planners living in this process (tests)."""
        arr = (C.c_void_p * len(planners))(*[p._h for p in planners])
        self._ck(load().kgmt_peer_attach_local(self._h, int(rank), len(planners), arr))

    def peer_expand_begin(self):
        self._ck(load().kgmt_peer_expand_begin(self._h))

    def peer_expand_end(self):
        s = IterStats()
        self._ck(load().kgmt_peer_expand_end(self._h, C.byref(s)))
        self.treeSize_, self.costToGoal_ = s.tree_size, s.cost_to_goal
        return s.as_dict()

    def peer_iterate(self):
        """One iteration across the attached ranks (every rank calls it)."""
        self.peer_expand_begin()
        return self.peer_expand_end()

    def peer_race(self, initial, goal, race_id):
        """Portfolio race with the attached ranks: this rank's plan of the query (its own seed), stopped early with
        stop == 5 when another rank reaches the goal first."""
        i, g = _f32(initial, 7), _f32(goal, 7)
        r = Result()
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_peer_race(self._h, i.ctypes.data_as(f32p), g.ctypes.data_as(f32p), int(race_id), C.byref(r)))
        self.treeSize_, self.costToGoal_ = r.tree_size, r.cost_to_goal
        return r.as_dict()

    def peer_iterate_fused(self, count=1):
        """Up to `count` sharded iterations in ONE persistent launch (compute + exchange fused); every rank calls it."""
        st = IterStats()
        self._ck(load().kgmt_peer_expand_iterations(self._h, int(count), C.byref(st)))
        self.treeSize_, self.costToGoal_ = st.tree_size, st.cost_to_goal
        return st.as_dict()

    def peer_plan(self, initial, goal):
        """KGMT::plan with every iteration's candidates split over the attached ranks, one persistent launch per rank."""
        i, g = _f32(initial, 7), _f32(goal, 7)
        r = Result()
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_peer_plan(self._h, i.ctypes.data_as(f32p), g.ctypes.data_as(f32p), C.byref(r)))
        self.treeSize_, self.costToGoal_ = r.tree_size, r.cost_to_goal
        return r.as_dict()

    # ------------------------------------------------------------------ communicator (NCCL inside the library)
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (rank 0 creates it; hand it to every rank by any transport)."""
        buf = (C.c_ubyte * 128)()
        if load().kgmt_comm_unique_id(buf) != OK:
            raise KgmtError("kgmt_comm_unique_id failed (NCCL not loadable?)")
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._ck(load().kgmt_comm_init(self._h, int(rank), int(world), buf))

    def comm_destroy(self):
        self._ck(load().kgmt_comm_destroy(self._h))

    def comm_barrier(self):
        self._ck(load().kgmt_comm_barrier(self._h))

    def expand_sharded(self, exchange=EXCHANGE_FUSED, timing=False):
        """One sharded iteration across the communicator's ranks; with timing also (compute_ms, exchange_ms, bytes)."""
        st = IterStats()
        ms = (C.c_float * 3)()
        self._ck(load().kgmt_expand_sharded(self._h, int(exchange), C.byref(st), ms if timing else None))
        self.treeSize_, self.costToGoal_ = st.tree_size, st.cost_to_goal
        d = st.as_dict()
        if timing:
            d.update(compute_ms=ms[0], exchange_ms=ms[1], exchange_bytes=ms[2])
        return d

    def plan_sharded(self, initial, goal):
        i, g = _f32(initial, 7), _f32(goal, 7)
        r = Result()
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_plan_sharded(self._h, i.ctypes.data_as(f32p), g.ctypes.data_as(f32p), C.byref(r)))
        self.treeSize_, self.costToGoal_ = r.tree_size, r.cost_to_goal
        return r.as_dict()

    def plan_batch_sharded(self, inits, goals, seeds, cluster_size=0, as_array=False):
        """Config 4 across the communicator: returns (results of ALL Q queries, max device ms over the ranks).
        as_array: the results as one numpy structured array over the kgmt_result rows the library filled (no per-query
        Python objects: building 1 024 dicts costs more than planning 1 024 queries on eight GPUs)."""
        a = _f32(inits).reshape(-1, 7)
        g = _f32(goals).reshape(-1, 7)
        sd = np.ascontiguousarray(seeds, dtype=np.uint32)
        Q = a.shape[0]
        res = (Result * Q)()
        ms = C.c_float()
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_plan_batch_sharded(self._h, a.ctypes.data_as(f32p), g.ctypes.data_as(f32p),
                                                sd.ctypes.data_as(C.POINTER(C.c_uint32)), Q, int(cluster_size), res, C.byref(ms)))
        if as_array:
            return np.ctypeslib.as_array(res), ms.value
        return [r.as_dict() for r in res], ms.value

    def plan_portfolio(self, initial, goal, base_seed, race_id, max_rows=256):
        """Same query, seed base_seed + rank per rank, first solution wins; returns (winner rank or -1, the winner's
        result dict, its path [L, 7]) on every rank."""
        i, g = _f32(initial, 7), _f32(goal, 7)
        r, win, plen = Result(), C.c_int(-1), C.c_int(0)
        path = np.zeros((max_rows, 7), dtype=np.float32)
        f32p = C.POINTER(C.c_float)
        self._ck(load().kgmt_plan_portfolio(self._h, i.ctypes.data_as(f32p), g.ctypes.data_as(f32p), int(base_seed) & 0xFFFFFFFF,
                                            int(race_id), C.byref(r), C.byref(win), path.ctypes.data_as(f32p), max_rows, C.byref(plen)))
        return win.value, r.as_dict(), path[:min(plen.value, max_rows)].copy()

    def peer_detach(self):
        self._ck(load().kgmt_peer_detach(self._h))

    # ------------------------------------------------------------------ data
    def export(self, array_id):
        dt, cols = _DTYPE[array_id]
        nbytes = load().kgmt_array_bytes(self._h, array_id)
        out = np.zeros(nbytes // np.dtype(dt).itemsize, dtype=dt)
        self._ck(load().kgmt_export(self._h, array_id, out.ctypes.data_as(C.c_void_p), nbytes))
        return out.reshape(-1, cols) if cols > 1 else out

    def import_(self, array_id, values):
        dt, _ = _DTYPE[array_id]
        a = np.ascontiguousarray(values, dtype=dt)
        self._ck(load().kgmt_import(self._h, array_id, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def dump_csv(self, directory):
        os.makedirs(directory, exist_ok=True)
        self._ck(load().kgmt_dump_csv(self._h, directory.encode()))

    def config(self):
        out = (C.c_int * 8)()
        self._ck(load().kgmt_get_config(self._h, out))
        keys = ("collide_backend", "cull_cells", "cull_items", "smem_bytes", "grid", "sms", "r1_hist", "K")
        return dict(zip(keys, list(out)))

    def iteration_log(self, enable=True):
        """Rows of 8 u64 (see kgmt_iteration_log) of the last plan; call once with enable to switch logging on."""
        rows = int(os.environ.get("KGMT_ITERLOG_ROWS", "256"))
        buf = (C.c_ulonglong * (8 * rows))()
        n = self._ck(load().kgmt_iteration_log(self._h, int(enable), buf, rows))
        return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 8)[:n].copy()

    @property
    def launch_count(self):
        return load().kgmt_launch_count(self._h)

    @property
    def stream(self):
        return load().kgmt_stream(self._h)
