"""TEST INFRASTRUCTURE ONLY — ctypes bindings of the two CPU checkers:

* ``Oracle``  -> oracle/libgbp_oracle.so  (plain-C restatement, oracle/gbp_oracle.c)
* ``Ref``     -> oracle/_ref/libgbp_ref.so (the UNMODIFIED reference core, oracle/ref_harness.cpp)
* ``RefPin``  -> oracle/_ref/libgbp_ref_pin.so (the same objects with the samplers served from the Philox stream:
                 the reference's own planner loops, oracle/ref_pin_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (global_body_planner_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libgbp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgbp_ref.so")
REF_PIN_SO = os.path.join(HERE, "_ref", "libgbp_ref_pin.so")
REFERENCE_ROOT = "/root/reference"

FORWARD, REVERSE = 0, 1
FLIGHT, STANCE = 0, 1
TRAPPED, ADVANCED, REACHED = 0, 1, 2
FLAG_OOG = 2


def build(ref=True):
    """Compile the checkers (building the checker is not using it)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _u8(a, n):
    if np.isscalar(a):
        return np.full(n, a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


class Terrain:
    """Host arrays of one terrain: strictly increasing axes, x-major [nx, ny] fp64 layers."""

    def __init__(self, x, y, z, dx=None, dy=None, dz=None):
        self.x, self.y = _f64(x), _f64(y)
        self.nx, self.ny = len(self.x), len(self.y)
        self.z = _f64(z, (self.nx, self.ny))
        self.dx = _f64(np.zeros_like(self.z) if dx is None else dx, (self.nx, self.ny))
        self.dy = _f64(np.zeros_like(self.z) if dy is None else dy, (self.nx, self.ny))
        self.dz = _f64(np.ones_like(self.z) if dz is None else dz, (self.nx, self.ny))

    @staticmethod
    def from_reference_csv(directory):
        """data/<name>/*.csv: rows = y, cols = x (terrain_map_publisher.cpp:330-370) -> x-major."""
        ld = lambda n: np.loadtxt(os.path.join(directory, n + "data.csv"), delimiter=",")
        X, Y = ld("x"), ld("y")
        return Terrain(X[0, :].copy(), Y[:, 0].copy(), ld("z").T, ld("dx").T, ld("dy").T, ld("dz").T)

    @staticmethod
    def from_npz(path):
        d = np.load(path)
        return Terrain(d["x"], d["y"], d["z"], d["dx"], d["dy"], d["dz"])

    def save_npz(self, path):
        np.savez_compressed(path, x=self.x, y=self.y, z=self.z, dx=self.dx, dy=self.dy, dz=self.dz)

    @staticmethod
    def synthetic(n=4096, pitch=0.05, noise_sigma=0.0, seed=1, fp32=True):
        """z = 0.05 sin(0.7x) cos(0.5y) (+ seeded noise), normals (0,0,1).
        Heights are rounded to fp32 (as the ROS ingest path does, fast_terrain_map.cpp:60-66)."""
        ax = np.arange(n, dtype=np.float64) * pitch
        z = 0.05 * np.sin(0.7 * ax)[:, None] * np.cos(0.5 * ax)[None, :]
        if noise_sigma > 0:
            z = z + np.random.default_rng(seed).normal(0.0, noise_sigma, z.shape)
        if fp32:
            z = z.astype(np.float32).astype(np.float64)
        return Terrain(ax, ax.copy(), z)


class _OrcTerrain(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
                ("dx", C.c_void_p), ("dy", C.c_void_p), ("dz", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("substates", C.c_longlong), ("lookups", C.c_longlong), ("nanprobes", C.c_longlong), ("flags", C.c_uint)]


class PlanParams(C.Structure):
    _fields_ = [("k_candidates", C.c_int), ("best_of_k", C.c_int), ("max_iters", C.c_int), ("max_vertices", C.c_int),
                ("adaptive", C.c_int), ("rrt_star", C.c_int), ("post_process", C.c_int),
                ("state_direction_sampling", C.c_int), ("state_direction_speed", C.c_int), ("action_direction_sampling", C.c_int),
                ("cost_add_yaw", C.c_int), ("state_direction_threshold", C.c_double), ("action_direction_threshold", C.c_double),
                ("cost_length_weight", C.c_double), ("cost_yaw_weight", C.c_double)]


class TreeDump(C.Structure):
    _fields_ = [("cap", C.c_int), ("n", C.c_int), ("states", C.c_void_p), ("actions", C.c_void_p), ("parent", C.c_void_p),
                ("g", C.c_void_p), ("y", C.c_void_p)]


def _tree_buffers(cap):
    b = dict(states=np.zeros((cap, 8)), actions=np.zeros((cap, 10)), parent=np.zeros(cap, np.int32), g=np.zeros(cap), yaw=np.zeros(cap))
    d = TreeDump(cap, 0, b["states"].ctypes.data, b["actions"].ctypes.data, b["parent"].ctypes.data, b["g"].ctypes.data, b["yaw"].ctypes.data)
    return b, d


def _tree_result(b, d):
    return {k: v[:d.n].copy() for k, v in b.items()}


class PlanStats(C.Structure):
    _fields_ = [("solved", C.c_int), ("iters", C.c_int), ("nv_a", C.c_int), ("nv_b", C.c_int), ("path_states", C.c_int),
                ("pad", C.c_int), ("path_length", C.c_double), ("path_yaw", C.c_double), ("path_duration", C.c_double),
                ("pair_checks", C.c_longlong), ("nn_queries", C.c_longlong), ("path_cost", C.c_double), ("reserved", C.c_longlong)]


PLAN_STATS_DTYPE = np.dtype([("solved", "i4"), ("iters", "i4"), ("nv_a", "i4"), ("nv_b", "i4"), ("path_states", "i4"),
                             ("pad", "i4"), ("path_length", "f8"), ("path_yaw", "f8"), ("path_duration", "f8"),
                             ("pair_checks", "i8"), ("nn_queries", "i8"), ("path_cost", "f8"), ("reserved", "i8")])


def _orc():
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    return C.CDLL(ORACLE_SO)


OWN_MAP_DEFAULT = dict(x_size=221, y_size=161, x_start=-0.5, y_start=-4.0, res=0.05)  # terrain_map_publisher.cpp:36-38


def own_map(seed, x_size=221, y_size=161, x_start=-0.5, y_start=-4.0, res=0.05, rects=None):
    """createOwnMap restated (terrain_map_publisher.cpp:34-231): float elevation layer in grid_map index order and
    (resolution, centre x, centre y); rects [n][6] = x1 y1 x2 y2 mu delta, None = the reference's table."""
    L = _orc()
    elev = np.zeros((x_size, y_size), np.float32); geom = np.zeros(3)
    r = None if rects is None else _f64(rects, (-1, 6))
    L.orc_own_map(C.c_uint64(seed), x_size, y_size, C.c_double(x_start), C.c_double(y_start), C.c_double(res),
                  0 if r is None else len(r), _p(r), _p(elev), _p(geom))
    return elev, geom


def own_map_axes(n, start, res):
    L = _orc(); ax = np.zeros(n)
    L.orc_own_map_axes(n, C.c_double(start), C.c_double(res), _p(ax))
    return ax


def own_map_range(xa, ya, rect):
    L = _orc(); xa, ya, rect = _f64(xa), _f64(ya), _f64(rect); g = np.zeros(4, np.int32)
    L.orc_own_map_range(_p(xa), len(xa), _p(ya), len(ya), _p(rect), _p(g))
    return g


def default_map():
    """createMap restated (terrain_map_publisher.cpp:253-286)"""
    L = _orc(); elev = np.zeros((60, 25), np.float32); geom = np.zeros(3)
    L.orc_default_map(_p(elev), _p(geom))
    return elev, geom


class Oracle:
    """The plain-C restatement."""

    def __init__(self, terrain=None):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.L = C.CDLL(ORACLE_SO)
        self.L.orc_ground_height.restype = C.c_double
        self.L.orc_det_log.restype = C.c_double
        self.L.orc_det_log.argtypes = [C.c_double]
        self.L.orc_near.restype = C.c_longlong
        self.t = None
        if terrain is not None:
            self.set_terrain(terrain)

    def set_terrain(self, terrain):
        self.terrain = terrain
        self.t = _OrcTerrain(terrain.nx, terrain.ny, _p(terrain.x), _p(terrain.y), _p(terrain.z), _p(terrain.dx),
                             _p(terrain.dy), _p(terrain.dz))

    # -- stream spec
    def philox(self, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.zeros(4, np.uint32)
        self.L.orc_philox4x32_10(_p(c), _p(k), _p(o))
        return o

    def uniforms(self, seed, stream, idx, purpose, first, n):
        u = np.zeros(n)
        self.L.orc_uniforms(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx), purpose, first, n, _p(u))
        return u

    def det_log(self, x):
        return self.L.orc_det_log(float(x))

    def det_sincos(self, x):
        s, c = C.c_double(), C.c_double()
        self.L.orc_det_sincos(C.c_double(x), C.byref(s), C.byref(c))
        return s.value, c.value

    # -- terrain
    def ground_height(self, x, y):
        x, y = _f64(x), _f64(y)
        out = np.zeros(len(x)); fl = np.zeros(len(x), np.uint8)
        for i in range(len(x)):
            f = C.c_uint(0)
            out[i] = self.L.orc_ground_height(C.byref(self.t), C.c_double(x[i]), C.c_double(y[i]), C.byref(f))
            fl[i] = f.value
        return out, fl

    def height_is_nan(self, x, y):
        x, y = _f64(x), _f64(y)
        return np.array([self.L.orc_height_is_nan(C.byref(self.t), C.c_double(a), C.c_double(b), None) for a, b in zip(x, y)], np.uint8)

    def surface_normal(self, x, y):
        x, y = _f64(x), _f64(y)
        out = np.zeros((len(x), 3))
        for i in range(len(x)):
            self.L.orc_surface_normal(C.byref(self.t), C.c_double(x[i]), C.c_double(y[i]), _p(out[i]), None)
        return out

    # -- primitives
    def _apply(self, fn, s, a, t):
        s = _f64(s, (-1, 8)); t = _f64(np.broadcast_to(t, (len(s),)))
        out = np.zeros_like(s)
        if a is not None:
            a = _f64(a, (-1, 10))
        for i in range(len(s)):
            if a is None:
                fn(_p(s[i]), C.c_double(t[i]), _p(out[i]))
            else:
                fn(_p(s[i]), _p(a[i]), C.c_double(t[i]), _p(out[i]))
        return out

    def apply_stance(self, s, a, t):
        return self._apply(self.L.orc_apply_stance, s, a, t)

    def apply_flight(self, s, t):
        return self._apply(self.L.orc_apply_flight, s, None, t)

    def apply_stance_reverse(self, s, a, t):
        return self._apply(self.L.orc_apply_stance_reverse, s, a, t)

    def rotate_grf(self, n, f):
        n, f = _f64(n, (-1, 3)), _f64(f, (-1, 3)); out = np.zeros_like(f)
        for i in range(len(f)):
            self.L.orc_rotate_grf(_p(n[i]), _p(f[i]), _p(out[i]))
        return out

    def is_valid_action(self, a):
        a = _f64(a, (-1, 10))
        return np.array([self.L.orc_is_valid_action(_p(r)) for r in a], np.uint8)

    def valid_states(self, s, phase):
        s = _f64(s, (-1, 8)); n = len(s); ph = _u8(phase, n)
        v = np.zeros(n, np.uint8); fl = np.zeros(n, np.uint8)
        self.L.orc_valid_states(C.byref(self.t), C.c_longlong(n), _p(s), _p(ph), _p(v), _p(fl))
        return v, fl

    def validate_pairs(self, s, a, direction, adaptive=False, nthreads=1):
        """-> verdict u8[n], flags u8[n], s_new [n,8], t_new [n], counters (k, L, nan probes) totals."""
        s, a = _f64(s, (-1, 8)), _f64(a, (-1, 10)); n = len(s); d = _u8(direction, n)
        v = np.zeros(n, np.uint8); fl = np.zeros(n, np.uint8); sn = np.zeros((n, 8)); tn = np.zeros(n)
        cnt = np.zeros(3, np.int64)
        self.L.orc_validate_pairs(C.byref(self.t), C.c_longlong(n), _p(s), _p(a), _p(d), int(adaptive), _p(v), _p(fl),
                                  _p(sn), _p(tn), _p(cnt), int(nthreads))
        return v, fl, sn, tn, cnt

    def distance(self, q1, q2, kind):
        q1, q2 = _f64(q1, (-1, 8)), _f64(q2, (-1, 8))
        fn = [self.L.orc_pose_distance, self.L.orc_state_distance, self.L.orc_yaw_distance][kind]
        fn.restype = C.c_double
        return np.array([fn(_p(a), _p(b)) for a, b in zip(q1, q2)])

    # -- samplers
    def sample_actions(self, seed, stream, idx0, n, normal=(0.0, 0.0, 1.0)):
        a = np.zeros((n, 10)); nrm = _f64(normal)
        self.L.orc_sample_actions(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.c_longlong(n), _p(nrm), _p(a))
        return a

    def sample_action_dir(self, seed, stream, idx, normal, thresh, s_from, s_to):
        a = np.zeros(10); nrm = _f64(normal); sf, st = _f64(s_from), _f64(s_to)
        self.L.orc_sample_action(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx), _p(nrm), 1, C.c_double(thresh),
                                 _p(sf), _p(st), _p(a))
        return a

    def sample_states(self, seed, stream, idx0, n):
        q = np.zeros((n, 8))
        self.L.orc_sample_states(C.byref(self.t), C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.c_longlong(n), _p(q))
        return q

    def sample_state_dir(self, seed, stream, idx, thresh, speed_dir, s_from, s_to):
        q = np.zeros(8); sf, st = _f64(s_from), _f64(s_to)
        self.L.orc_sample_state(C.byref(self.t), C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx), 1,
                                C.c_double(thresh), int(speed_dir), _p(sf), _p(st), _p(q))
        return q

    # -- tree queries
    def nearest(self, verts, q):
        verts, q = _f64(verts, (-1, 8)), _f64(q, (-1, 8))
        idx = np.zeros(len(q), np.int32); dist = np.zeros(len(q)); uniq = np.zeros(len(q), np.int32)
        for j in range(len(q)):
            d, u = C.c_double(), C.c_int()
            idx[j] = self.L.orc_nearest(_p(verts), C.c_longlong(len(verts)), _p(q[j]), C.byref(d), C.byref(u))
            dist[j], uniq[j] = d.value, u.value
        return idx, dist, uniq

    def near(self, verts, q, radius):
        verts, q = _f64(verts, (-1, 8)), _f64(q)
        ids = np.zeros(len(verts), np.int32)
        n = self.L.orc_near(_p(verts), C.c_longlong(len(verts)), _p(q), C.c_double(radius), _p(ids), C.c_longlong(len(verts)))
        return ids[:n].copy()

    def attempt_connect(self, s_existing, s, direction, adaptive=False):
        se, s = _f64(s_existing, (-1, 8)), _f64(s, (-1, 8)); n = len(s); d = _u8(direction, n)
        st = np.zeros(n, np.int32); sn = np.zeros((n, 8)); an = np.zeros((n, 10)); fl = np.zeros(n, np.uint8)
        for i in range(n):
            c = Counters()
            st[i] = self.L.orc_attempt_connect(C.byref(self.t), _p(se[i]), _p(s[i]), int(d[i]), int(adaptive), _p(sn[i]), _p(an[i]), C.byref(c))
            fl[i] = c.flags
        return st, sn, an, fl

    # -- Tier-2 planner
    def plan(self, start, goal, seed, query, params, path_cap=4096):
        st = PlanStats(); ps = np.zeros((path_cap, 8)); pa = np.zeros((path_cap, 10))
        s, g = _f64(start), _f64(goal)
        self.L.orc_plan(C.byref(self.t), _p(s), _p(g), C.c_uint64(seed), C.c_uint64(query), C.byref(params), C.byref(st),
                        _p(ps), _p(pa), path_cap)
        n = st.path_states
        return st, ps[:n].copy(), pa[:max(n - 1, 0)].copy()

    def plan_ex(self, start, goal, seed, query, params, path_cap=4096):
        """orc_plan with both trees copied out -> stats, path states, path actions, tree A, tree B"""
        st = PlanStats(); ps = np.zeros((path_cap, 8)); pa = np.zeros((path_cap, 10))
        s, g = _f64(start), _f64(goal)
        ba, da = _tree_buffers(params.max_vertices); bb, db = _tree_buffers(params.max_vertices)
        self.L.orc_plan_ex(C.byref(self.t), _p(s), _p(g), C.c_uint64(seed), C.c_uint64(query), C.byref(params), C.byref(st),
                           _p(ps), _p(pa), path_cap, C.byref(da), C.byref(db))
        n = st.path_states
        return st, ps[:n].copy(), pa[:max(n - 1, 0)].copy(), _tree_result(ba, da), _tree_result(bb, db)

    def plan_batch(self, starts, goals, seed, query0, params, nthreads=1):
        s, g = _f64(starts, (-1, 8)), _f64(goals, (-1, 8)); nq = len(s)
        st = np.zeros(nq, PLAN_STATS_DTYPE)
        self.L.orc_plan_batch(C.byref(self.t), C.c_longlong(nq), _p(s), _p(g), C.c_uint64(seed), C.c_uint64(query0),
                              C.byref(params), _p(st), int(nthreads))
        return st

    def interp_path(self, states, actions, dt=0.05, cap=1 << 16):
        s, a = _f64(states, (-1, 8)), _f64(actions, (-1, 10))
        os_, ot, op = np.zeros((cap, 8)), np.zeros(cap), np.zeros(cap, np.int32)
        self.L.orc_interp_path.restype = C.c_longlong
        m = self.L.orc_interp_path(len(a), _p(s), _p(a), C.c_double(dt), C.c_longlong(cap), _p(os_), _p(ot), _p(op))
        return os_[:m].copy(), ot[:m].copy(), op[:m - 1].copy()

    def max_curvature(self, states):
        s = _f64(states, (-1, 8))
        self.L.orc_max_curvature.restype = C.c_double
        return self.L.orc_max_curvature(C.c_longlong(len(s)), _p(s))

    def post_process_path(self, states, actions, adaptive=False):
        s, a = _f64(states, (-1, 8)).copy(), _f64(actions, (-1, 10)).copy()
        a = np.concatenate([a, np.zeros((1, 10))])
        st3 = np.zeros(3)
        m = self.L.orc_post_process_path(C.byref(self.t), len(s), _p(s), _p(a), int(adaptive), _p(st3))
        return s[:m].copy(), a[:m - 1].copy(), st3


class PinParams(C.Structure):
    _fields_ = [("star", C.c_int), ("max_iters", C.c_int), ("adaptive", C.c_int), ("sort_near", C.c_int),
                ("state_direction_sampling", C.c_int), ("state_direction_speed", C.c_int), ("action_direction_sampling", C.c_int),
                ("cost_add_yaw", C.c_int), ("state_direction_threshold", C.c_double), ("action_direction_threshold", C.c_double),
                ("cost_length_weight", C.c_double), ("cost_yaw_weight", C.c_double)]


class PinResult(C.Structure):
    _fields_ = [("solved", C.c_int), ("cells_used", C.c_int), ("nv_a", C.c_int), ("nv_b", C.c_int), ("path_states", C.c_int),
                ("budget_hit", C.c_int), ("path_length", C.c_double), ("path_yaw", C.c_double), ("path_cost", C.c_double),
                ("oog_height", C.c_longlong), ("oog_nan", C.c_longlong), ("near_sets", C.c_longlong), ("near_reordered", C.c_longlong)]


class RefPin:
    """The unmodified reference's planner loops (runRRTConnect, extend, newConfig, connect, RRT* extend) on the shared Philox
    stream: oracle/ref_pin_harness.cpp.  What the Tier-2 oracle planner (Oracle.plan_ex) is pinned against."""

    @staticmethod
    def available():
        return os.path.exists(REF_PIN_SO)

    def __init__(self, terrain):
        self.L = C.CDLL(REF_PIN_SO)
        self.L.pin_terrain_create.restype = C.c_void_p
        self.terrain = t = terrain
        self.h = C.c_void_p(self.L.pin_terrain_create(t.nx, t.ny, _p(t.x), _p(t.y), _p(t.z), _p(t.dx), _p(t.dy), _p(t.dz)))
        self.L.pin_terrain_set(t.nx, t.ny, _p(t.x), _p(t.y), _p(t.z), _p(t.dx), _p(t.dy), _p(t.dz))

    def close(self):
        if getattr(self, "h", None):
            self.L.pin_terrain_clear()
            self.L.pin_terrain_destroy(self.h)
            self.h = None

    __del__ = close

    def run(self, start, goal, seed, query, params, cap=4096, path_cap=4096):
        """-> PinResult, tree A, tree B, path states, path actions (the stitched path before postProcessPath)"""
        t = self.terrain
        self.L.pin_terrain_set(t.nx, t.ny, _p(t.x), _p(t.y), _p(t.z), _p(t.dx), _p(t.dy), _p(t.dz))  # one terrain is current at a time
        s, g = _f64(start), _f64(goal)
        ba, _ = _tree_buffers(cap); bb, _ = _tree_buffers(cap)
        ps = np.zeros((path_cap, 8)); pa = np.zeros((path_cap, 10))
        out = PinResult()
        self.L.pin_run(self.h, _p(s), _p(g), C.c_uint64(seed), C.c_uint64(query), C.byref(params), C.byref(out), cap,
                       _p(ba["states"]), _p(ba["actions"]), _p(ba["parent"]), _p(ba["g"]), _p(ba["yaw"]),
                       _p(bb["states"]), _p(bb["actions"]), _p(bb["parent"]), _p(bb["g"]), _p(bb["yaw"]), path_cap, _p(ps), _p(pa))
        if out.nv_a > cap or out.nv_b > cap:
            raise RuntimeError("tree capacity of the dump exceeded")
        n = out.path_states
        return (out, {k: v[:out.nv_a].copy() for k, v in ba.items()}, {k: v[:out.nv_b].copy() for k, v in bb.items()},
                ps[:n].copy(), pa[:max(n - 1, 0)].copy())


class Ref:
    """The unmodified reference behind oracle/ref_harness.cpp."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, terrain=None):
        self.L = C.CDLL(REF_SO)
        self.L.ref_terrain_create.restype = C.c_void_p
        self.L.ref_terrain_create_gridmap.restype = C.c_void_p
        self.L.ref_near.restype = C.c_longlong
        self.h = None
        if terrain is not None:
            self.set_terrain(terrain)

    def set_terrain(self, t):
        self.close()
        self.terrain = t
        self.h = C.c_void_p(self.L.ref_terrain_create(t.nx, t.ny, _p(t.x), _p(t.y), _p(t.z), _p(t.dx), _p(t.dy), _p(t.dz)))

    def set_terrain_gridmap(self, nx, ny, res, cx, cy, elev, dxl=None, dyl=None, dzl=None):
        self.close()
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        e, a, b, c = f32(elev), f32(dxl), f32(dyl), f32(dzl)
        self.h = C.c_void_p(self.L.ref_terrain_create_gridmap(nx, ny, C.c_double(res), C.c_double(cx), C.c_double(cy),
                                                              _p(e), _p(a), _p(b), _p(c)))

    def axes(self):
        r = self.L.ref_terrain_axes(self.h, None, None)
        x, y = np.zeros(r & 0xffff), np.zeros(r >> 16)
        self.L.ref_terrain_axes(self.h, _p(x), _p(y))
        return x, y

    def close(self):
        if self.h is not None:
            self.L.ref_terrain_destroy(self.h)
            self.h = None

    def ground_height(self, x, y):
        x, y = _f64(x), _f64(y); out = np.zeros(len(x))
        self.L.ref_ground_height(self.h, C.c_longlong(len(x)), _p(x), _p(y), _p(out))
        return out

    def height_is_nan(self, x, y):
        x, y = _f64(x), _f64(y); out = np.zeros(len(x), np.uint8)
        self.L.ref_height_is_nan(self.h, C.c_longlong(len(x)), _p(x), _p(y), _p(out))
        return out

    def surface_normal(self, x, y):
        x, y = _f64(x), _f64(y); out = np.zeros((len(x), 3))
        self.L.ref_surface_normal(self.h, C.c_longlong(len(x)), _p(x), _p(y), _p(out))
        return out

    def apply_stance(self, s, a, t):
        s, a = _f64(s, (-1, 8)), _f64(a, (-1, 10)); t = _f64(np.broadcast_to(t, (len(s),))); out = np.zeros_like(s)
        self.L.ref_apply_stance(C.c_longlong(len(s)), _p(s), _p(a), _p(t), _p(out))
        return out

    def apply_flight(self, s, t):
        s = _f64(s, (-1, 8)); t = _f64(np.broadcast_to(t, (len(s),))); out = np.zeros_like(s)
        self.L.ref_apply_flight(C.c_longlong(len(s)), _p(s), _p(t), _p(out))
        return out

    def apply_stance_reverse(self, s, a, t):
        s, a = _f64(s, (-1, 8)), _f64(a, (-1, 10)); t = _f64(np.broadcast_to(t, (len(s),))); out = np.zeros_like(s)
        self.L.ref_apply_stance_reverse(C.c_longlong(len(s)), _p(s), _p(a), _p(t), _p(out))
        return out

    def rotate_grf(self, n, f):
        n, f = _f64(n, (-1, 3)), _f64(f, (-1, 3)); out = np.zeros_like(f)
        self.L.ref_rotate_grf(C.c_longlong(len(f)), _p(n), _p(f), _p(out))
        return out

    def is_valid_action(self, a):
        a = _f64(a, (-1, 10)); out = np.zeros(len(a), np.uint8)
        self.L.ref_is_valid_action(C.c_longlong(len(a)), _p(a), _p(out))
        return out

    def valid_states(self, s, phase):
        s = _f64(s, (-1, 8)); ph = _u8(phase, len(s)); out = np.zeros(len(s), np.uint8)
        self.L.ref_is_valid_state(self.h, C.c_longlong(len(s)), _p(s), _p(ph), _p(out))
        return out

    def validate_pairs(self, s, a, direction, adaptive=False, nthreads=1):
        s, a = _f64(s, (-1, 8)), _f64(a, (-1, 10)); n = len(s); d = _u8(direction, n)
        v = np.zeros(n, np.uint8); sn = np.zeros((n, 8)); tn = np.zeros(n)
        self.L.ref_validate_pairs(self.h, C.c_longlong(n), _p(s), _p(a), _p(d), int(adaptive), _p(v), _p(sn), _p(tn), int(nthreads))
        return v, sn, tn

    def distance(self, q1, q2, kind):
        q1, q2 = _f64(q1, (-1, 8)), _f64(q2, (-1, 8)); out = np.zeros(len(q1))
        self.L.ref_distance(C.c_longlong(len(q1)), _p(q1), _p(q2), kind, _p(out))
        return out

    def nearest(self, verts, q):
        verts, q = _f64(verts, (-1, 8)), _f64(q, (-1, 8))
        idx = np.zeros(len(q), np.int32); dist = np.zeros(len(q))
        self.L.ref_nearest(C.c_longlong(len(verts)), _p(verts), C.c_longlong(len(q)), _p(q), _p(idx), _p(dist))
        return idx, dist

    def near(self, verts, q, radius):
        verts, q = _f64(verts, (-1, 8)), _f64(q); ids = np.zeros(len(verts), np.int32)
        n = self.L.ref_near(C.c_longlong(len(verts)), _p(verts), _p(q), C.c_double(radius), _p(ids), C.c_longlong(len(verts)))
        return ids[:n].copy()

    def tree_gy(self, verts, parent):
        verts = _f64(verts, (-1, 8)); parent = np.ascontiguousarray(parent, dtype=np.int32)
        g, y = np.zeros(len(verts)), np.zeros(len(verts))
        self.L.ref_tree_gy(C.c_longlong(len(verts)), _p(verts), _p(parent), _p(g), _p(y))
        return g, y

    def attempt_connect(self, s_existing, s, direction, adaptive=False):
        se, s = _f64(s_existing, (-1, 8)), _f64(s, (-1, 8)); n = len(s); d = _u8(direction, n)
        st = np.zeros(n, np.int32); sn = np.zeros((n, 8)); an = np.zeros((n, 10))
        self.L.ref_attempt_connect(self.h, C.c_longlong(n), _p(se), _p(s), _p(d), int(adaptive), _p(st), _p(sn), _p(an))
        return st, sn, an

    def interp_path(self, states, actions, dt=0.05, cap=1 << 16):
        s, a = _f64(states, (-1, 8)), _f64(actions, (-1, 10))
        os_, ot, op = np.zeros((cap, 8)), np.zeros(cap), np.zeros(cap, np.int32)
        self.L.ref_interp_path.restype = C.c_longlong
        m = self.L.ref_interp_path(len(a), _p(s), _p(a), C.c_double(dt), C.c_longlong(cap), _p(os_), _p(ot), _p(op))
        return os_[:m].copy(), ot[:m].copy(), op[:m - 1].copy()

    def max_curvature(self, states):
        s = _f64(states, (-1, 8))
        self.L.ref_max_curvature.restype = C.c_double
        return self.L.ref_max_curvature(C.c_longlong(len(s)), _p(s))

    def post_process_path(self, states, actions):
        s = _f64(states, (-1, 8)).copy(); a = np.concatenate([_f64(actions, (-1, 10)), np.zeros((1, 10))])
        st3 = np.zeros(3)
        m = self.L.ref_post_process_path(self.h, len(s), _p(s), _p(a), len(s), _p(st3))
        return s[:m].copy(), a[:m - 1].copy(), st3

    def plan(self, algorithm, start, goal, max_time, adaptive=False, cap=8192):
        out = np.zeros(7); pc, nn = C.c_longlong(), C.c_longlong()
        ss, aa = np.zeros((cap, 8)), np.zeros((cap, 10)); s, g = _f64(start), _f64(goal)
        n = self.L.ref_plan(self.h, algorithm, _p(s), _p(g), C.c_double(max_time), int(adaptive), _p(out), C.byref(pc),
                            C.byref(nn), _p(ss), _p(aa), cap)
        return dict(plan_time=out[0], success=int(out[1]), vertices=int(out[2]), time_to_first=out[3], cost=out[4],
                    path_duration=out[5], n_states=n, pair_checks=pc.value, nn_queries=nn.value), ss[:n].copy(), aa[:max(n - 1, 0)].copy()
