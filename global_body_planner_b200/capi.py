"""ctypes binding of libgbp_b200.so (include/gbp_b200.h).  Marshalling only — every number comes from
the CUDA kernels.  Mirrors the reference's operator names where one exists (getGroundHeight,
isValidStateActionPair, getNearestNeighbor ...) through the C entry point that replaces it."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("GBP_B200_SO", os.path.join(HERE, "libgbp_b200.so"))

FORWARD, REVERSE = 0, 1
FLIGHT, STANCE = 0, 1
TRAPPED, ADVANCED, REACHED = 0, 1, 2
FLAG_VALID, FLAG_OOG, FLAG_NEAR = 1, 2, 4


class GbpError(RuntimeError):
    pass


class PlanParams(C.Structure):
    _fields_ = [("k_candidates", C.c_int), ("best_of_k", C.c_int), ("max_iters", C.c_int), ("max_vertices", C.c_int),
                ("adaptive", C.c_int), ("rrt_star", C.c_int), ("post_process", C.c_int), ("stop_after_solved", C.c_int),
                ("state_direction_sampling", C.c_int), ("state_direction_speed", C.c_int), ("action_direction_sampling", C.c_int),
                ("cost_add_yaw", C.c_int), ("state_direction_threshold", C.c_double), ("action_direction_threshold", C.c_double),
                ("cost_length_weight", C.c_double), ("cost_yaw_weight", C.c_double)]


class SvParams(C.Structure):
    """gbp_sv_params (include/gbp_b200.h): the per-call part of a sample + validate batch."""
    _fields_ = [("seed", C.c_uint64), ("stream", C.c_uint64), ("idx0", C.c_uint64), ("normal", C.c_double * 3),
                ("adaptive", C.c_int), ("direction0", C.c_int), ("action_direction_sampling", C.c_int),
                ("action_direction_threshold", C.c_double), ("target", C.c_double * 8), ("row0", C.c_int64),
                ("start_states_valid", C.c_int), ("direction_in_row", C.c_int)]


class SvResult(C.Structure):
    _fields_ = [("n_valid", C.c_int64), ("substates", C.c_int64), ("lookups", C.c_int64), ("nanprobes", C.c_int64),
                ("oog", C.c_int64), ("near", C.c_int64), ("reserved", C.c_int64 * 2)]


def sv_params(seed, stream, idx0, normal=(0.0, 0.0, 1.0), adaptive=False, direction0=0, target=None, thresh=0.0, row0=0, states_valid=False, direction_in_row=False):
    p = SvParams()
    p.direction_in_row = int(direction_in_row)
    p.start_states_valid = int(states_valid)
    p.seed, p.stream, p.idx0 = seed, stream, idx0
    p.normal[:] = [float(v) for v in normal]
    p.adaptive, p.direction0, p.row0 = int(adaptive), int(direction0), int(row0)
    p.action_direction_sampling = 0 if target is None else 1
    p.action_direction_threshold = float(thresh)
    p.target[:] = [0.0] * 8 if target is None else [float(v) for v in target]
    return p


def pack_rows(state_idx, direction):
    """The one-word wire format of gbp_sample_validate (gbp_sv_params.direction_in_row): row number | direction << 31."""
    r = np.ascontiguousarray(state_idx, dtype=np.int64)
    assert r.min(initial=0) >= 0 and r.max(initial=0) < 2 ** 31
    return (r | (np.asarray(direction, dtype=np.int64) << 31)).astype(np.uint32).view(np.int32)


PLAN_STATS_DTYPE = np.dtype([("solved", "i4"), ("iters", "i4"), ("nv_a", "i4"), ("nv_b", "i4"), ("path_states", "i4"),
                             ("pad", "i4"), ("path_length", "f8"), ("path_yaw", "f8"), ("path_duration", "f8"),
                             ("pair_checks", "i8"), ("nn_queries", "i8"), ("path_cost", "f8"), ("reserved", "i8")])

_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise GbpError(f"{SO} is missing: run `python -m global_body_planner_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO)
        L.gbp_last_error.restype = C.c_char_p
        L.gbp_version.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise GbpError(f"gbp error {rc}: {lib().gbp_last_error().decode()}")


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _u8(a, n):
    if np.isscalar(a):
        return np.full(n, a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def version():
    return lib().gbp_version().decode()


def device_count():
    n = C.c_int()
    lib().gbp_device_count(C.byref(n))
    return n.value


def set_device(i):
    _check(lib().gbp_set_device(int(i)))


def propagate(kind, states, actions, t):
    """applyStance (kind 0) / applyFlight (1) / applyStanceReverse (2), planning_utils.cpp:237-370."""
    s = _f64(states, (-1, 8)); n = len(s); t = _f64(np.broadcast_to(t, (n,))); out = np.zeros_like(s)
    a = None if actions is None else _f64(actions, (-1, 10))
    _check(lib().gbp_propagate(kind, C.c_int64(n), _p(s), _p(a), _p(t), _p(out)))
    return out


def rotate_grf(normals, forces):
    """rotate_grf, planning_utils.cpp:198-231 (n triples)."""
    nn, f = _f64(normals, (-1, 3)), _f64(forces, (-1, 3)); out = np.zeros_like(f)
    _check(lib().gbp_rotate_grf(C.c_int64(len(f)), _p(nn), _p(f), _p(out)))
    return out


def curvature(points6):
    """calculateCurvature, planning_utils.cpp:884-899 (n sextuples x1 y1 x2 y2 x3 y3)."""
    p = _f64(points6, (-1, 6)); out = np.zeros(len(p))
    _check(lib().gbp_curvature(C.c_int64(len(p)), _p(p), _p(out)))
    return out


def valid_actions(actions):
    """isValidAction, planning_utils.cpp:519-556."""
    a = _f64(actions, (-1, 10)); out = np.zeros(len(a), np.uint8)
    _check(lib().gbp_valid_actions(C.c_int64(len(a)), _p(a), _p(out)))
    return out


def distance(kind, q1, q2):
    """poseDistance (0) / stateDistance (1) / stateYawDistance (2)."""
    q1, q2 = _f64(q1, (-1, 8)), _f64(q2, (-1, 8)); out = np.zeros(len(q1))
    _check(lib().gbp_distance(kind, C.c_int64(len(q1)), _p(q1), _p(q2), _p(out)))
    return out


def interp_path(states, actions, dt=0.05, cap=None):
    """getInterpPath (planning_utils.cpp:175-192) -> interpolated states [m, 8], times [m], phases [m - 1]."""
    s, a = _f64(states, (-1, 8)), _f64(actions, (-1, 10))
    if len(s) != len(a) + 1:
        raise GbpError("interp_path needs len(states) == len(actions) + 1")
    m = C.c_int64()
    if cap is None:  # size query: cap = 0 writes nothing and returns the count
        _check(lib().gbp_interp_path(len(a), _p(s), _p(a), C.c_double(dt), C.c_int64(0), None, None, None, C.byref(m)))
        cap = m.value
    os_, ot, op = np.zeros((cap, 8)), np.zeros(cap), np.zeros(cap, np.int32)
    _check(lib().gbp_interp_path(len(a), _p(s), _p(a), C.c_double(dt), C.c_int64(cap), _p(os_), _p(ot), _p(op), C.byref(m)))
    n = min(m.value, cap)
    return os_[:n].copy(), ot[:n].copy(), op[:max(n - 1, 0)].copy()


def max_curvature(states):
    """calculateMaxCurvature (planning_utils.cpp:900-909)."""
    s = _f64(states, (-1, 8)); out = C.c_double()
    _check(lib().gbp_max_curvature(C.c_int64(len(s)), _p(s), C.byref(out)))
    return out.value


def sample_actions(seed, stream, idx0, n, normal=(0.0, 0.0, 1.0), s_from=None, s_to=None, thresh=0.0):
    """getRandomAction / getRandomActionDirection on the Philox stream."""
    a = np.zeros((n, 10)); nrm = _f64(normal)
    sf = None if s_from is None else _f64(s_from); st = None if s_to is None else _f64(s_to)
    _check(lib().gbp_sample_actions(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.c_int64(n), _p(nrm), _p(sf),
                                    _p(st), C.c_double(thresh), _p(a)))
    return a


def own_map_layer(seed, x_size=221, y_size=161, x_start=-0.5, y_start=-4.0, res=0.05, rects=None):
    """The float elevation layer createOwnMap fills (grid_map index order) and (resolution, centre x, centre y)."""
    r = None if rects is None else _f64(rects, (-1, 6))
    elev = np.zeros((x_size, y_size), np.float32); geom = np.zeros(3)
    _check(lib().gbp_own_map_layer(C.c_uint64(seed), x_size, y_size, C.c_double(x_start), C.c_double(y_start), C.c_double(res),
                                   0 if r is None else len(r), _p(r), _p(elev), _p(geom)))
    return elev, geom


class States:
    """Device-resident table of start states (gbp_states): the vertices sample + validate candidates start from."""

    def __init__(self, states):
        s = _f64(states, (-1, 8))
        h = C.c_void_p()
        _check(lib().gbp_states_create(C.c_int64(len(s)), _p(s), C.byref(h)))
        self.h, self.rows = h, len(s)

    def close(self):
        if getattr(self, "h", None):
            lib().gbp_states_destroy(self.h)
            self.h = None

    __del__ = close


def unpack_bits(bits, n):
    """verdict words -> one 0/1 byte per candidate"""
    return np.unpackbits(np.ascontiguousarray(bits, dtype="<u4").view(np.uint8), bitorder="little")[:n]


class Terrain:
    """Device-resident FastTerrainMap (fast_terrain_map.h).  x-major [nx, ny] layers."""

    def __init__(self, x, y, z, dx=None, dy=None, dz=None):
        x, y = _f64(x), _f64(y)
        z = _f64(z, (len(x), len(y)))
        lay = [None if a is None else _f64(a, z.shape) for a in (dx, dy, dz)]
        h = C.c_void_p()
        _check(lib().gbp_terrain_create(len(x), len(y), _p(x), _p(y), _p(z), _p(lay[0]), _p(lay[1]), _p(lay[2]), C.byref(h)))
        self.h = h
        self.nx, self.ny = len(x), len(y)

    @classmethod
    def from_gridmap(cls, nx, ny, res, cx, cy, elevation, dx=None, dy=None, dz=None):
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        e, a, b, c = f32(elevation), f32(dx), f32(dy), f32(dz)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _check(lib().gbp_terrain_create_gridmap(nx, ny, C.c_double(res), C.c_double(cx), C.c_double(cy), _p(e), _p(a), _p(b),
                                                _p(c), C.byref(h)))
        self.h, self.nx, self.ny = h, nx, ny
        return self

    @classmethod
    def from_csv(cls, directory, via_gridmap=False):
        """<directory>/{x,y,z,dx,dy,dz}data.csv in the reference's format (terrain_map_publisher.cpp:289-370)."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        _check(lib().gbp_terrain_create_csv(os.fsencode(directory), int(bool(via_gridmap)), C.byref(h)))
        self.h = h
        a, b = C.c_int(), C.c_int()
        _check(lib().gbp_terrain_dims(h, C.byref(a), C.byref(b), None))
        self.nx, self.ny = a.value, b.value
        return self

    @classmethod
    def own_map(cls, seed, x_size=221, y_size=161, x_start=-0.5, y_start=-4.0, res=0.05, rects=None):
        """TerrainMapPublisher::createOwnMap (terrain_map_publisher.cpp:34-231) on the Philox stream; rects [n][6] =
        x1 y1 x2 y2 mu delta, None = the reference's table."""
        r = None if rects is None else _f64(rects, (-1, 6))
        self = cls.__new__(cls)
        h = C.c_void_p()
        _check(lib().gbp_terrain_create_own_map(C.c_uint64(seed), x_size, y_size, C.c_double(x_start), C.c_double(y_start),
                                                C.c_double(res), 0 if r is None else len(r), _p(r), C.byref(h)))
        self.h, self.nx, self.ny = h, x_size, y_size
        return self

    @classmethod
    def default_map(cls):
        """TerrainMapPublisher::createMap (terrain_map_publisher.cpp:253-286)."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        _check(lib().gbp_terrain_create_default_map(C.byref(h)))
        self.h, self.nx, self.ny = h, 60, 25
        return self

    def close(self):
        if getattr(self, "h", None):
            lib().gbp_terrain_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def cell_bytes(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _check(lib().gbp_terrain_dims(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return c.value

    def flags(self):
        u, m = C.c_int(), C.c_int()
        _check(lib().gbp_terrain_flags(self.h, C.byref(u), C.byref(m)))
        g = C.c_int()
        _check(lib().gbp_terrain_fetch_path(self.h, C.byref(g)))
        return dict(uniform_axes=bool(u.value), mixed_precision=bool(m.value), texture_gather=bool(g.value))

    def axes(self):
        x, y = np.zeros(self.nx), np.zeros(self.ny)
        _check(lib().gbp_terrain_axes(self.h, _p(x), _p(y)))
        return x, y

    def ground_height(self, x, y):
        x, y = _f64(x), _f64(y); h = np.zeros(len(x)); fl = np.zeros(len(x), np.uint8)
        _check(lib().gbp_ground_height(self.h, C.c_int64(len(x)), _p(x), _p(y), _p(h), _p(fl)))
        return h, fl

    def height_is_nan(self, x, y):
        x, y = _f64(x), _f64(y); out = np.zeros(len(x), np.uint8)
        _check(lib().gbp_height_is_nan(self.h, C.c_int64(len(x)), _p(x), _p(y), _p(out)))
        return out

    def surface_normal(self, x, y):
        x, y = _f64(x), _f64(y); out = np.zeros((len(x), 3))
        _check(lib().gbp_surface_normal(self.h, C.c_int64(len(x)), _p(x), _p(y), _p(out)))
        return out

    def valid_states(self, states, phase):
        s = _f64(states, (-1, 8)); n = len(s); ph = _u8(phase, n); v = np.zeros(n, np.uint8); fl = np.zeros(n, np.uint8)
        _check(lib().gbp_valid_states(self.h, C.c_int64(n), _p(s), _p(ph), _p(v), _p(fl)))
        return v, fl

    def valid_states_dev(self, n, states_ptr, phase_ptr, verdict_ptr, flags_ptr=0, stream=0):
        vp = lambda p: C.c_void_p(p) if p else None
        _check(lib().gbp_valid_states_dev(self.h, C.c_int64(n), vp(states_ptr), vp(phase_ptr), vp(verdict_ptr), vp(flags_ptr), vp(stream)))

    def validate_pairs(self, states, actions, direction, adaptive=False, variant=0):
        """isValidStateActionPair[Reverse] on HOST arrays -> verdict, flags, s_new, t_new."""
        s, a = _f64(states, (-1, 8)), _f64(actions, (-1, 10)); n = len(s); d = _u8(direction, n)
        v = np.zeros(n, np.uint8); fl = np.zeros(n, np.uint8); sn = np.zeros((n, 8)); tn = np.zeros(n)
        _check(lib().gbp_validate_pairs(self.h, C.c_int64(n), _p(s), _p(a), _p(d), int(adaptive), int(variant), _p(v), _p(fl),
                                        _p(sn), _p(tn)))
        return v, fl, sn, tn

    def validate_pairs_dev(self, n, states_ptr, actions_ptr, dir_ptr, adaptive, variant, verdict_ptr, flags_ptr, snew_ptr,
                           tnew_ptr, stream=0):
        """Device pointers (ints) in, enqueue only."""
        vp = lambda p: C.c_void_p(p) if p else None
        _check(lib().gbp_validate_pairs_dev(self.h, C.c_int64(n), vp(states_ptr), vp(actions_ptr), vp(dir_ptr), int(adaptive),
                                            int(variant), vp(verdict_ptr), vp(flags_ptr), vp(snew_ptr), vp(tnew_ptr), vp(stream)))

    def sample_validate(self, table, n, params, state_idx=None, direction=None, valid_cap=None, want_flags=True, rows=True):
        """newConfig's unit of work (rrt.cpp:34-50) on HOST arrays through the narrow wire format: -> dict with the
        verdict words, the unpacked verdicts, optional flags, the compact rows of the valid candidates and the result."""
        idx = None if state_idx is None else np.ascontiguousarray(state_idx, dtype=np.int32)
        d = None if direction is None else _u8(direction, n)
        cap = n if valid_cap is None else int(valid_cap)
        bits = np.zeros((n + 31) // 32, np.uint32)
        fl = np.zeros(n, np.uint8) if want_flags else None
        vi = np.zeros(cap, np.int32)
        sn = np.zeros((cap, 8)) if rows else None
        tn = np.zeros(cap) if rows else None
        ac = np.zeros((cap, 10)) if rows else None
        res = SvResult()
        _check(lib().gbp_sample_validate(self.h, table.h, C.c_int64(n), _p(idx), _p(d), C.byref(params), _p(bits), _p(fl), C.c_int64(cap),
                                         _p(vi), _p(sn), _p(tn), _p(ac), C.byref(res)))
        m = min(res.n_valid, cap)
        out = dict(bits=bits, verdict=unpack_bits(bits, n), flags=fl, index=vi[:m], n_valid=res.n_valid,
                   counters=dict(substates=res.substates, lookups=res.lookups, nanprobes=res.nanprobes, oog=res.oog, near=res.near))
        if rows:
            out.update(s_new=sn[:m], t_new=tn[:m], action=ac[:m])
        return out

    def sample_validate_dev(self, states_ptr, table_rows, n, params, idx_ptr, dir_ptr, bits_ptr, flags_ptr, valid_cap, index_ptr, snew_ptr, tnew_ptr,
                            action_ptr, result_ptr, stream=0):
        """Device pointers (ints) in, enqueue only."""
        vp = lambda q: C.c_void_p(q) if q else None
        _check(lib().gbp_sample_validate_dev(self.h, vp(states_ptr), C.c_int64(table_rows), C.c_int64(n), vp(idx_ptr), vp(dir_ptr), C.byref(params), vp(bits_ptr),
                                             vp(flags_ptr), C.c_int64(valid_cap), vp(index_ptr), vp(snew_ptr), vp(tnew_ptr), vp(action_ptr),
                                             vp(result_ptr), vp(stream)))

    def sample_validate_walk_dev(self, states_ptr, table_rows, n, params, idx_ptr, dir_ptr, bits_ptr, counters_ptr, stream=0):
        vp = lambda q: C.c_void_p(q) if q else None
        _check(lib().gbp_sample_validate_walk_dev(self.h, vp(states_ptr), C.c_int64(table_rows), C.c_int64(n), vp(idx_ptr), vp(dir_ptr),
                                                  C.byref(params), vp(bits_ptr), vp(counters_ptr), vp(stream)))

    def validate_counters(self):
        c = np.zeros(6, np.int64)
        _check(lib().gbp_validate_counters(self.h, _p(c)))
        return dict(substates=int(c[0]), lookups=int(c[1]), nanprobes=int(c[2]), oog=int(c[3]), near=int(c[4]), valid=int(c[5]))

    def sample_states(self, seed, stream, idx0, n, s_from=None, s_to=None, thresh=0.0, speed_dir=False):
        q = np.zeros((n, 8))
        sf = None if s_from is None else _f64(s_from); st = None if s_to is None else _f64(s_to)
        _check(lib().gbp_sample_states(self.h, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.c_int64(n), _p(sf), _p(st),
                                       C.c_double(thresh), int(speed_dir), _p(q)))
        return q

    def sample_states_dev(self, seed, stream, idx0, n, out_ptr, cuda_stream=0):
        _check(lib().gbp_sample_states_dev(self.h, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.c_int64(n),
                                           C.c_void_p(out_ptr), C.c_void_p(cuda_stream) if cuda_stream else None))

    def attempt_connect(self, s_existing, s, direction, adaptive=False):
        se, s = _f64(s_existing, (-1, 8)), _f64(s, (-1, 8)); n = len(s); d = _u8(direction, n)
        st = np.zeros(n, np.int32); sn = np.zeros((n, 8)); an = np.zeros((n, 10)); fl = np.zeros(n, np.uint8)
        _check(lib().gbp_attempt_connect(self.h, C.c_int64(n), _p(se), _p(s), _p(d), int(adaptive), _p(st), _p(sn), _p(an), _p(fl)))
        return st, sn, an, fl

    def plan_batch(self, starts, goals, seed, query0, params, path_cap=0):
        s, g = _f64(starts, (-1, 8)), _f64(goals, (-1, 8)); nq = len(s)
        st = np.zeros(nq, PLAN_STATS_DTYPE)
        ps = np.zeros((nq, path_cap, 8)) if path_cap else None
        pa = np.zeros((nq, path_cap, 10)) if path_cap else None
        _check(lib().gbp_plan_batch(self.h, C.c_int64(nq), _p(s), _p(g), C.c_uint64(seed), C.c_uint64(query0), C.byref(params),
                                    _p(st), _p(ps), _p(pa), int(path_cap)))
        return (st, ps, pa) if path_cap else st

    def plan_batch_form(self, params, nq):
        """0 megakernel, 1 pipelined, 2 stepped, 3 device-wide: the form gbp_plan_batch takes for these parameters and this batch size"""
        f = C.c_int()
        _check(lib().gbp_plan_batch_form(self.h, C.byref(params), C.c_int64(nq), C.byref(f)))
        return ("megakernel", "pipelined", "stepped", "device-wide")[f.value]

    def plan_batch_trees(self, starts, goals, seed, query0, params, path_cap, tree_cap):
        """gbp_plan_batch_trees -> stats, path states, path actions, list of (tree A, tree B) dicts per query"""
        s, g = _f64(starts, (-1, 8)), _f64(goals, (-1, 8)); nq = len(s)
        st = np.zeros(nq, PLAN_STATS_DTYPE)
        ps, pa = np.zeros((nq, path_cap, 8)), np.zeros((nq, path_cap, 10))
        ts, ta = np.zeros((nq, 2, tree_cap, 8)), np.zeros((nq, 2, tree_cap, 10))
        tp = np.zeros((nq, 2, tree_cap), np.int32); tg, ty = np.zeros((nq, 2, tree_cap)), np.zeros((nq, 2, tree_cap))
        _check(lib().gbp_plan_batch_trees(self.h, C.c_int64(nq), _p(s), _p(g), C.c_uint64(seed), C.c_uint64(query0), C.byref(params),
                                          _p(st), _p(ps), _p(pa), int(path_cap), int(tree_cap), _p(ts), _p(ta), _p(tp), _p(tg), _p(ty)))
        trees = []
        for q in range(nq):
            pair = []
            for w, n in ((0, int(st["nv_a"][q])), (1, int(st["nv_b"][q]))):
                n = min(n, tree_cap)
                pair.append(dict(states=ts[q, w, :n].copy(), actions=ta[q, w, :n].copy(), parent=tp[q, w, :n].copy(), g=tg[q, w, :n].copy(),
                                 yaw=ty[q, w, :n].copy()))
            trees.append(tuple(pair))
        return st, ps, pa, trees

    def plan_batch_dev(self, nq, starts_ptr, goals_ptr, seed, query0, params, stats_ptr, stream=0):
        _check(lib().gbp_plan_batch_dev(self.h, C.c_int64(nq), C.c_void_p(starts_ptr), C.c_void_p(goals_ptr), C.c_uint64(seed),
                                        C.c_uint64(query0), C.byref(params), C.c_void_p(stats_ptr), None, None, 0,
                                        C.c_void_p(stream) if stream else None))


def sample_actions_dev(seed, stream, idx0, n, out_ptr, normal=(0.0, 0.0, 1.0), cuda_stream=0):
    nrm = _f64(normal)
    _check(lib().gbp_sample_actions_dev(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.c_int64(n), _p(nrm),
                                        C.c_void_p(out_ptr), C.c_void_p(cuda_stream) if cuda_stream else None))


class Tree:
    """Device-resident GraphClass / PlannerClass store."""

    def __init__(self, capacity, root=None):
        h = C.c_void_p()
        _check(lib().gbp_tree_create(int(capacity), C.byref(h)))
        self.h = h
        if root is not None:
            self.init(root)

    def close(self):
        if getattr(self, "h", None):
            lib().gbp_tree_destroy(self.h)
            self.h = None

    __del__ = close

    def init(self, root):
        r = _f64(root)
        _check(lib().gbp_tree_init(self.h, _p(r)))

    def size(self):
        n = C.c_int()
        _check(lib().gbp_tree_size(self.h, C.byref(n)))
        return n.value

    def append(self, parent, state, action):
        s, a = _f64(state), _f64(action); i = C.c_int()
        _check(lib().gbp_tree_append(self.h, int(parent), _p(s), _p(a), C.byref(i)))
        return i.value

    def load(self, states, actions=None, parent=None):
        s = _f64(states, (-1, 8)); n = len(s)
        a = None if actions is None else _f64(actions, (-1, 10))
        p = np.ascontiguousarray(np.arange(-1, n - 1) if parent is None else parent, dtype=np.int32)
        _check(lib().gbp_tree_load(self.h, n, _p(s), _p(a), _p(p)))

    def read(self, first=0, n=None):
        n = self.size() - first if n is None else n
        s = np.zeros((n, 8)); a = np.zeros((n, 10)); p = np.zeros(n, np.int32); g = np.zeros(n); y = np.zeros(n)
        _check(lib().gbp_tree_read(self.h, first, n, _p(s), _p(a), _p(p), _p(g), _p(y)))
        return dict(states=s, actions=a, parent=p, g=g, yaw=y)

    def nearest(self, queries):
        q = _f64(queries, (-1, 8)); idx = np.zeros(len(q), np.int32); dist = np.zeros(len(q))
        _check(lib().gbp_nearest(self.h, C.c_int64(len(q)), _p(q), _p(idx), _p(dist)))
        return idx, dist

    def near(self, query, radius, cap=None):
        cap = self.size() if cap is None else cap
        ids = np.zeros(max(cap, 1), np.int32); cnt = C.c_int()
        _check(lib().gbp_near(self.h, _p(_f64(query)), C.c_double(radius), _p(ids), int(cap), C.byref(cnt)))
        return ids[:min(cnt.value, cap)].copy(), cnt.value

    def extend(self, terrain, target, direction, k, best_of_k, seed, stream, idx0, adaptive=False, dir_thresh=-1.0):
        st, nid, chk = C.c_int(), C.c_int(), C.c_int64()
        _check(lib().gbp_extend(self.h, terrain.h, _p(_f64(target)), int(direction), int(k), int(best_of_k), int(adaptive),
                                C.c_double(dir_thresh), C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0), C.byref(st), C.byref(nid), C.byref(chk)))
        return st.value, nid.value, chk.value

    def connect(self, terrain, target, direction, adaptive=False):
        st, nid = C.c_int(), C.c_int()
        _check(lib().gbp_connect(self.h, terrain.h, _p(_f64(target)), int(direction), int(adaptive), C.byref(st), C.byref(nid)))
        return st.value, nid.value
