"""ctypes binding of the C ABI in include/icp_b200.h (libicp_b200.so).

Used by tests/, bench.py and __graft_entry__.py only — the product is the shared library and the
drop-in executables; Python is not on the product path. There is no fallback: if the library
cannot be loaded, importing this module raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "libicp_b200.so"))

DIST_SQ, DIST_SQRT, DIST_STD = 0, 1, 2
POINT_TO_POINT, POINT_TO_PLANE = 0, 1
NN_BRUTE, NN_GRID, NN_BRUTE_DIRECT = 0, 1, 2
FLAG_FIX_REFLECTION, FLAG_PROFILE, FLAG_GRAPH, FLAG_REJECT_UNMATCHED = 1, 2, 4, 8


class Params(C.Structure):
    _fields_ = [("metric", C.c_int), ("dist_mode", C.c_int), ("nn_method", C.c_int), ("max_iter", C.c_int),
                ("stop_early", C.c_int), ("sync_every", C.c_int), ("sentinel", C.c_float), ("tol", C.c_double), ("flags", C.c_int)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("iterations_run", C.c_int), ("R", C.c_double * 9), ("t", C.c_double * 3),
                ("last_R", C.c_float * 9), ("last_T", C.c_float * 3), ("elapsed_ms", C.c_float), ("match_ms", C.c_float),
                ("nn_pairs", C.c_double), ("minimize_ms", C.c_float), ("transform_ms", C.c_float)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise OSError("libicp_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "or `make -C fast-point-cloud-registration-with-gpus_b200` (%s)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    sig = {
        "icpb_version": (C.c_int, []),
        "icpb_status_string": (C.c_char_p, [C.c_int]),
        "icpb_default_params": (None, [C.POINTER(Params)]),
        "icpb_device_count": (C.c_int, [ip]),
        "icpb_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "icpb_nccl_unique_id": (C.c_int, [vp]),
        "icpb_create_dist": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]),
        "icpb_dist_info": (C.c_int, [vp, ip, ip, ip]),
        "icpb_destroy": (C.c_int, [vp]),
        "icpb_last_error": (C.c_char_p, [vp]),
        "icpb_device_info": (C.c_int, [vp, ip, ip, C.c_char_p]),
        "icpb_set_target": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "icpb_set_source": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "icpb_get_source": (C.c_int, [vp, vp, C.c_int]),
        "icpb_get_correspondences": (C.c_int, [vp, vp, C.c_int]),
        "icpb_get_min_distances": (C.c_int, [vp, vp, C.c_int]),
        "icpb_match": (C.c_int, [vp, C.c_int, C.c_int, C.c_float]),
        "icpb_minimize": (C.c_int, [vp, C.c_int, fp, fp]),
        "icpb_transform": (C.c_int, [vp, fp]),
        "icpb_set_transform": (C.c_int, [vp, fp, fp]),
        "icpb_get_moments": (C.c_int, [vp, dp, C.c_int]),
        "icpb_estimate_normals": (C.c_int, [vp, C.c_int, fp]),
        "icpb_estimate_normals_ex": (C.c_int, [vp, C.c_int, C.c_int, fp]),
        "icpb_lidar_convert": (C.c_int, [vp, vp, C.c_int, C.c_ulonglong, fp, fp, C.c_int, C.c_int, C.c_int, vp, C.c_int, fp]),
        "icpb_apply_transform": (C.c_int, [vp, fp, fp, vp, C.c_int, vp, C.c_int, fp]),
        "icpb_scale_cloud": (C.c_int, [vp, C.c_float, vp, C.c_int, C.c_int]),
        "icpb_get_neighbors": (C.c_int, [vp, vp, C.c_int]),
        "icpb_get_normals": (C.c_int, [vp, vp, C.c_int]),
        "icpb_set_normals": (C.c_int, [vp, vp, C.c_int]),
        "icpb_run": (C.c_int, [vp, C.POINTER(Params), fp, C.POINTER(Result)]),
        "icpb_iterate_host": (C.c_int, [vp, C.POINTER(Params), vp, C.c_int, vp, C.c_int, vp, fp, fp, fp]),
        "icpb_run_batched": (C.c_int, [vp, C.POINTER(Params), C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, fp]),
        "icpb_measure_fp32_peak": (C.c_int, [vp, dp]),
        "icpb_time_match": (C.c_int, [vp, C.c_int, C.c_int, C.c_float, C.c_int, fp, fp]),
        "icpb_get_grid_stats": (C.c_int, [vp, dp, ip, ip, fp]),
        "icpb_get_filter_stats": (C.c_int, [vp, dp, dp]),
        "icpb_get_filter_config": (C.c_int, [vp, ip, ip, ip, dp]),
        "icpb_get_filter_tc_config": (C.c_int, [vp, ip, ip]),
        "icpb_get_filter_tc_order": (C.c_int, [vp, ip]),
        "icpb_launch_count": (C.c_longlong, [vp]),
        "icpb_host_alloc": (C.c_int, [C.POINTER(vp), C.c_ulonglong]),
        "icpb_host_free": (C.c_int, [vp]),
        "icpb_group_create": (C.c_int, [C.POINTER(vp), ip, C.c_int]),
        "icpb_group_destroy": (C.c_int, [vp]),
        "icpb_group_size": (C.c_int, [vp]),
        "icpb_group_ctx": (vp, [vp, C.c_int]),
        "icpb_group_last_error": (C.c_char_p, [vp]),
        "icpb_group_info": (C.c_int, [vp, ip, ip, ip]),
        "icpb_group_set_target": (C.c_int, [vp, vp, C.c_int]),
        "icpb_group_set_source": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "icpb_group_estimate_normals": (C.c_int, [vp, C.c_int, C.c_int, fp]),
        "icpb_group_run": (C.c_int, [vp, C.POINTER(Params), fp, C.POINTER(Result)]),
        "icpb_group_get_source": (C.c_int, [vp, vp]),
        "icpb_group_get_correspondences": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)     # AttributeError here = the library does not export what the header declares
        fn.restype, fn.argtypes = res, args
    return lib, sorted(sig)


lib, EXPORTS = _load()


class IcpError(RuntimeError):
    pass


def default_params(**kw):
    p = Params()
    lib.icpb_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Thin RAII wrapper around icpb_ctx."""

    def __init__(self, device=0, rank=0, world=1, nccl_id=None):
        self.h = C.c_void_p()
        if world > 1:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            rc = lib.icpb_create_dist(C.byref(self.h), device, rank, world, C.cast(buf, C.c_void_p))
        else:
            rc = lib.icpb_create(C.byref(self.h), device)
        if rc != 0:
            raise IcpError("icpb_create: %s" % lib.icpb_status_string(rc).decode())
        self.n = 0
        self.m = 0

    def close(self):
        if self.h:
            lib.icpb_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise IcpError("%s: %s (%s)" % (what, lib.icpb_status_string(rc).decode(), lib.icpb_last_error(self.h).decode()))

    def device_info(self):
        sm, khz, name = C.c_int(), C.c_int(), C.create_string_buffer(64)
        self._ck(lib.icpb_device_info(self.h, C.byref(sm), C.byref(khz), name), "device_info")
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "name": name.value.decode()}

    def set_target(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, 3)
        self.m = q.shape[0]
        self._ck(lib.icpb_set_target(self.h, _ptr(q), self.m, 0), "set_target")

    def set_source(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        self.n = p.shape[0]
        self._ck(lib.icpb_set_source(self.h, _ptr(p), self.n, 0), "set_source")

    def get_source(self, out=None):
        """`out`: optional C-contiguous float32 [n,3] array to fill (e.g. a view of pinned host memory)."""
        if out is None:
            out = np.empty((self.n, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.size == 3 * self.n
        self._ck(lib.icpb_get_source(self.h, _ptr(out), 0), "get_source")
        return out

    def correspondences(self):
        out = np.empty(self.n, dtype=np.int32)
        self._ck(lib.icpb_get_correspondences(self.h, _ptr(out), 0), "get_correspondences")
        return out

    def min_distances(self):
        out = np.empty(self.n, dtype=np.float32)
        self._ck(lib.icpb_get_min_distances(self.h, _ptr(out), 0), "get_min_distances")
        return out

    def match(self, dist_mode=DIST_SQ, nn_method=NN_BRUTE, sentinel=100000.0):
        self._ck(lib.icpb_match(self.h, dist_mode, nn_method, sentinel), "match")
        return self.correspondences()

    def minimize(self, metric=POINT_TO_POINT):
        R = np.zeros(9, dtype=np.float32)
        T = np.zeros(3, dtype=np.float32)
        self._ck(lib.icpb_minimize(self.h, metric, R.ctypes.data_as(C.POINTER(C.c_float)), T.ctypes.data_as(C.POINTER(C.c_float))), "minimize")
        return R, T

    def transform(self):
        rms = C.c_float()
        self._ck(lib.icpb_transform(self.h, C.byref(rms)), "transform")
        return rms.value

    def set_transform(self, R, T):
        R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(3)
        fpp = C.POINTER(C.c_float)
        self._ck(lib.icpb_set_transform(self.h, R.ctypes.data_as(fpp), T.ctypes.data_as(fpp)), "set_transform")

    def moments(self, count=16):
        out = np.zeros(count, dtype=np.float64)
        self._ck(lib.icpb_get_moments(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), count), "get_moments")
        return out

    def estimate_normals(self, k=4, knn_dist_mode=DIST_SQRT):
        ms = C.c_float()
        self._ck(lib.icpb_estimate_normals_ex(self.h, k, knn_dist_mode, C.byref(ms)), "estimate_normals")
        return ms.value

    def lidar_convert(self, ranges, encoder_count, altitude, azimuth, ticks_per_block=88, ticks_per_rev=90112):
        ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        altitude = np.ascontiguousarray(altitude, dtype=np.float32)
        azimuth = np.ascontiguousarray(azimuth, dtype=np.float32)
        out = np.empty((ranges.shape[0], 3), dtype=np.float32)
        ms = C.c_float()
        fpp = C.POINTER(C.c_float)
        self._ck(lib.icpb_lidar_convert(self.h, _ptr(ranges), ranges.shape[0], int(encoder_count), altitude.ctypes.data_as(fpp),
                                        azimuth.ctypes.data_as(fpp), altitude.shape[0], ticks_per_block, ticks_per_rev, _ptr(out), 0, C.byref(ms)), "lidar_convert")
        return out, ms.value

    def apply_transform(self, R, T, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(3)
        out = np.empty_like(xyz)
        fpp = C.POINTER(C.c_float)
        self._ck(lib.icpb_apply_transform(self.h, R.ctypes.data_as(fpp), T.ctypes.data_as(fpp), _ptr(xyz), xyz.shape[0], _ptr(out), 0, None), "apply_transform")
        return out

    def scale_cloud(self, alpha, xyz):
        out = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3).copy()
        self._ck(lib.icpb_scale_cloud(self.h, float(alpha), _ptr(out), out.shape[0], 0), "scale_cloud")
        return out

    def neighbors(self, k=4):
        out = np.empty((self.m, k + 1), dtype=np.int32)
        self._ck(lib.icpb_get_neighbors(self.h, _ptr(out), 0), "get_neighbors")
        return out

    def normals(self):
        out = np.empty((self.m, 3), dtype=np.float32)
        self._ck(lib.icpb_get_normals(self.h, _ptr(out), 0), "get_normals")
        return out

    def set_normals(self, nrm):
        nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        self._ck(lib.icpb_set_normals(self.h, _ptr(nrm), 0), "set_normals")

    def run(self, params):
        errors = np.zeros(params.max_iter + 1, dtype=np.float32)
        res = Result()
        self._ck(lib.icpb_run(self.h, C.byref(params), errors.ctypes.data_as(C.POINTER(C.c_float)), C.byref(res)), "run")
        return errors, res

    def iterate_host(self, params, p, q, want_idx=True, idx_out=None):
        """`idx_out`: optional int32 [n] array to receive the correspondences (e.g. a view of pinned host memory)."""
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, 3)
        self.n, self.m = p.shape[0], q.shape[0]
        idx = (idx_out if idx_out is not None else np.empty(self.n, dtype=np.int32)) if want_idx else None
        assert idx is None or (idx.dtype == np.int32 and idx.flags["C_CONTIGUOUS"] and idx.size == self.n)
        R = np.zeros(9, dtype=np.float32)
        T = np.zeros(3, dtype=np.float32)
        rms = C.c_float()
        fpp = C.POINTER(C.c_float)
        self._ck(lib.icpb_iterate_host(self.h, C.byref(params), _ptr(p), self.n, _ptr(q), self.m,
                                       _ptr(idx) if want_idx else None, R.ctypes.data_as(fpp), T.ctypes.data_as(fpp), C.byref(rms)), "iterate_host")
        return idx, R, T, rms.value

    def run_batched(self, params, sources, targets):
        sources = np.ascontiguousarray(sources, dtype=np.float32)
        targets = np.ascontiguousarray(targets, dtype=np.float32)
        batch, n = sources.shape[0], sources.shape[1]
        m = targets.shape[1]
        errors = np.zeros((batch, params.max_iter + 1), dtype=np.float32)
        iters = np.zeros(batch, dtype=np.int32)
        R = np.zeros((batch, 9), dtype=np.float64)
        t = np.zeros((batch, 3), dtype=np.float64)
        ms = C.c_float()
        self._ck(lib.icpb_run_batched(self.h, C.byref(params), batch, _ptr(sources), n, _ptr(targets), m,
                                      _ptr(errors), _ptr(iters), _ptr(R), _ptr(t), C.byref(ms)), "run_batched")
        return errors, iters, R, t, ms.value

    def run_batched_ptr(self, params, batch, n, m, sources_ptr, targets_ptr):
        """Same, with the clouds given as raw addresses (device pointers are used in place: the device-resident timing)."""
        errors = np.zeros((batch, params.max_iter + 1), dtype=np.float32)
        iters = np.zeros(batch, dtype=np.int32)
        R = np.zeros((batch, 9), dtype=np.float64)
        t = np.zeros((batch, 3), dtype=np.float64)
        ms = C.c_float()
        self._ck(lib.icpb_run_batched(self.h, C.byref(params), batch, C.c_void_p(sources_ptr), n, C.c_void_p(targets_ptr), m,
                                      _ptr(errors), _ptr(iters), _ptr(R), _ptr(t), C.byref(ms)), "run_batched")
        return errors, iters, R, t, ms.value

    def fp32_peak_tflops(self):
        v = C.c_double()
        self._ck(lib.icpb_measure_fp32_peak(self.h, C.byref(v)), "measure_fp32_peak")
        return v.value

    def time_match(self, dist_mode=DIST_SQ, nn_method=NN_BRUTE, sentinel=100000.0, reps=3):
        mean, mn = C.c_float(), C.c_float()
        self._ck(lib.icpb_time_match(self.h, dist_mode, nn_method, sentinel, reps, C.byref(mean), C.byref(mn)), "time_match")
        return mean.value, mn.value

    def grid_stats(self):
        cand, opened, dims, cell = C.c_double(), C.c_int(), (C.c_int * 3)(), C.c_float()
        self._ck(lib.icpb_get_grid_stats(self.h, C.byref(cand), C.byref(opened), dims, C.byref(cell)), "get_grid_stats")
        return {"candidates_visited": cand.value, "last_open_sources": opened.value, "dims": list(dims), "cell": cell.value}

    def filter_stats(self):
        a, b = C.c_double(), C.c_double()
        self._ck(lib.icpb_get_filter_stats(self.h, C.byref(a), C.byref(b)), "get_filter_stats")
        return {"subtile_tests": a.value, "subtile_exact": b.value}

    def filter_config(self):
        a, b, d, f = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self._ck(lib.icpb_get_filter_config(self.h, C.byref(a), C.byref(b), C.byref(d), C.byref(f)), "get_filter_config")
        return {"dims_next": a.value, "dims_last": b.value, "drop_axis": d.value, "last_exact_fraction": f.value}

    def filter_tc_config(self):
        a, b = C.c_int(), C.c_int()
        self._ck(lib.icpb_get_filter_tc_config(self.h, C.byref(a), C.byref(b)), "get_filter_tc_config")
        o = C.c_int()
        self._ck(lib.icpb_get_filter_tc_order(self.h, C.byref(o)), "get_filter_tc_order")
        return {"enabled": bool(a.value), "targets_per_column": b.value, "morton_order": bool(o.value)}

    def dist_info(self):
        r, w, p = C.c_int(), C.c_int(), C.c_int()
        self._ck(lib.icpb_dist_info(self.h, C.byref(r), C.byref(w), C.byref(p)), "dist_info")
        return {"rank": r.value, "world": w.value, "peer_exchange": bool(p.value)}

    def launch_count(self):
        return int(lib.icpb_launch_count(self.h))


class Group:
    """Several GPUs driven from this one process through icpb_group_* (the C/C++ multi-GPU host path)."""

    def __init__(self, ndev, devices=None):
        self.h = C.c_void_p()
        dev = (C.c_int * ndev)(*devices) if devices is not None else None
        rc = lib.icpb_group_create(C.byref(self.h), dev, ndev)
        if rc != 0:
            raise IcpError("icpb_group_create: %s" % lib.icpb_status_string(rc).decode())
        self.ndev, self.n, self.m = ndev, 0, 0

    def close(self):
        if self.h:
            lib.icpb_group_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise IcpError("%s: %s (%s)" % (what, lib.icpb_status_string(rc).decode(), lib.icpb_group_last_error(self.h).decode()))

    def info(self):
        n, p, d = C.c_int(), C.c_int(), (C.c_int * self.ndev)()
        self._ck(lib.icpb_group_info(self.h, C.byref(n), C.byref(p), d), "group_info")
        return {"ndev": n.value, "peer_exchange": bool(p.value), "devices": list(d)}

    def set_target(self, q):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, 3)
        self.m = q.shape[0]
        self._ck(lib.icpb_group_set_target(self.h, _ptr(q), self.m), "group_set_target")

    def set_source(self, p, block=2048):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        self.n = p.shape[0]
        self._ck(lib.icpb_group_set_source(self.h, _ptr(p), self.n, block), "group_set_source")

    def estimate_normals(self, k=4, knn_dist_mode=DIST_SQRT):
        ms = C.c_float()
        self._ck(lib.icpb_group_estimate_normals(self.h, k, knn_dist_mode, C.byref(ms)), "group_estimate_normals")
        return ms.value

    def run(self, params):
        errors = np.zeros(params.max_iter + 1, dtype=np.float32)
        res = Result()
        self._ck(lib.icpb_group_run(self.h, C.byref(params), errors.ctypes.data_as(C.POINTER(C.c_float)), C.byref(res)), "group_run")
        return errors, res

    def get_source(self):
        out = np.empty((self.n, 3), dtype=np.float32)
        self._ck(lib.icpb_group_get_source(self.h, _ptr(out)), "group_get_source")
        return out

    def correspondences(self):
        out = np.empty(self.n, dtype=np.int32)
        self._ck(lib.icpb_group_get_correspondences(self.h, _ptr(out)), "group_get_correspondences")
        return out

    def launch_count(self, rank=0):
        return int(lib.icpb_launch_count(lib.icpb_group_ctx(self.h, rank)))


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib.icpb_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise IcpError("icpb_nccl_unique_id: %s" % lib.icpb_status_string(rc).decode())
    return bytes(buf.raw)
