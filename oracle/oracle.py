"""TEST INFRASTRUCTURE — ctypes loader for oracle/liboracle.so (the CPU restatement in icp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")

MODE_SQ, MODE_SQRT, MODE_STD = 0, 1, 2


def build():
    """Compile the restatement (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _load():
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return C.CDLL(LIB_PATH)


_lib = _load()
_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_i = C.POINTER(C.c_int)


def _fp(a): return a.ctypes.data_as(_f)
def _dp(a): return a.ctypes.data_as(_d)
def _ip(a): return a.ctypes.data_as(_i)


_lib.orc_num_threads.restype = C.c_int
_lib.orc_set_num_threads.restype = C.c_int
_lib.orc_icp_p2p_f32.restype = C.c_int
_lib.orc_icp_p2plane_f32.restype = C.c_int
_lib.orc_icp_p2plane_mode_f32.restype = C.c_int
_lib.orc_read_cloud_text.restype = C.c_int
_lib.orc_lidar_parse_packets.restype = C.c_int
_lib.orc_lidar_read_beams.restype = C.c_int
_lib.orc_icp_cpu_f64.restype = C.c_int
_lib.orc_rms.restype = C.c_double
_lib.orc_plane_rt.restype = C.c_int
_lib.orc_solve6_spd.restype = C.c_int


def num_threads():
    return _lib.orc_num_threads()


def set_num_threads(n=0):
    """n <= 0: every processor OpenMP sees (not OMP_NUM_THREADS, which torchrun sets to 1). Returns the count in effect."""
    return _lib.orc_set_num_threads(C.c_int(int(n)))


def synth_p2p(width, npts=None):
    npts = width * width if npts is None else npts
    D = np.zeros((npts, 3), np.float32); M = np.zeros((npts, 3), np.float32)
    _lib.orc_synth_p2p_f32(C.c_int(width), C.c_int(npts), _fp(D), _fp(M))
    return D, M


def synth_source(width, npts=None):
    npts = width * width if npts is None else npts
    D = np.zeros((npts, 3), np.float32)
    _lib.orc_synth_source_f32(C.c_int(width), C.c_int(npts), _fp(D))
    return D


def euler_matrix(ri):
    r = np.zeros(9, np.float32)
    _lib.orc_euler_matrix_f32(_fp(np.asarray(ri, np.float32)), _fp(r))
    return r


def rigid_move(D, h_r, t):
    D = np.ascontiguousarray(D, np.float32)
    M = np.zeros_like(D)
    _lib.orc_rigid_move_f32(_fp(D), C.c_int(D.shape[0]), _fp(np.asarray(h_r, np.float32)), _fp(np.asarray(t, np.float32)), _fp(M))
    return M


def synth_standard(width):
    n = width * width
    D = np.zeros((n, 3), np.float32); M = np.zeros((n, 3), np.float32)
    _lib.orc_synth_standard_f32(C.c_int(width), _fp(D), _fp(M))
    return D, M


def synth_cpu_f64(width):
    n = width * width
    D = np.zeros(3 * n, np.float64); M = np.zeros(3 * n, np.float64)
    _lib.orc_synth_cpu_f64(C.c_int(width), _dp(D), _dp(M))
    return D, M


def match(P, Q, mode=MODE_SQ, sentinel=100000.0, idx0=None):
    P = np.ascontiguousarray(P, np.float32); Q = np.ascontiguousarray(Q, np.float32)
    idx = np.zeros(P.shape[0], np.int32) if idx0 is None else np.ascontiguousarray(idx0, np.int32).copy()
    _lib.orc_match_f32(_fp(P), C.c_int(P.shape[0]), _fp(Q), C.c_int(Q.shape[0]), C.c_int(mode), C.c_float(sentinel), _ip(idx))
    return idx


def pair_distances(P, Q, mode=MODE_SQ):
    """d[i] = the reference kernel's distance between P[i] and Q[i] (same arithmetic as `match`)."""
    P = np.ascontiguousarray(P, np.float32).reshape(-1, 3); Q = np.ascontiguousarray(Q, np.float32).reshape(-1, 3)
    d = np.zeros(P.shape[0], np.float32)
    _lib.orc_pair_dist_f32(_fp(P), _fp(Q), C.c_int(P.shape[0]), C.c_int(mode), _fp(d))
    return d


def moments(P, Q, idx):
    P = np.ascontiguousarray(P, np.float32); Q = np.ascontiguousarray(Q, np.float32); idx = np.ascontiguousarray(idx, np.int32)
    mom = np.zeros(16, np.float64)
    _lib.orc_moments(_fp(P), _fp(Q), _ip(idx), C.c_int(P.shape[0]), _dp(mom))
    return mom


def rt_from_moments(mom):
    R = np.zeros(9, np.float64); T = np.zeros(3, np.float64)
    _lib.orc_rt_from_moments(_dp(np.ascontiguousarray(mom, np.float64)), _dp(R), _dp(T))
    return R, T


def polar_rotation(W):
    R = np.zeros(9, np.float64)
    _lib.orc_polar_rotation(_dp(np.ascontiguousarray(W, np.float64)), _dp(R))
    return R


def transform(P, R, T):
    P = np.ascontiguousarray(P, np.float32).copy()
    _lib.orc_transform_f32(_fp(P), C.c_int(P.shape[0]), _fp(np.asarray(R, np.float32)), _fp(np.asarray(T, np.float32)))
    return P


def rms(P, Q, idx):
    P = np.ascontiguousarray(P, np.float32); Q = np.ascontiguousarray(Q, np.float32); idx = np.ascontiguousarray(idx, np.int32)
    return float(_lib.orc_rms(_fp(P), _fp(Q), _ip(idx), C.c_int(P.shape[0])))


def icp_p2p(P, Q, mode=MODE_SQ, sentinel=100000.0, max_iter=40, tol=1e-6, stop_early=True):
    P = np.ascontiguousarray(P, np.float32).copy(); Q = np.ascontiguousarray(Q, np.float32)
    n = P.shape[0]
    errors = np.zeros(max_iter + 1, np.float32); idx = np.zeros(n, np.int32)
    R = np.zeros(9); t = np.zeros(3); run = C.c_int()
    it = _lib.orc_icp_p2p_f32(_fp(P), C.c_int(n), _fp(Q), C.c_int(Q.shape[0]), C.c_int(mode), C.c_float(sentinel), C.c_int(max_iter),
                              C.c_double(tol), C.c_int(1 if stop_early else 0), _fp(errors), _ip(idx), _dp(R), _dp(t), C.byref(run))
    return {"iterations": it, "iterations_run": run.value, "errors": errors, "idx": idx, "R": R, "t": t, "P": P}


def icp_cpu_f64(width=100, max_iter=200, tol=1e-5):
    D, M = synth_cpu_f64(width)
    n = width * width
    pt = D.copy(); E = np.zeros(max_iter + 1); idx = np.zeros(n, np.int32); R = np.zeros(9); t = np.zeros(3)
    it = _lib.orc_icp_cpu_f64(_dp(pt), C.c_int(n), _dp(M), C.c_int(n), C.c_int(max_iter), C.c_double(tol), _dp(E), _ip(idx), _dp(R), _dp(t))
    return {"iterations": it, "errors": E, "idx": idx, "R": R, "t": t, "D": D, "M": M, "pt": pt}


def knn(Q, k1=5, mode=MODE_SQRT):
    Q = np.ascontiguousarray(Q, np.float32)
    nbr = np.zeros((Q.shape[0], k1), np.int32)
    _lib.orc_knn_mode_f32(_fp(Q), C.c_int(Q.shape[0]), C.c_int(k1), C.c_int(mode), _ip(nbr))
    return nbr


def normals(Q, nbr, k=4):
    Q = np.ascontiguousarray(Q, np.float32); nbr = np.ascontiguousarray(nbr, np.int32)
    out = np.zeros((Q.shape[0], 3), np.float32)
    _lib.orc_normals_f32(_fp(Q), C.c_int(Q.shape[0]), _ip(nbr), C.c_int(k), _fp(out))
    return out


def cxb(P, Q, idx, nrm):
    P = np.ascontiguousarray(P, np.float32); Q = np.ascontiguousarray(Q, np.float32)
    idx = np.ascontiguousarray(idx, np.int32); nrm = np.ascontiguousarray(nrm, np.float32)
    Cm = np.zeros(36); b = np.zeros(6)
    _lib.orc_cxb(_fp(P), _fp(Q), _ip(idx), _fp(nrm), C.c_int(P.shape[0]), _dp(Cm), _dp(b))
    return Cm, b


def plane_rt(Cm, b):
    R = np.zeros(9, np.float32); T = np.zeros(3, np.float32)
    info = _lib.orc_plane_rt(_dp(np.ascontiguousarray(Cm, np.float64)), _dp(np.ascontiguousarray(b, np.float64)), _fp(R), _fp(T))
    return info, R, T


def icp_p2plane(P, Q, nrm, sentinel=100000.0, max_iter=50, tol=1e-6, mode=MODE_SQRT):
    P = np.ascontiguousarray(P, np.float32).copy(); Q = np.ascontiguousarray(Q, np.float32); nrm = np.ascontiguousarray(nrm, np.float32)
    n = P.shape[0]
    errors = np.zeros(max_iter + 1, np.float32); idx = np.zeros(n, np.int32)
    R = np.zeros(9); t = np.zeros(3); run = C.c_int()
    it = _lib.orc_icp_p2plane_mode_f32(_fp(P), C.c_int(n), _fp(Q), C.c_int(Q.shape[0]), _fp(nrm), C.c_int(mode), C.c_float(sentinel), C.c_int(max_iter),
                                  C.c_double(tol), _fp(errors), _ip(idx), _dp(R), _dp(t), C.byref(run))
    return {"iterations": it, "iterations_run": run.value, "errors": errors, "idx": idx, "R": R, "t": t, "P": P}


# ---- dataset front ends ----------------------------------------------------
def read_cloud_text(path, max_points=1 << 20):
    buf = np.zeros(3 * max_points, np.float32)
    n = _lib.orc_read_cloud_text(os.fsencode(path), _fp(buf), C.c_int(buf.size))
    if n < 0:
        raise OSError("cannot open %s" % path)
    return buf[:n].copy()


def lidar_parse_packets(path, n=16384):
    r = np.zeros(n, np.float32)
    enc = C.c_ulonglong()
    k = _lib.orc_lidar_parse_packets(os.fsencode(path), _fp(r), C.c_int(n), C.byref(enc))
    if k < 0:
        raise OSError("cannot open %s" % path)
    return r[:k].copy(), int(enc.value)


def lidar_read_beams(path):
    alt = np.zeros(16, np.float32); az = np.zeros(16, np.float32)
    if _lib.orc_lidar_read_beams(os.fsencode(path), _fp(alt), _fp(az)) != 0:
        raise OSError("cannot open %s" % path)
    return alt, az


def lidar_convert(ranges, encoder_count, altitude, azimuth):
    ranges = np.ascontiguousarray(ranges, np.float32)
    out = np.zeros((ranges.shape[0], 3), np.float32)
    _lib.orc_lidar_convert(_fp(ranges), C.c_int(ranges.shape[0]), C.c_ulonglong(encoder_count),
                           _fp(np.ascontiguousarray(altitude, np.float32)), _fp(np.ascontiguousarray(azimuth, np.float32)), _fp(out))
    return out


def lidar_clouds(data_dir, n=16384):
    """Read_data of the LiDAR programs (src/CUDA/GPU_point_to_point_real.cu:432-620) + the mm -> m scaling (:169-171):
    source = converted scan, target = RyT(scan) with t = (0.001,-0.0202,0.02), r = (0.01,-0.003,0.05)."""
    r, enc = lidar_parse_packets(os.path.join(data_dir, "Donut_1024x16.csv"), n)
    alt, az = lidar_read_beams(os.path.join(data_dir, "beam_intrinsics.csv"))
    P = lidar_convert(r, enc, alt, az)
    Q = transform(P, euler_matrix([0.01, -0.003, 0.05]), np.array([0.001, -0.0202, 0.02], np.float32))
    a = np.float32(1.0 / 1000.0)
    return P * a, Q * a, (r, enc, alt, az, P, Q)


def bunny_clouds(data_dir, n=8171):
    """src/CUDA/GPU_point_to_point_bunny.cu:109-166: the data cloud and its copy moved by t = (0.01,-0.04,0.02),
    r = (0.15,-0.1,0.05)."""
    D = read_cloud_text(os.path.join(data_dir, "Bunny_res.csv"))[:3 * n].reshape(n, 3)
    M = rigid_move(D, euler_matrix([0.15, -0.1, 0.05]), np.array([0.01, -0.04, 0.02], np.float32))
    return D, M
