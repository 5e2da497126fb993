"""numpy restatement of the reference's synthetic clouds (src/ICP_point_to_point.cu:103-190): the
saddle z = x^2 - y^2 on a WxW grid over [-2,2]^2 and its rigidly moved copy. Every operation is an
element-wise float32 numpy op (one rounding each, no FMA), in the reference's order, so the clouds are
bit-identical to what the reference host code builds (tests/test_synth.py checks this against the
oracle's C generator). Used by bench.py; the C++ twin for the executables is apps/synth.h."""
import numpy as np

F = np.float32


def source(width, npts=None):
    npts = width * width if npts is None else npts
    i = np.arange(width, dtype=np.float32)
    lin = F(-2.0) + (i * F(4.0)) / (F(width) - F(1.0))
    k = np.arange(npts)
    x = lin[k // width]
    y = lin[k % width]
    z = (x.astype(np.float64) ** 2 - y.astype(np.float64) ** 2).astype(np.float32)
    return np.ascontiguousarray(np.stack([x, y, z], axis=1))


def euler_matrix(ri):
    cx, cy, cz = (F(np.cos(np.float64(F(a)))) for a in ri)
    sx, sy, sz = (F(np.sin(np.float64(F(a)))) for a in ri)
    r = np.zeros(9, np.float32)
    r[0] = cy * cz; r[1] = (cz * sx * sy) + (cx * sz); r[2] = -(cx * cz * sy) + (sx * sz)
    r[3] = -cy * sz; r[4] = (cx * cz) - (sx * sy * sz); r[5] = (cx * sy * sz) + (cz * sx)
    r[6] = sy; r[7] = -cy * sx; r[8] = cx * cy
    return r


def rigid_move(D, r, t):
    x, y, z = D[:, 0], D[:, 1], D[:, 2]
    cols = []
    for j in range(3):
        acc = r[j] * x            # temp = 0 + A[j,0]*B[0]
        acc = acc + r[j + 3] * y
        acc = acc + r[j + 6] * z
        cols.append(acc + F(t[j]))
    return np.ascontiguousarray(np.stack(cols, axis=1).astype(np.float32))


def p2p_clouds(width, npts=None):
    """Defaults of ICP_point_to_point.cu: t = (0.8,-0.3,0.2), r = (0.2,-0.2,0.05) rad."""
    D = source(width, npts)
    M = rigid_move(D, euler_matrix([0.2, -0.2, 0.05]), [0.8, -0.3, 0.2])
    return D, M


def cpu_clouds(width=100):
    """The clouds of src/ICP_CPU.c:51-149 (BASELINE.json config 0/1: double precision, its own pose t = (1,-0.3,0.2),
    r = (1,-0.5,0.05) with transposed-sign elementary rotations, R = rx*ry*rz), rounded to the float32 AoS layout the
    GPU engine takes."""
    i = np.arange(width, dtype=np.float64)
    lin = -2.0 + i * 4.0 / (float(width) - 1.0)
    k = np.arange(width * width)
    x, y = lin[k // width], lin[k % width]
    D = np.stack([x, y, x ** 2 - y ** 2], axis=0)
    a, b, c = 1.0, -0.5, 0.05
    rx = np.array([[1, 0, 0], [0, np.cos(a), np.sin(a)], [0, -np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, -np.sin(b)], [0, 1, 0], [np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), np.sin(c), 0], [-np.sin(c), np.cos(c), 0], [0, 0, 1]])
    M = (rx @ ry @ rz) @ D + np.array([1.0, -0.3, 0.2])[:, None]
    return np.ascontiguousarray(D.T, np.float32), np.ascontiguousarray(M.T, np.float32)


# ---- BASELINE.json config 5: batched independent pairs ---------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(state):
    """One step of splitmix64 on an array of uint64 states -> (new_state, output)."""
    with np.errstate(over="ignore"):
        state = state + np.uint64(0x9E3779B97F4A7C15)
        z = state.copy()
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return state, z


def batched_poses(batch, seed=20240):
    """Pose of pair b from the counter-based generator: stream b of splitmix64(seed): r = 0.2*u, u in U(-1,1)^3;
    t = (0.8,-0.3,0.2) * (0.5 + 0.5 v), v in U(0,1)^3."""
    state = (np.uint64(seed) << np.uint64(32)) | np.arange(batch, dtype=np.uint64)
    draws = []
    for _ in range(6):
        state, z = _splitmix64(state)
        draws.append((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))
    u = np.stack(draws[:3], axis=1) * 2.0 - 1.0
    v = np.stack(draws[3:], axis=1)
    r = (0.2 * u).astype(np.float32)
    t = (np.array([0.8, -0.3, 0.2]) * (0.5 + 0.5 * v)).astype(np.float32)
    return r, t


def batched_pairs(batch, n=2048, width=46, seed=20240):
    """sources [batch,n,3] (first n points of the width x width saddle, identical for every pair) and
    targets [batch,n,3] (each moved by its own pose)."""
    D = source(width, n)
    r, t = batched_poses(batch, seed)
    S = np.ascontiguousarray(np.broadcast_to(D, (batch, n, 3)))
    T = np.empty((batch, n, 3), np.float32)
    for b in range(batch):
        T[b] = rigid_move(D, euler_matrix(r[b]), t[b])
    return S, T, r, t
