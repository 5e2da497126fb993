"""CPU suite: known answers and edge cases of the oracle's building blocks."""
import numpy as np
import pytest


def test_tie_breaks_to_lowest_index_and_sentinel(orc):
    Q = np.array([[1, 0, 0], [0, 1, 0], [1, 0, 0], [-1, 0, 0]], np.float32)   # 0 and 2 are duplicates
    P = np.array([[1, 0, 0], [0, 0, 0], [500, 0, 0]], np.float32)
    for mode in (orc.MODE_SQ, orc.MODE_SQRT, orc.MODE_STD):
        idx = orc.match(P, Q, mode)
        assert idx[0] == 0          # exact duplicate: first wins
        assert idx[1] == 0          # four equidistant targets: first wins
    # nothing closer than the sentinel: the previous correspondence is kept (src/ICP_point_to_point.cu:51-55)
    prev = np.array([7, 7, 7], np.int32)
    idx = orc.match(P, Q, orc.MODE_SQ, sentinel=100000.0, idx0=prev)
    assert idx[2] == 7              # d^2 = 249001 > 100000 for every target
    idx = orc.match(P, Q, orc.MODE_SQRT, sentinel=100000.0, idx0=prev)
    assert idx[2] in (0, 2) and idx[2] == 0   # sqrt mode compares ~499 against 100000: matched


def test_sqrt_merges_distinct_squares(orc):
    """sqrt.rn maps several float d^2 to one float d, which moves ties to a lower index."""
    base = np.float32(1.0)
    up = np.nextafter(base, np.float32(2.0))
    # target 0 at squared distance `up`, target 1 at squared distance 1.0: SQ picks 1, SQRT picks 0 iff sqrt merges
    P = np.zeros((1, 3), np.float32)
    Q = np.array([[np.sqrt(np.float64(up)), 0, 0], [1, 0, 0]], np.float32)
    d0 = np.float32(Q[0, 0]) * np.float32(Q[0, 0])
    sq = orc.match(P, Q, orc.MODE_SQ)[0]
    sr = orc.match(P, Q, orc.MODE_SQRT)[0]
    assert sq == (0 if d0 <= np.float32(1.0) else 1)
    assert sr == (0 if np.sqrt(d0, dtype=np.float32) <= np.float32(1.0) else 1)


def test_centroid_known_answer(orc):
    """src/tests/centroid.cu:68-73: 2048 points of (1,2,3) -> sums (2048, 4096, 6144)."""
    P = np.tile(np.array([1, 2, 3], np.float32), (2048, 1))
    Q = P.copy()
    mom = orc.moments(P, Q, np.arange(2048, dtype=np.int32))
    assert np.array_equal(mom[:3], [2048.0, 4096.0, 6144.0])
    assert np.array_equal(mom[3:6], [2048.0, 4096.0, 6144.0])
    assert mom[15] == 2048


def test_polar_rotation_vs_numpy(orc):
    rng = np.random.default_rng(1)
    for _ in range(50):
        W = rng.normal(size=(3, 3))
        R = orc.polar_rotation(W.T.reshape(-1)).reshape(3, 3).T      # column-major in, column-major out
        U, _, Vt = np.linalg.svd(W)
        assert np.allclose(R, U @ Vt, atol=1e-12)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)


def test_rt_recovers_known_motion(orc):
    rng = np.random.default_rng(2)
    P = rng.normal(size=(500, 3)).astype(np.float32)
    h_r = orc.euler_matrix([0.3, -0.2, 0.1])
    t = np.array([0.5, -1.0, 2.0], np.float32)
    Q = orc.rigid_move(P, h_r, t)
    mom = orc.moments(P, Q, np.arange(500, dtype=np.int32))
    R, T = orc.rt_from_moments(mom)
    assert np.allclose(R, h_r.astype(np.float64), atol=1e-6)
    assert np.allclose(T, t, atol=1e-5)


def test_plane_solve_vs_numpy(orc):
    rng = np.random.default_rng(3)
    A = rng.normal(size=(6, 6))
    S = A @ A.T + 6 * np.eye(6)
    b = rng.normal(size=6) * 0.01
    Cm = np.triu(S).T.reshape(-1)     # column-major with only the upper triangle filled
    info, R, T = orc.plane_rt(Cm, b)
    assert info == 0
    x = np.linalg.solve(S.astype(np.float32).astype(np.float64), b.astype(np.float32).astype(np.float64))
    assert np.allclose(T, x[3:], rtol=1e-5, atol=1e-7)
    cx, cy, cz = np.cos(x[:3]); sx, sy, sz = np.sin(x[:3])
    assert np.isclose(R[0], cy * cz, atol=1e-6) and np.isclose(R[2], -sy, atol=1e-6) and np.isclose(R[5], cy * sx, atol=1e-6)
    # not positive definite -> potrf's devInfo
    info, _, _ = orc.plane_rt(-Cm, b)
    assert info == 1


def test_generators(orc):
    D, M = orc.synth_p2p(128)
    assert D.shape == (16384, 3)
    assert D[0, 0] == -2.0 and D[0, 1] == -2.0 and D[0, 2] == 0.0
    assert D[127, 1] == 2.0 and D[128, 0] > -2.0          # y fast, x slow
    assert np.allclose(D[:, 2], D[:, 0].astype(np.float64) ** 2 - D[:, 1].astype(np.float64) ** 2, atol=1e-6)
    Dt, Mt = orc.synth_p2p(317, 100000)                     # "100k = first 100 000 points of the W=317 grid"
    assert Dt.shape == (100000, 3)
    Dfull, _ = orc.synth_p2p(317)
    assert np.array_equal(Dt, Dfull[:100000])
    Ds, Ms = orc.synth_standard(32)
    assert np.allclose(Ms[0], np.array([0.876485812, -0.37591464, 0.300767018]) * -2 + np.array([-0.04386084, 0.559789799, 0.827473024]) * -2 + np.array([1, -0.3, 0.2]), atol=1e-5)


def test_ground_truth_recovery_small(orc):
    """the composed transform equals the generating pose, idx == i at convergence."""
    D, M = orc.synth_p2p(32)
    o = orc.icp_p2p(D, M, max_iter=40)
    assert o["iterations_run"] == 15
    assert np.array_equal(o["idx"], np.arange(1024))
    h_r = orc.euler_matrix([0.2, -0.2, 0.05]).astype(np.float64)
    assert np.abs(o["R"] - h_r).max() < 1e-5
    assert np.abs(o["t"] - np.array([0.8, -0.3, 0.2])).max() < 1e-5
