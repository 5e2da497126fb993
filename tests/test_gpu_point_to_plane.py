"""GPU suite (-m gpu): k-NN, PCA normals, 6x6 normal equations and the point-to-plane loop through the
C ABI, against the oracle and against the golden outputs of the reference's own knn / Normals / Cxb kernels."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200")


def test_knn_bit_exact_vs_reference_kernel_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_knn_normals.npz"))
    for W in (32, 64):
        ctx.set_target(g["Q_W%d" % W])
        ctx.estimate_normals(4)
        assert np.array_equal(ctx.neighbors(4), g["nbr_W%d" % W]), "neighbour lists differ from the reference knn kernel (W=%d)" % W


@pytest.mark.parametrize("m", [5, 9, 257, 1000, 4097])
def test_knn_ties_and_ragged_vs_oracle(ctx, orc, m):
    """Lattice clouds: many exactly equal distances, duplicates, sizes not multiple of anything."""
    rng = np.random.default_rng(m)
    Q = (rng.integers(-6, 7, size=(m, 3)) * 0.5).astype(np.float32)
    ctx.set_target(Q)
    ctx.estimate_normals(4)
    assert np.array_equal(ctx.neighbors(4), orc.knn(Q, 5))


def test_knn_sqrt_merge_certificate_fallback(ctx, orc):
    """Neighbours whose squared distances differ by an ulp but whose float square roots coincide, arranged so
    that more than the kept candidates tie: the certificate must fail and the exact path must take over."""
    b = np.float32(1.0)
    e = np.sqrt(np.float64(np.spacing(b))).astype(np.float32)
    pts = [[0, 0, 0]]
    for k in range(12):                      # 12 points on a circle of radius 1 (+ tiny offsets along z)
        a = 2 * np.pi * k / 12
        pts.append([np.cos(a), np.sin(a), e if k % 2 else 0.0])
    Q = np.array(pts, np.float32)
    ctx.set_target(Q)
    ctx.estimate_normals(4)
    assert np.array_equal(ctx.neighbors(4), orc.knn(Q, 5))


def test_normals_vs_oracle_and_reference_covariance(ctx, orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_knn_normals.npz"))
    for W in (32, 64):
        Q = g["Q_W%d" % W]
        ctx.set_target(Q)
        ctx.estimate_normals(4)
        n_gpu = ctx.normals().astype(np.float64)
        n_orc = orc.normals(Q, g["nbr_W%d" % W], 4).astype(np.float64)
        assert np.allclose(np.linalg.norm(n_gpu, axis=1), 1.0, atol=1e-5)
        cos = np.abs(np.sum(n_gpu * n_orc, axis=1))
        # sign is free; direction within 1e-5 wherever the smallest eigenvalue is separated
        A = g["A_W%d" % W].astype(np.float64)
        full = np.zeros((W * W, 3, 3))
        full[:, 0, 0], full[:, 0, 1], full[:, 0, 2] = A[:, 0], A[:, 1], A[:, 2]
        full[:, 1, 1], full[:, 1, 2], full[:, 2, 2] = A[:, 4], A[:, 5], A[:, 8]
        full[:, 1, 0], full[:, 2, 0], full[:, 2, 1] = A[:, 1], A[:, 2], A[:, 5]
        w = np.sort(np.abs(np.linalg.eigvalsh(full)), axis=1)
        well = (w[:, 1] - w[:, 0]) > 1e-4 * w[:, 2]
        assert well.mean() > 0.9 and np.all(cos[well] > 1 - 1e-5)


def test_normal_equations_vs_reference_cxb_golden(ctx, ib, orc, golden_dir):
    c = np.load(os.path.join(golden_dir, "ref_cxb.npz"))
    ctx.set_target(c["Q"]); ctx.set_source(c["P"]); ctx.set_normals(c["normals"])
    idx = ctx.match(ib.DIST_SQRT)
    assert np.array_equal(idx, c["idx"])
    R, T = ctx.minimize(ib.POINT_TO_PLANE)
    mom = ctx.moments(28)
    Cref = c["C"].reshape(6, 6).T                      # column-major 6x6 -> [row, col]
    k = 0
    for r in range(6):
        for col in range(r, 6):
            assert abs(mom[k] - Cref[r, col]) <= 2e-6 * np.abs(Cref).max(), (r, col)
            k += 1
    assert np.abs(mom[21:27] - c["b"]).max() <= 2e-6 * np.abs(c["b"]).max()
    assert mom[27] == 1024
    Cm, b = orc.cxb(c["P"], c["Q"], c["idx"], c["normals"])
    info, oR, oT = orc.plane_rt(Cm, b)
    assert info == 0
    assert np.abs(R - oR).max() < 1e-6 and np.abs(T - oT).max() < 1e-6


def test_full_point_to_plane_run_vs_oracle_and_reference_stdout(ctx, ib, orc, golden_dir):
    D, M = orc.synth_p2p(128)
    ctx.set_target(M); ctx.set_source(D)
    ctx.estimate_normals(4)
    nbr = ctx.neighbors(4)
    assert np.array_equal(nbr, orc.knn(M, 5))
    p = ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT, max_iter=50)
    err, res = ctx.run(p)
    o = orc.icp_p2plane(D, M, orc.normals(M, nbr, 4), max_iter=50)
    assert res.iterations == o["iterations"] and res.iterations_run == o["iterations_run"] == 5
    k = res.iterations_run + 1
    assert np.all(np.abs(err[:k] - o["errors"][:k]) <= 1e-5 * np.abs(o["errors"][:k]) + 2e-7)
    assert np.abs(np.array(res.R[:]) - o["R"]).max() <= 1e-5 and np.abs(np.array(res.t[:]) - o["t"]).max() <= 1e-5
    # the reference binary's own printout on a B200
    ref = [float(x) for x in re.findall(r"Current error \(\d+\): (-?\d+\.\d+)", open(os.path.join(golden_dir, "ref_p2l_stdout.txt")).read())]
    assert len(ref) == res.iterations_run
    assert np.abs(np.array(ref) - err[1:k]).max() <= 1.01e-4


def test_point_to_plane_executable_stdout(golden_dir):
    out = subprocess.run([os.path.join(PKG, "apps", "icp_point_to_plane")], capture_output=True, text=True, check=True).stdout
    ref = open(os.path.join(golden_dir, "ref_p2l_stdout.txt")).read()

    def norm(s):
        s = re.sub(r"calculated in [\d.]+ ms", "calculated in T ms", s)
        return re.sub(r"Elapsed time: [\d.]+ ms", "Elapsed time: T ms", s)
    ol, rl = norm(out).split("\n"), norm(ref).split("\n")
    assert len(ol) == len(rl)
    for a, b in zip(ol, rl):
        if a == b:
            continue
        fa, fb = re.findall(r"-?\d+\.\d+", a), re.findall(r"-?\d+\.\d+", b)
        assert re.sub(r"-?\d+\.\d+", "#", a) == re.sub(r"-?\d+\.\d+", "#", b), (a, b)
        assert all(abs(float(x) - float(y)) <= 1.01e-4 for x, y in zip(fa, fb)), (a, b)


def test_point_to_plane_needs_normals(ctx, ib, orc):
    D, M = orc.synth_p2p(32)
    ctx.set_target(M); ctx.set_source(D)
    with pytest.raises(ib.IcpError, match="normals"):
        ctx.run(ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT))


def test_100k_point_to_plane_properties(ctx, ib, orc):
    """BASELINE.json config 3: 100 000 points, k-NN PCA normals + 6x6 solve."""
    D, M = orc.synth_p2p(317, 100000)
    ctx.set_target(M); ctx.set_source(D)
    ctx.estimate_normals(4)
    nbr = ctx.neighbors(4)
    assert np.array_equal(nbr[:, 0], np.arange(100000))
    sel = np.arange(0, 100000, 97)
    Qd = M.astype(np.float32)
    for i in sel[:200]:
        d = Qd[i] - Qd
        d2 = (d[:, 2] * d[:, 2] + (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])).astype(np.float32)
        order = np.lexsort((np.arange(100000), np.sqrt(d2, dtype=np.float32)))[:5]
        assert list(order) == list(nbr[i]), i
    err, res = ctx.run(ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT, max_iter=50))
    assert res.iterations_run <= 10
    assert np.array_equal(ctx.correspondences(), np.arange(100000))
    assert np.abs(np.array(res.R[:]) - orc.euler_matrix([0.2, -0.2, 0.05])).max() < 2e-5
    assert np.abs(np.array(res.t[:]) - [0.8, -0.3, 0.2]).max() < 2e-5


def test_knn_through_the_pyramid_equals_the_tiled_scan(ib, orc, golden_dir):
    """K5's neighbour lists by best-first descent of the grid's occupancy pyramid (the default since round 2;
    ICPB_KNN_PYRAMID=0 selects the tiled brute-force scan) — same lists as the tiled scan, the reference knn kernel's
    golden output and the oracle, in both ranking modes, on ties / ragged sizes / the sqrt-merge case / clouds spread
    beyond the 10000 cut-off."""
    os.environ["ICPB_KNN_PYRAMID"] = "1"
    try:
        c = ib.Context(0)
    finally:
        del os.environ["ICPB_KNN_PYRAMID"]
    try:
        g = np.load(os.path.join(golden_dir, "ref_knn_normals.npz"))
        for W in (32, 64):
            c.set_target(g["Q_W%d" % W]); c.estimate_normals(4)
            assert np.array_equal(c.neighbors(4), g["nbr_W%d" % W])
        rng = np.random.default_rng(2)
        for m in (5, 9, 257, 4097):
            Q = (rng.integers(-6, 7, size=(m, 3)) * 0.5).astype(np.float32)
            for mode, omode in ((ib.DIST_SQRT, orc.MODE_SQRT), (ib.DIST_SQ, orc.MODE_SQ)):
                c.set_target(Q); c.estimate_normals(4, mode)
                assert np.array_equal(c.neighbors(4), orc.knn(Q, 5, omode)), (m, mode)
        Q = (rng.normal(size=(700, 3)) * 6000.0).astype(np.float32)
        c.set_target(Q); c.estimate_normals(4)
        assert np.array_equal(c.neighbors(4), orc.knn(Q, 5))
        D, M = orc.synth_p2p(317, 100000)
        c.set_target(M); ms = c.estimate_normals(4)
        os.environ["ICPB_KNN_PYRAMID"] = "0"
        try:
            ref = ib.Context(0)
        finally:
            del os.environ["ICPB_KNN_PYRAMID"]
        with ref:
            ref.set_target(M); ms_ref = ref.estimate_normals(4)
            assert np.array_equal(c.neighbors(4), ref.neighbors(4))
        print("normals at 100k points: pyramid %.3f ms, tiled scan %.3f ms" % (ms, ms_ref))
    finally:
        c.close()


def test_same_target_uploaded_again_keeps_its_normals(ib, orc):
    """icpb_set_target with a cloud that is bit for bit the current target keeps what was derived from it (a host-driven
    loop re-uploads its target at every step); one differing bit makes it a new target."""
    D, M = orc.synth_p2p(48)
    with ib.Context(0) as c:
        c.set_target(M); c.set_source(D)
        c.estimate_normals(4)
        p = ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT, max_iter=4, stop_early=0)
        e1, r1 = c.run(p)
        c.set_target(M.copy()); c.set_source(D)                  # identical bits: the normals are still there
        e2, r2 = c.run(p)
        assert np.array_equal(e1, e2) and list(r1.R) == list(r2.R)
        M2 = M.copy(); M2[5, 1] = np.nextafter(M2[5, 1], np.float32(9.0))
        c.set_target(M2); c.set_source(D)                        # a new target: point-to-plane needs its normals first
        with pytest.raises(ib.IcpError):
            c.run(p)
        c.estimate_normals(4)
        c.run(p)
