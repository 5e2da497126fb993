"""GPU suite (-m gpu): dataset front ends and dataset programs through the C ABI and the
drop-in executables, against the golden outputs of the reference's own kernels / programs and against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200")
REFDIR = os.path.join(ROOT, "oracle", "_ref")

need_bunny = pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "Bunny_res.csv")), reason="oracle/_ref/Bunny_res.csv not built")
need_lidar = pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "Donut_1024x16.csv")), reason="oracle/_ref/Donut_1024x16.csv not built")


def parse_errors(text):
    m = re.search(r"Error:\n((?:\d+: -?[\d.]+\n)+)", text)
    return np.array([float(l.split(":")[1]) for l in m.group(1).strip().splitlines()])


def run_app(name, *args):
    r = subprocess.run([os.path.join(PKG, "apps", name), "--data", REFDIR] + list(args), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_lidar_convert_bit_exact_vs_reference_kernel(ctx, golden_dir):
    """Same device cosf/sinf, same operation order as Conversion<<<>>>: the clouds must be bitwise the reference's."""
    g = np.load(os.path.join(golden_dir, "ref_lidar.npz"))
    P, ms = ctx.lidar_convert(g["ranges"], int(g["encoder_count"]), g["altitude"], g["azimuth"])
    assert np.array_equal(P.view(np.uint32), g["P_mm"].view(np.uint32))
    assert ms >= 0


def test_lidar_convert_ragged_and_wraparound(ctx, orc):
    """n not a multiple of the beam count or of the block size; encoder count wrapping past 90112; other beam counts."""
    rng = np.random.default_rng(3)
    for n, beams, enc in ((1, 16, 0), (1000, 16, 90000), (4097, 64, 45055), (16385, 16, 90111)):
        r = rng.integers(0, 2 ** 20, size=n).astype(np.float32)
        alt = np.linspace(16.6, -16.6, beams).astype(np.float32); az = np.linspace(3.1, -3.1, beams).astype(np.float32)
        P, _ = ctx.lidar_convert(r, enc, alt, az)
        if beams == 16:
            ref = orc.lidar_convert(r, enc, alt, az)
            assert (np.abs(P - ref) <= 2e-6 * np.maximum(r, 1.0)[:, None]).all()
        assert np.allclose(np.linalg.norm(P.astype(np.float64), axis=1), r, rtol=1e-6, atol=1e-3)


def test_apply_transform_and_scale_bit_exact(ctx, orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_lidar.npz"))
    R = orc.euler_matrix([0.01, -0.003, 0.05]); T = np.array([0.001, -0.0202, 0.02], np.float32)
    Q = ctx.apply_transform(R, T, g["P_mm"])
    assert np.array_equal(Q.view(np.uint32), g["Q_mm"].view(np.uint32))
    a = np.float32(1.0 / 1000.0)
    assert np.array_equal(ctx.scale_cloud(a, g["P_mm"]).view(np.uint32), (g["P_mm"] * a).view(np.uint32))
    # ragged size
    X = np.random.default_rng(1).standard_normal((1001, 3)).astype(np.float32)
    assert np.array_equal(ctx.apply_transform(R, T, X).view(np.uint32), orc.transform(X, R, T).view(np.uint32))


def test_lidar_first_matching_pass_vs_reference_kernel(ctx, ib, golden_dir):
    """Sentinel 1e6, 27 % of the points collapsed onto the origin (massive exact ties): lowest index must win."""
    g = np.load(os.path.join(golden_dir, "ref_lidar.npz"))
    a = np.float32(1.0 / 1000.0)
    ctx.set_target(g["Q_mm"] * a); ctx.set_source(g["P_mm"] * a)
    for nn in (ib.NN_BRUTE, ib.NN_BRUTE_DIRECT, ib.NN_GRID):
        assert np.array_equal(ctx.match(ib.DIST_SQ, nn, 1e6), g["idx_first"]), "nn_method %d" % nn


@need_bunny
def test_bunny_knn_sq_and_matching_vs_reference_kernels(ctx, ib, orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_bunny.npz"))
    D, M = orc.bunny_clouds(REFDIR)
    ctx.set_target(M); ctx.set_source(D)
    ctx.estimate_normals(4, ib.DIST_SQ)
    assert np.array_equal(ctx.neighbors(4), g["nbr"])
    for nn in (ib.NN_BRUTE, ib.NN_BRUTE_DIRECT, ib.NN_GRID):
        assert np.array_equal(ctx.match(ib.DIST_SQ, nn, 100000.0), g["idx_first"]), "nn_method %d" % nn
    # normals against the oracle's (sign free)
    n_gpu = ctx.normals().astype(np.float64); n_orc = orc.normals(M, g["nbr"], 4).astype(np.float64)
    cos = np.abs((n_gpu * n_orc).sum(axis=1))
    assert np.mean(cos > 1 - 1e-6) > 0.99


@pytest.mark.parametrize("m", [5, 9, 257, 4097])
def test_knn_sq_ties_vs_oracle(ctx, ib, orc, m):
    rng = np.random.default_rng(m)
    Q = (rng.integers(-6, 7, size=(m, 3)) * 0.5).astype(np.float32)
    ctx.set_target(Q)
    ctx.estimate_normals(4, ib.DIST_SQ)
    assert np.array_equal(ctx.neighbors(4), orc.knn(Q, 5, orc.MODE_SQ))


@need_bunny
def test_bunny_programs_vs_reference_stdout_and_oracle(orc, golden_dir):
    D, M = orc.bunny_clouds(REFDIR)
    for app, gold, plane in (("icp_bunny_point_to_point", "ref_bunny_p2p_stdout.txt", False), ("icp_bunny_point_to_plane", "ref_bunny_p2l_stdout.txt", True)):
        out = run_app(app)
        ref_text = open(os.path.join(golden_dir, gold)).read()
        got, ref = parse_errors(out), parse_errors(ref_text)
        k = min(len(got), len(ref))
        assert abs(len(got) - len(ref)) <= 1, (app, len(got), len(ref))
        assert np.abs(got[:k] - ref[:k]).max() <= 1.01e-4, app
        # same banner lines as the reference (timings differ)
        for line in ref_text.splitlines():
            if line.startswith("Grid Size") or line.startswith("For "):
                assert line in out.splitlines(), (app, line)
        for phase in ("matching", "minimization", "transformation", "error estimation"):
            assert re.search(r"The %s step represents the [\d.]+%% of the total time with [\d.]+ ms" % phase, out), (app, phase)
        # and exactly the oracle's trajectory
        if plane:
            o = orc.icp_p2plane(D, M, orc.normals(M, orc.knn(M, 5, orc.MODE_SQ), 4), max_iter=40, mode=orc.MODE_SQ)
        else:
            o = orc.icp_p2p(D, M, max_iter=40)
        assert len(got) == o["iterations"] + 1
        assert np.abs(got - np.round(o["errors"][:len(got)].astype(np.float64), 4)).max() <= 1.01e-4


@need_lidar
def test_lidar_programs_vs_reference_stdout(golden_dir):
    for app, gold in (("icp_lidar_point_to_point", "ref_lidar_p2p_stdout.txt"), ("icp_lidar_point_to_plane", "ref_lidar_p2l_stdout.txt")):
        out = run_app(app)
        ref_text = open(os.path.join(golden_dir, gold)).read()
        got, ref = parse_errors(out), parse_errors(ref_text)
        k = min(len(got), len(ref))
        assert abs(len(got) - len(ref)) <= 1, (app, len(got), len(ref))
        assert np.abs(got[:k] - ref[:k]).max() <= 1.01e-4, app
        assert "Conversion kernel's elapsed time" in out
        assert ("RyT kernel's elapsed time" in out) == ("RyT kernel's elapsed time" in ref_text)


@need_lidar
def test_lidar_registration_recovers_the_pose(ctx, ib, orc):
    """The target is the scan moved by a known pose (in the millimetre frame, then scaled): the accumulated transform
    must be that pose — rotation r = (0.01,-0.003,0.05), translation t/1000."""
    P, Q, _ = orc.lidar_clouds(REFDIR)
    ctx.set_target(Q); ctx.set_source(P)
    err, res = ctx.run(ib.default_params(max_iter=100, sentinel=1e6))
    R_true = orc.euler_matrix([0.01, -0.003, 0.05]).astype(np.float64)
    assert np.abs(np.array(res.R[:]) - R_true).max() < 1e-4
    assert np.abs(np.array(res.t[:]) - np.array([0.001, -0.0202, 0.02]) / 1000.0).max() < 1e-4
    assert err[res.iterations + 1] < 1e-4
