"""GPU suite (-m gpu): the CUDA engine, called through the C ABI, against the oracle, against the golden
vectors of the reference's own kernels, and — at BASELINE.json's full sizes — through size-independent
properties. Index work is compared bit-exactly; floating point within the tolerance stated in each test.
"""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200")


def _loaded_native():
    return any("libicp_b200.so" in l for l in open("/proc/self/maps"))


def test_native_library_is_what_runs(ctx):
    assert _loaded_native()
    info = ctx.device_info()
    assert info["sm_count"] >= 100 and "B200" in info["name"]


# ------------------------------------------------------------------------------------------------
# matching: bit-exact indices
# ------------------------------------------------------------------------------------------------
MODES = {"p2p": 0, "p2l": 1, "std": 2}


def test_matching_vs_reference_kernel_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_matching.npz"))
    n = 0
    for k in g.files:
        if not k.startswith("match_"):
            continue
        parts = k.split("_")
        if "lattice" in k:
            P, Q = g["P_lattice"], g["Q_lattice"]
        elif "standard_clouds" in k:
            P, Q = g["P_standard"], g["Q_standard"]
        else:
            P, Q = g["P_%s_%s" % (parts[2], parts[3])], g["Q_%s" % parts[2]]
        ctx.set_target(Q); ctx.set_source(P)
        idx = ctx.match(MODES[parts[1]])
        assert np.array_equal(idx, g[k]), k
        n += 1
    assert n == 15


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n,m", [(1, 1), (1, 1500), (33, 7), (2049, 1023), (2048, 1024), (5000, 1025), (4097, 3000)])
def test_matching_ragged_sizes_vs_oracle(ctx, orc, mode, n, m):
    """Sizes that are not multiples of any block/tile; random clouds snapped to a coarse lattice so that
    exact ties (equal distances at different indices) are common."""
    rng = np.random.default_rng(n * 7919 + m)
    Q = (rng.integers(-8, 9, size=(m, 3)) * 0.25).astype(np.float32)
    P = (rng.integers(-16, 17, size=(n, 3)) * 0.125).astype(np.float32)
    ctx.set_target(Q); ctx.set_source(P)
    idx = ctx.match(mode)
    ref = orc.match(P, Q, mode)
    assert np.array_equal(idx, ref)
    # winning distance is the oracle's too
    d = ctx.min_distances()
    diff = P - Q[ref]
    if mode == 0:
        dx, dy, dz = diff[:, 0], diff[:, 1], diff[:, 2]
        assert np.allclose(d, dz * dz + dx * dx + dy * dy, rtol=1e-6)


@pytest.mark.parametrize("mode", [0, 1])
def test_matching_random_float_clouds_vs_oracle(ctx, orc, mode):
    rng = np.random.default_rng(42 + mode)
    P = rng.normal(size=(3000, 3)).astype(np.float32) * 3
    Q = rng.normal(size=(7000, 3)).astype(np.float32) * 3
    Q[5000:5100] = Q[100:200]                     # exact duplicates at higher indices must never win
    ctx.set_target(Q); ctx.set_source(P)
    assert np.array_equal(ctx.match(mode), orc.match(P, Q, mode))


def test_matching_sentinel_keeps_previous_correspondence(ctx, orc):
    Q = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0]], np.float32)
    P = np.array([[0.1, 0, 0], [1000, 0, 0], [0, 1.9, 0]], np.float32)
    ctx.set_target(Q); ctx.set_source(P)
    idx = ctx.match(0)                             # squared: 1e6 > sentinel for source 1 -> stays at its initial 0
    assert list(idx) == [0, 0, 2]
    assert np.array_equal(idx, orc.match(P, Q, 0))
    idx = ctx.match(1)                             # sqrt: 999 < 100000 -> matched
    assert np.array_equal(idx, orc.match(P, Q, 1)) and idx[1] == 1
    idx = ctx.match(0, sentinel=0.5)
    assert np.array_equal(idx, orc.match(P, Q, 0, sentinel=0.5, idx0=orc.match(P, Q, 1)))


def test_matching_sqrt_class_merging(ctx, orc):
    """Squared distances 1 ulp apart whose float square roots coincide: squared mode takes the smaller
    square (higher index here), sqrt mode sees a tie and must take the LOWER index — without ever taking
    a square root in the inner loop. Target 2i sits at (b, e, 10i) with e*e ~ 1 ulp of b*b, target 2i+1 at
    (b, 0, 10i); source i at (0, 0, 10i)."""
    rng = np.random.default_rng(5)
    b = rng.uniform(0.6, 1.9, size=600).astype(np.float32)
    ulp = np.spacing((b * b).astype(np.float32))
    e = np.sqrt(ulp.astype(np.float64) * 1.0).astype(np.float32)
    z = (np.arange(600) * 10.0).astype(np.float32)
    Q = np.zeros((1200, 3), np.float32)
    Q[0::2, 0], Q[0::2, 1], Q[0::2, 2] = b, e, z
    Q[1::2, 0], Q[1::2, 2] = b, z
    P = np.zeros((600, 3), np.float32)
    P[:, 2] = z
    ctx.set_target(Q); ctx.set_source(P)
    sq, sr = ctx.match(0), ctx.match(1)
    assert np.array_equal(sq, orc.match(P, Q, 0))
    assert np.array_equal(sr, orc.match(P, Q, 1))
    assert (sq != sr).sum() > 50, "the case must actually exercise the merge"
    assert np.all(sr[sq != sr] % 2 == 0) and np.all(sq[sq != sr] % 2 == 1)


def test_matching_100k_slice_vs_oracle(ctx, orc):
    """BASELINE.json config 2 size (100 000 x 100 000): every source is matched on the GPU; a 4096-source
    slice (spread over the cloud) is checked against the oracle's full scan."""
    D, M = orc.synth_p2p(317, 100000)
    P = orc.icp_p2p(D[::25], M[::25], max_iter=3, stop_early=False)      # a plausible intermediate pose
    Pm = orc.transform(D, np.asarray(P["R"], np.float32), np.asarray(P["t"], np.float32))
    ctx.set_target(M); ctx.set_source(Pm)
    for mode in (0, 1):
        idx = ctx.match(mode)
        sel = np.arange(0, 100000, 25)[:4096]
        ref = orc.match(Pm[sel], M, mode)
        assert np.array_equal(idx[sel], ref)
        assert idx.min() >= 0 and idx.max() < 100000


# ------------------------------------------------------------------------------------------------
# minimisation / transform / error: tolerance 1e-6 relative (FP64 moments on both sides)
# ------------------------------------------------------------------------------------------------
def test_centroid_known_answer_on_gpu(ctx):
    """src/tests/centroid.cu:68-73"""
    P = np.tile(np.array([1, 2, 3], np.float32), (2048, 1))
    ctx.set_target(P); ctx.set_source(P)
    ctx.match(0)
    ctx.minimize(0)
    mom = ctx.moments(16)
    assert np.array_equal(mom[:6], [2048, 4096, 6144, 2048, 4096, 6144]) and mom[15] == 2048


def test_step_by_step_vs_oracle(ctx, orc, golden_dir):
    D, M = orc.synth_p2p(64)
    ctx.set_target(M); ctx.set_source(D)
    P = D.copy()
    for it in range(4):
        idx = ctx.match(0)
        ref_idx = orc.match(P, M, 0)
        assert np.array_equal(idx, ref_idx)
        R, T = ctx.minimize(0)
        mom = ctx.moments(16)
        omom = orc.moments(P, M, ref_idx)
        assert np.allclose(mom, omom, rtol=1e-12, atol=1e-9)
        oR, oT = orc.rt_from_moments(omom)
        assert np.abs(R - oR).max() < 2e-7 and np.abs(T - oT).max() < 2e-7
        rms = ctx.transform()
        P = orc.transform(P, R, T)                 # same float arithmetic as RyT: bit-exact
        got = ctx.get_source()
        assert np.array_equal(got.view(np.uint32), P.view(np.uint32))
        assert abs(rms - orc.rms(P, M, ref_idx)) <= 1e-6 * max(rms, 1e-3)
    # and against the reference's RyT kernel itself (tests/golden/ref_ryt.npz = its output on a B200): K4 in the loop
    # (transform_kernel, with R,T injected) and the stand-alone entry point must both reproduce it bit for bit
    r = np.load(os.path.join(golden_dir, "ref_ryt.npz"))
    ctx.set_target(r["P"]); ctx.set_source(r["P"])
    assert np.array_equal(ctx.match(0), np.arange(r["P"].shape[0]))
    ctx.set_transform(r["R"], r["T"])
    rms = ctx.transform()
    got = ctx.get_source()
    assert np.array_equal(got.view(np.uint32), r["out"].view(np.uint32))
    assert abs(rms - orc.rms(r["out"], r["P"], np.arange(r["P"].shape[0], dtype=np.int32))) <= 1e-6 * max(rms, 1e-3)
    assert np.array_equal(ctx.apply_transform(r["R"], r["T"], r["P"]).view(np.uint32), r["out"].view(np.uint32))


@pytest.mark.parametrize("W,expect_run", [(32, 15), (128, 27)])
def test_full_p2p_run_vs_oracle(ctx, ib, orc, W, expect_run):
    D, M = orc.synth_p2p(W)
    ctx.set_target(M); ctx.set_source(D)
    err, res = ctx.run(ib.default_params(max_iter=40))
    o = orc.icp_p2p(D, M, max_iter=40)
    assert res.iterations == o["iterations"] and res.iterations_run == o["iterations_run"] == expect_run
    k = res.iterations + 2
    # tolerance: 1e-5 relative (north star), floor 1e-7 absolute (float noise of the residual itself)
    assert np.all(np.abs(err[:k] - o["errors"][:k]) <= 1e-5 * np.abs(o["errors"][:k]) + 1e-7)
    assert np.abs(np.array(res.R[:]) - o["R"]).max() <= 1e-5
    assert np.abs(np.array(res.t[:]) - o["t"]).max() <= 1e-5
    assert np.array_equal(ctx.correspondences(), np.arange(W * W))
    # ground truth: the generating pose
    assert np.abs(np.array(res.R[:]) - orc.euler_matrix([0.2, -0.2, 0.05])).max() < 1e-5
    assert np.abs(np.array(res.t[:]) - [0.8, -0.3, 0.2]).max() < 1e-5
    # sync cadence must not change anything
    ctx.set_source(D)
    err2, res2 = ctx.run(ib.default_params(max_iter=40, sync_every=7))
    assert np.array_equal(err, err2) and res2.iterations == res.iterations and list(res2.R) == list(res.R)


def test_standard_program_run_vs_oracle(ctx, ib, orc):
    D, M = orc.synth_standard(32)
    ctx.set_target(M); ctx.set_source(D)
    err, res = ctx.run(ib.default_params(dist_mode=ib.DIST_STD, max_iter=40, stop_early=0, sync_every=40))
    o = orc.icp_p2p(D, M, mode=orc.MODE_STD, max_iter=40, stop_early=False)
    assert res.iterations == 40 and res.iterations_run == 40
    assert np.all(np.abs(err - o["errors"]) <= 1e-5 * np.abs(o["errors"]) + 1e-7)


def test_icp_cpu_config0_vs_reference_cpu_golden(ctx, ib, orc, golden_dir):
    """BASELINE.json config 0: src/ICP_CPU.c's own clouds (cast to float) through the GPU engine with its
    MAX_ITER 200 / tol 1e-5. The CPU program works in double and stops on |dE| < 1e-5 in a slowly creeping
    local minimum, so the comparison is on the common prefix of the printed trajectory: 1e-4 absolute on
    values ~0.83 (float clouds + float transform arithmetic accumulate over 60 iterations)."""
    gold = [float(x) for x in re.findall(r"-?\d+\.\d+", open(os.path.join(golden_dir, "icp_cpu_stdout.txt")).read().split("\n")[1])]
    Dd, Md = orc.synth_cpu_f64(100)
    D = np.ascontiguousarray(Dd.reshape(3, -1).T, np.float32)
    M = np.ascontiguousarray(Md.reshape(3, -1).T, np.float32)
    ctx.set_target(M); ctx.set_source(D)
    err, res = ctx.run(ib.default_params(max_iter=200, tol=1e-5))
    k = min(len(gold), res.iterations + 1)
    assert k >= 40
    assert np.abs(err[:k] - np.array(gold[:k])).max() < 1e-4
    assert abs(res.iterations - 61) <= 12


def test_executables_stdout(golden_dir):
    """The drop-in programs print what the reference programs print (time line aside). Point-to-point: same
    lines, values within 1 unit of the 4th printed decimal, at most one extra/missing final line (the
    reference's float cuBLAS/cuSOLVER noise floor sits at its 1e-6 stop threshold)."""
    out = subprocess.run([os.path.join(PKG, "apps", "icp_point_to_point")], capture_output=True, text=True, check=True).stdout
    ref = open(os.path.join(golden_dir, "ref_p2p_stdout.txt")).read()
    ol, rl = out.split("\n"), ref.split("\n")
    assert ol[0] == rl[0] == "Grid Size: 16, Block Size: 1024" and ol[1] == rl[1] == "Error:"
    ov = [l for l in ol if re.match(r"^\d+: ", l)]
    rv = [l for l in rl if re.match(r"^\d+: ", l)]
    assert abs(len(ov) - len(rv)) <= 1
    for a, b in zip(ov, rv):
        assert a.split(":")[0] == b.split(":")[0]
        assert abs(float(a.split()[1]) - float(b.split()[1])) <= 1.01e-4
    assert "ICP converged successfully!\n\nElapsed time: " in out and out.endswith(" ms\n")
    out = subprocess.run([os.path.join(PKG, "apps", "icp_standard")], capture_output=True, text=True, check=True).stdout
    ol = out.split("\n")
    assert ol[0] == "Grid Size: 8, Block Size: 128" and ol[1] == "Error:"
    vals = ol[2].split()
    assert len(vals) == 40 and vals[:3] == ["1.0061", "0.9449", "0.9127"] and vals[33] == "0.0000"
    assert ol[3].startswith("Elapsed time: ")


# ------------------------------------------------------------------------------------------------
# the reference's kernels, run side by side on the GPU box (oracle/_ref travels with the snapshot)
# ------------------------------------------------------------------------------------------------
def test_live_reference_matching_kernel(ctx, orc, tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_kernels")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ref_kernels not built (needs /root/reference at build time)")
    rng = np.random.default_rng(11)
    n, m = 4096, 8192
    P = (rng.normal(size=(n, 3)) * 2).astype(np.float32)
    Q = np.round(rng.normal(size=(m, 3)) * 2, 1).astype(np.float32)
    P.tofile(tmp_path / "P.bin"); Q.tofile(tmp_path / "Q.bin")
    ctx.set_target(Q); ctx.set_source(P)
    for which, mode in (("p2p", 0), ("p2l", 1)):
        subprocess.run([ref, "match", which, str(n), str(m), str(tmp_path / "P.bin"), str(tmp_path / "Q.bin"), str(tmp_path / "idx.bin")], check=True)
        want = np.fromfile(tmp_path / "idx.bin", np.int32)
        assert np.array_equal(ctx.match(mode), want)


# ------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs 2 and 4)
# ------------------------------------------------------------------------------------------------
def test_100k_registration_properties(ctx, ib, orc):
    D, M = orc.synth_p2p(317, 100000)
    ctx.set_target(M); ctx.set_source(D)
    err, res = ctx.run(ib.default_params(max_iter=64))
    assert res.iterations_run < 64
    assert np.array_equal(ctx.correspondences(), np.arange(100000))
    assert np.abs(np.array(res.R[:]) - orc.euler_matrix([0.2, -0.2, 0.05])).max() < 1e-5
    assert np.abs(np.array(res.t[:]) - [0.8, -0.3, 0.2]).max() < 1e-5
    e = err[1: res.iterations_run + 1]
    assert np.all(np.diff(e) < 1e-6), "the RMS error never increases"
    assert e[-1] < 1e-5
    # idempotence: registering the registered cloud is a fixed point (identity transform, same matches)
    err2, res2 = ctx.run(ib.default_params(max_iter=5))
    assert res2.iterations_run <= 2
    assert np.abs(np.array(res2.R[:]) - np.eye(3).reshape(-1)).max() < 1e-6


def test_1m_single_match_sampled_vs_oracle(ctx, orc):
    """BASELINE.json config 4 size: one 1M x 1M matching pass; 512 sampled sources checked against the
    oracle's full scan, and the winning distance of EVERY source re-derived from its index."""
    D, M = orc.synth_p2p(1000)
    ctx.set_target(M); ctx.set_source(D)
    idx = ctx.match(0)
    d = ctx.min_distances()
    sel = np.random.default_rng(3).choice(1000000, 512, replace=False)
    assert np.array_equal(idx[sel], orc.match(D[sel], M, 0))
    diff = (D - M[idx]).astype(np.float32)
    chain = diff[:, 2] * diff[:, 2] + (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1])
    assert np.allclose(d, chain, rtol=2e-6, atol=1e-12)


def test_reflection_fix_flag(ctx, ib, orc):
    """SURVEY.md §8 f-4. Sources = small clusters strung along y (so that matching is the identity), targets =
    the same clusters mirrored in x: the reference's R = U*V^T is a reflection (det = -1); with
    ICPB_FLAG_FIX_REFLECTION the engine returns the best proper rotation (Kabsch), checked against numpy."""
    rng = np.random.default_rng(12)
    off = (rng.normal(size=(400, 3)) * [1e-3, 1e-3, 5e-4]).astype(np.float32)     # well inside the 0.05 spacing
    line = (np.arange(400, dtype=np.float32) * np.float32(0.05))[:, None] * np.array([0, 1, 0], np.float32)
    P = (off + line).astype(np.float32)
    Q = (off * np.array([-1, 1, 1], np.float32) + line).astype(np.float32)
    ctx.set_target(Q)
    out = {}
    for flags in (0, ib.FLAG_FIX_REFLECTION):
        ctx.set_source(P)
        err, res = ctx.run(ib.default_params(max_iter=1, stop_early=0, flags=flags))
        assert np.array_equal(ctx.correspondences(), np.arange(400))
        out[flags] = np.array(res.R[:]).reshape(3, 3).T
    assert np.linalg.det(out[0]) < -0.99, "the reference keeps the reflection"
    assert np.linalg.det(out[ib.FLAG_FIX_REFLECTION]) > 0.99
    pc, qc = P.astype(np.float64) - P.astype(np.float64).mean(0), Q.astype(np.float64) - Q.astype(np.float64).mean(0)
    U, S, Vt = np.linalg.svd(qc.T @ pc)
    assert np.abs(out[0] - U @ Vt).max() < 1e-5
    assert np.abs(out[ib.FLAG_FIX_REFLECTION] - U @ np.diag([1, 1, -1]) @ Vt).max() < 1e-5


@pytest.mark.parametrize("metric", [0, 1])
def test_graph_replay_equals_plain_loop(ctx, ib, orc, metric):
    """ICPB_FLAG_GRAPH replays batches of iterations from a CUDA graph; the default is the plain loop.
    Same errors, iteration counts and transform, bit for bit; repeated runs reuse the instantiated graph."""
    D, M = orc.synth_p2p(128)
    ctx.set_target(M)
    mode = ib.DIST_SQRT if metric else ib.DIST_SQ
    if metric:
        ctx.estimate_normals(4)
    out = []
    G = ib.FLAG_GRAPH
    for flags, sync in ((G, 0), (G, 0), (0, 0), (G, 3), (G | ib.FLAG_PROFILE, 1)):
        ctx.set_source(D)
        l0 = ctx.launch_count()
        e, r = ctx.run(ib.default_params(metric=metric, dist_mode=mode, max_iter=50, flags=flags, sync_every=sync))
        out.append((e.copy(), r.iterations, r.iterations_run, list(r.R), list(r.t), r.match_ms, ctx.launch_count() - l0))
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and o[1:5] == out[0][1:5]
    assert out[0][5] == 0.0 and out[2][5] > 0.0, "match_ms is only measured in the plain loop"
    assert out[0][2] == (5 if metric else 27)


def test_1m_full_registration_properties(ctx, ib, orc):
    """BASELINE.json config 4 on one GPU: the whole 1M x 1M registration (45 iterations, more than the reference's
    MAX_ITER 40): identity correspondences, ground-truth pose, monotone error."""
    import icp_synth
    D, M = icp_synth.p2p_clouds(1000)
    ctx.set_target(M); ctx.set_source(D)
    err, res = ctx.run(ib.default_params(max_iter=64))
    assert res.iterations_run == 45
    assert np.array_equal(ctx.correspondences(), np.arange(1000000))
    assert np.abs(np.array(res.R[:]) - orc.euler_matrix([0.2, -0.2, 0.05])).max() < 1e-5
    assert np.abs(np.array(res.t[:]) - [0.8, -0.3, 0.2]).max() < 1e-5
    e = err[1: res.iterations_run + 1]
    assert np.all(np.diff(e) < 1e-6) and e[-1] < 1e-5


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("nn", ["brute", "grid"])
def test_reject_unmatched_equals_registration_of_the_inliers(ib, orc, metric, nn):
    """ICPB_FLAG_REJECT_UNMATCHED (the reference's LiDAR programs only raise the sentinel,
    src/CUDA/GPU_point_to_point_real.cu:18,44): sources with no target below the sentinel are dropped from the moments
    and from the RMS and report correspondence -1. With outliers that stay beyond the sentinel for the whole run, the
    registration must be the registration of the inliers alone: same iteration count, same trajectory and transform (to
    FP64 summation-order noise), same correspondences. Without the flag the same outliers keep correspondence 0 and
    drag the transform (the reference's behaviour), which the test also checks."""
    D, M = orc.synth_p2p(96)
    rng = np.random.default_rng(31)
    out_at = np.sort(rng.choice(D.shape[0], 300, replace=False))
    P = D.copy()
    P[out_at] += np.float32(60.0) + rng.random((300, 3)).astype(np.float32)          # far from everything, forever
    keep = np.setdiff1d(np.arange(D.shape[0]), out_at)
    mode = ib.DIST_SQRT if metric else ib.DIST_SQ
    sentinel = 8.0 if metric else 64.0                                                # same radius in both distance domains
    nnm = ib.NN_GRID if nn == "grid" else ib.NN_BRUTE
    base = dict(metric=metric, dist_mode=mode, nn_method=nnm, max_iter=60, sentinel=sentinel)
    with ib.Context(0) as ctx:
        ctx.set_target(M)
        if metric:
            ctx.estimate_normals(4)
        ctx.set_source(P)
        e_r, r_r = ctx.run(ib.default_params(flags=ib.FLAG_REJECT_UNMATCHED, **base))
        idx_r = ctx.correspondences()
        ctx.set_source(np.ascontiguousarray(P[keep]))
        e_i, r_i = ctx.run(ib.default_params(**base))
        idx_i = ctx.correspondences()
        if metric == 0:
            ctx.set_source(P)
            e_n, r_n = ctx.run(ib.default_params(**base))
            idx_n = ctx.correspondences()
    assert (r_r.iterations, r_r.iterations_run) == (r_i.iterations, r_i.iterations_run)
    k = r_i.iterations + 2
    assert np.all(np.abs(e_r[:k] - e_i[:k]) <= 1e-6 * np.abs(e_i[:k]) + 1e-7)
    assert np.abs(np.array(r_r.R[:]) - np.array(r_i.R[:])).max() < 1e-8 and np.abs(np.array(r_r.t[:]) - np.array(r_i.t[:])).max() < 1e-8
    assert np.array_equal(idx_r[keep], idx_i) and np.all(idx_r[out_at] == -1)
    # the reference's rule: unmatched sources keep their previous correspondence (0 after the upload) and stay in the sums
    if metric == 0:
        assert np.all(idx_n[out_at] == 0)
        assert np.abs(np.array(r_n.R[:]) - np.array(r_i.R[:])).max() > 1e-3
