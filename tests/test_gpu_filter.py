"""GPU suite (-m gpu): the lower-bound-filter matching kernel (ICPB_NN_BRUTE) against the direct kernel
(ICPB_NN_BRUTE_DIRECT) and the oracle. The filter only decides which 128-target sub-tiles may be skipped; the
indices and winning distances must be identical bit for bit, whatever the warm start."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _context(ib, **env):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return ib.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


# every test runs with the bound chosen automatically (full on cold passes, planar on warm ones, switched by the
# measured exact-pass rate), with the planar bound forced for each choice of the dropped axis, and with the full one
@pytest.fixture(scope="module", params=["tc-auto", "auto", "planar-x", "planar-y", "planar-z", "full", "tc", "tc-split", "tc-split16", "tc-split2", "tc-pair", "tc-pair1", "tc-quad", "tc-oct", "tc-hex",
                                        "tc-u8x16", "tc-u16x16", "tc-u8x8", "tc-u8x2", "tc-d16", "tc-d8", "tc-d4", "tc-d2",
                                        "tc-sorted", "tc-sorted16", "tc-sorted2"])
def ctx(ib, request):
    """A context that sends EVERY brute-force pass through the filter kernel (by default passes below 1e9 pairs
    use the direct kernel, which would make most of these small cases vacuous)."""
    env = {"ICPB_K1_FILTER_MIN_PAIRS": 0, "ICPB_K1_TC": 0}       # K1F variants: the FP32 filter kernel
    if request.param == "tc-auto":                 # the default: K1T, group size chosen by the exact-pass-rate policy
        env.update(ICPB_K1_TC=1)
    elif request.param.startswith("planar"):
        env.update(ICPB_KF_DIMS=2, ICPB_KF_DROP="xyz".index(request.param[-1]))
    elif request.param == "full":
        env.update(ICPB_KF_DIMS=3)
    elif request.param.startswith("tc-sorted"):    # K1T over the targets in Morton order, whatever their scan order (policy-chosen, 16 and 2 per column)
        env.update(ICPB_K1_TC=1, ICPB_KT_SORT=1)
        if request.param != "tc-sorted":
            env.update(ICPB_KT_VAR={"tc-sorted16": 14, "tc-sorted2": 20}[request.param])
    elif request.param.startswith("tc"):           # K1T: the 3-D bound evaluated by tcgen05.mma kind::tf32 (csrc/nn_filter_tc.cu)
        env.update(ICPB_K1_TC=1, ICPB_KT_VAR={"tc-auto": -1, "tc": 0, "tc-split": 5, "tc-split16": 6, "tc-split2": 7, "tc-pair": 8, "tc-pair1": 9, "tc-quad": 10, "tc-oct": 11, "tc-hex": 12,
                                                  # exact-pass units of 8 / 16 columns at 16 / 8 / 2 targets per column; d*: two 128-column accumulators per group
                                                  "tc-u16x16": 13, "tc-u8x16": 14, "tc-u8x8": 16, "tc-u8x2": 20, "tc-d16": 21, "tc-d8": 22, "tc-d4": 23, "tc-d2": 24}[request.param])
    c = _context(ib, **env)
    c.variant = request.param
    yield c
    c.close()


def _check(ctx, ib, orc, P, Q, modes=(0, 1), sentinel=100000.0, oracle=True, expect_filter=True):
    ctx.set_target(Q); ctx.set_source(P)
    for mode in modes:
        for rep in range(2):                      # second pass is warm-started from the first pass's indices
            s0 = ctx.filter_stats()["subtile_tests"]
            a = ctx.match(mode, ib.NN_BRUTE, sentinel); da = ctx.min_distances()
            assert (ctx.filter_stats()["subtile_tests"] > s0) == expect_filter, "which kernel ran is part of the test"
            if expect_filter:
                cfg = ctx.filter_config()
                want = {"planar": 2, "full": 3, "tc": 4}.get(ctx.variant.split("-")[0])
                if ctx.variant == "auto":
                    assert cfg["dims_last"] in (2, 3)
                assert want is None or cfg["dims_last"] == want, (ctx.variant, cfg)
                if ctx.variant.startswith("planar"):
                    assert cfg["drop_axis"] == "xyz".index(ctx.variant[-1])
            b = ctx.match(mode, ib.NN_BRUTE_DIRECT, sentinel); db = ctx.min_distances()
            assert np.array_equal(a, b), (mode, rep)
            assert np.array_equal(da.view(np.uint32), db.view(np.uint32)), (mode, rep)
        if oracle:
            assert np.array_equal(a, orc.match(P, Q, mode, sentinel))


def test_filter_equals_direct_on_registration_stages(ctx, ib, orc):
    D, M = orc.synth_p2p(100)
    for iters in (0, 1, 4, 9, 20):
        P = D if iters == 0 else orc.icp_p2p(D, M, max_iter=iters, stop_early=False)["P"]
        _check(ctx, ib, orc, P, M)
    st = ctx.filter_stats()
    assert st["subtile_tests"] > 0 and 0 < st["subtile_exact"] < st["subtile_tests"]


def test_filter_with_adversarial_warm_start(ctx, ib, orc):
    """The seed index may be anything: best possible, worst possible, stale from another cloud."""
    rng = np.random.default_rng(0)
    Q = rng.normal(size=(5000, 3)).astype(np.float32)
    P = rng.normal(size=(3000, 3)).astype(np.float32)
    ctx.set_target(Q); ctx.set_source(P)
    ref = orc.match(P, Q, 0)
    assert np.array_equal(ctx.match(0, ib.NN_BRUTE), ref)          # seeds = 0 everywhere
    assert np.array_equal(ctx.match(0, ib.NN_BRUTE), ref)          # seeds = exact answer
    P2 = (P[::-1] * 1.7).astype(np.float32)                        # same buffers, now stale seeds
    ctx.set_target(Q); ctx.set_source(P2)
    assert np.array_equal(ctx.match(1, ib.NN_BRUTE), orc.match(P2, Q, 1))


def test_filter_ties_lattice_and_duplicates(ctx, ib, orc):
    rng = np.random.default_rng(21)
    Q = (rng.integers(-8, 9, size=(9000, 3)) * 0.25).astype(np.float32)
    Q[7000:7500] = Q[200:700]
    P = (rng.integers(-16, 17, size=(4100, 3)) * 0.125).astype(np.float32)
    _check(ctx, ib, orc, P, Q)


def test_filter_far_offsets_and_scales(ctx, ib, orc):
    """Clouds far from the origin (centering matters), tiny and huge scales, a source far outside the target."""
    rng = np.random.default_rng(8)
    base_q = rng.normal(size=(3000, 3)).astype(np.float32)
    base_p = rng.normal(size=(1500, 3)).astype(np.float32)
    for shift, scale in ((1000.0, 1.0), (0.0, 1e-4), (-5e4, 30.0), (3.0, 1e5)):
        Q = (base_q * np.float32(scale) + np.float32(shift)).astype(np.float32)
        P = (base_p * np.float32(scale) + np.float32(shift)).astype(np.float32)
        P[:10] += np.float32(50 * scale)
        _check(ctx, ib, orc, P, Q, sentinel=3e38)


def test_filter_denormal_squares_use_the_direct_kernel(ctx, ib, orc):
    rng = np.random.default_rng(9)
    Q = (rng.normal(size=(1200, 3)) * 1e-21).astype(np.float32)
    P = (rng.normal(size=(500, 3)) * 1e-21).astype(np.float32)
    _check(ctx, ib, orc, P, Q, expect_filter=False)


def test_filter_sentinel_and_unmatched(ctx, ib, orc):
    rng = np.random.default_rng(4)
    Q = rng.normal(size=(2000, 3)).astype(np.float32)
    P = (rng.normal(size=(700, 3)) * 3).astype(np.float32)
    for s in (0.02, 0.5, 4.0):
        _check(ctx, ib, orc, P, Q, sentinel=s, oracle=False)
        ctx.set_target(Q); ctx.set_source(P)
        a = ctx.match(0, ib.NN_BRUTE, s)
        assert np.array_equal(a, orc.match(P, Q, 0, s, idx0=np.zeros(700, np.int32)))


def test_filter_non_finite_inputs_fall_back_consistently(ctx, ib, orc):
    rng = np.random.default_rng(6)
    Q = rng.normal(size=(1500, 3)).astype(np.float32)
    P = rng.normal(size=(600, 3)).astype(np.float32)
    P[5] = np.nan; P[17, 1] = np.inf
    _check(ctx, ib, orc, P, Q)                      # NaN/inf sources: never matched, by either kernel or the reference
    Q2 = Q.copy(); Q2[3] = np.inf; Q2[40, 2] = np.nan
    _check(ctx, ib, orc, P, Q2, expect_filter=False)  # non-finite targets: the filter steps aside (direct kernel)
    Q3 = (Q * np.float32(1e17)).astype(np.float32)
    _check(ctx, ib, orc, (P * np.float32(1e17)).astype(np.float32), Q3, sentinel=3e38, oracle=False, expect_filter=False)


def test_full_run_filter_is_bitwise_the_direct_run(ctx, ib, orc):
    D, M = orc.synth_p2p(317, 100000)
    ctx.set_target(M)
    out = []
    for nn in (ib.NN_BRUTE, ib.NN_BRUTE_DIRECT):
        ctx.set_source(D)
        e, r = ctx.run(ib.default_params(max_iter=64, nn_method=nn))
        out.append((e.copy(), r.iterations, list(r.R), list(r.t), ctx.correspondences(), r.match_ms))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1:4] == out[1][1:4]
    assert np.array_equal(out[0][4], out[1][4])


def test_planar_bound_policy(ib, orc):
    """Automatic choice: on the 1M-point saddle the planar bound stays in use once the pass is warm (few sub-tiles
    reach the exact chain); on a small volumetric cloud whose thresholds are of the order of the point spacing its
    exact-pass rate is high and the engine goes back to the full bound. Same indices either way."""
    c = _context(ib, ICPB_K1_FILTER_MIN_PAIRS=0, ICPB_K1_TC=0)
    try:
        D, M = orc.synth_p2p(1000)
        c.set_target(M); c.set_source(D)
        e, r = c.run(ib.default_params(max_iter=4, stop_early=0, sync_every=1))
        cfg = c.filter_config()
        assert cfg["drop_axis"] in (0, 1, 2) and cfg["dims_last"] == 2 and cfg["dims_next"] == 2, cfg
        assert 0 < cfg["last_exact_fraction"] < 0.10
        idx_planar = c.correspondences()
        c.set_source(D)
        e2, r2 = c.run(ib.default_params(max_iter=4, stop_early=0, sync_every=1, nn_method=ib.NN_BRUTE_DIRECT))
        assert np.array_equal(idx_planar, c.correspondences()) and np.array_equal(e, e2) and list(r.R) == list(r2.R)
        rng = np.random.default_rng(5)
        Q = rng.random((20000, 3)).astype(np.float32)
        P = rng.random((20000, 3)).astype(np.float32)          # unrelated to Q: thresholds of the order of the point spacing
        c.set_target(Q); c.set_source(P)
        e, r = c.run(ib.default_params(max_iter=6, stop_early=0, sync_every=1))
        cfg = c.filter_config()
        assert cfg["dims_next"] == 3 and cfg["dims_last"] == 3, cfg
        c.set_source(P)
        assert np.array_equal(c.match(0, ib.NN_BRUTE), orc.match(P, Q, 0))
    finally:
        c.close()


def test_grouped_columns_on_rasters_no_group_size_divides(ctx, ib, orc):
    """K1T's grouped columns never span a jump of the scan (a row end, a gap in a sweep): rasters whose row length no group
    size divides, a cloud whose scan is cut by gaps of random length, and a raster followed by scattered points."""
    for w in (37, 61):
        D, M = orc.synth_p2p(w)
        _check(ctx, ib, orc, D, M)
        _check(ctx, ib, orc, orc.icp_p2p(D, M, max_iter=12, stop_early=False)["P"], M, oracle=False)
    rng = np.random.default_rng(77)
    t = np.cumsum(np.where(rng.random(6000) < 0.02, rng.uniform(0.5, 3.0, 6000), 0.01)).astype(np.float32)      # a sweep with gaps
    Q = np.stack([np.cos(t) * (1 + 0.05 * t), np.sin(t) * (1 + 0.05 * t), 0.1 * np.sin(7 * t)], axis=1).astype(np.float32)
    P = (Q[rng.permutation(6000)[:3000]] + rng.normal(scale=0.004, size=(3000, 3))).astype(np.float32)
    _check(ctx, ib, orc, P, Q)
    D, M = orc.synth_p2p(50)
    Q2 = np.concatenate([M, rng.normal(size=(777, 3)).astype(np.float32) * 3.0, M[:333] + np.float32(0.001)]).astype(np.float32)
    _check(ctx, ib, orc, D, Q2)


def test_tc_group_size_policy(ib, orc):
    """Automatic choice of K1T's targets per column: the power of two at or above sqrt(m) / 64 on a scan-ordered cloud, one at once for a cloud in
    arbitrary order (mean step of the scan far above the point spacing) — which is then grouped along its Morton order —
    and again from the start for the next target."""
    c = _context(ib, ICPB_K1_FILTER_MIN_PAIRS=0)
    try:
        for w, want in ((128, 2), (317, 8)):
            D, M = orc.synth_p2p(w)
            c.set_target(M); c.set_source(D)
            a = c.match(0, ib.NN_BRUTE)
            # (the exact-pass rate of this cold first pass may already have halved it once)
            assert c.filter_config()["dims_last"] == 4 and c.filter_tc_config()["targets_per_column"] in (want, want // 2), c.filter_tc_config()
            assert np.array_equal(a, c.match(0, ib.NN_BRUTE_DIRECT))
        rng = np.random.default_rng(3)
        Q = rng.random((20000, 3)).astype(np.float32); P = rng.random((5000, 3)).astype(np.float32)
        c.set_target(Q); c.set_source(P)
        a = c.match(0, ib.NN_BRUTE)
        cfg = c.filter_tc_config()
        assert cfg["morton_order"] and cfg["targets_per_column"] >= 2, cfg      # groups over the Morton order instead of one target per column
        assert np.array_equal(a, orc.match(P, Q, 0))
        assert np.array_equal(c.match(1, ib.NN_BRUTE), orc.match(P, Q, 1))
        D, M = orc.synth_p2p(317)
        c.set_target(M); c.set_source(D)
        c.match(0, ib.NN_BRUTE)
        assert c.filter_tc_config()["targets_per_column"] in (8, 4) and not c.filter_tc_config()["morton_order"]
    finally:
        c.close()
