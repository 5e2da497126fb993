"""GPU suite (-m gpu): batched registration (BASELINE.json config 5) — every pair must behave exactly like a
stand-alone registration: oracle trajectory / iteration count / transform, and pose recovery for all pairs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batched_pairs_vs_oracle(ctx, ib, orc):
    import icp_synth
    S, T, r, t = icp_synth.batched_pairs(48)
    p = ib.default_params(max_iter=40)
    errors, iters, R, tt, ms = ctx.run_batched(p, S, T)
    for b in (0, 1, 7, 19, 33, 47):
        o = orc.icp_p2p(S[b], T[b], max_iter=40)
        assert iters[b] == o["iterations"], b
        k = o["iterations"] + 2
        assert np.all(np.abs(errors[b, :k] - o["errors"][:k]) <= 1e-5 * np.abs(o["errors"][:k]) + 1e-7), b
        assert np.abs(R[b] - o["R"]).max() <= 1e-5 and np.abs(tt[b] - o["t"]).max() <= 1e-5, b
    # every pair recovers its generating pose (ground truth)
    for b in range(48):
        assert np.abs(R[b] - icp_synth.euler_matrix(r[b]).astype(np.float64)).max() < 2e-5, b
        assert np.abs(tt[b] - t[b]).max() < 2e-5, b
        assert errors[b, iters[b] + 1] < 1e-5


def test_batched_equals_streaming_engine(ctx, ib):
    import icp_synth
    S, T, r, t = icp_synth.batched_pairs(6, n=1500)          # n not a multiple of the block
    p = ib.default_params(max_iter=40, dist_mode=ib.DIST_SQRT)
    errors, iters, R, tt, ms = ctx.run_batched(p, S, T)
    for b in range(6):
        ctx.set_target(T[b]); ctx.set_source(S[b])
        e1, res = ctx.run(p)
        assert res.iterations == iters[b]
        k = res.iterations + 2
        assert np.all(np.abs(errors[b, :k] - e1[:k]) <= 1e-6 * np.abs(e1[:k]) + 1e-7)
        assert np.abs(R[b] - np.array(res.R[:])).max() < 1e-6


def test_batched_argument_checks(ctx, ib):
    import icp_synth
    S, T, _, _ = icp_synth.batched_pairs(2)
    with pytest.raises(ib.IcpError, match="point-to-point"):
        ctx.run_batched(ib.default_params(metric=ib.POINT_TO_PLANE), S, T)
    big = np.zeros((1, 4097, 3), np.float32)
    with pytest.raises(ib.IcpError, match="larger"):
        ctx.run_batched(ib.default_params(), big, big)


def test_batched_filter_is_bitwise_the_direct_kernel(ib):
    """K9F (lower-bound filter + exact refine, the default) against K9 (exact chain on every pair, ICPB_K9_FILTER=0):
    identical error trajectories, iteration counts and accumulated transforms, bit for bit — in both distance modes,
    for ragged sizes (n, m not multiples of anything, n != m), lattice targets full of exact ties, and a degenerate
    target (non-finite point) that must take the no-filter path."""
    import os
    import icp_synth
    rng = np.random.default_rng(11)
    cases = []
    S, T, _, _ = icp_synth.batched_pairs(24)
    cases.append((S, T))
    S2, T2, _, _ = icp_synth.batched_pairs(5, n=1500)
    cases.append((S2, T2[:, :1333].copy()))                                       # n != m, ragged
    lat_t = (rng.integers(-6, 7, size=(3, 3000, 3)) * 0.25).astype(np.float32)
    lat_s = (rng.integers(-12, 13, size=(3, 700, 3)) * 0.125 + 0.03).astype(np.float32)
    cases.append((lat_s, lat_t))
    bad_t = T[:2].copy(); bad_t[1, 17] = np.inf
    cases.append((S[:2].copy(), bad_t))
    out = {}
    for flt in ("1", "0"):
        os.environ["ICPB_K9_FILTER"] = flt
        try:
            with ib.Context(0) as c:
                out[flt] = [c.run_batched(ib.default_params(max_iter=30, dist_mode=mode), s, t)[:4]
                            for (s, t) in cases for mode in (ib.DIST_SQ, ib.DIST_SQRT)]
        finally:
            del os.environ["ICPB_K9_FILTER"]
    for a, b in zip(out["1"], out["0"]):
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_batched_full_config5_4096_pairs(ctx, ib, orc):
    """BASELINE.json config 5 at full size: 4096 independent pairs of 2048-point clouds, poses from the counter-based
    generator of SURVEY.md 8(d) (splitmix64, seed 20240, stream b). Every pair must recover its generating pose; a
    sample is compared with the oracle's whole trajectory; and the batch is split-invariant (the 8-GPU run is 8
    replicas of 512 pairs each: a pair's result must not depend on which batch it sits in)."""
    import icp_synth
    B = 4096
    S, T, r, t = icp_synth.batched_pairs(B)
    assert S.shape == (B, 2048, 3) and T.shape == (B, 2048, 3)
    p = ib.default_params(max_iter=40)
    errors, iters, R, tt, ms = ctx.run_batched(p, S, T)
    Rtrue = np.stack([icp_synth.euler_matrix(r[b]).astype(np.float64) for b in range(B)])
    assert np.abs(R - Rtrue).max() < 2e-5
    assert np.abs(tt - t.astype(np.float64)).max() < 2e-5
    final = errors[np.arange(B), iters + 1]
    assert final.max() < 1e-5 and iters.min() >= 3 and iters.max() < 40
    for b in (0, 511, 512, 1777, 2048, 3000, 4095):
        o = orc.icp_p2p(S[b], T[b], max_iter=40)
        assert iters[b] == o["iterations"], b
        k = o["iterations"] + 2
        assert np.all(np.abs(errors[b, :k] - o["errors"][:k]) <= 1e-5 * np.abs(o["errors"][:k]) + 1e-7), b
        assert np.abs(R[b] - o["R"]).max() <= 1e-5 and np.abs(tt[b] - o["t"]).max() <= 1e-5, b
    # rank 3 of 8 holds pairs [1536, 2048): identical bits to the same pairs inside the full batch
    lo, hi = 3 * 512, 4 * 512
    e2, i2, R2, t2, _ = ctx.run_batched(p, S[lo:hi], T[lo:hi])
    assert np.array_equal(i2, iters[lo:hi]) and np.array_equal(e2.view(np.uint32), errors[lo:hi].view(np.uint32))
    assert np.array_equal(R2.view(np.uint64), R[lo:hi].view(np.uint64)) and np.array_equal(t2.view(np.uint64), tt[lo:hi].view(np.uint64))
