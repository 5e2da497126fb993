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
