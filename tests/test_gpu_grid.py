"""GPU suite (-m gpu): the exact uniform-grid nearest-neighbour variant must give the SAME indices as the
brute-force kernel (hence as the reference), including ties, far-away sources and the sentinel rule."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _both(ctx, ib, P, Q, mode, sentinel=100000.0):
    ctx.set_target(Q); ctx.set_source(P)
    a = ctx.match(mode, ib.NN_BRUTE, sentinel)
    da = ctx.min_distances()
    b = ctx.match(mode, ib.NN_GRID, sentinel)
    db = ctx.min_distances()
    return a, b, da, db


@pytest.mark.parametrize("mode", [0, 1])
def test_grid_equals_brute_on_registration_stages(ctx, ib, orc, mode):
    D, M = orc.synth_p2p(64)
    for iters in (0, 2, 6, 14):
        P = D if iters == 0 else orc.icp_p2p(D, M, max_iter=iters, stop_early=False)["P"]
        a, b, da, db = _both(ctx, ib, P, M, mode)
        assert np.array_equal(a, b), iters
        assert np.array_equal(da, db)
        assert np.array_equal(a, orc.match(P, M, mode))
    st = ctx.grid_stats()
    assert st["candidates_visited"] > 0 and min(st["dims"]) >= 1


@pytest.mark.parametrize("mode", [0, 1])
def test_grid_equals_brute_ties_duplicates_and_outliers(ctx, ib, orc, mode):
    rng = np.random.default_rng(17 + mode)
    Q = (rng.integers(-10, 11, size=(6000, 3)) * 0.25).astype(np.float32)          # lattice: massive ties + duplicates
    P = (rng.integers(-20, 21, size=(3000, 3)) * 0.125).astype(np.float32)
    P[:50] += np.float32(40.0)                                                      # far outside the target's box
    P[50:60] -= np.float32(1e3)
    a, b, da, db = _both(ctx, ib, P, Q, mode)
    assert np.array_equal(a, b) and np.array_equal(da, db)
    assert np.array_equal(a, orc.match(P, Q, mode))


def test_grid_respects_sentinel(ctx, ib, orc):
    rng = np.random.default_rng(3)
    Q = rng.normal(size=(2000, 3)).astype(np.float32)
    P = (rng.normal(size=(500, 3)) * 3).astype(np.float32)
    for mode in (0, 1):
        a, b, da, db = _both(ctx, ib, P, Q, mode, sentinel=0.05)
        assert np.array_equal(a, b) and np.array_equal(da, db)


def test_grid_flat_and_tiny_clouds(ctx, ib, orc):
    Q = np.zeros((300, 3), np.float32); Q[:, 0] = np.linspace(0, 1, 300)           # a line: degenerate bounding box
    P = Q[::7] + np.float32(0.001)
    a, b, _, _ = _both(ctx, ib, P, Q, 0)
    assert np.array_equal(a, b)
    Q1 = np.array([[1, 2, 3]], np.float32)
    a, b, _, _ = _both(ctx, ib, np.array([[0, 0, 0], [5, 5, 5]], np.float32), Q1, 0)
    assert list(a) == list(b) == [0, 0]


def test_full_run_grid_is_bitwise_the_brute_force_run(ctx, ib, orc):
    D, M = orc.synth_p2p(317, 100000)
    ctx.set_target(M); ctx.set_source(D)
    e1, r1 = ctx.run(ib.default_params(max_iter=64))
    idx1 = ctx.correspondences()
    ctx.set_source(D)
    e2, r2 = ctx.run(ib.default_params(max_iter=64, nn_method=ib.NN_GRID))
    assert r1.iterations == r2.iterations
    assert np.array_equal(e1, e2) and list(r1.R) == list(r2.R) and list(r1.t) == list(r2.t)
    assert np.array_equal(idx1, ctx.correspondences())


@pytest.mark.parametrize("pyramid", [1, 0])
def test_grid_search_variants_equal_brute(ib, orc, pyramid):
    """The two ways the grid is searched — best-first descent of the occupancy pyramid over its cells
    (csrc/grid_tree.cuh, the default) and rings 0-2 with a brute-force fallback (ICPB_GRID_PYRAMID=0) — under the same
    contract: identical indices and distances in both modes, on every stage of a registration (far and near field), on
    tie-heavy lattices with outliers, under a small sentinel, on degenerate boxes; with the pyramid no source is ever
    left to the brute-force kernel."""
    import os
    os.environ["ICPB_GRID_PYRAMID"] = str(pyramid)
    try:
        c = ib.Context(0)
    finally:
        del os.environ["ICPB_GRID_PYRAMID"]
    try:
        D, M = orc.synth_p2p(64)
        for mode in (0, 1):
            for iters in (0, 2, 6, 14):
                P = D if iters == 0 else orc.icp_p2p(D, M, max_iter=iters, stop_early=False)["P"]
                a, b, da, db = _both(c, ib, P, M, mode)
                assert np.array_equal(a, b) and np.array_equal(da, db), (mode, iters)
                assert not pyramid or c.grid_stats()["last_open_sources"] == 0
                b2 = c.match(mode, ib.NN_GRID)                                    # warm-started from the pass before
                assert np.array_equal(b2, a)
            rng = np.random.default_rng(17 + mode)
            Q = (rng.integers(-10, 11, size=(6000, 3)) * 0.25).astype(np.float32)
            P = (rng.integers(-20, 21, size=(3000, 3)) * 0.125).astype(np.float32)
            P[:50] += np.float32(40.0); P[50:60] -= np.float32(1e3)
            a, b, da, db = _both(c, ib, P, Q, mode)
            assert np.array_equal(a, b) and np.array_equal(da, db)
            assert np.array_equal(a, orc.match(P, Q, mode))
            Qn = rng.normal(size=(2000, 3)).astype(np.float32)
            Pn = (rng.normal(size=(500, 3)) * 3).astype(np.float32)
            a, b, da, db = _both(c, ib, Pn, Qn, mode, sentinel=0.05)
            assert np.array_equal(a, b) and np.array_equal(da, db)
        Q = np.zeros((300, 3), np.float32); Q[:, 0] = np.linspace(0, 1, 300)
        a, b, _, _ = _both(c, ib, Q[::7] + np.float32(0.001), Q, 0)
        assert np.array_equal(a, b)
        a, b, _, _ = _both(c, ib, np.array([[0, 0, 0], [5, 5, 5]], np.float32), np.array([[1, 2, 3]], np.float32), 0)
        assert list(a) == list(b) == [0, 0]
        # a whole registration: bitwise the brute-force trajectory
        D, M = orc.synth_p2p(200)
        c.set_target(M)
        out = []
        for nn in (ib.NN_GRID, ib.NN_BRUTE_DIRECT):
            c.set_source(D)
            e, r = c.run(ib.default_params(max_iter=64, nn_method=nn))
            out.append((e.copy(), r.iterations, list(r.R), c.correspondences()))
        assert np.array_equal(out[0][0], out[1][0]) and out[0][1:3] == out[1][1:3] and np.array_equal(out[0][3], out[1][3])
    finally:
        c.close()
