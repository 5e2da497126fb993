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


# ---- cold starts and adversarial seeds (VERDICT r1 weak #1 / ADVICE r1) ------------------------------------------
# _both() above runs the brute-force kernel first, and icpb_match leaves its result in the seed array the grid descent
# starts from: an over-pruning descent would still return the right index. The tests below never run a brute-force pass
# before the grid pass they check: the reference is the CPU oracle only.

def _fresh(ib, **env):
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


def _cold_cases(orc):
    rng = np.random.default_rng(123)
    D, M = orc.synth_p2p(64)
    yield "saddle initial pose", D, M, 100000.0
    yield "saddle after 6 iterations", orc.icp_p2p(D, M, max_iter=6, stop_early=False)["P"], M, 100000.0
    Pn = M.copy(); Pn[:, 0] = np.nextafter(Pn[:, 0], np.float32(np.inf)); Pn[::7, 2] += np.float32(1e-3)
    yield "saddle near field (ulps off the target)", Pn, M, 100000.0
    Q = (rng.integers(-10, 11, size=(6000, 3)) * 0.25).astype(np.float32)
    P = (rng.integers(-20, 21, size=(3000, 3)) * 0.125).astype(np.float32)
    P[:50] += np.float32(40.0); P[50:60] -= np.float32(1e3)
    yield "lattice: ties, duplicates, outliers", P, Q, 100000.0
    Qn = rng.normal(size=(2000, 3)).astype(np.float32)
    yield "random cloud, sentinel 0.05", (rng.normal(size=(500, 3)) * 3).astype(np.float32), Qn, 0.05
    Qu = (Qn + np.array([1e6, -2.5e5, 3e3], np.float32)).astype(np.float32)
    yield "unit cloud a million units from the origin", (Qu[:700] + np.float32(0.25)).astype(np.float32), Qu, 3e38
    # one axis that would need > 2^17 cells at the nominal cell size (368334 x 1 x 1, 18 levels, wrong neighbours cold)
    L = np.zeros((100000, 3), np.float32); L[:, 0] = np.arange(100000, dtype=np.float32) * np.float32(1e-3)
    PL = L[::331].copy(); PL[:, 0] += np.float32(4e-4); PL[::2, 1] += np.float32(0.01); PL[::3, 2] -= np.float32(0.02)
    yield "100k collinear target", PL, L, 100000.0
    T = np.stack([200.0 * rng.random(100000), 0.01 * rng.random(100000), 0.01 * rng.random(100000)], axis=1).astype(np.float32)
    PT = T[::407].copy(); PT[:, 0] += np.float32(1e-3); PT[:, 1] -= np.float32(2e-3)
    yield "100k thin rod", PT, T, 100000.0
    F = np.stack([3000.0 * rng.random(60000), 3000.0 * rng.random(60000), np.full(60000, 7.0)], axis=1).astype(np.float32)
    PF = F[::263].copy(); PF[:, 0] += np.float32(0.5); PF[:, 2] += np.float32(1.5)
    yield "60k flat sheet of large extent", PF, F, 100000.0


@pytest.mark.parametrize("seeded", [0, 1])
def test_grid_cold_first_pass_equals_oracle(ib, orc, seeded):
    """NN_GRID is the FIRST matching pass of a fresh context (seeded=1: the seed array holds its reset value 0 for every
    source, i.e. target 0, which is just some candidate; seeded=0: ICPB_K1_SEED=0, the descent starts with no bound at all)."""
    c = _fresh(ib, ICPB_K1_SEED=seeded)
    try:
        for name, P, Q, sentinel in _cold_cases(orc):
            for mode in (0, 1):
                c.set_target(Q); c.set_source(P)
                got = c.match(mode, ib.NN_GRID, sentinel)
                d = c.min_distances()
                want = orc.match(P, Q, mode, sentinel, idx0=np.zeros(P.shape[0], np.int32))
                assert np.array_equal(got, want), (name, mode)
                # the winning distance is the reference chain's value of the winning pair, bit for bit
                hit = d < np.float32(sentinel)
                ref_d = orc.pair_distances(P[hit], Q[got[hit]], mode)
                assert np.array_equal(d[hit].view(np.uint32), ref_d.view(np.uint32)), (name, mode)
            assert c.grid_stats()["last_open_sources"] == 0
            assert max(c.grid_stats()["dims"]) <= (1 << 17), name
    finally:
        c.close()


def test_grid_with_scrambled_and_stale_seeds(ib, orc):
    """Seeds are only ever an upper bound: the correspondences of ANOTHER cloud of the same size (a permutation, a scaled
    and reversed copy) must not change the result. No brute-force pass runs in this context."""
    c = _fresh(ib)
    try:
        rng = np.random.default_rng(77)
        for name, P, Q, sentinel in _cold_cases(orc):
            for mode in (0, 1):
                c.set_target(Q)
                c.set_source(np.ascontiguousarray(P[rng.permutation(P.shape[0])]))
                c.match(mode, ib.NN_GRID, sentinel)                       # leaves the permuted cloud's correspondences as seeds
                c.set_source(P)                                             # same size: the seeds survive the upload
                assert np.array_equal(c.match(mode, ib.NN_GRID, sentinel), orc.match(P, Q, mode, sentinel, idx0=np.zeros(P.shape[0], np.int32))), (name, mode, "permuted")
                P2 = (P[::-1] * np.float32(1.7) + np.float32(0.3)).astype(np.float32)
                c.set_source(P2)                                            # stale seeds from P
                want2 = orc.match(P2, Q, mode, sentinel, idx0=np.zeros(P2.shape[0], np.int32))
                assert np.array_equal(c.match(mode, ib.NN_GRID, sentinel), want2), (name, mode, "stale")
    finally:
        c.close()


def test_whole_registration_grid_first_then_brute(ib, orc):
    """A whole registration with the grid variant in a context that has never run the brute-force kernel, against the
    oracle's trajectory; only then the brute-force run for the bitwise comparison."""
    c = _fresh(ib)
    try:
        D, M = orc.synth_p2p(128)
        c.set_target(M); c.set_source(D)
        e, r = c.run(ib.default_params(max_iter=40, nn_method=ib.NN_GRID))
        idx = c.correspondences()
        o = orc.icp_p2p(D, M, max_iter=40)
        assert r.iterations == o["iterations"]
        k = r.iterations + 2
        assert np.all(np.abs(e[:k] - o["errors"][:k]) <= 1e-5 * np.abs(o["errors"][:k]) + 1e-7)
        assert np.array_equal(idx, np.arange(128 * 128))
        c.set_source(D)
        e2, r2 = c.run(ib.default_params(max_iter=40, nn_method=ib.NN_BRUTE_DIRECT))
        assert np.array_equal(e, e2) and list(r.R) == list(r2.R) and list(r.t) == list(r2.t)
    finally:
        c.close()
