"""CPU suite: the tie rule of K1T's Morton form (csrc/nn_filter_tc.cu), restated in numpy.

With the tiles cut from the targets in Morton order the slots of a unit are no longer in index order, and the units are
visited by several warps in whatever order the pipeline delivers them. The kernel keeps, per source, one 64-bit key
(bits of the exact threshold << 32 | smallest ORIGINAL index that attains it) updated with atomicMin, and offers a unit
whenever its minimum is <= the threshold read before the exact pass. Claim: whatever the visiting order, whatever the
starting threshold (the seed's distance, which some target attains), and however aggressively units are skipped (here: a
unit is skipped exactly when none of its targets is within the current threshold — the tightest filter any valid lower
bound can be), the final key is the oracle's answer: the smallest distance, and among equal distances the lowest index.
The lower bound's arithmetic is covered by test_filter_bound_cpu.py; this test covers the bookkeeping."""
import numpy as np
import pytest

F = np.float32


def chain(P, Q):
    """float32 chain of the reference kernel for all pairs: fma(dz, dz, fma(dx, dx, dy * dy))."""
    d = (P[:, None, :] - Q[None, :, :]).astype(np.float32)
    t = (d[..., 1] * d[..., 1]).astype(np.float32)
    t = (d[..., 0].astype(np.float64) * d[..., 0].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
    return (d[..., 2].astype(np.float64) * d[..., 2].astype(np.float64) + t.astype(np.float64)).astype(np.float32)


def spread21(v):
    x = v.astype(np.uint64) & np.uint64(0x1fffff)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_order(Q):
    lo, hi = Q.min(axis=0), Q.max(axis=0)
    ctr = 0.5 * lo + 0.5 * hi
    rq = np.sqrt(((Q - ctr) ** 2).sum(axis=1).max()) * 1.0001 + 1e-30
    f = np.clip((Q - (ctr - rq)) * (0.5 / rq), 0.0, 1.0)
    u = np.minimum((f * 2097152.0).astype(np.uint64), np.uint64(2097151))
    key = spread21(u[:, 0]) | (spread21(u[:, 1]) << np.uint64(1)) | (spread21(u[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable")


def test_morton_keys_keep_neighbours_together():
    rng = np.random.default_rng(1)
    Q = rng.uniform(-1, 1, size=(20000, 3)).astype(np.float32)
    perm = morton_order(Q)
    assert np.array_equal(np.sort(perm), np.arange(Q.shape[0]))
    step_sorted = np.linalg.norm(Q[perm][1:] - Q[perm][:-1], axis=1).mean()
    step_raw = np.linalg.norm(Q[1:] - Q[:-1], axis=1).mean()
    assert step_sorted < 0.1 * step_raw                      # consecutive positions are neighbours again


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_key_with_smallest_original_index_is_order_independent(orc, seed):
    rng = np.random.default_rng(seed)
    m, n, unit = 1500, 160, 16
    Q = (rng.integers(-5, 6, size=(m, 3)) * 0.5).astype(np.float32)          # lattice: many exactly equal distances
    Q[1200:1300] = Q[100:200]                                               # duplicates at higher indices must never win
    P = (rng.integers(-10, 11, size=(n, 3)) * 0.25).astype(np.float32)
    ref = orc.match(P, Q, 0)
    perm = morton_order(Q)                                                  # slot -> original index
    Qs = Q[perm]
    D = chain(P, Qs)                                                        # [n, slots]
    nunits = (m + unit - 1) // unit
    for i in range(n):
        seed_j = int(rng.integers(0, m))                                    # any target is a valid seed
        th = chain(P[i:i + 1], Q[seed_j:seed_j + 1])[0, 0]
        best = (np.float32(np.inf), np.int64(2 ** 31 - 1))                  # the key: (threshold, original index); nothing offered yet
        key_th = th                                                         # threshold part of the key: starts at the seed's distance
        have = False
        for u in rng.permutation(nunits):                                   # any visiting order
            sl = slice(u * unit, min(m, (u + 1) * unit))
            du = D[i, sl]
            th_read = key_th                                                # read before the exact pass
            if not (du.min() <= th_read):                                   # tightest valid filter: nothing of this unit is within the threshold
                continue
            mm = du.min()
            attaining = perm[sl][du <= mm]                                  # every slot that attains the unit's minimum
            cand = (mm, attaining.min())
            cur = (key_th, best[1]) if have else (key_th, np.int64(2 ** 31 - 1))
            if cand < cur:                                                  # atomicMin on (threshold bits, index): distances are non-negative floats
                key_th, best, have = mm, cand, True
        assert have, "the seed itself is always offered"
        assert best[1] == ref[i], (i, best, ref[i])
