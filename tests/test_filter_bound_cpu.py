"""CPU suite: the arithmetic of K1F's lower-bound filter (csrc/nn_filter.cu), restated in numpy.

The kernel skips a 128-target sub-tile for a source when  min_j e~_j > tau(thr)  and then never looks at those targets
again, so correctness rests on one inequality per (source, target) pair:

        e~_j  <=  tau(d_chain_j)                       (if target j is at chain distance d, no threshold >= d lets it be skipped)

with every quantity computed exactly as the kernel computes it in float32 — centring, FMA chains, directed roundings.
This test evaluates both sides for millions of pairs (full 3-FMA bound and the planar 2-FMA bound for each dropped
axis) on clouds chosen to stress the error analysis: large offsets from the origin (cancellation), tiny and huge
scales, coincident and nearly coincident points (relative error of tiny distances), lattices, sources far outside the
target. float32 FMAs are emulated in extended precision (exact product, one rounding)."""
import numpy as np
import pytest

F = np.float32
U = F(2.0 ** -24)
ONE8U = F(1.0) + F(8.0) * U


def fma(a, b, c):
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(np.float32)


def _dir(x64, up):
    """Round float64 values to float32 toward +inf (up) or -inf."""
    x64 = np.atleast_1d(np.asarray(x64, np.float64))
    x32 = x64.astype(np.float32)
    back = x32.astype(np.float64)
    if up:
        bad = back < x64
        x32[bad] = np.nextafter(x32[bad], F(np.inf))
    else:
        bad = back > x64
        x32[bad] = np.nextafter(x32[bad], F(-np.inf))
    return x32


def mul_ru(a, b): return _dir(np.asarray(a, np.float64) * np.asarray(b, np.float64), True)
def mul_rd(a, b): return _dir(np.asarray(a, np.float64) * np.asarray(b, np.float64), False)
def add_ru(a, b): return _dir(np.asarray(a, np.float64) + np.asarray(b, np.float64), True)
def sub_ru(a, b): return _dir(np.asarray(a, np.float64) - np.asarray(b, np.float64), True)
def fma_ru(a, b, c): return _dir((np.asarray(a, np.longdouble) * np.asarray(b, np.longdouble) + np.asarray(c, np.longdouble)).astype(np.float64), True)
def sqrt_ru(a): return _dir(np.sqrt(np.asarray(a, np.float64)), True)


def chain(dx, dy, dz):
    return fma(dz, dz, fma(dx, dx, dy * dy))


def check_cloud(P, Q, axes_sets=((0, 1, 2), (1, 2), (0, 2), (0, 1))):
    P = np.ascontiguousarray(P, np.float32); Q = np.ascontiguousarray(Q, np.float32)
    ctr = (F(0.5) * Q.min(axis=0) + F(0.5) * Q.max(axis=0)).astype(np.float32)          # kf_center
    qc = (Q - ctr).astype(np.float32)
    w3 = chain(qc[:, 0], qc[:, 1], qc[:, 2])
    rq = np.nextafter(F(np.sqrt(F(w3.max())) * ONE8U), F(np.inf))                        # kf_rq
    pc = (P - ctr).astype(np.float32)
    # exact chain distances, all pairs
    d = chain((P[:, None, 0] - Q[None, :, 0]).astype(np.float32), (P[:, None, 1] - Q[None, :, 1]).astype(np.float32),
              (P[:, None, 2] - Q[None, :, 2]).astype(np.float32))
    worst = {}
    for axes in axes_sets:
        if len(axes) == 3:
            p2 = chain(pc[:, 0], pc[:, 1], pc[:, 2])
            w = w3
            e = fma(-2 * pc[:, None, 0], qc[None, :, 0], fma(-2 * pc[:, None, 1], qc[None, :, 1],
                    fma(-2 * pc[:, None, 2], qc[None, :, 2], np.broadcast_to(w[None, :], d.shape))))
        else:
            a, b = axes
            p2 = fma(pc[:, a], pc[:, a], pc[:, b] * pc[:, b])
            w = fma(qc[:, a], qc[:, a], qc[:, b] * qc[:, b])
            e = fma(-2 * pc[:, None, a], qc[None, :, a], fma(-2 * pc[:, None, b], qc[None, :, b], np.broadcast_to(w[None, :], d.shape)))
        p2lo = mul_rd(p2, F(1.0) - F(8.0) * U)
        rp = mul_ru(sqrt_ru(p2), ONE8U)
        eps = mul_ru(F(8.0) * rq, rq) * np.ones_like(rp)
        eps = fma_ru(F(10.0) * rp, rq * np.ones_like(rp), eps)
        eps = fma_ru(F(2.0) * rp, rp, eps)
        eps = mul_ru(eps, F(1.05) * U)
        kk = sub_ru(eps, p2lo)
        tau = add_ru(mul_ru(d, ONE8U), np.broadcast_to(kk[:, None], d.shape))          # tau(thr = d_chain_j)
        ok = e <= tau
        assert ok.all(), "lower bound violated for axes %s at %d pairs (first: %s)" % (axes, (~ok).sum(), np.argwhere(~ok)[0])
        worst[axes] = float(np.min((tau.astype(np.float64) - e.astype(np.float64)) / np.maximum(np.abs(tau.astype(np.float64)), 1e-300)))
    return worst


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bound_holds_on_random_clouds_offsets_and_scales(seed):
    rng = np.random.default_rng(seed)
    base_q = rng.normal(size=(700, 3)); base_p = rng.normal(size=(500, 3))
    for shift, scale in ((0.0, 1.0), (1000.0, 1.0), (-5e4, 30.0), (3.0, 1e5), (0.0, 1e-4), (7.0, 1e-3), (123456.0, 0.5)):
        Q = (base_q * scale + shift).astype(np.float32)
        P = (base_p * scale + shift).astype(np.float32)
        P[:10] += np.float32(50 * scale)                      # far outside the target
        check_cloud(P, Q)


def test_bound_holds_for_coincident_and_nearly_coincident_points(orc):
    """ICP at convergence: sources sit on targets (d = 0 exactly, or a few ulps) — the regime where a relative bound
    on a difference of large numbers is most easily wrong."""
    D, M = orc.synth_p2p(40)
    check_cloud(M[:600], M)                                                    # exact coincidence
    P = M[:600].copy()
    P = np.nextafter(P, np.float32(np.inf)); P[::2] = np.nextafter(P[::2], np.float32(np.inf))
    check_cloud(P, M)                                                          # 1-2 ulps away
    P = orc.icp_p2p(D, M, max_iter=30)["P"]                                    # a registration's last stage
    check_cloud(P[:800], M)
    check_cloud(D[:800], M)                                                    # and its first


def test_bound_holds_on_lattices_and_flat_clouds():
    rng = np.random.default_rng(3)
    Q = (rng.integers(-8, 9, size=(900, 3)) * 0.25).astype(np.float32)
    P = (rng.integers(-16, 17, size=(600, 3)) * 0.125).astype(np.float32)
    check_cloud(P, Q)
    Q[:, 2] = 0.0; P[:, 2] = np.float32(1e-3)                                  # planar target: one centred coordinate is all zeros
    check_cloud(P, Q)
    Q[:, 1] = Q[:, 0]                                                          # collinear
    check_cloud(P, Q)


# ---------------------------------------------------------------------------------------------------------------------
# K1T (csrc/nn_filter_tc.cu): the same bound with the bracket evaluated by tcgen05.mma kind::tf32 — operands as round-to-
# nearest TF32 hi + lo pairs, products hi*hi + hi*lo + lo*hi per coordinate, w as hi + lo — and eps multiplied by 16; and
# the grouped form, where one column stands for up to TPC consecutive targets through their centroid g and radius h:
#       e~_g - 2 s0 h - h^2  <=  tau(d_chain_k)      for every member k and every s0 >= sqrt(d_chain_k (1 + 8u))
# (s0 = the root of the threshold the source starts a sweep with; a member only matters while the threshold is >= its
# distance). The -h^2 sits in w, -2 s0 h is a twelfth TF32 product. The tensor core's own accumulation rounding is not
# specified; the kernel budgets 60 u of sum |terms| for it (measured on B200: 2.2 u, tools/ubench_tc_filter check). Here the
# products are summed exactly and the WORST case of that budget is added to e~ before the comparison, so the test covers
# any accumulation order within the budget.
# ---------------------------------------------------------------------------------------------------------------------
def tf32_rn(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xfff + ((u >> 13) & 1)) & 0xffffe000
    return u.astype(np.uint32).view(np.float32)


def tf32_ru_pos(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1fff) & 0xffffe000).astype(np.uint32).view(np.float32)


def tf32_split(x):
    hi = tf32_rn(x)
    lo = tf32_rn((np.asarray(x, np.float32) - hi).astype(np.float32))
    return hi, lo


def sub_rd(a, b): return _dir(np.asarray(a, np.float64) - np.asarray(b, np.float64), False)


def tc_bracket(pc, mc, w):
    """(e, mag) for sources pc [n,3] against columns mc [g,3], w [g]: exact sum of the 11 TF32 products and the sum of their magnitudes."""
    a = (F(-2.0) * pc).astype(np.float32)
    e = np.zeros((pc.shape[0], mc.shape[0]), np.float64); mag = np.zeros_like(e)
    for c in range(3):
        ah, al = tf32_split(a[:, c]); qh, ql = tf32_split(mc[:, c])
        for x, y in ((ah, qh), (ah, ql), (al, qh)):
            t = x.astype(np.float64)[:, None] * y.astype(np.float64)[None, :]
            e += t; mag += np.abs(t)
    wh, wl = tf32_split(w)
    for y in (wh, wl):
        e += y.astype(np.float64)[None, :]; mag += np.abs(y.astype(np.float64))[None, :]
    return e, mag


def tc_columns(Q, tpc):
    """Column starts as build_filter_tc_data makes them: runs of the scan between jumps (a step > 8 x the mean step), cut into groups of tpc."""
    m = Q.shape[0]
    if tpc == 1 or m < 2:
        return np.arange(m + 1)
    step = np.sqrt(((Q[1:] - Q[:-1]).astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    jump = F(8.0) * F(step.astype(np.float64).sum() / (m - 1))
    brk = np.concatenate([[True], ~(step <= jump)])
    runstart = np.maximum.accumulate(np.where(brk, np.arange(m), 0))
    flag = (np.arange(m) - runstart) % tpc == 0
    return np.concatenate([np.nonzero(flag)[0], [m]])


def check_cloud_tc(P, Q, tpc, s0_factors=(1.0, 3.0, 1000.0)):
    P = np.ascontiguousarray(P, np.float32); Q = np.ascontiguousarray(Q, np.float32)
    m = Q.shape[0]
    ctr = (F(0.5) * Q.min(axis=0) + F(0.5) * Q.max(axis=0)).astype(np.float32)
    qc_all = (Q - ctr).astype(np.float32)
    rq = np.nextafter(F(np.sqrt(F(chain(qc_all[:, 0], qc_all[:, 1], qc_all[:, 2]).max())) * ONE8U), F(np.inf))       # kf_rq over the TARGETS
    cs = tc_columns(Q, tpc)
    ncol = len(cs) - 1
    lens = np.diff(cs)
    assert lens.min() >= 1 and lens.max() <= tpc
    # members padded to tpc per column (index -1 = no member)
    member = np.where(np.arange(tpc)[None, :] < lens[:, None], cs[:-1, None] + np.arange(tpc)[None, :], -1)
    G = np.zeros((ncol, 3), np.float32)
    inv = (F(1.0) / lens.astype(np.float32)).astype(np.float32)
    for k in range(tpc):
        has = member[:, k] >= 0
        G[has] = (G[has] + Q[member[has, k]] * inv[has, None]).astype(np.float32)       # the kernel's centroid, same operation order
    if tpc == 1:
        G = Q.copy()
    dd = np.zeros(ncol, np.float32)
    for k in range(tpc):
        has = member[:, k] >= 0
        ex = (Q[member[has, k]] - G[has]).astype(np.float32)
        dk = fma_ru(ex[:, 2], ex[:, 2], fma_ru(ex[:, 0], ex[:, 0], mul_ru(ex[:, 1], ex[:, 1])))
        dd[has] = np.maximum(dd[has], dk)
    H = tf32_ru_pos(mul_ru(sqrt_ru(dd), F(1.0) + F(16.0) * U)) if tpc > 1 else np.zeros(ncol, np.float32)
    hmax = F(H.max())
    gc = (G - ctr).astype(np.float32)
    w = chain(gc[:, 0], gc[:, 1], gc[:, 2])
    if tpc > 1:
        w = sub_rd(w, mul_ru(H, H))
    pc = (P - ctr).astype(np.float32)
    p2 = chain(pc[:, 0], pc[:, 1], pc[:, 2])
    p2lo = mul_rd(p2, F(1.0) - F(8.0) * U)
    rp = mul_ru(sqrt_ru(p2), ONE8U)
    e0, mag0 = tc_bracket(pc, gc, w)
    n = P.shape[0]
    for k in range(tpc):
        has = member[:, k] >= 0
        Qk = Q[np.maximum(member[:, k], 0)]
        d = chain((P[:, None, 0] - Qk[None, :, 0]).astype(np.float32), (P[:, None, 1] - Qk[None, :, 1]).astype(np.float32),
                  (P[:, None, 2] - Qk[None, :, 2]).astype(np.float32))
        for fac in (s0_factors if tpc > 1 else (1.0,)):
            # the smallest s0 the kernel can hold while this member still matters (threshold == its distance), and larger ones
            th = (d * F(fac) * F(fac)).astype(np.float32)
            s0 = tf32_ru_pos(mul_ru(sqrt_ru(mul_ru(th, ONE8U)), F(1.0) + F(4.0) * U)).reshape(d.shape) if tpc > 1 else np.zeros_like(d)
            eps = np.broadcast_to((mul_ru(F(8.0) * rq, rq) * np.ones_like(rp))[:, None], d.shape)
            eps = fma_ru(np.broadcast_to((F(10.0) * rp)[:, None], d.shape), rq * np.ones_like(d), eps).reshape(d.shape)
            eps = fma_ru(np.broadcast_to((F(2.0) * rp)[:, None], d.shape), np.broadcast_to(rp[:, None], d.shape), eps).reshape(d.shape)
            if tpc > 1:
                eps = fma_ru(F(8.0) * hmax * np.ones_like(d), hmax * np.ones_like(d), eps).reshape(d.shape)
                eps = fma_ru(F(4.0) * s0, hmax * np.ones_like(d), eps).reshape(d.shape)
            eps = mul_ru(eps, F(1.05) * F(16.0) * U).reshape(d.shape)                                  # TC_EPS_SCALE = 16
            kk = sub_ru(eps, np.broadcast_to(p2lo[:, None], d.shape)).reshape(d.shape)
            t12 = s0.astype(np.float64) * (F(-2.0) * H).astype(np.float64)[None, :]                    # exact: both TF32
            e = e0 + t12 + 60.0 * float(U) * (mag0 + np.abs(t12))                                      # worst case of the accumulation budget
            tau = add_ru(mul_ru(d, ONE8U), kk).reshape(d.shape)                                        # tau(threshold = this member's distance)
            ok = (e <= tau.astype(np.float64)) | ~has[None, :]
            assert ok.all(), "K1T bound violated (tpc %d, member %d, s0 x %g) at %d pairs" % (tpc, k, fac, (~ok).sum())
    return lens


@pytest.mark.parametrize("tpc", [1, 2, 4, 8, 16])
def test_tensor_core_bound_single_and_grouped_columns(orc, tpc):
    rng = np.random.default_rng(40 + tpc)
    base_q = rng.normal(size=(640, 3)); base_p = rng.normal(size=(300, 3))
    for shift, scale in ((0.0, 1.0), (1000.0, 1.0), (-5e4, 30.0), (0.0, 1e-4), (123456.0, 0.5)):
        Q = (base_q * scale + shift).astype(np.float32)
        P = (base_p * scale + shift).astype(np.float32)
        P[:10] += np.float32(50 * scale)
        check_cloud_tc(P, Q, tpc)
    D, M = orc.synth_p2p(40)                                                   # raster order: consecutive targets are neighbours
    lens = check_cloud_tc(M[:400], M, tpc)                                     # exact coincidence
    if tpc > 1:                                                                # rows of 40: no column spans a row end
        assert np.all(np.diff(np.concatenate([[0], np.cumsum(lens)]))[:] <= tpc) and np.all(np.cumsum(lens)[np.cumsum(lens) % 40 == 0].size >= 40)
    check_cloud_tc(np.nextafter(M[:400], np.float32(np.inf)), M, tpc)
    check_cloud_tc(orc.icp_p2p(D, M, max_iter=30)["P"][:500], M, tpc)
    check_cloud_tc(D[:500], M, tpc)
    D37, M37 = orc.synth_p2p(37)                                               # rows that no group size divides
    check_cloud_tc(D37[:400], M37, tpc)
    L = (rng.integers(-8, 9, size=(640, 3)) * 0.25).astype(np.float32)         # lattice with duplicates: groups of identical points
    check_cloud_tc((rng.integers(-16, 17, size=(300, 3)) * 0.125).astype(np.float32), L, tpc)
