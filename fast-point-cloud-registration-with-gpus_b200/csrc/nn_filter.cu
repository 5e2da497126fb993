// nn_filter.cu — K1F: exact brute-force matching with a rigorous lower-bound filter in front of the
// reference's distance chain. Same contract and same bits as K1 (nn_bruteforce.cu): every (source, target)
// pair is examined, idx[i] = argmin_j d_chain(P_i,Q_j) with the lowest j on ties.
//
// Why: the reference's chain costs 6 FP32-pipe operations per pair, which caps a direct evaluation at 66.7 %
// of the FFMA peak. |p-q|^2 = |p|^2 + (|q|^2 - 2 p.q) needs only 3 FMAs per pair for the bracket
//     e~_j = fma(ax, qx_j, fma(ay, qy_j, fma(az, qz_j, w_j))),   a = -2 (p - c),  q_j <- q_j - c,  w_j = |q_j - c|^2
// (c = centre of the target's bounding box, removed to keep magnitudes — and hence rounding — small).
// e~ is NOT the reference's arithmetic, so it is only used to PROVE that a sub-tile of 128 targets cannot
// contain the answer:
//     d_chain_j >= (|pc|^2 + e~_j - eps)(1 - 5u)      u = 2^-24, eps = u (8 Rq^2 + 10 Rp Rq + 2 Rp^2) (1.05)
//   (Rq = max |q_j - c|, Rp = |p - c|; derivation in DESIGN.md §K1F), hence
//     min_j e~_j  >  tau := thr (1 + 8u) - |pc|^2_lo + eps   ==>   every d_chain_j in the sub-tile > thr,
// where thr is the running exact threshold of K1 (best exact distance so far, in the squared domain; the
// factor 1+8u also covers the floats that sqrt.rn merges with thr in sqrt mode). Sub-tiles that fail the
// test — the few that can hold the nearest neighbour — are evaluated by the whole warp with the exact
// packed chain of K1 and update thr / the remembered sub-tile exactly as K1 does (strict `<` in ascending
// sub-tile order => lowest index). thr starts from the exact distance to the previous iteration's
// correspondence (ICP warm start; any index is a valid upper bound, so this only affects speed).
// Everything the answer depends on is computed by the exact chain; the filter only decides what to skip.
//
// DIMS = 2 (the default once a registration is warm): |p-q|^2 >= the squared distance of the projections onto two
// coordinate axes, so the same bound with one coordinate dropped is still a rigorous lower bound and costs 2 FMAs per
// pair instead of 3 (measured on B200, tools/ubench_filter.cu: 13.5-14.3 cycles per source x 4 targets per SMSP
// instead of 17.2-17.5). The derivation is the 3-D one restricted to the kept plane: e~_j = fma(aA, qA_j, fma(aB, qB_j,
// w2_j)), w2 = |(q-c)_AB|^2, |pc|^2 taken in the plane; Rq (3-D) and eps over-estimate the planar quantities. What it
// gives up is tightness: more sub-tiles reach the exact pass (all those whose SHADOW on the kept plane comes within
// thr of the source). The dropped axis is chosen per target cloud as the one whose projection keeps the sub-tiles
// best separated (kf_score_kernel), and the host watches the exact-pass rate (kf_policy_update): above 10 % it goes
// back to the 3-D bound, below 1 % it tries the planar one again. Either way the returned indices are those of the
// exact chain.
#include "common.cuh"
#include "k1_device.cuh"
#include <cmath>

namespace icpb {

constexpr int KF_TT     = 512;                 // targets per shared-memory tile (8 float arrays: X Y Z | Xc Yc Zc | W3 W2)
constexpr int KF_TRK    = 128;                 // filter / tracking sub-tile
constexpr int KF_STAGES = 3;
constexpr int KF_TILE_FLOATS = 8 * KF_TT;
constexpr int KF_TILE_BYTES  = KF_TILE_FLOATS * 4;
constexpr float KF_U = 5.9604644775390625e-08f;            // 2^-24

struct KFParams {
	const float* px; const float* py; const float* pz;
	const float* tiles7;      // [nt][8][KF_TT]
	const float4* q4;         // for the warm-start gather
	const int*   seed_idx;    // previous correspondences (may hold anything in [0, m))
	u64*         keys;
	int          n, m, nt;
	long long    total_units;   // source blocks x nt: (source block, target tile) work units, source-block major
	int          min_chunk, max_chunk;   // tiles per grab: guided self-scheduling between these bounds
	int          gss_div;                // a grab takes 1/(gss_div x grid) of what is left
	unsigned long long* work_counter;    // zeroed before the launch; CTAs draw ranges of units from it
	float        thr0;
	float        cx, cy, cz;  // centre removed from both clouds for the filter quantities
	float        rq;          // upper bound of max_j |q_j - c|
	int          drop;        // DIMS == 2: the coordinate axis (0,1,2) left out of the bound
	const int*   done;
	unsigned long long* stats;   // [0] sub-tile x warp x source filter tests, [1] of which evaluated exactly
};

__global__ void kf_bbox_kernel(const float4* __restrict__ q4, int m, unsigned* mm)
{
	float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
	for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
		const float4 q = q4[j];
		lo[0] = fminf(lo[0], q.x); lo[1] = fminf(lo[1], q.y); lo[2] = fminf(lo[2], q.z);
		hi[0] = fmaxf(hi[0], q.x); hi[1] = fmaxf(hi[1], q.y); hi[2] = fmaxf(hi[2], q.z);
	}
	for (int k = 0; k < 3; k++) {
		for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
		if ((threadIdx.x & 31) == 0) {
			unsigned a = __float_as_uint(lo[k]); a = (a & 0x80000000u) ? ~a : (a | 0x80000000u);
			unsigned b = __float_as_uint(hi[k]); b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
			atomicMin(mm + k, a); atomicMax(mm + 3 + k, b);
		}
	}
}

// tiles7 + the radius bound (max of the float chain value of |q-c|^2, as ordered uint)
__global__ void kf_pack_kernel(const float4* __restrict__ q4, int m, int m_pad, float cx, float cy, float cz, int drop, float* __restrict__ tiles7, unsigned* __restrict__ r2max)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m_pad) return;
	float x = __int_as_float(0x7f800000), y = x, z = x, xc = 1e18f, yc = 1e18f, zc = 1e18f, w = 3e36f, w2 = 3e36f;
	if (j < m) {
		const float4 q = q4[j];
		x = q.x; y = q.y; z = q.z;
		xc = __fsub_rn(x, cx); yc = __fsub_rn(y, cy); zc = __fsub_rn(z, cz);
		w = __fmaf_rn(zc, zc, __fmaf_rn(xc, xc, __fmul_rn(yc, yc)));
		const float a = (drop == 0) ? yc : xc, b = (drop == 2) ? yc : zc;     // the two kept axes, in axis order
		w2 = __fmaf_rn(a, a, __fmul_rn(b, b));
		atomicMax(r2max, __float_as_uint(w));        // w >= 0: uint order = float order
	}
	if (tiles7 == nullptr) return;                 // K1T only needs the radius bound (its own tiles: nn_filter_tc.cu)
	float* t = tiles7 + (size_t)(j / KF_TT) * KF_TILE_FLOATS + (j % KF_TT);
	t[0] = x; t[KF_TT] = y; t[2 * KF_TT] = z; t[3 * KF_TT] = xc; t[4 * KF_TT] = yc; t[5 * KF_TT] = zc; t[6 * KF_TT] = w; t[7 * KF_TT] = w2;
}

// Which axis to leave out of the planar bound: for each of the three choices, count how many (sub-tile, sample point)
// pairs have the sample inside the sub-tile's bounding rectangle on the kept plane — an estimate of how many sub-tiles
// a source lying on the cloud cannot be separated from. 512 samples (every m/512-th target); one block per sub-tile.
__global__ void __launch_bounds__(KF_TRK) kf_score_kernel(const float4* __restrict__ q4, int m, unsigned long long* __restrict__ score /* [3] */)
{
	__shared__ float lo[3], hi[3];
	__shared__ unsigned cnt[3];
	const int j = blockIdx.x * KF_TRK + threadIdx.x;
	float4 q = make_float4(INFINITY, INFINITY, INFINITY, 0.f), r = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
	if (j < m) { q = q4[j]; r = q; }
	float l[3] = { q.x, q.y, q.z }, h[3] = { r.x, r.y, r.z };
	if (threadIdx.x < 3) cnt[threadIdx.x] = 0;
	for (int k = 0; k < 3; k++) {
		for (int o = 16; o > 0; o >>= 1) { l[k] = fminf(l[k], __shfl_xor_sync(0xffffffffu, l[k], o)); h[k] = fmaxf(h[k], __shfl_xor_sync(0xffffffffu, h[k], o)); }
	}
	__shared__ float wl[KF_TRK / 32][3], wh[KF_TRK / 32][3];
	if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; k++) { wl[threadIdx.x >> 5][k] = l[k]; wh[threadIdx.x >> 5][k] = h[k]; }
	__syncthreads();
	if (threadIdx.x < 3) {
		float a = INFINITY, b = -INFINITY;
		for (int w = 0; w < KF_TRK / 32; w++) { a = fminf(a, wl[w][threadIdx.x]); b = fmaxf(b, wh[w][threadIdx.x]); }
		lo[threadIdx.x] = a; hi[threadIdx.x] = b;
	}
	__syncthreads();
	const int stride = max(1, m / 512);
	unsigned c0 = 0, c1 = 0, c2 = 0;
	for (int s = threadIdx.x; s * (long long)stride < m && s < 512; s += KF_TRK) {
		const float4 v = q4[(size_t)s * stride];
		const bool ix = v.x >= lo[0] && v.x <= hi[0], iy = v.y >= lo[1] && v.y <= hi[1], iz = v.z >= lo[2] && v.z <= hi[2];
		c0 += (iy && iz); c1 += (ix && iz); c2 += (ix && iy);
	}
	atomicAdd(&cnt[0], c0); atomicAdd(&cnt[1], c1); atomicAdd(&cnt[2], c2);
	__syncthreads();
	if (threadIdx.x < 3) atomicAdd(score + threadIdx.x, (unsigned long long)cnt[threadIdx.x]);
}

template <int S, int THREADS, int MODE, int MINB, int DIMS>
__global__ void __launch_bounds__(THREADS, MINB) k1_filter(const KFParams p)
{
	static_assert(DIMS == 2 || DIMS == 3, "planar or full bound");
	constexpr int SB   = S * THREADS;
	constexpr int SUBS = KF_TT / KF_TRK;
	if (p.done != nullptr && *p.done) return;

	extern __shared__ __align__(128) unsigned char kf_smem[];
	__shared__ __align__(8) uint64_t full_bar[KF_STAGES];
	float* ring = reinterpret_cast<float*>(kf_smem);
	const int tid = threadIdx.x;
	// per-thread slots behind the ring, [s][tid]: exact threshold, remembered sub-tile, original source coordinates
	float* thr_s  = reinterpret_cast<float*>(kf_smem + (size_t)KF_STAGES * KF_TILE_BYTES + 64) + tid;
	int*   best_s = reinterpret_cast<int*>(thr_s - tid + S * THREADS) + tid;
	float* ox_s   = thr_s + 2 * S * THREADS;
	float* oy_s   = thr_s + 3 * S * THREADS;
	float* oz_s   = thr_s + 4 * S * THREADS;

	// Work = ranges of consecutive (source block, target tile) units drawn from a global counter: sub-tiles that need
	// the exact pass cluster where a source block's neighbours live, so a static split would leave the blocks that own
	// those tiles running long after the others (measured: 12 % at 125k sources per GPU). Guided self-scheduling: a
	// grab takes 1/(4 x grid) of what is left, between min_chunk and max_chunk tiles — large ranges while there is
	// plenty (the per-range source reload and index re-scan stay negligible), small ones at the end (the tail is a few
	// tiles, not a whole chunk: at 125k sources per GPU fixed 8-tile chunks left 4 % on the table).
	__shared__ unsigned long long s_u0;
	__shared__ int s_len;
	if (tid == 0) {
		for (int s = 0; s < KF_STAGES; s++) mbar_init(&full_bar[s], 1);
		fence_mbar_init();
	}
	__syncthreads();
	int it = 0;                  // tiles consumed so far by this CTA: ring stage and mbarrier parity follow it across chunks

	float ax[S], ay[S], az[S], tau[S], kk[S];     // DIMS == 2: (ax, ay) hold the two kept axes, az is unused
	const int axA = (DIMS == 2 && p.drop == 0) ? 1 : 0, axB = (DIMS == 2 && p.drop != 2) ? 2 : 1;
	unsigned long long n_tests = 0, n_exact = 0;
	const float inf = __int_as_float(0x7f800000);
	const float one8u = 1.0f + 8.0f * KF_U;

	auto flush = [&](int sb) {
#pragma unroll
		for (int s = 0; s < S; s++) {
			const int i = sb * SB + s * THREADS + tid;
			const int bs = best_s[s * THREADS];
			if (i < p.n && bs >= 0) {
				const float th = thr_s[s * THREADS];
				const float sx = ox_s[s * THREADS], sy = oy_s[s * THREADS], sz = oz_s[s * THREADS];
				const float* gx = p.tiles7 + (size_t)(bs / SUBS) * KF_TILE_FLOATS + (size_t)(bs % SUBS) * KF_TRK;
				const float4* X4 = reinterpret_cast<const float4*>(gx);
				const float4* Y4 = reinterpret_cast<const float4*>(gx + KF_TT);
				const float4* Z4 = reinterpret_cast<const float4*>(gx + 2 * KF_TT);
				const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(th) : th;
				int found = -1;
				for (int j = 0; j < KF_TRK / 4 && found < 0; j++) {
					const float4 X = __ldg(X4 + j), Y = __ldg(Y4 + j), Z = __ldg(Z4 + j);
					float d0 = dist_chain(sx, sy, sz, X.x, Y.x, Z.x);
					float d1 = dist_chain(sx, sy, sz, X.y, Y.y, Z.y);
					float d2 = dist_chain(sx, sy, sz, X.z, Y.z, Z.z);
					float d3 = dist_chain(sx, sy, sz, X.w, Y.w, Z.w);
					if (MODE == ICPB_DIST_SQRT) { d0 = __fsqrt_rn(d0); d1 = __fsqrt_rn(d1); d2 = __fsqrt_rn(d2); d3 = __fsqrt_rn(d3); }
					if (d0 <= target) found = 4 * j;
					else if (d1 <= target) found = 4 * j + 1;
					else if (d2 <= target) found = 4 * j + 2;
					else if (d3 <= target) found = 4 * j + 3;
				}
				if (found >= 0) {
					const u64 key = ((u64)__float_as_uint(target) << 32) | (u64)(uint32_t)(bs * KF_TRK + found);
					atomicMin(p.keys + i, key);
				}
			}
		}
	};

	while (true) {
		__syncthreads();                                   // previous range fully consumed (s_u0, s_len)
		if (tid == 0) {
			const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(p.work_counter);   // may be stale: only sizes the grab
			const long long left = p.total_units - (long long)seen;
			long long len = left / ((long long)p.gss_div * (long long)gridDim.x);
			if (len < p.min_chunk) len = p.min_chunk;
			if (len > p.max_chunk) len = p.max_chunk;
			s_len = (int)len;
			s_u0 = atomicAdd(p.work_counter, (unsigned long long)len);
		}
		__syncthreads();
		if ((long long)s_u0 >= p.total_units) break;
		long long u = (long long)s_u0;
		const long long u_end = min(p.total_units, u + (long long)s_len);
		// a range may run over the end of a source block: one segment per source block
		while (u < u_end) {
		const int sb = (int)(u / p.nt);
		const int t0 = (int)(u - (long long)sb * p.nt);
		const int t1 = (int)min((long long)p.nt, (long long)t0 + (u_end - u));
		u += t1 - t0;
		__syncthreads();                                   // previous segment fully consumed (ring, per-thread slots)
		int next_load = t0;                                // producer cursor (thread 0 only)
		if (tid == 0) {
			for (int k = 0; k < KF_STAGES - 1 && next_load < t1; k++, next_load++) {
				const int ls = (it + k) % KF_STAGES;
				mbar_expect_tx(&full_bar[ls], KF_TILE_BYTES);
				tma_load_1d(ring + (size_t)ls * KF_TILE_FLOATS, p.tiles7 + (size_t)next_load * KF_TILE_FLOATS, KF_TILE_BYTES, &full_bar[ls]);
			}
		}
		{
#pragma unroll
			for (int s = 0; s < S; s++) {
				const int i = sb * SB + s * THREADS + tid;
				const float x = p.px[i], y = p.py[i], z = p.pz[i];
				ox_s[s * THREADS] = x; oy_s[s * THREADS] = y; oz_s[s * THREADS] = z;
				const float pcx = __fsub_rn(x, p.cx), pcy = __fsub_rn(y, p.cy), pcz = __fsub_rn(z, p.cz);
				float p2;
				if (DIMS == 3) {
					ax[s] = -2.0f * pcx; ay[s] = -2.0f * pcy; az[s] = -2.0f * pcz;
					p2 = __fmaf_rn(pcz, pcz, __fmaf_rn(pcx, pcx, __fmul_rn(pcy, pcy)));
				} else {
					const float pa = (axA == 0) ? pcx : pcy, pb = (axB == 2) ? pcz : pcy;
					ax[s] = -2.0f * pa; ay[s] = -2.0f * pb; az[s] = 0.0f;
					p2 = __fmaf_rn(pa, pa, __fmul_rn(pb, pb));
				}
				const float p2lo = __fmul_rd(p2, 1.0f - 8.0f * KF_U);
				const float rp = __fmul_ru(__fsqrt_ru(p2), 1.0f + 8.0f * KF_U);
				// eps = 1.05 u (8 Rq^2 + 10 Rp Rq + 2 Rp^2), every operation rounded up
				float e = __fmul_ru(8.0f * p.rq, p.rq);
				e = __fmaf_ru(10.0f * rp, p.rq, e);
				e = __fmaf_ru(2.0f * rp, rp, e);
				e = __fmul_ru(e, 1.05f * KF_U);
				kk[s] = __fsub_ru(e, p2lo);
				// warm start: exact distance to the previous correspondence is an upper bound of the minimum
				float th = p.thr0;
				if (p.seed_idx != nullptr && i < p.n) {
					const int j0 = p.seed_idx[i];
					if (j0 >= 0 && j0 < p.m) {
						const float4 q = __ldg(p.q4 + j0);
						const float us = dist_chain(x, y, z, q.x, q.y, q.z);
						// strictly above the seed value so that the seed itself (or an equal, lower-indexed target) is
						// found by the exact pass; in sqrt mode also above every float sharing its square root
						const float up = (MODE == ICPB_DIST_SQRT) ? __fmul_ru(us, one8u) : us;
						const float nx = __uint_as_float(__float_as_uint(up) + 1u);
						if (us == us && nx < th) th = nx;
					}
				}
				thr_s[s * THREADS] = th; best_s[s * THREADS] = -1;
				tau[s] = __fadd_ru(__fmul_ru(th, one8u), kk[s]);
			}
		}
		for (int t = t0; t < t1; t++, it++) {
		const int stage = it % KF_STAGES;
		const uint32_t parity = (uint32_t)((it / KF_STAGES) & 1);

		__syncthreads();                                   // every thread has finished tile it-1: its slot may be refilled
		if (tid == 0 && next_load < t1) {
			const int ls = (it + KF_STAGES - 1) % KF_STAGES;
			mbar_expect_tx(&full_bar[ls], KF_TILE_BYTES);
			tma_load_1d(ring + (size_t)ls * KF_TILE_FLOATS, p.tiles7 + (size_t)next_load * KF_TILE_FLOATS, KF_TILE_BYTES, &full_bar[ls]);
			next_load++;
		}
		mbar_wait(&full_bar[stage], parity);

		const float* tile = ring + (size_t)stage * KF_TILE_FLOATS;
		const float4* X4  = reinterpret_cast<const float4*>(tile);
		const float4* Y4  = X4 + KF_TT / 4;
		const float4* Z4  = Y4 + KF_TT / 4;
		// filter operands: DIMS == 3: Xc Yc Zc W3;  DIMS == 2: the two kept centred coordinates and W2
		const float4* XC4 = Z4 + (size_t)(1 + axA) * (KF_TT / 4);
		const float4* YC4 = Z4 + (size_t)(1 + axB) * (KF_TT / 4);
		const float4* ZC4 = Z4 + 3 * (KF_TT / 4);
		const float4* W4  = Z4 + (size_t)(DIMS == 3 ? 4 : 5) * (KF_TT / 4);

#pragma unroll 1
		for (int sub = 0; sub < SUBS; sub++) {
			const int j0 = sub * (KF_TRK / 4), j1 = j0 + KF_TRK / 4;
			float em[S];
#pragma unroll
			for (int s = 0; s < S; s++) em[s] = inf;
			float4 X = XC4[j0], Y = YC4[j0], Z = (DIMS == 3) ? ZC4[j0] : make_float4(0.f, 0.f, 0.f, 0.f), W = W4[j0];
#pragma unroll 2
			for (int j = j0; j < j1; j++) {
				// prefetch of the next quad: j + 1 may be the first quad of the next array (never past the ring + pad)
				const float4 Xn = XC4[j + 1], Yn = YC4[j + 1], Wn = W4[j + 1];
				float4 Zn = Z;
				if (DIMS == 3) Zn = ZC4[j + 1];
				const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
				const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
				const u64 w01 = pack2(W.x, W.y), w23 = pack2(W.z, W.w);
				if (DIMS == 3) {
					const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
					for (int s = 0; s < S; s++) {
						const u64 AX = bcast2v(ax[s]), AY = bcast2v(ay[s]), AZ = bcast2v(az[s]);
						u64 e = fma2(AX, x01, fma2(AY, y01, fma2(AZ, z01, w01)));
						float a, b;
						unpack2(e, a, b);
						em[s] = min3(em[s], a, b);
						e = fma2(AX, x23, fma2(AY, y23, fma2(AZ, z23, w23)));
						unpack2(e, a, b);
						em[s] = min3(em[s], a, b);
					}
				} else {
#pragma unroll
					for (int s = 0; s < S; s++) {
						const u64 AX = bcast2v(ax[s]), AY = bcast2v(ay[s]);
						u64 e = fma2(AX, x01, fma2(AY, y01, w01));
						float a, b;
						unpack2(e, a, b);
						em[s] = min3(em[s], a, b);
						e = fma2(AX, x23, fma2(AY, y23, w23));
						unpack2(e, a, b);
						em[s] = min3(em[s], a, b);
					}
				}
				X = Xn; Y = Yn; Z = Zn; W = Wn;
			}
			// which (warp, source) pairs cannot rule this sub-tile out?
			unsigned need = 0;
#pragma unroll
			for (int s = 0; s < S; s++) {
				const unsigned bal = __ballot_sync(0xffffffffu, em[s] <= tau[s]);
				if (bal) need |= (1u << s);
			}
			n_tests += S;
			if (need) {
				const int gsub = t * SUBS + sub;
#pragma unroll
				for (int s = 0; s < S; s++) {
					if (need & (1u << s)) {          // warp-uniform
						n_exact += 1;
						const float sx = ox_s[s * THREADS], sy = oy_s[s * THREADS], sz = oz_s[s * THREADS];
						const float th = thr_s[s * THREADS];
						const u64 PX = pack2(sx, sx), PY = pack2(sy, sy), PZ = pack2(sz, sz);
						float mm = th;
#pragma unroll 4
						for (int j = j0; j < j1; j++) {
							const float4 Xo = X4[j], Yo = Y4[j], Zo = Z4[j];
							u64 dx = sub2(PX, pack2(Xo.x, Xo.y)), dy = sub2(PY, pack2(Yo.x, Yo.y)), dz = sub2(PZ, pack2(Zo.x, Zo.y));
							u64 d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
							float a, b;
							unpack2(d, a, b);
							mm = min3(mm, a, b);
							dx = sub2(PX, pack2(Xo.z, Xo.w)); dy = sub2(PY, pack2(Yo.z, Yo.w)); dz = sub2(PZ, pack2(Zo.z, Zo.w));
							d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
							unpack2(d, a, b);
							mm = min3(mm, a, b);
						}
						if (mm < th) {
							const float nt_ = lower_threshold<MODE>(mm);
							thr_s[s * THREADS] = nt_; best_s[s * THREADS] = gsub;
							tau[s] = __fadd_ru(__fmul_ru(nt_, one8u), kk[s]);
						}
					}
				}
			}
		}
		}
		flush(sb);
		}
	}
	if (p.stats != nullptr) {
		for (int o = 16; o > 0; o >>= 1) { n_tests += __shfl_xor_sync(0xffffffffu, n_tests, o); n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o); }
		if ((tid & 31) == 0) { atomicAdd(p.stats, n_tests / 32); atomicAdd(p.stats + 1, n_exact / 32); }
	}
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
// Filter data of the current target: bounding box -> centre, the axis the planar bound leaves out, the 8-array tiles
// and the radius bound. Every buffer is kept across targets of the same size (a host-driven loop uploads the target at
// every step: no cudaMalloc / cudaFree there — they cost up to hundreds of milliseconds when they hit).
static int build_filter_data(Ctx* c)
{
	const int m = c->m;
	const int nt = (m + KF_TT - 1) / KF_TT;
	if (!c->kf_scratch) ICPB_CUDA(c, cudaMalloc((void**)&c->kf_scratch, 16 * sizeof(unsigned long long)));
	unsigned* scratch = reinterpret_cast<unsigned*>(c->kf_scratch);                    // [0..5] bbox, [6] r2max
	unsigned long long* score = c->kf_scratch + 8;                                     // [3]
	unsigned init[16];
	memset(init, 0, sizeof init);
	init[0] = init[1] = init[2] = 0xffffffffu;
	ICPB_CUDA(c, cudaMemcpyAsync(scratch, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
	ICPB_CUDA(c, cudaMemsetAsync(score, 0, 3 * sizeof(unsigned long long), c->stream));
	kf_bbox_kernel<<<c->sm_count, 256, 0, c->stream>>>(c->q4, m, scratch);
	c->launches++;
	// K1T (the default) needs the centre and the radius bound only; the planar axis score and K1F's 32-byte-per-target tiles
	// are built when the FP32 filter kernel is the one in use
	const bool fp32_tiles = !c->k1_use_tc;
	// the axis the planar bound leaves out: the projection that keeps the sub-tiles best separated
	if (fp32_tiles) {
		kf_score_kernel<<<(m + KF_TRK - 1) / KF_TRK, KF_TRK, 0, c->stream>>>(c->q4, m, score);
		c->launches++;
	}
	unsigned h[8];
	unsigned long long hs[3] = { 0, 0, 0 };
	ICPB_CUDA(c, cudaMemcpyAsync(h, scratch, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaMemcpyAsync(hs, score, sizeof hs, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	auto dec = [](unsigned u) { unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
	for (int k = 0; k < 3; k++) {
		const float lo = dec(h[k]), hi = dec(h[3 + k]);
		float ctr = 0.5f * lo + 0.5f * hi;
		if (!std::isfinite(ctr)) ctr = 0.0f;       // non-finite coordinates: any centre is valid, only the bound's tightness changes
		c->kf_center[k] = ctr;
	}
	int drop = 2;
	if (hs[1] < hs[drop]) drop = 1;
	if (hs[0] < hs[drop]) drop = 0;
	if (c->kf_drop_forced >= 0 && c->kf_drop_forced <= 2) drop = c->kf_drop_forced;
	c->kf_drop = drop;
	for (int k = 0; k < 3; k++) c->kf_score[k] = (double)hs[k];
	if (fp32_tiles && nt > c->kf_tiles_cap) {
		cudaFree(c->kf_tiles7); c->kf_tiles7 = nullptr; c->kf_tiles_cap = 0;
		ICPB_CUDA(c, cudaMalloc((void**)&c->kf_tiles7, sizeof(float) * (size_t)nt * KF_TILE_FLOATS));
		c->kf_tiles_cap = nt;
	}
	kf_pack_kernel<<<(nt * KF_TT + 255) / 256, 256, 0, c->stream>>>(c->q4, m, nt * KF_TT, c->kf_center[0], c->kf_center[1], c->kf_center[2], c->kf_drop,
	                                                                 fp32_tiles ? c->kf_tiles7 : nullptr, scratch + 6);
	c->kf_tiles_built = fp32_tiles;
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	ICPB_CUDA(c, cudaMemcpyAsync(h, scratch, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	float r2; memcpy(&r2, &h[6], 4);
	// Rq >= max |q - c|: the float chain value is within (1 +- 3u) of the real one
	c->kf_rq = nextafterf(sqrtf(r2) * (1.0f + 8.0f * KF_U), INFINITY);
	c->kf_nt = nt;
	c->kf_dims = 2; c->kf_bounces = 0; c->kf_hold = 0;     // a new target: optimistic again
	// K1T's group size starts over for a NEW target; the same cloud uploaded again (a host-driven loop re-uploads the target
	// at every step: same size, bit-identical centre and radius) keeps what the exact-pass rate has taught. The start:
	// the exact pass works on quarters of 32 TPC consecutive targets, which should stay a fraction of a scan line —
	// measured best on the raster saddle with exact-pass units of 8 columns: 2 at 128^2, 8 at 317^2, 16 at 1000^2 points, i.e.
	// the power of two at or above sqrt(m) / 64
	const float fp[4] = { c->kf_center[0], c->kf_center[1], c->kf_center[2], c->kf_rq };
	if (c->m != c->kt_policy_m || memcmp(fp, c->kt_policy_fp, sizeof fp) != 0) {
		int t = 1;
		while (t < c->kt_tpc_start && (double)t * 64.0 < sqrt((double)c->m)) t *= 2;
		c->kt_tpc_auto = c->kt_tpc_forced_start ? c->kt_tpc_start : t;
		c->kt_policy_m = c->m; memcpy(c->kt_policy_fp, fp, sizeof fp);
	}
	if (!c->kf_stats) { ICPB_CUDA(c, cudaMalloc((void**)&c->kf_stats, 4 * sizeof(unsigned long long))); ICPB_CUDA(c, cudaMemsetAsync(c->kf_stats, 0, 4 * sizeof(unsigned long long), c->stream)); }
	c->kf_ready = true;
	return ICPB_OK;
}

// Builds whatever the matching method of the coming run needs, OUTSIDE the timed loop (icpb_run calls this before it
// records its first event; launch_match_filter still builds lazily for the step-wise API).
int prepare_match_filter(Ctx* c)
{
	if (c->kf_ready || c->m <= 0) return ICPB_OK;
	return build_filter_data(c);
}

static float sqrt_domain_threshold_f(float sentinel)
{
	if (!(sentinel > 0.0f)) return 0.0f;
	float y = sentinel * sentinel;
	if (std::isinf(y)) return y;
	while (sqrtf(y) < sentinel) y = nextafterf(y, INFINITY);
	while (y > 0.0f && sqrtf(nextafterf(y, 0.0f)) >= sentinel) y = nextafterf(y, 0.0f);
	return y;
}

// S sources per thread: 8 (two CTAs per SM, <= 128 registers) or 16 (one CTA per SM, the register file to itself;
// ICPB_KF_S=16 — the bare planar loop is ~5 % faster with 16 in tools/ubench_filter.cu).
template <int S, int MINB> static int launch_filter_cfg(Ctx* c, int dist_mode, float sentinel)
{
	constexpr int THREADS = 256;
	constexpr int SB = S * THREADS;
	KFParams p;
	p.px = c->px; p.py = c->py; p.pz = c->pz;
	p.tiles7 = c->kf_tiles7; p.q4 = c->q4; p.seed_idx = c->kf_use_seed ? c->seed : nullptr; p.keys = c->keys;
	p.n = c->n; p.m = c->m; p.nt = c->kf_nt;
	const int nb = (c->n + SB - 1) / SB;
	p.thr0 = (dist_mode == ICPB_DIST_SQRT) ? sqrt_domain_threshold_f(sentinel) : sentinel;
	p.cx = c->kf_center[0]; p.cy = c->kf_center[1]; p.cz = c->kf_center[2]; p.rq = c->kf_rq;
	p.done = &c->st->done;
	p.stats = c->kf_stats;
	p.drop = c->kf_drop;
	// planar bound once the thresholds are warm (a pass that starts from the sentinel sends ~6 % of the tests to the exact
	// chain even with the full bound); ICPB_KF_DIMS=2|3 forces one of them
	int dims = (c->kf_use_seed && c->kf_seeded) ? c->kf_dims : 3;
	if (c->kf_dims_forced == 2 || c->kf_dims_forced == 3) dims = c->kf_dims_forced;
	c->kf_dims_last = dims;
	if (dims == 3 && c->kf_hold > 0) c->kf_hold--;
	c->kf_seeded = true;
	c->pairs_acc += (double)c->n * (double)c->m;
	const size_t smem = (size_t)KF_STAGES * KF_TILE_BYTES + 64 + (size_t)5 * S * THREADS * 4;
	const bool sq = dist_mode == ICPB_DIST_SQRT;
	auto kern = (dims == 3) ? (sq ? k1_filter<S, THREADS, ICPB_DIST_SQRT, MINB, 3> : k1_filter<S, THREADS, ICPB_DIST_SQ, MINB, 3>)
	                        : (sq ? k1_filter<S, THREADS, ICPB_DIST_SQRT, MINB, 2> : k1_filter<S, THREADS, ICPB_DIST_SQ, MINB, 2>);
	static int cached_per_sm[4][64] = {};     // one table per <S, MINB> instantiation
	int& per_sm = cached_per_sm[(sq ? 1 : 0) + (dims == 3 ? 2 : 0)][c->device & 63];
	if (per_sm == 0) {
		ICPB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		ICPB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
		if (per_sm < 1) per_sm = 1;
	}
	long long grid = (long long)c->sm_count * per_sm;
	if (c->k1_grid_override > 0) grid = c->k1_grid_override;
	// guided self-scheduling between 2 and 32 tiles per grab (measured at 1M x 1M with fixed chunks: 16-64 tiles 118 ms,
	// 256: 120 ms, 1024: 126 ms); ICPB_KF_CHUNK=k pins both bounds to k (fixed chunks, for experiments)
	p.total_units = (long long)nb * p.nt;
	p.min_chunk = 2; p.max_chunk = 32; p.gss_div = 4;      // tools/sweep_gss.py: flat within 1 % around these; 128+ tiles per grab hurts small shards
	if (c->kf_gss[0] > 0) { p.min_chunk = c->kf_gss[0]; p.max_chunk = c->kf_gss[1]; p.gss_div = c->kf_gss[2]; }     // ICPB_KF_GSS=min,max,div
	if (c->kf_chunk_override > 0) p.min_chunk = p.max_chunk = c->kf_chunk_override;
	if (p.max_chunk > p.nt) p.max_chunk = p.nt;
	if (p.min_chunk > p.max_chunk) p.min_chunk = p.max_chunk;
	const long long max_ctas = (p.total_units + p.min_chunk - 1) / p.min_chunk;
	if (grid > max_ctas) grid = max_ctas;
	if (!c->kf_work_counter) ICPB_CUDA(c, cudaMalloc((void**)&c->kf_work_counter, sizeof(unsigned long long)));
	p.work_counter = c->kf_work_counter;
	ICPB_CUDA(c, cudaMemsetAsync(c->kf_work_counter, 0, sizeof(unsigned long long), c->stream));
	kern<<<(unsigned)grid, THREADS, smem, c->stream>>>(p);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

int launch_match_filter(Ctx* c, int dist_mode, float sentinel)
{
	if (c->n <= 0 || c->m <= 0) return ICPB_OK;
	int rc;
	if (!c->kf_ready) { if ((rc = build_filter_data(c)) != ICPB_OK) return rc; }
	// The bound's error analysis is relative (u = 2^-24 per operation): it needs finite magnitudes whose squares stay
	// in the normal range. Anything else (non-finite coordinates, clouds of radius > 1e15 or < 1e-15) goes through
	// the direct kernel.
	if (!std::isfinite(c->kf_rq) || c->kf_rq > 1e15f || c->kf_rq < 1e-15f) return launch_match_brute(c, dist_mode, sentinel);
	if (c->k1_use_tc) return launch_match_filter_tc(c, dist_mode, sentinel);      // K1T: the same bound on the tensor cores (nn_filter_tc.cu)
	if (!c->kf_tiles_built) { if ((rc = build_filter_data(c)) != ICPB_OK) return rc; }
	return (c->kf_s == 16) ? launch_filter_cfg<16, 1>(c, dist_mode, sentinel) : launch_filter_cfg<8, 2>(c, dist_mode, sentinel);
}

// Called by the engine whenever it has synchronised with the stream anyway (read_state): looks at the share of
// (warp, source, sub-tile) tests that needed the exact chain since the last look and picks the bound for the next
// launches. Planar -> full above 10 %; full -> planar below 1 %, but only after 16, 32, 64 launches on the full bound
// and at most 3 times per target: a cloud for which the planar bound keeps failing stays on the full one. Results
// never depend on this choice.
int kf_policy_update(Ctx* c)
{
	if (!c->kf_ready || !c->kf_stats) return ICPB_OK;
	unsigned long long h[4] = { 0, 0, 0, 0 };
	ICPB_CUDA(c, cudaMemcpyAsync(h, c->kf_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->k1_use_tc) { const int rc = filter_tc_check(c); if (rc != ICPB_OK) return rc; }
	const double dt = (double)(h[0] - c->kf_stats_seen[0]), de = (double)(h[1] - c->kf_stats_seen[1]);
	const double dn = (double)(h[2] - c->kf_stats_seen[2]), ds = (double)(h[3] - c->kf_stats_seen[3]);
	c->kf_stats_seen[0] = h[0]; c->kf_stats_seen[1] = h[1]; c->kf_stats_seen[2] = h[2]; c->kf_stats_seen[3] = h[3];
	if (dt <= 0.0) return ICPB_OK;
	const double frac = de / dt;
	c->kf_last_frac = frac;
	if (c->kf_dims_last == 4) {
		// K1T: grouped columns (up to 16 consecutive targets per MMA column) pay off while consecutive targets are neighbours
		// in space. A quarter test costs the same whatever the group size, the exact pass it may ask for covers 32 x TPC
		// targets: the group size is halved for this target once exact passes cost about as much as the tests (0.8 / TPC of them)
		// — but only when the groups are what asks for them: the exact pass is wanted wherever a group reaches into the ball of
		// radius s0 around a source, and while the registration is far from converged (or on a cold pass) s0 is much larger
		// than any group; smaller groups would not help then. The kernel counts the sources whose s0 is below the largest
		// group radius: the groups drive the exact-pass rate when that is most of them.
		const bool groups_matter = dn > 0.0 && ds > 0.5 * dn;
		if (c->kt_variant < 0 && c->kt_tpc_auto > 1 && groups_matter && frac * (double)c->kt_tpc_auto > 0.8) c->kt_tpc_auto /= 2;
		return ICPB_OK;
	}
	if (c->kf_dims_last == 2 && frac > 0.10) { c->kf_dims = 3; c->kf_bounces++; c->kf_hold = 8 << c->kf_bounces; }
	else if (c->kf_dims_last == 3 && c->kf_dims == 3 && frac < 0.01 && c->kf_bounces < 3 && c->kf_hold <= 0) c->kf_dims = 2;
	return ICPB_OK;
}

} // namespace icpb

extern "C" int icpb_get_filter_config(icpb_ctx* ctx, int* dims_next, int* dims_last, int* drop_axis, double* last_exact_fraction)
{
	using namespace icpb;
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = reinterpret_cast<Ctx*>(ctx);
	if (dims_next) *dims_next = (c->kf_dims_forced == 2 || c->kf_dims_forced == 3) ? c->kf_dims_forced : c->kf_dims;
	if (dims_last) *dims_last = c->kf_dims_last;
	if (drop_axis) *drop_axis = c->kf_ready ? c->kf_drop : -1;
	if (last_exact_fraction) *last_exact_fraction = c->kf_last_frac;
	return ICPB_OK;
}

extern "C" int icpb_get_filter_tc_config(icpb_ctx* ctx, int* enabled, int* targets_per_column)
{
	using namespace icpb;
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = reinterpret_cast<Ctx*>(ctx);
	if (enabled) *enabled = (c->k1_use_tc && c->k1_use_filter) ? 1 : 0;
	if (targets_per_column) *targets_per_column = (c->kt_variant >= 0) ? c->kt_tpc : c->kt_tpc_auto;
	return ICPB_OK;
}

extern "C" int icpb_get_filter_tc_order(icpb_ctx* ctx, int* morton_order)
{
	using namespace icpb;
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = reinterpret_cast<Ctx*>(ctx);
	if (morton_order) *morton_order = (c->kt_ready && c->kt_sorted) ? 1 : 0;
	return ICPB_OK;
}

extern "C" int icpb_get_filter_stats(icpb_ctx* ctx, double* subtile_tests, double* subtile_exact)
{
	using namespace icpb;
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = reinterpret_cast<Ctx*>(ctx);
	ICPB_CUDA(c, cudaSetDevice(c->device));
	unsigned long long h[2] = { 0, 0 };
	if (c->kf_stats) {
		ICPB_CUDA(c, cudaMemcpyAsync(h, c->kf_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	if (subtile_tests) *subtile_tests = (double)h[0];
	if (subtile_exact) *subtile_exact = (double)h[1];
	return ICPB_OK;
}
