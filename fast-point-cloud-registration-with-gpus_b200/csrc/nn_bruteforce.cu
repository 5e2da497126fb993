// nn_bruteforce.cu — K1: exact brute-force nearest-neighbour matching for sm_100a.
//
// Replaces `Matching<<<>>>` of the reference (src/ICP_point_to_point.cu:31-57; sqrt variant
// src/ICP_point_to_plane.cu:163-181; double pow/sqrt variant src/ICP_standard.cu:21-39) with
// bit-identical correspondences: idx[i] = argmin_j d(P_i, Q_j), lowest j on ties, nothing written
// when no distance is below the sentinel.
//
// Design:
//  * FP32-pipe bound by construction: the reference's distance is the rounded chain
//        dx = xp-xq, dy = yp-yq, dz = zp-zq,  d = fma(dz,dz, fma(dx,dx, dy*dy))
//    (6 FP32 operations per pair); it is evaluated with the packed sm_100 instructions
//    FADD2 / FMUL2 / FFMA2 (two targets per instruction, source coordinate broadcast), which are
//    IEEE-identical per element to the scalar ops but halve the issue slots, leaving room for the
//    min on the ALU pipe.
//  * The inner loop keeps ONLY a running minimum per source (one FMNMX3 per two pairs). Which
//    target produced it is recovered later: after every sub-tile of K1_TRK targets the thread
//    checks `m < thr` (strict, so an equal distance in a later sub-tile never replaces an earlier
//    one — the reference's tie rule) and remembers the sub-tile; when a block of sources is done
//    each thread re-scans its one remembered sub-tile with the same arithmetic to find the first
//    index that attains the minimum.
//  * sqrt mode never takes a square root in the inner loop: sqrt.rn is monotone, so the running
//    compare `sqrt(d) < smin` is done in the squared domain against thr = the smallest float whose
//    sqrt.rn equals smin (recomputed only when the minimum improves).
//  * Targets stream through a 3-stage shared-memory ring filled by 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier); every lane reads the same float4 (broadcast LDS.128).
//  * Work = (source block x target tile) units, flattened and split evenly over a persistent grid
//    of (SMs x resident CTAs) blocks (stream-K style), so small clouds still fill 148 SMs; partial
//    results of CTAs that share a source block are merged with one 64-bit atomicMin per source on
//    (distance bits << 32 | index), which is exactly "smallest distance, then lowest index".
#include "common.cuh"
#include "k1_device.cuh"
#include <cmath>

namespace icpb {

// ------------------------------------------------------------------------------------------------
// K1 main kernel
// ------------------------------------------------------------------------------------------------
// VAR bits (tuning variants, selected at run time through the config table):
//   1: re-materialise the broadcast source operand inside the loop (asm volatile), which makes ptxas use
//      the scalar-broadcast operand form (R.F32) instead of keeping duplicated register pairs
//   2: software-prefetch the next target quad from shared memory into registers
//   4: keep thr[] / best[] in shared memory (touched once per sub-tile) instead of registers
template <int S, int THREADS, int MODE, int MINB, int VAR>
__global__ void __launch_bounds__(THREADS, MINB) k1_match(const K1Params p)
{
	constexpr int SB   = S * THREADS;          // sources per block
	constexpr int SUBS = K1_TT / K1_TRK;       // tracking sub-tiles per tile
	if (p.done != nullptr && *p.done) return;

	extern __shared__ __align__(128) unsigned char k1_smem[];
	__shared__ __align__(8) uint64_t full_bar[K1_STAGES];
	float* ring = reinterpret_cast<float*>(k1_smem);

	const int tid = threadIdx.x;
	int n_eff = p.n;
	long long units = p.units;
	if (p.count_dev != nullptr) {            // compact source list whose length is only known on the device
		n_eff = *p.count_dev;
		units = (long long)((n_eff + SB - 1) / SB) * p.nt;
	}
	const long long u0 = (units * (long long)blockIdx.x) / gridDim.x;
	const long long u1 = (units * (long long)(blockIdx.x + 1)) / gridDim.x;
	if (u0 >= u1) return;

	if (tid == 0) {
		for (int s = 0; s < K1_STAGES; s++) mbar_init(&full_bar[s], 1);
		fence_mbar_init();
	}
	__syncthreads();

	long long next_load = u0;   // producer cursor (thread 0 only)
	if (tid == 0) {
		for (int k = 0; k < K1_STAGES - 1 && next_load < u1; k++, next_load++) {
			const int t = (int)(next_load % p.nt);
			mbar_expect_tx(&full_bar[k], K1_TILE_BYTES);
			tma_load_1d(ring + (size_t)k * 3 * K1_TT, p.qtiles + (size_t)t * 3 * K1_TT, K1_TILE_BYTES, &full_bar[k]);
		}
	}

	float sx[S], sy[S], sz[S], m[S];
	float thr_r[(VAR & 4) ? 1 : S];
	int   best_r[(VAR & 4) ? 1 : S];
	// VAR&4: per-thread thr/best slots live behind the ring in dynamic shared memory, [s][tid] (conflict free)
	float* thr_s  = reinterpret_cast<float*>(k1_smem + (size_t)K1_STAGES * K1_TILE_BYTES + 64) + tid;
	int*   best_s = reinterpret_cast<int*>(thr_s - tid + S * THREADS) + tid;
	auto get_thr  = [&](int s) -> float { if constexpr ((VAR & 4) != 0) return thr_s[s * THREADS]; else return thr_r[s]; };
	auto get_best = [&](int s) -> int { if constexpr ((VAR & 4) != 0) return best_s[s * THREADS]; else return best_r[s]; };
	auto set_thr  = [&](int s, float v) { if constexpr ((VAR & 4) != 0) thr_s[s * THREADS] = v; else thr_r[s] = v; };
	auto set_best = [&](int s, int v) { if constexpr ((VAR & 4) != 0) best_s[s * THREADS] = v; else best_r[s] = v; };
	int cur_sb = -1;

	auto flush = [&](int sb) {
#pragma unroll
		for (int s = 0; s < S; s++) {
			const int i = sb * SB + s * THREADS + tid;
			const int bs = get_best(s);
			if (i < n_eff && bs >= 0) {
				const float th = get_thr(s);
				const float* gx = p.qtiles + (size_t)(bs / SUBS) * 3 * K1_TT + (size_t)(bs % SUBS) * K1_TRK;
				const float4* X4 = reinterpret_cast<const float4*>(gx);
				const float4* Y4 = reinterpret_cast<const float4*>(gx + K1_TT);
				const float4* Z4 = reinterpret_cast<const float4*>(gx + 2 * K1_TT);
				const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(th) : th;
				int found = -1;
				for (int j = 0; j < K1_TRK / 4 && found < 0; j++) {
					const float4 X = __ldg(X4 + j), Y = __ldg(Y4 + j), Z = __ldg(Z4 + j);
					float d0 = dist_chain(sx[s], sy[s], sz[s], X.x, Y.x, Z.x);
					float d1 = dist_chain(sx[s], sy[s], sz[s], X.y, Y.y, Z.y);
					float d2 = dist_chain(sx[s], sy[s], sz[s], X.z, Y.z, Z.z);
					float d3 = dist_chain(sx[s], sy[s], sz[s], X.w, Y.w, Z.w);
					if (MODE == ICPB_DIST_SQRT) { d0 = __fsqrt_rn(d0); d1 = __fsqrt_rn(d1); d2 = __fsqrt_rn(d2); d3 = __fsqrt_rn(d3); }
					if (d0 <= target) found = 4 * j;
					else if (d1 <= target) found = 4 * j + 1;
					else if (d2 <= target) found = 4 * j + 2;
					else if (d3 <= target) found = 4 * j + 3;
				}
				if (found >= 0) {
					const u64 key = ((u64)__float_as_uint(target) << 32) | (u64)(uint32_t)(bs * K1_TRK + found);
					atomicMin(p.keys + (p.remap != nullptr ? p.remap[i] : i), key);
				}
			}
		}
	};

	for (long long u = u0; u < u1; u++) {
		const int it    = (int)(u - u0);
		const int stage = it % K1_STAGES;
		const uint32_t parity = (uint32_t)((it / K1_STAGES) & 1);
		const int sb = (int)(u / p.nt);
		const int t  = (int)(u % p.nt);

		if (sb != cur_sb) {
			if (cur_sb >= 0) flush(cur_sb);
#pragma unroll
			for (int s = 0; s < S; s++) {
				const int i = sb * SB + s * THREADS + tid;
				int gi = i;                                   // identity: the arrays are padded to a block multiple
				if (p.remap != nullptr) gi = (i < n_eff) ? p.remap[i] : 0;
				sx[s] = p.px[gi]; sy[s] = p.py[gi]; sz[s] = p.pz[gi];
				m[s] = p.thr0; set_thr(s, p.thr0); set_best(s, -1);
			}
			cur_sb = sb;
		}

		__syncthreads();   // every thread has finished tile it-1: its ring slot may be refilled
		if (tid == 0 && next_load < u1) {
			const int ls = (it + K1_STAGES - 1) % K1_STAGES;
			const int lt = (int)(next_load % p.nt);
			mbar_expect_tx(&full_bar[ls], K1_TILE_BYTES);
			tma_load_1d(ring + (size_t)ls * 3 * K1_TT, p.qtiles + (size_t)lt * 3 * K1_TT, K1_TILE_BYTES, &full_bar[ls]);
			next_load++;
		}
		mbar_wait(&full_bar[stage], parity);

		const float4* X4 = reinterpret_cast<const float4*>(ring + (size_t)stage * 3 * K1_TT);
		const float4* Y4 = X4 + K1_TT / 4;
		const float4* Z4 = Y4 + K1_TT / 4;

#pragma unroll 1
		for (int sub = 0; sub < SUBS; sub++) {
			const int j0 = sub * (K1_TRK / 4), j1 = j0 + K1_TRK / 4;
			float4 X = X4[j0], Y = Y4[j0], Z = Z4[j0];
#pragma unroll 2
			for (int j = j0; j < j1; j++) {
				float4 Xn, Yn, Zn;
				if constexpr ((VAR & 2) != 0) { Xn = X4[j + 1]; Yn = Y4[j + 1]; Zn = Z4[j + 1]; }   // (reads 16 B past the sub-tile: still inside the ring + pad)
				else if (j > j0) { X = X4[j]; Y = Y4[j]; Z = Z4[j]; }
				const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
				const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
				const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
				for (int s = 0; s < S; s++) {
					u64 PX, PY, PZ;
					if constexpr ((VAR & 1) != 0) { PX = bcast2v(sx[s]); PY = bcast2v(sy[s]); PZ = bcast2v(sz[s]); }
					else { PX = pack2(sx[s], sx[s]); PY = pack2(sy[s], sy[s]); PZ = pack2(sz[s], sz[s]); }
					u64 dx = sub2(PX, x01), dy = sub2(PY, y01), dz = sub2(PZ, z01);
					u64 d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
					float a, b;
					unpack2(d, a, b);
					m[s] = min3(m[s], a, b);
					dx = sub2(PX, x23); dy = sub2(PY, y23); dz = sub2(PZ, z23);
					d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
					unpack2(d, a, b);
					m[s] = min3(m[s], a, b);
				}
				if constexpr ((VAR & 2) != 0) { X = Xn; Y = Yn; Z = Zn; }
			}
			const int gsub = t * SUBS + sub;
#pragma unroll
			for (int s = 0; s < S; s++) {
				if (m[s] < get_thr(s)) {
					const float nt_ = lower_threshold<MODE>(m[s]);
					set_best(s, gsub);
					set_thr(s, nt_);
					m[s] = nt_;
				}
			}
		}
	}
	flush(cur_sb);
}

// ------------------------------------------------------------------------------------------------
// ICP_standard's formula: float differences, squares/sum/sqrt in double, result cast to float
// (src/ICP_standard.cu:31). FP64 throughout; only ever used at that program's 1024 points, so a plain
// one-thread-per-source kernel with shared-memory target tiles is all it needs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k1_match_std(const K1Params p)
{
	if (p.done != nullptr && *p.done) return;
	__shared__ float tx[256], ty[256], tz[256];
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = i < p.n;
	const float xp = valid ? p.px[i] : 0.f, yp = valid ? p.py[i] : 0.f, zp = valid ? p.pz[i] : 0.f;
	float mn = p.sentinel;
	int best = -1;
	for (int t = 0; t < p.nt; t++) {
		for (int c = 0; c < K1_TT; c += 256) {
			__syncthreads();
			for (int k = threadIdx.x; k < 256; k += blockDim.x) {
				const float* g = p.qtiles + (size_t)t * 3 * K1_TT + c + k;
				tx[k] = g[0]; ty[k] = g[K1_TT]; tz[k] = g[2 * K1_TT];
			}
			__syncthreads();
			for (int k = 0; k < 256; k++) {
				const float dx = __fsub_rn(xp, tx[k]), dy = __fsub_rn(yp, ty[k]), dz = __fsub_rn(zp, tz[k]);
				const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)),
				                           __dmul_rn((double)dz, (double)dz));
				const float d = (float)sqrt(s);
				if (d < mn) { mn = d; best = t * K1_TT + c + k; }
			}
		}
	}
	if (valid && best >= 0) p.keys[i] = ((u64)__float_as_uint(mn) << 32) | (u64)(uint32_t)best;
}

// ------------------------------------------------------------------------------------------------
// Small problems. k1_match works in units of (1024..2048 sources) x (1024 targets) and recovers indices with a scan of the
// winning sub-tile from L2 at the end: at the reference's own sizes (1 024 points = ONE unit = one SM busy, 56 us per pass
// measured, most of it the serial index scan) that is all latency. Here: one source per thread, a slice of the targets
// per block staged in shared memory, distance and index tracked together, one atomicMin on the (distance, index) key per
// thread — the keys give the lowest index among equal distances whatever order the slices finish in.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) k1_match_small(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                                                      const float4* __restrict__ q4, u64* keys, int n, int m, int ts, float sentinel, const int* done)
{
	if (done != nullptr && *done) return;
	extern __shared__ float4 k1s_q[];
	const int t0 = blockIdx.y * ts;
	const int cnt = min(m - t0, ts);
	for (int k = threadIdx.x; k < cnt; k += 128) k1s_q[k] = q4[t0 + k];
	__syncthreads();
	const int i = blockIdx.x * 128 + threadIdx.x;
	if (i >= n) return;
	const float x = px[i], y = py[i], z = pz[i];
	float best = sentinel;
	int bj = -1;
#pragma unroll 4
	for (int k = 0; k < cnt; k++) {
		const float4 q = k1s_q[k];
		float d;
		if (MODE == ICPB_DIST_STD) {                     // ICP_standard's formula (src/ICP_standard.cu:31), as k1_match_std evaluates it
			const float dx = __fsub_rn(x, q.x), dy = __fsub_rn(y, q.y), dz = __fsub_rn(z, q.z);
			d = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)), __dmul_rn((double)dz, (double)dz)));
		} else {
			d = dist_chain(x, y, z, q.x, q.y, q.z);
			if (MODE == ICPB_DIST_SQRT) d = __fsqrt_rn(d);
		}
		if (d < best) { best = d; bj = t0 + k; }        // strict: the first of equal distances stays (src/ICP_point_to_point.cu:47-55)
	}
	if (bj >= 0) atomicMin(keys + i, ((u64)__float_as_uint(best) << 32) | (u64)(uint32_t)bj);
}

__global__ void key_reset_kernel(u64* keys, int n, const int* done)
{
	if (done != nullptr && *done) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) keys[i] = KEY_UNMATCHED;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
struct K1Config { int S, threads, minb, var; };
// index = ICPB_K1_CFG (tuning); entry 0 is the default chosen from measurements on B200 (profiles/)
#define K1_CONFIGS(X) \
	X(0, 8, 256, 2, 0) X(1, 8, 512, 1, 0) X(2, 4, 256, 3, 0) X(3, 4, 512, 1, 0) \
	X(4, 8, 256, 2, 1) X(5, 8, 256, 2, 3) X(6, 8, 256, 2, 7) X(7, 8, 256, 3, 7) \
	X(8, 4, 256, 4, 7) X(9, 4, 512, 2, 7) X(10, 8, 512, 1, 7) X(11, 4, 256, 3, 3) \
	X(12, 8, 384, 1, 7) X(13, 6, 256, 3, 7) X(14, 4, 512, 1, 3) X(15, 8, 256, 2, 5)
static const K1Config k1_table[] = {
#define X(i, s, t, b, v) { s, t, b, v },
	K1_CONFIGS(X)
#undef X
};
constexpr int K1_NUM_CFG = (int)(sizeof(k1_table) / sizeof(k1_table[0]));

int k1_max_block_sources()
{
	int mx = 0;
	for (int k = 0; k < K1_NUM_CFG; k++) { int sbs = k1_table[k].S * k1_table[k].threads; if (sbs > mx) mx = sbs; }
	if (mx < 16 * 256) mx = 16 * 256;       // the filter kernel's largest source block (nn_filter.cu, S = 16)
	return mx;
}

// Smallest float y whose correctly-rounded square root is >= sentinel: `sqrt(d) < sentinel` <=> `d < y`.
static float sqrt_domain_threshold(float sentinel)
{
	if (!(sentinel > 0.0f)) return 0.0f;
	float y = sentinel * sentinel;
	if (std::isinf(y)) return y;
	while (sqrtf(y) < sentinel) y = nextafterf(y, INFINITY);
	while (y > 0.0f && sqrtf(nextafterf(y, 0.0f)) >= sentinel) y = nextafterf(y, 0.0f);
	return y;
}

template <int S, int THREADS, int MINB, int VAR>
static int launch_cfg(Ctx* c, int mode, const K1Params& base)
{
	K1Params p = base;
	constexpr int SB = S * THREADS;
	const int nb = (c->n + SB - 1) / SB;
	p.units = (long long)nb * p.nt;
	const size_t smem = (size_t)K1_STAGES * K1_TILE_BYTES + 64 + ((VAR & 4) ? (size_t)2 * S * THREADS * 4 : 0);
	auto kern = (mode == ICPB_DIST_SQRT) ? k1_match<S, THREADS, ICPB_DIST_SQRT, MINB, VAR> : k1_match<S, THREADS, ICPB_DIST_SQ, MINB, VAR>;
	// per (kernel, device) launch attributes are queried once: they sit on the host path of every iteration
	static int cached_per_sm[2][64] = {};
	int& per_sm = cached_per_sm[mode == ICPB_DIST_SQRT ? 1 : 0][c->device & 63];
	if (per_sm == 0) {
		ICPB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		ICPB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
		if (per_sm < 1) per_sm = 1;
	}
	long long grid = (long long)c->sm_count * per_sm;
	if (c->k1_grid_override > 0) grid = c->k1_grid_override;
	if (grid > p.units) grid = p.units;
	if (grid < 1) return ICPB_OK;
	kern<<<(unsigned)grid, THREADS, smem, c->stream>>>(p);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

int launch_key_reset(Ctx* c)
{
	if (c->n <= 0) return ICPB_OK;
	key_reset_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->keys, c->n, nullptr);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

static int launch_match_brute_impl(Ctx* c, int dist_mode, float sentinel, const int* remap, const int* count_dev)
{
	if (c->n <= 0 || c->m <= 0) return ICPB_OK;
	K1Params p;
	p.remap = remap; p.count_dev = count_dev;
	p.px = c->px; p.py = c->py; p.pz = c->pz;
	p.qtiles = c->qtiles; p.keys = c->keys;
	p.n = c->n; p.nt = c->nt; p.units = 0;
	p.sentinel = sentinel;
	p.thr0 = (dist_mode == ICPB_DIST_SQRT) ? sqrt_domain_threshold(sentinel) : sentinel;
	p.done = &c->st->done;
	if (remap == nullptr) c->pairs_acc += (double)c->n * (double)c->m;
	// too few (source block, tile) units to fill the GPU: the small-problem kernel
	const bool small = remap == nullptr && !c->k1_cfg_forced && (long long)((c->n + 1023) / 1024) * c->nt < 2LL * c->sm_count;
	if (dist_mode == ICPB_DIST_STD && !small) {
		k1_match_std<<<(c->n + 127) / 128, 128, 0, c->stream>>>(p);
		c->launches++;
		ICPB_CUDA(c, cudaGetLastError());
		return ICPB_OK;
	}
	if (small) {
		const int gx = (c->n + 127) / 128;
		int gy = (4 * c->sm_count + gx - 1) / gx;
		if (gy > (c->m + 63) / 64) gy = (c->m + 63) / 64;
		if (gy < 1) gy = 1;
		int ts = ((c->m + gy - 1) / gy + 63) / 64 * 64;
		if (ts > 3072) ts = 3072;                             // 48 KB of shared memory
		gy = (c->m + ts - 1) / ts;
		const dim3 grid((unsigned)gx, (unsigned)gy);
		if (dist_mode == ICPB_DIST_STD) k1_match_small<ICPB_DIST_STD><<<grid, 128, (size_t)ts * sizeof(float4), c->stream>>>(c->px, c->py, c->pz, c->q4, c->keys, c->n, c->m, ts, sentinel, &c->st->done);
		else if (dist_mode == ICPB_DIST_SQRT) k1_match_small<ICPB_DIST_SQRT><<<grid, 128, (size_t)ts * sizeof(float4), c->stream>>>(c->px, c->py, c->pz, c->q4, c->keys, c->n, c->m, ts, sentinel, &c->st->done);
		else k1_match_small<ICPB_DIST_SQ><<<grid, 128, (size_t)ts * sizeof(float4), c->stream>>>(c->px, c->py, c->pz, c->q4, c->keys, c->n, c->m, ts, sentinel, &c->st->done);
		c->launches++;
		ICPB_CUDA(c, cudaGetLastError());
		return ICPB_OK;
	}
	// small problems (< 1e9 pairs) are dominated by per-block prologue/epilogue: 4 sources per thread gives twice
	// the blocks per source and a cheaper index re-scan (measured: tools/sweep_small.py); ICPB_K1_CFG overrides
	int cfg = c->k1_cfg;
	if (!c->k1_cfg_forced && (double)c->n * (double)c->m < 1e9) cfg = 8;
	switch (cfg) {
#define X(i, s, t, b, v) case i: return launch_cfg<s, t, b, v>(c, dist_mode, p);
	K1_CONFIGS(X)
#undef X
	default: return launch_cfg<8, 256, 2, 0>(c, dist_mode, p);
	}
}

int launch_match_brute(Ctx* c, int dist_mode, float sentinel) { return launch_match_brute_impl(c, dist_mode, sentinel, nullptr, nullptr); }
int launch_match_brute_remap(Ctx* c, int dist_mode, float sentinel, const int* remap, const int* count_dev)
{
	return launch_match_brute_impl(c, dist_mode, sentinel, remap, count_dev);
}

} // namespace icpb
