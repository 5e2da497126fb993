// grid_tree.cuh — K1g far field: exact nearest neighbour by best-first descent of an occupancy pyramid built over the
// uniform grid (an octree whose leaves are the grid's cells). Host/device code: the same traversal is compiled for the
// GPU kernel (grid_nn.cu) and for the CPU harness that checks it against brute force (tools/grid_tree_host_test.cu).
//
// Why: the ring search of grid_nn.cu closes a source only if its neighbour lies within 2 cells; sources farther away
// (the first iterations of a registration) were handed to the brute-force kernel. The pyramid closes every source,
// whatever its distance, in O(depth) node tests plus the points of the few leaves that survive pruning.
//
// Exactness (same contract as every matching kernel: argmin of the reference's float chain, lowest index on ties,
// nothing at or above the sentinel): a node is skipped only if  lb * 0.998 > bound,  where lb is the squared distance
// from the source to the node's box reduced by `slack` per axis (covers the rounding of the cell assignment
// floor((v - o) * inv_h) and of the box corners; both are evaluated in the grid's own frame, coordinates minus the grid
// origin, so they are relative to the grid's extent and a cloud far from the origin prunes as well as a centred one),
// 0.998 covers the
// rounding of the float chain (relative 3 * 2^-24) and of lb itself, and `bound` is the squared-domain value of the
// best key so far widened to its whole sqrt class in sqrt mode. Equality never prunes, so an equally distant target
// with a lower index is still found. Candidates are compared as (distance bits << 32 | index) keys, independent of the
// visiting order. `nodes` (optional) counts node visits.
#pragma once
#include "common.cuh"
#include <cmath>

namespace icpb {

constexpr int GP_STACK = 7 * GP_MAX_L + 2;      // GridGeom / GridPyramid / GP_MAX_L: common.cuh

// Cell size: <= ~4 cells per point, <= 16M cells and <= GP_MAX_DIM cells along any axis (the pyramid has GP_MAX_L
// levels and packs node coordinates in 18 bits: a collinear or very thin cloud must not exceed that on its long axis).
// h grows until both limits hold — a flat cloud of large extent needs many more than a handful of steps because its
// flat axis is clamped to an absolute 1e-6 and the cube-root estimate is then far too small.
constexpr int GP_MAX_DIM = 1 << (GP_MAX_L - 1);
inline GridGeom compute_grid_geom(const float lo[3], const float hi[3], int m)
{
	float ext[3];
	for (int k = 0; k < 3; k++) { ext[k] = fmaxf(hi[k] - lo[k], 1e-6f); if (!(ext[k] < 3e38f)) ext[k] = 3e38f; }
	const double cells_max = fmin(fmax(4.0 * m, 4096.0), 16.0 * 1024 * 1024);
	double h = cbrt((double)ext[0] * ext[1] * ext[2] / cells_max);
	const double emax = fmax((double)ext[0], fmax((double)ext[1], (double)ext[2]));
	h = fmax(h, emax / (double)(GP_MAX_DIM - 1));          // the long axis alone
	if (!(h > 0.0)) h = 1e-6;
	for (int it = 0; it < 4096; it++) {
		const double dx = floor(ext[0] / h) + 1, dy = floor(ext[1] / h) + 1, dz = floor(ext[2] / h) + 1;
		if (dx * dy * dz <= cells_max && fmax(dx, fmax(dy, dz)) <= (double)GP_MAX_DIM) break;
		h *= 1.26;
	}
	GridGeom g;
	g.h = (float)h; g.inv_h = (float)(1.0 / h);
	g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
	g.nx = (int)floor(ext[0] / h) + 1; g.ny = (int)floor(ext[1] / h) + 1; g.nz = (int)floor(ext[2] / h) + 1;
	// h >= emax / (GP_MAX_DIM - 1) from the start, so every axis has at most GP_MAX_DIM cells: the pyramid reaches a single root
	return g;
}

// Level dimensions / offsets of the pyramid over a grid; returns the total number of occupancy bytes.
inline long long pyramid_layout(const GridGeom& g, GridPyramid& py)
{
	int nx = g.nx, ny = g.ny, nz = g.nz, L = 0;
	long long total = 0;
	while (true) {
		py.nx[L] = nx; py.ny[L] = ny; py.nz[L] = nz; py.off[L] = total;
		total += (long long)nx * ny * nz;
		if ((nx == 1 && ny == 1 && nz == 1) || L == GP_MAX_L - 1) break;
		nx = (nx + 1) >> 1; ny = (ny + 1) >> 1; nz = (nz + 1) >> 1; L++;
	}
	py.levels = L + 1;
	// box distances are evaluated in the grid's own frame (coordinates minus the grid origin, as the cell assignment
	// is), so every rounding involved is relative to the grid's EXTENT, not to how far the cloud is from the origin
	const float ext = fmaxf((float)g.nx, fmaxf((float)g.ny, (float)g.nz)) * g.h;
	py.slack = 4e-6f * (ext + g.h);
	return total;
}

__host__ __device__ __forceinline__ int gt_cell_coord(float v, float o, float inv_h) { return (int)floorf((v - o) * inv_h); }

__host__ __device__ __forceinline__ float gt_chain(float xp, float yp, float zp, float xq, float yq, float zq)
{
#ifdef __CUDA_ARCH__
	const float dx = __fsub_rn(xp, xq), dy = __fsub_rn(yp, yq), dz = __fsub_rn(zp, zq);
	return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
#else
	const float dx = xp - xq, dy = yp - yq, dz = zp - zq;
	return fmaf(dz, dz, fmaf(dx, dx, dy * dy));       // host harness: compile with -ffp-contract=off semantics (explicit fmaf)
#endif
}

__host__ __device__ __forceinline__ float gt_sqrt(float v)
{
#ifdef __CUDA_ARCH__
	return __fsqrt_rn(v);
#else
	return sqrtf(v);
#endif
}

// squared-domain bound below which (inclusive) a node may still hold a candidate that beats or ties `best`
template <int MODE> __host__ __device__ __forceinline__ float gt_bound(u64 best)
{
	if (best == KEY_UNMATCHED) return INFINITY;
	unsigned hi32 = (unsigned)(best >> 32);
	float f;
#ifdef __CUDA_ARCH__
	f = __uint_as_float(hi32);
#else
	memcpy(&f, &hi32, 4);
#endif
	if (MODE != ICPB_DIST_SQRT) return f;
	const float t = f * f;
	return t + t * 2e-6f + 1e-37f;                      // every square whose sqrt.rn equals f, and then some
}

// (x, y, z): the source in the grid's frame, i.e. minus the grid origin
__host__ __device__ __forceinline__ float gt_box_lb(float x, float y, float z, float w, int cx, int cy, int cz, float slack)
{
	const float lx = (float)cx * w, ly = (float)cy * w, lz = (float)cz * w;
	float dx = fmaxf(fmaxf(lx - x, x - (lx + w)), 0.0f), dy = fmaxf(fmaxf(ly - y, y - (ly + w)), 0.0f), dz = fmaxf(fmaxf(lz - z, z - (lz + w)), 0.0f);
	dx = fmaxf(dx - slack, 0.0f); dy = fmaxf(dy - slack, 0.0f); dz = fmaxf(dz - slack, 0.0f);
	return (dz * dz + dx * dx + dy * dy) * 0.998f;
}

// Exact nearest neighbour of (x,y,z) among the grid's points. `best`: KEY_UNMATCHED or the key of any real candidate
// (warm start). `seen` (optional) accumulates the number of points examined.
template <int MODE>
__host__ __device__ inline u64 grid_tree_nn(float x, float y, float z, const GridGeom& g, const GridPyramid& py, const int* __restrict__ cell_start,
                                            const float4* __restrict__ sorted4, float thr0, u64 best, unsigned long long* seen,
                                            unsigned long long* nodes = nullptr)
{
	if (!(x == x) || !(y == y) || !(z == z)) return best;           // NaN source: no distance compares below anything
	const float xr = x - g.ox, yr = y - g.oy, zr = z - g.oz;        // the grid's frame (exact, or rounded relative to the distance)
	float bound = gt_bound<MODE>(best);
	unsigned long long pts = 0, visited = 0;
	// Near field (every iteration of a registration but the first few): the warm start already bounds the answer to a
	// ball that touches a handful of cells. Those cells are enumerated directly — whole x-runs are contiguous in the
	// sorted array — instead of descending ~10 levels from the root with 8 box tests per level (ncu, r2: the descent
	// spent 85k warp instructions per warp on 41 nodes + 55 points per source). The cube below contains every cell whose
	// box the descent would NOT have pruned (lb * 0.998 <= bound with `slack` per axis), so the result is the same key.
	if (bound < INFINITY) {
		const float r = gt_sqrt(bound) * 1.0011f + py.slack;                  // 1/sqrt(0.998) = 1.001002
		const float fx0 = floorf((xr - r) * g.inv_h), fx1 = floorf((xr + r) * g.inv_h);
		const float fy0 = floorf((yr - r) * g.inv_h), fy1 = floorf((yr + r) * g.inv_h);
		const float fz0 = floorf((zr - r) * g.inv_h), fz1 = floorf((zr + r) * g.inv_h);
		// compare as floats first: the products can be far outside the int range for sources far from the cloud
		if (fx1 - fx0 < 8.0f && fy1 - fy0 < 8.0f && fz1 - fz0 < 8.0f && (fx1 - fx0 + 1.0f) * (fy1 - fy0 + 1.0f) * (fz1 - fz0 + 1.0f) <= 64.0f) {
			const float nxm = (float)(g.nx - 1), nym = (float)(g.ny - 1), nzm = (float)(g.nz - 1);
			const int x0 = (int)fminf(fmaxf(fx0, 0.0f), nxm), x1 = (int)fminf(fmaxf(fx1, 0.0f), nxm);
			const int y0 = (int)fminf(fmaxf(fy0, 0.0f), nym), y1 = (int)fminf(fmaxf(fy1, 0.0f), nym);
			const int z0 = (int)fminf(fmaxf(fz0, 0.0f), nzm), z1 = (int)fminf(fmaxf(fz1, 0.0f), nzm);
			for (int cz = z0; cz <= z1; cz++)
				for (int cy = y0; cy <= y1; cy++) {
					const int row = g.nx * (cy + g.ny * cz);
					const int k0 = cell_start[row + x0], k1 = cell_start[row + x1 + 1];
					for (int k = k0; k < k1; k++) {
						const float4 q = sorted4[k];
						float d = gt_chain(x, y, z, q.x, q.y, q.z);
						if (d < thr0) {
							if (MODE == ICPB_DIST_SQRT) d = gt_sqrt(d);
							unsigned db, ib;
#ifdef __CUDA_ARCH__
							db = __float_as_uint(d); ib = (unsigned)__float_as_int(q.w);
#else
							memcpy(&db, &d, 4); memcpy(&ib, &q.w, 4);
#endif
							const u64 key = ((u64)db << 32) | (u64)ib;
							if (key < best) best = key;
						}
					}
					pts += (unsigned long long)(k1 - k0);
				}
			if (seen) *seen += pts;
			if (nodes) *nodes += 1;
			return best;
		}
	}
	u64 stack[GP_STACK];
	int sp = 0;
	stack[sp++] = (u64)(py.levels - 1) << 54;                       // root: level | cx << 36 | cy << 18 | cz
	while (sp > 0) {
		const u64 node = stack[--sp];
		visited++;
		const int L = (int)(node >> 54), cx = (int)((node >> 36) & 0x3ffff), cy = (int)((node >> 18) & 0x3ffff), cz = (int)(node & 0x3ffff);
		const float w = g.h * (float)(1 << L);
		const float lb = gt_box_lb(xr, yr, zr, w, cx, cy, cz, py.slack);
		if (lb > bound || lb >= thr0) continue;
		if (L == 0) {
			const int c = cx + g.nx * (cy + g.ny * cz);
			const int k0 = cell_start[c], k1 = cell_start[c + 1];
			for (int k = k0; k < k1; k++) {
				const float4 q = sorted4[k];
				float d = gt_chain(x, y, z, q.x, q.y, q.z);
				if (d < thr0) {
					if (MODE == ICPB_DIST_SQRT) d = gt_sqrt(d);
					unsigned db, ib;
#ifdef __CUDA_ARCH__
					db = __float_as_uint(d); ib = (unsigned)__float_as_int(q.w);
#else
					memcpy(&db, &d, 4); memcpy(&ib, &q.w, 4);
#endif
					const u64 key = ((u64)db << 32) | (u64)ib;
					if (key < best) { best = key; bound = gt_bound<MODE>(best); }
				}
			}
			pts += (unsigned long long)(k1 - k0);
			continue;
		}
		// children at level L-1, the octant nearest to the source pushed last (popped first)
		const int Lc = L - 1;
		const float wc = g.h * (float)(1 << Lc);
		const float mx = (float)(2 * cx + 1) * wc, my = (float)(2 * cy + 1) * wc, mz = (float)(2 * cz + 1) * wc;
		const int nearo = (xr >= mx ? 1 : 0) | (yr >= my ? 2 : 0) | (zr >= mz ? 4 : 0);
		// one byte per internal node: bit k set = child k (x + 2y + 4z) holds points (children outside the grid never do)
		const unsigned mask = py.occ[py.off[L] + (long long)cx + (long long)py.nx[L] * ((long long)cy + (long long)py.ny[L] * cz)];
		for (int i = 7; i >= 0; i--) {
			const int o = i ^ nearo;
			if (!((mask >> o) & 1u)) continue;
			const int ccx = 2 * cx + (o & 1), ccy = 2 * cy + ((o >> 1) & 1), ccz = 2 * cz + ((o >> 2) & 1);
			const float clb = gt_box_lb(xr, yr, zr, wc, ccx, ccy, ccz, py.slack);
			if (clb > bound || clb >= thr0) continue;
			stack[sp++] = ((u64)Lc << 54) | ((u64)ccx << 36) | ((u64)ccy << 18) | (u64)ccz;
		}
	}
	if (seen) *seen += pts;
	if (nodes) *nodes += visited;
	return best;
}

// ------------------------------------------------------------------------------------------------
// k nearest neighbours by the same descent (K5: the neighbour lists PCA normals are estimated from).
// Contract of the reference's knn + minimum (src/ICP_point_to_plane.cu:30-70): the K1 smallest entries under the
// order (distance, index) — distance = sqrt.rn of the chain in SQRT mode, the chain itself in SQ mode — among the
// entries below 10000.0; the query point itself (distance 0) is an ordinary entry. `kd/ki` come back sorted; slots that
// found nothing hold ki = -1. The pruning bound is the K1-th best key so far (its whole sqrt class in SQRT mode);
// equality never prunes, so equally distant targets with lower indices are still found.
// ------------------------------------------------------------------------------------------------
constexpr int GT_KMAX = 8;

template <int MODE>
__host__ __device__ inline void grid_tree_knn(float x, float y, float z, int k1, const GridGeom& g, const GridPyramid& py, const int* __restrict__ cell_start,
                                              const float4* __restrict__ sorted4, float kd[GT_KMAX], int ki[GT_KMAX])
{
	for (int p = 0; p < GT_KMAX; p++) { kd[p] = INFINITY; ki[p] = -1; }
	if (!(x == x) || !(y == y) || !(z == z)) return;
	const float limit = 10000.0f;                                    // `minimum` starts from 10000.0 and uses strict `<`
	const float thr0 = (MODE == ICPB_DIST_SQRT) ? 1.0e8f : limit;    // squared-domain cut: sqrt(d) < 1e4  <=>  d < 1e8 (1e8 is a perfect square of a float)
	const float xr = x - g.ox, yr = y - g.oy, zr = z - g.oz;
	u64 stack[GP_STACK];
	int sp = 0;
	stack[sp++] = (u64)(py.levels - 1) << 54;
	float bound = INFINITY;                                          // squared-domain bound of the K1-th best so far
	while (sp > 0) {
		const u64 node = stack[--sp];
		const int L = (int)(node >> 54), cx = (int)((node >> 36) & 0x3ffff), cy = (int)((node >> 18) & 0x3ffff), cz = (int)(node & 0x3ffff);
		const float w = g.h * (float)(1 << L);
		const float lb = gt_box_lb(xr, yr, zr, w, cx, cy, cz, py.slack);
		if (lb > bound || lb >= thr0) continue;
		if (L == 0) {
			const int c = cx + g.nx * (cy + g.ny * cz);
			const int k0 = cell_start[c], k1e = cell_start[c + 1];
			for (int k = k0; k < k1e; k++) {
				const float4 q = sorted4[k];
				float d = gt_chain(x, y, z, q.x, q.y, q.z);
				if (MODE == ICPB_DIST_SQRT) d = gt_sqrt(d);
				if (!(d < limit)) continue;
				int idx;
#ifdef __CUDA_ARCH__
				idx = __float_as_int(q.w);
#else
				memcpy(&idx, &q.w, 4);
#endif
				// insert (d, idx) if it beats the current K1-th entry
				const int last = k1 - 1;
				if (ki[last] >= 0 && !(d < kd[last] || (d == kd[last] && idx < ki[last]))) continue;
				int pos = last;
				while (pos > 0 && (ki[pos - 1] < 0 || d < kd[pos - 1] || (d == kd[pos - 1] && idx < ki[pos - 1]))) { kd[pos] = kd[pos - 1]; ki[pos] = ki[pos - 1]; pos--; }
				kd[pos] = d; ki[pos] = idx;
				if (ki[last] >= 0) {
					const float f = kd[last];
					if (MODE == ICPB_DIST_SQRT) { const float t = f * f; bound = t + t * 2e-6f + 1e-37f; } else bound = f;
				}
			}
			continue;
		}
		const int Lc = L - 1;
		const float wc = g.h * (float)(1 << Lc);
		const float mx = (float)(2 * cx + 1) * wc, my = (float)(2 * cy + 1) * wc, mz = (float)(2 * cz + 1) * wc;
		const int nearo = (xr >= mx ? 1 : 0) | (yr >= my ? 2 : 0) | (zr >= mz ? 4 : 0);
		// one byte per internal node: bit k set = child k (x + 2y + 4z) holds points (children outside the grid never do)
		const unsigned mask = py.occ[py.off[L] + (long long)cx + (long long)py.nx[L] * ((long long)cy + (long long)py.ny[L] * cz)];
		for (int i = 7; i >= 0; i--) {
			const int o = i ^ nearo;
			if (!((mask >> o) & 1u)) continue;
			const int ccx = 2 * cx + (o & 1), ccy = 2 * cy + ((o >> 1) & 1), ccz = 2 * cz + ((o >> 2) & 1);
			const float clb = gt_box_lb(xr, yr, zr, wc, ccx, ccy, ccz, py.slack);
			if (clb > bound || clb >= thr0) continue;
			stack[sp++] = ((u64)Lc << 54) | ((u64)ccx << 36) | ((u64)ccy << 18) | (u64)ccz;
		}
	}
}

} // namespace icpb
