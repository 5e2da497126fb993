// nn_filter_tc.cu — K1T: the lower-bound filter of K1F evaluated on the 5th-generation tensor cores.
//
// Same contract and same bits as K1 / K1F (nn_bruteforce.cu, nn_filter.cu): idx[i] = argmin_j d_chain(P_i,Q_j) with the
// lowest j on ties; every index and every distance that is returned comes from the reference's exact FP32 chain. What
// changes is WHO evaluates the bracket  e~_j = |q_j - c|^2 - 2 (p - c).(q_j - c)  whose minimum over a 128-target
// sub-tile decides whether the sub-tile can be skipped:
//     K1F: 2-3 FFMA2 per source and target pair on the FP32 pipe + half an FMNMX3 on the ALU pipe   (37 pairs/clk/SM)
//     K1T: one tcgen05.mma kind::tf32 128x256x16 per [128 sources x 256 targets] into TMEM; the FP32 pipe is left with
//          the exact pass only, the ALU pipe with the running minimum of what tcgen05.ld brings back (tools/
//          ubench_tc_filter.cu: 65-75 pairs/clk/SM for tcgen05.ld + FMNMX3; the MMA itself would do 128-256).
// North star: "Tensor cores are used only if ncu shows the |p|^2+|q|^2-2p.q contraction beats the FP32 FMA path at K=3".
//
// TF32 keeps 11 significant bits, so each operand goes in as a hi + lo pair (round-to-nearest splits, residual
// <= 2^-22 |x|) and the product p.q as the three terms hi*hi + hi*lo + lo*hi per coordinate; with w = |q-c|^2 as hi + lo
// against a constant 1 that is 11 of the 16 K slots of two K=8 instructions:
//     A row (source):  ax_hi ax_hi ax_lo | ay_hi ay_hi ay_lo | az_hi az_hi az_lo | 1    1    | 0 ...      a = -2 (p - c)
//     B row (target):  qx_hi qx_lo qx_hi | qy_hi qy_lo qy_hi | qz_hi qz_lo qz_hi | w_hi w_lo | 0 ...      q = q_j - c
// Products of TF32 values are exact in FP32; the accumulation inside the tensor core is not IEEE round-to-nearest, its
// error relative to sum |a_k b_k| is measured by `tools/ubench_tc_filter check` (a few 2^-24). The bound used here is
// K1F's with eps multiplied by TC_EPS_SCALE = 16:  16 u (8 Rq^2 + 10 Rp Rq + 2 Rp^2) covers the split residuals
// (2^-22 = 4u per term: 4u (Rq^2 + 6 Rp Rq)), w's own rounding (3u Rq^2) and an accumulation error of up to 60 u of
// sum |terms| <= Rq^2 + 2 Rp Rq. The price of the larger eps is a few more exact passes (the test fails for sub-tiles
// within sqrt(thr + eps) instead of sqrt(thr)); measured exact-pass rate in bench.py's line.
//
// One MMA column stands for a GROUP of up to 16 targets that are neighbours in space (consecutive in scan order, or along
// the Morton order for large or shuffled clouds): the group's slack rides in the MMA as a twelfth K slot, see "grouped
// form" below. Measured at 1M x 1M points on one B200: K1 direct 185.6 ms per pass, K1F 94.3, K1T one target per column
// 67.4, K1T grouped 6.7 (1.5e14 pairs/s) — identical correspondences throughout.
//
// Kernels. k1_filter_tc_split (the default): one CTA per SM, 18 warps.
//   warps 0-15: epilogue — two groups of 8 (two warps per TMEM lane quarter, each reading one 128-column half of an
//               accumulator): tcgen05.ld 16 columns at a time, partial minima per unit of QC columns, one vote per half on
//               their minimum, K1's exact packed chain over the units that cannot be excluded (originals from the ring)
//   warp 16   : TMA producer — streams target tiles (B operand block in the canonical K-major no-swizzle UMMA layout +
//               the original coordinates of the tile's 256 x TPC target slots, one cp.async.bulk each) through a 2-3 stage ring
//   warp 17   : MMA issuer — for every tile and every one of the 8 source slabs (128 sources each, A operand built in
//               shared memory by the CTA): two tcgen05.mma (K = 2 x 8) into one of two 256-column TMEM accumulators,
//               tcgen05.commit to an mbarrier
// A source's state is one 64-bit word (threshold bits << 32 | unit) updated with atomicMin in shared memory: the tie rule
// is K1's whatever order the warps get there in. Every mbarrier wait is bounded: a protocol error ends the kernel with a
// flag instead of hanging the GPU. k1_filter_tc (ICPB_KT_VAR 0-4) is the first pipeline shape (10 warps, one target per
// column), kept for the comparison in DESIGN.md.
#include "common.cuh"
#include "k1_device.cuh"
#include <cmath>
#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace icpb {

constexpr int TC_TN         = 256;                      // targets per tile = MMA N
constexpr int TC_K          = 16;                       // K slots (two K=8 TF32 instructions)
constexpr int TC_SLABS      = 8;                        // 128-source slabs per CTA pass
constexpr int TC_SB         = 128 * TC_SLABS;           // sources per source block
constexpr int TC_STAGES     = 3;
constexpr int TC_B_BYTES    = TC_TN * TC_K * 4;         // 16384
constexpr int TC_B_FLOATS   = TC_B_BYTES / 4;
constexpr int TC_TILE_BYTES = TC_B_BYTES + 3 * TC_TN * 4;   // + X | Y | Z originals = 19456
constexpr int TC_TILE_FLOATS = TC_TILE_BYTES / 4;
constexpr int TC_A_BYTES    = 128 * TC_K * 4;           // 8192 per slab
constexpr int TC_THREADS    = 320;
constexpr int TC_TRK        = 128;                      // filter / tracking sub-tile: two per tile
constexpr float TC_U        = 5.9604644775390625e-08f;  // 2^-24
constexpr float TC_EPS_SCALE = 16.0f;
constexpr size_t TC_SMEM    = (size_t)TC_SLABS * TC_A_BYTES + (size_t)TC_STAGES * TC_TILE_BYTES + (size_t)6 * TC_SB * 4 + 1024;

struct KTParams {
	const float* px; const float* py; const float* pz;
	const float* tiles;       // [nt][TC_TILE_FLOATS]
	const float4* q4;
	const int*   seed_idx;
	u64*         keys;
	int          n, m, nt;
	long long    total_units;        // source blocks x nt
	int          min_chunk, max_chunk, gss_div;
	unsigned long long* work_counter;
	unsigned long long* exit_counter;   // split form: CTAs that are through; the last one re-arms both counters (no memset launch per pass)
	float        thr0;
	float        cx, cy, cz, rq;
	const int*   done;
	unsigned long long* stats;       // [0] sub-tile tests, [1] exact passes
	int*         fail;               // set on a protocol time-out
	const int*   colstart;           // grouped form: index of the first target of every column
	float        hmax;               // grouped form: the largest group radius h
	const int*   slot_index;         // sorted form: original index of the target in every slot of every tile (-1: unused slot)
	const int*   src_perm;           // sources in arbitrary order: position -> source index, Morton order (nullptr: identity)
};

// element (row r, slot k) of an operand block in the canonical K-major no-swizzle layout: core matrices of 8 rows x 16 B,
// the 4 core matrices of a row group contiguous (LBO = 128 B), row groups 512 B apart (SBO)
__host__ __device__ __forceinline__ int tc_elem(int r, int k) { return (r >> 3) * 128 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ float tf32_rn(float x)
{
	unsigned u = __float_as_uint(x);
	u += 0xfffu + ((u >> 13) & 1u);
	return __uint_as_float(u & 0xffffe000u);
}
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) { hi = tf32_rn(x); lo = tf32_rn(__fsub_rn(x, hi)); }

__global__ void tc_pack_kernel(const float4* __restrict__ q4, int m, int m_pad, float cx, float cy, float cz, float* __restrict__ tiles)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m_pad) return;
	float* tile = tiles + (size_t)(j / TC_TN) * TC_TILE_FLOATS;
	const int r = j % TC_TN;
	float v[TC_K];
#pragma unroll
	for (int k = 0; k < TC_K; k++) v[k] = 0.0f;
	float x = __int_as_float(0x7f800000), y = x, z = x;
	if (j < m) {
		const float4 q = q4[j];
		x = q.x; y = q.y; z = q.z;
		const float xc = __fsub_rn(x, cx), yc = __fsub_rn(y, cy), zc = __fsub_rn(z, cz);
		const float w = __fmaf_rn(zc, zc, __fmaf_rn(xc, xc, __fmul_rn(yc, yc)));
		float h, l;
		tf32_split(xc, h, l); v[0] = h; v[1] = l; v[2] = h;
		tf32_split(yc, h, l); v[3] = h; v[4] = l; v[5] = h;
		tf32_split(zc, h, l); v[6] = h; v[7] = l; v[8] = h;
		tf32_split(w, h, l);  v[9] = h; v[10] = l;
	} else {
		v[9] = 3.0e38f;                 // padding: e~ = 3e38 for every source, never below any tau; originals +inf
	}
#pragma unroll
	for (int k = 0; k < TC_K; k++) tile[tc_elem(r, k)] = v[k];
	tile[TC_B_FLOATS + r] = x; tile[TC_B_FLOATS + TC_TN + r] = y; tile[TC_B_FLOATS + 2 * TC_TN + r] = z;
}

// ---- grouped form: one MMA column per GROUP of consecutive targets --------------------------------------------------
// K1T is bound by what the epilogue can read out of TMEM and min-reduce, and by the MMA hand-shake per accumulator —
// both per COLUMN. A column can stand for several targets: for a point g (the float centroid of the group) and
// h >= max_k |q_k - g|, the triangle inequality gives |p - g| <= |p - q_k| + h. A member can only matter while its
// distance is within the source's current threshold, |p - q_k|^2 <= tau <= s0^2 (s0 = the square root of the threshold the
// source starts the sweep with), so for such a member
//     |p - g|^2 <= tau + 2 sqrt(tau) h + h^2 <= tau + 2 s0 h + h^2,
// i.e. the whole group is excluded when
//     |p - g|^2 - 2 s0 h - h^2 > tau.
// The left side is still ONE inner product per (source, column): the bracket of g with |g - c|^2 - h^2 in the constant
// slots and one more K slot holding s0 (source row, rounded UP to TF32) against -2 h (column, h rounded UP to TF32; the
// product of two TF32 values is exact). No per-sub-tile slack is left for the epilogue to add, and the slack that is paid
// is of the order (s0 + h)^2 - s0^2: the exact pass is asked for only where a group really reaches into the ball of
// radius s0 around the source. (Round 2's first grouped form bounded 2 |p - g| h by its largest value over a sub-tile,
// which grows with the distance to the sub-tile's far end: 6 % of the quarters went to the exact pass on the 1M raster
// at 4 targets per column and one group straddling a row end spoilt its whole sub-tile; now 8-16 targets per column work.)
//   A row (source):  ax_hi ax_hi ax_lo | ay.. | az.. | 1    1    | s0   | 0 ...
//   B row (column):  gx_hi gx_lo gx_hi | gy.. | gz.. | w_hi w_lo | -2 h | 0 ...          w = |g - c|^2 - h^2, rounded down
// Groups never span a JUMP of the scan: consecutive targets are neighbours in space for scan-ordered clouds (the
// reference's raster saddle, LiDAR sweeps) except where a row or a sweep ends; a step longer than TCG_JUMP times the mean
// step closes the column early (its unused slots hold +inf originals and are never attained), so columns cover a variable
// number (1 .. TPC) of targets and `colstart` maps a column back to the index of its first target. For clouds in arbitrary
// order every h is of the order of the cloud, every quarter goes to the exact pass, and the host's exact-pass-rate policy
// (kf_policy_update) halves TPC down to one target per column.
// Tile = 256 columns: B block | X Y Z originals in slot order (slot = column * TPC + member).
__host__ __device__ constexpr int tcg_tile_targets(int tpc) { return TC_TN * tpc; }
__host__ __device__ constexpr int tcg_tile_floats(int tpc) { return tpc == 1 ? TC_TILE_FLOATS : TC_B_FLOATS + 3 * TC_TN * tpc; }
constexpr float TCG_JUMP = 8.0f;
constexpr int   TCG_SORT_MIN = 200000;     // targets from which the Morton form is used whatever the scan order

__device__ __forceinline__ float tf32_ru_pos(float x)      // smallest TF32 value >= x, x >= 0 and finite
{
	return __uint_as_float((__float_as_uint(x) + 0x1fffu) & 0xffffe000u);
}

// sum of the steps |q_i - q_(i-1)| of the scan (double; the order of the atomic adds only moves the jump threshold by rounding)
__global__ void tcg_step_sum_kernel(const float4* __restrict__ q4, int m, double* sum)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float d = 0.0f;
	if (i > 0 && i < m) {
		const float4 a = q4[i - 1], b = q4[i];
		const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
		d = sqrtf(ex * ex + ey * ey + ez * ez);
		if (!(d == d) || d == __int_as_float(0x7f800000)) d = 0.0f;
	}
	double v = d;
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	__shared__ double s[8];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s[w]; atomicAdd(sum, t); }
}
// mark[i] = i where a column must start because the scan jumps (or i == 0), else 0: an inclusive max-scan gives every
// target the start of its run
__global__ void tcg_break_kernel(const float4* __restrict__ q4, int m, const double* sum, int* mark)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	int v = 0;
	if (i > 0) {
		const float jump = TCG_JUMP * (float)(*sum / (double)(m > 1 ? m - 1 : 1));
		const float4 a = q4[i - 1], b = q4[i];
		const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
		const float d = sqrtf(ex * ex + ey * ey + ez * ez);
		if (!(d <= jump)) v = i;
	}
	mark[i] = v;
}
// flag[i] = 1 where a column starts: every TPC-th target of a run
__global__ void tcg_colflag_kernel(const int* __restrict__ runstart, int m, int tpc, int* flag)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m) flag[i] = ((i - runstart[i]) % tpc == 0) ? 1 : 0;
}
__global__ void tcg_colstart_kernel(const int* __restrict__ runstart, const int* __restrict__ colid, int m, int tpc, int* colstart)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m && (i - runstart[i]) % tpc == 0) colstart[colid[i] - 1] = i;
}
struct TcgMax { __host__ __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; } };

// ---- sorted form: clouds in arbitrary order ----------------------------------------------------------------------------
// Grouping needs consecutive targets to be neighbours in space. Where the scan order does not give that (mean step far
// above the point spacing: a shuffled cloud, a mesh's vertex list) the tiles are built over the targets in MORTON order
// instead (21 bits per axis over the cube around the cloud, cub radix sort): consecutive positions are close again, the
// jumps of the Z curve close the columns like row ends do. Slots are then no longer in index order, so the lowest-index
// tie rule cannot come from the slot order: the exact pass, whenever it offers a unit, looks up the ORIGINAL index of
// every slot that attains the unit's minimum (slot_index[], read from L2 for those slots only) and the per-source key
// carries (threshold, smallest original index) — atomicMin keeps the reference's answer whatever the visiting order.
__device__ __forceinline__ unsigned long long morton_spread21(unsigned v)
{
	unsigned long long x = v & 0x1fffffu;
	x = (x | (x << 32)) & 0x1f00000000ffffull;
	x = (x | (x << 16)) & 0x1f0000ff0000ffull;
	x = (x | (x << 8))  & 0x100f00f00f00f00full;
	x = (x | (x << 4))  & 0x10c30c30c30c30c3ull;
	x = (x | (x << 2))  & 0x1249249249249249ull;
	return x;
}
__global__ void tcg_morton_kernel(const float4* __restrict__ q4, int m, float x0, float y0, float z0, float inv, unsigned long long* keys, int* vals)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const float4 q = q4[i];
	const float fx = fminf(fmaxf((q.x - x0) * inv, 0.0f), 1.0f), fy = fminf(fmaxf((q.y - y0) * inv, 0.0f), 1.0f), fz = fminf(fmaxf((q.z - z0) * inv, 0.0f), 1.0f);
	const unsigned ux = min(2097151u, (unsigned)(fx * 2097152.0f)), uy = min(2097151u, (unsigned)(fy * 2097152.0f)), uz = min(2097151u, (unsigned)(fz * 2097152.0f));
	keys[i] = morton_spread21(ux) | (morton_spread21(uy) << 1) | (morton_spread21(uz) << 2);
	vals[i] = i;
}
__global__ void tcg_step_sum_soa_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz, int n, double* sum)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float d = 0.0f;
	if (i > 0 && i < n) {
		const float ex = px[i] - px[i - 1], ey = py[i] - py[i - 1], ez = pz[i] - pz[i - 1];
		d = sqrtf(ex * ex + ey * ey + ez * ez);
		if (!(d == d) || d == __int_as_float(0x7f800000)) d = 0.0f;
	}
	double v = d;
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	__shared__ double s[8];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s[w]; atomicAdd(sum, t); }
}
__global__ void tcg_morton_soa_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz, int n,
                                      float x0, float y0, float z0, float inv, unsigned long long* keys, int* vals)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float fx = (px[i] - x0) * inv, fy = (py[i] - y0) * inv, fz = (pz[i] - z0) * inv;
	fx = (fx == fx) ? fminf(fmaxf(fx, 0.0f), 1.0f) : 0.0f; fy = (fy == fy) ? fminf(fmaxf(fy, 0.0f), 1.0f) : 0.0f; fz = (fz == fz) ? fminf(fmaxf(fz, 0.0f), 1.0f) : 0.0f;
	const unsigned ux = min(2097151u, (unsigned)(fx * 2097152.0f)), uy = min(2097151u, (unsigned)(fy * 2097152.0f)), uz = min(2097151u, (unsigned)(fz * 2097152.0f));
	keys[i] = morton_spread21(ux) | (morton_spread21(uy) << 1) | (morton_spread21(uz) << 2);
	vals[i] = i;
}
__global__ void tcg_gather_kernel(const float4* __restrict__ q4, const int* __restrict__ perm, int m, float4* out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m) out[i] = q4[perm[i]];
}

// one block of 128 threads per sub-tile (128 columns)
template <int TPC>
__global__ void __launch_bounds__(128) tc_pack_group_kernel(const float4* __restrict__ q4, int m, const int* __restrict__ colstart, int ncols,
                                                            float cx, float cy, float cz, float* __restrict__ tiles, float* hmax,
                                                            const int* __restrict__ perm, int* __restrict__ slot_index)
{
	constexpr int TT = TC_TN * TPC;
	const int sub = blockIdx.x;                      // global sub-tile
	const int col = threadIdx.x;                     // column inside the sub-tile
	float* tile = tiles + (size_t)(sub >> 1) * tcg_tile_floats(TPC);
	const int r = (sub & 1) * 128 + col;             // B row inside the tile
	const int gc = sub * 128 + col;                  // global column
	const float inf = __int_as_float(0x7f800000);
	int j0 = m, len = 0;
	if (gc < ncols) { j0 = colstart[gc]; len = ((gc + 1 < ncols) ? colstart[gc + 1] : m) - j0; }      // 1 .. TPC members
	float v[TC_K];
#pragma unroll
	for (int k = 0; k < TC_K; k++) v[k] = 0.0f;
	float4 q[TPC];
#pragma unroll
	for (int k = 0; k < TPC; k++) q[k] = (k < len) ? q4[j0 + k] : make_float4(inf, inf, inf, 0.f);
	if (len > 0) {
		// g: any float point works; the centroid of the members
		float gx = 0.f, gy = 0.f, gz = 0.f;
		const float inv = 1.0f / (float)len;
#pragma unroll
		for (int k = 0; k < TPC; k++) if (k < len) { gx += q[k].x * inv; gy += q[k].y * inv; gz += q[k].z * inv; }
		// h >= max |q_k - g|: differences rounded (relative u), squares and sums rounded up, a safety factor, then up to TF32
		float d = 0.f;
#pragma unroll
		for (int k = 0; k < TPC; k++) {
			if (k < len) {
				const float ex = q[k].x - gx, ey = q[k].y - gy, ez = q[k].z - gz;
				d = fmaxf(d, __fmaf_ru(ez, ez, __fmaf_ru(ex, ex, __fmul_ru(ey, ey))));
			}
		}
		const float h = tf32_ru_pos(__fmul_ru(__fsqrt_ru(d), 1.0f + 16.0f * TC_U));
		const float xc = __fsub_rn(gx, cx), yc = __fsub_rn(gy, cy), zc = __fsub_rn(gz, cz);
		const float w = __fsub_rd(__fmaf_rn(zc, zc, __fmaf_rn(xc, xc, __fmul_rn(yc, yc))), __fmul_ru(h, h));
		float hi, lo;
		tf32_split(xc, hi, lo); v[0] = hi; v[1] = lo; v[2] = hi;
		tf32_split(yc, hi, lo); v[3] = hi; v[4] = lo; v[5] = hi;
		tf32_split(zc, hi, lo); v[6] = hi; v[7] = lo; v[8] = hi;
		tf32_split(w, hi, lo);  v[9] = hi; v[10] = lo;
		v[11] = -2.0f * h;
		atomicMax(reinterpret_cast<int*>(hmax), __float_as_int(h));          // h >= 0: the bit patterns order like the values
	} else {
		v[9] = 3.0e38f;                              // padding column: never below any tau
	}
#pragma unroll
	for (int k = 0; k < TC_K; k++) tile[tc_elem(r, k)] = v[k];
	float* X = tile + TC_B_FLOATS + (sub & 1) * 128 * TPC + col * TPC;
#pragma unroll
	for (int k = 0; k < TPC; k++) { X[k] = q[k].x; X[TT + k] = q[k].y; X[2 * TT + k] = q[k].z; }
	if (slot_index != nullptr) {                     // sorted form: q4 is the cloud in Morton order, perm its original indices
		int* S = slot_index + (size_t)(sub >> 1) * TT + (sub & 1) * 128 * TPC + col * TPC;
#pragma unroll
		for (int k = 0; k < TPC; k++) S[k] = (k < len) ? perm[j0 + k] : -1;
	}
}

// ---- small PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity)
{
	for (int spin = 0; spin < (1 << 26); spin++) {
		uint32_t ok;
		asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (ok) return true;
	}
	return false;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
	// start >> 4 [0,14), LBO = 128 B >> 4 at [16,30), SBO = 512 B >> 4 at [32,46), descriptor version 1 at [46,48), no swizzle
	return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
	uint32_t* u = reinterpret_cast<uint32_t*>(v);
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
	             : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]),
	               "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]),
	               "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
	             : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void min32(const float (&v)[32], float& m0, float& m1)
{
#pragma unroll
	for (int k = 0; k < 32; k += 4) { m0 = min3(m0, v[k], v[k + 1]); m1 = min3(m1, v[k + 2], v[k + 3]); }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
	uint32_t* u = reinterpret_cast<uint32_t*>(v);
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	             : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]),
	               "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
	             : "r"(taddr));
}
__device__ __forceinline__ void min16(const float (&v)[16], float& m0, float& m1)
{
#pragma unroll
	for (int k = 0; k < 16; k += 4) { m0 = min3(m0, v[k], v[k + 1]); m1 = min3(m1, v[k + 2], v[k + 3]); }
}
// minimum over the 128 columns of this thread's TMEM lane starting at taddr (double-buffered loads of LDW columns)
template <int LDW> __device__ __forceinline__ float subtile_min(uint32_t taddr)
{
	const float inf = __int_as_float(0x7f800000);
	float m0 = inf, m1 = inf;
	if constexpr (LDW == 32) {
		float va[32], vb[32];
		tmem_ld32(taddr, va);
		tmem_wait_ld(); tmem_ld32(taddr + 32, vb); min32(va, m0, m1);
		tmem_wait_ld(); tmem_ld32(taddr + 64, va); min32(vb, m0, m1);
		tmem_wait_ld(); tmem_ld32(taddr + 96, vb); min32(va, m0, m1);
		tmem_wait_ld(); min32(vb, m0, m1);
	} else {
		float va[16], vb[16];
		tmem_ld16(taddr, va);
#pragma unroll
		for (int c = 0; c < 128; c += 32) {
			tmem_wait_ld(); tmem_ld16(taddr + c + 16, vb); min16(va, m0, m1);
			tmem_wait_ld(); if (c + 32 < 128) tmem_ld16(taddr + c + 32, va);
			min16(vb, m0, m1);
		}
	}
	return fminf(m0, m1);
}

// 128 / QC partial minima of the 128 columns starting at taddr: columns [QC q, QC q + QC), 16-column loads double-buffered
template <int QC, int NCOL> __device__ __forceinline__ void subtile_mins(uint32_t taddr, float (&mq)[NCOL / QC])
{
	static_assert(QC == 32 || QC == 16 || QC == 8, "columns per exact-pass unit");
	static_assert(NCOL == 128 || NCOL == 64, "columns a warp reads per unit");
	constexpr int NQ = NCOL / 32;
	const float inf = __int_as_float(0x7f800000);
	float va[16], vb[16];
	tmem_ld16(taddr, va);
#pragma unroll
	for (int q = 0; q < NQ; q++) {
		tmem_wait_ld(); tmem_ld16(taddr + 32 * q + 16, vb);
		if constexpr (QC == 32) {
			float m0 = inf, m1 = inf;
			min16(va, m0, m1);
			tmem_wait_ld(); if (q < NQ - 1) tmem_ld16(taddr + 32 * q + 32, va);
			min16(vb, m0, m1);
			mq[q] = fminf(m0, m1);
		} else if constexpr (QC == 16) {
			float m0 = inf, m1 = inf;
			min16(va, m0, m1);
			mq[2 * q] = fminf(m0, m1);
			tmem_wait_ld(); if (q < NQ - 1) tmem_ld16(taddr + 32 * q + 32, va);
			m0 = inf; m1 = inf;
			min16(vb, m0, m1);
			mq[2 * q + 1] = fminf(m0, m1);
		} else {
			mq[4 * q]     = fminf(min3(va[0], va[1], va[2]), min3(va[3], va[4], va[5])); mq[4 * q] = min3(mq[4 * q], va[6], va[7]);
			mq[4 * q + 1] = fminf(min3(va[8], va[9], va[10]), min3(va[11], va[12], va[13])); mq[4 * q + 1] = min3(mq[4 * q + 1], va[14], va[15]);
			tmem_wait_ld(); if (q < NQ - 1) tmem_ld16(taddr + 32 * q + 32, va);
			mq[4 * q + 2] = fminf(min3(vb[0], vb[1], vb[2]), min3(vb[3], vb[4], vb[5])); mq[4 * q + 2] = min3(mq[4 * q + 2], vb[6], vb[7]);
			mq[4 * q + 3] = fminf(min3(vb[8], vb[9], vb[10]), min3(vb[11], vb[12], vb[13])); mq[4 * q + 3] = min3(mq[4 * q + 3], vb[14], vb[15]);
		}
	}
}

// NS       sub-tiles (128 targets) per MMA: 2 = one 256-column instruction per (slab, tile), 1 = two 128-column ones
// NACC_G   TMEM accumulators per epilogue group (1: the group waits out every MMA; 2: the next unit's MMA runs meanwhile)
// SLABS    128-source slabs per CTA pass, STAGES ring depth, MINB CTAs per SM (2: each CTA takes half the SM's TMEM)
template <int MODE, int NS, int NACC_G, int SLABS, int STAGES, int MINB, int LDW>
__global__ void __launch_bounds__(TC_THREADS, MINB) k1_filter_tc(const KTParams p)
{
	constexpr int SUBS = TC_TN / TC_TRK;      // 2 sub-tiles per tile
	constexpr int UPT = SUBS / NS;            // MMA units per (slab, tile)
	constexpr int UNIT_COLS = NS * TC_TRK;
	constexpr int TMEM_COLS = 2 * NACC_G * UNIT_COLS;
	constexpr int SBN = 128 * SLABS;          // sources per source block
	static_assert(TMEM_COLS * MINB <= 512 && (TMEM_COLS == 256 || TMEM_COLS == 512), "TMEM budget");
	static_assert(SLABS % 2 == 0 && SLABS <= TC_SLABS, "slabs are dealt to two epilogue groups");
	if (p.done != nullptr && *p.done) return;
	extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
	// 1 KB alignment by hand (the dynamic segment is only guaranteed 16 B)
	unsigned char* tc_smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
	__shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[4], tempty_bar[4];
	__shared__ uint32_t tmem_base_s;
	__shared__ unsigned long long s_u0;
	__shared__ int s_len, s_fail;

	float* a_slabs = reinterpret_cast<float*>(tc_smem);
	float* ring    = reinterpret_cast<float*>(tc_smem + (size_t)SLABS * TC_A_BYTES);
	float* thr_s   = reinterpret_cast<float*>(tc_smem + (size_t)SLABS * TC_A_BYTES + (size_t)STAGES * TC_TILE_BYTES);
	int*   best_s  = reinterpret_cast<int*>(thr_s + SBN);
	float* ox_s    = thr_s + 2 * SBN;
	float* oy_s    = thr_s + 3 * SBN;
	float* oz_s    = thr_s + 4 * SBN;
	float* kk_s    = thr_s + 5 * SBN;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const bool is_epi = warp >= 2;
	const int ew = warp - 2;                  // epilogue warp 0..7
	const int grp = ew >> 2;                  // epilogue group: slab parity and accumulator set this warp serves
	const int quarter = warp & 3;             // TMEM lane quarter a warp may read: warp index mod 4
	const int row = quarter * 32 + lane;      // source row inside a slab

	if (tid == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
		for (int a = 0; a < 4; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
		s_fail = 0;
		fence_mbar_init();
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = tmem_base_s;
	// kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UNIT_COLS >> 3) << 17) | ((128u >> 4) << 24);

	int it = 0;                   // tiles this CTA has streamed so far: ring stage and parities follow it across segments
	unsigned long long n_tests = 0, n_exact = 0;
	const float inf = __int_as_float(0x7f800000);
	const float one8u = 1.0f + 8.0f * TC_U;
	float tau[SLABS / 2];                       // per owned slab (epilogue threads)
#pragma unroll
	for (int q = 0; q < SLABS / 2; q++) tau[q] = -inf;
	bool failed = false;

	while (true) {
		__syncthreads();
		if (tid == 0) {
			const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(p.work_counter);
			const long long left = p.total_units - (long long)seen;
			long long len = left / ((long long)p.gss_div * (long long)gridDim.x);
			if (len < p.min_chunk) len = p.min_chunk;
			if (len > p.max_chunk) len = p.max_chunk;
			s_len = (int)len;
			s_u0 = atomicAdd(p.work_counter, (unsigned long long)len);
		}
		__syncthreads();
		if ((long long)s_u0 >= p.total_units || s_fail) break;
		long long u = (long long)s_u0;
		const long long u_end = min(p.total_units, u + (long long)s_len);
		while (u < u_end) {
			const int sb = (int)(u / p.nt);
			const int t0 = (int)(u - (long long)sb * p.nt);
			const int t1 = (int)min((long long)p.nt, (long long)t0 + (u_end - u));
			u += t1 - t0;
			__syncthreads();                       // previous segment fully consumed (A slabs, per-source state, ring)
			// ---- per-source state and the A operand of the segment's sources (epilogue threads: SLABS / 2 sources each) ----
			if (is_epi) {
#pragma unroll
				for (int q = 0; q < SLABS / 2; q++) {
					const int a = grp + 2 * q;
					const int sidx = a * 128 + row;
					const int i = sb * SBN + sidx;
					const float x = p.px[i], y = p.py[i], z = p.pz[i];       // arrays are padded beyond n
					ox_s[sidx] = x; oy_s[sidx] = y; oz_s[sidx] = z;
					const float pcx = __fsub_rn(x, p.cx), pcy = __fsub_rn(y, p.cy), pcz = __fsub_rn(z, p.cz);
					const float p2 = __fmaf_rn(pcz, pcz, __fmaf_rn(pcx, pcx, __fmul_rn(pcy, pcy)));
					const float p2lo = __fmul_rd(p2, 1.0f - 8.0f * TC_U);
					const float rp = __fmul_ru(__fsqrt_ru(p2), 1.0f + 8.0f * TC_U);
					float e = __fmul_ru(8.0f * p.rq, p.rq);
					e = __fmaf_ru(10.0f * rp, p.rq, e);
					e = __fmaf_ru(2.0f * rp, rp, e);
					e = __fmul_ru(e, 1.05f * TC_EPS_SCALE * TC_U);
					const float kk = __fsub_ru(e, p2lo);
					kk_s[sidx] = kk;
					float th = p.thr0;
					if (p.seed_idx != nullptr && i < p.n) {
						const int j0 = p.seed_idx[i];
						if (j0 >= 0 && j0 < p.m) {
							const float4 qq = __ldg(p.q4 + j0);
							const float us = dist_chain(x, y, z, qq.x, qq.y, qq.z);
							const float up = (MODE == ICPB_DIST_SQRT) ? __fmul_ru(us, one8u) : us;
							const float nx = __uint_as_float(__float_as_uint(up) + 1u);
							if (us == us && nx < th) th = nx;
						}
					}
					thr_s[sidx] = th; best_s[sidx] = -1;
					// sources past the end of the cloud, and non-finite ones, never ask for an exact pass
					const bool live = (i < p.n) && (p2 == p2) && (p2 < inf);
					tau[q] = live ? __fadd_ru(__fmul_ru(th, one8u), kk) : -inf;
					// A row: a = -2 (p - c) as hi + lo; slots ax_hi ax_hi ax_lo | ay.. | az.. | 1 1 | 0...
					float* A = a_slabs + (size_t)a * (TC_A_BYTES / 4);
					float h, l;
					const float ax = live ? -2.0f * pcx : 0.0f, ay = live ? -2.0f * pcy : 0.0f, az = live ? -2.0f * pcz : 0.0f;
					tf32_split(ax, h, l); A[tc_elem(row, 0)] = h; A[tc_elem(row, 1)] = h; A[tc_elem(row, 2)] = l;
					tf32_split(ay, h, l); A[tc_elem(row, 3)] = h; A[tc_elem(row, 4)] = h; A[tc_elem(row, 5)] = l;
					tf32_split(az, h, l); A[tc_elem(row, 6)] = h; A[tc_elem(row, 7)] = h; A[tc_elem(row, 8)] = l;
					A[tc_elem(row, 9)] = 1.0f; A[tc_elem(row, 10)] = 1.0f;
#pragma unroll
					for (int k = 11; k < TC_K; k++) A[tc_elem(row, k)] = 0.0f;
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes of A -> visible to the tensor core
			}
			__syncthreads();

			if (warp == 0) {
				// ---- TMA producer ----
				if (lane == 0) {
					for (int t = t0; t < t1 && !failed; t++) {
						const int k = it + (t - t0), st = k % STAGES, use = k / STAGES;
						if (use >= 1 && !mbar_wait_bounded(&empty_bar[st], (uint32_t)((use - 1) & 1))) { failed = true; break; }
						mbar_expect_tx(&full_bar[st], TC_TILE_BYTES);
						tma_load_1d(ring + (size_t)st * TC_TILE_FLOATS, p.tiles + (size_t)t * TC_TILE_FLOATS, TC_TILE_BYTES, &full_bar[st]);
					}
				}
			} else if (warp == 1) {
				// ---- MMA issuer ----
				if (lane == 0) {
					for (int t = t0; t < t1 && !failed; t++) {
						const int k = it + (t - t0), st = k % STAGES, use = k / STAGES;
						if (!mbar_wait_bounded(&full_bar[st], (uint32_t)(use & 1))) { failed = true; break; }
						const uint32_t sB = smem_u32(ring + (size_t)st * TC_TILE_FLOATS);
#pragma unroll 1
						for (int a = 0; a < SLABS && !failed; a++) {
							const uint32_t sA = smem_u32(a_slabs + (size_t)a * (TC_A_BYTES / 4));
#pragma unroll 1
							for (int un = 0; un < UPT; un++) {
								// unit (slab a, MMA un) of group g = a & 1; the group's units rotate through its NACC_G accumulators
								const int g = a & 1;
								const int v = (k * (SLABS / 2) + (a >> 1)) * UPT + un;            // units of group g so far
								const int acc = g * NACC_G + (v % NACC_G);
								const int j = v / NACC_G;                                        // use counter of accumulator `acc`
								if (j >= 1 && !mbar_wait_bounded(&tempty_bar[acc], (uint32_t)((j - 1) & 1))) { failed = true; break; }
								asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
								for (int kk2 = 0; kk2 < 2; kk2++) {
									// K advances by two 128 B core matrices; unit `un` starts at row un * UNIT_COLS of the B block (row groups of 512 B)
									const uint64_t dA = umma_desc(sA + kk2 * 256), dB = umma_desc(sB + un * (UNIT_COLS / 8) * 512 + kk2 * 256);
									asm volatile("{\n.reg .pred pacc;\nsetp.ne.b32 pacc, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pacc;\n}\n"
									             :: "r"(tmem_base + (uint32_t)(acc * UNIT_COLS)), "l"(dA), "l"(dB), "r"(idesc), "r"((uint32_t)kk2));
								}
								asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&tfull_bar[acc])) : "memory");
							}
						}
					}
				}
			} else {
				// ---- epilogue: running minimum per sub-tile, test against tau, exact pass where it cannot be excluded ----
				for (int t = t0; t < t1 && !failed; t++) {
					const int k = it + (t - t0), st = k % STAGES, use = k / STAGES;
					if (!mbar_wait_bounded(&full_bar[st], (uint32_t)(use & 1))) { failed = true; break; }
					const float* tile = ring + (size_t)st * TC_TILE_FLOATS;
					const float4* X4 = reinterpret_cast<const float4*>(tile + TC_B_FLOATS);
					const float4* Y4 = X4 + TC_TN / 4;
					const float4* Z4 = Y4 + TC_TN / 4;
#pragma unroll
					for (int q = 0; q < SLABS / 2; q++) {
						if (failed) break;
						const int a = grp + 2 * q;
						const int sidx = a * 128 + row;
#pragma unroll 1
						for (int un = 0; un < UPT; un++) {
							const int v = (k * (SLABS / 2) + q) * UPT + un;
							const int acc = grp * NACC_G + (v % NACC_G);
							const int j = v / NACC_G;
							if (!mbar_wait_bounded(&tfull_bar[acc], (uint32_t)(j & 1))) { failed = true; break; }
							asm volatile("tcgen05.fence::after_thread_sync;");
							const uint32_t taddr = tmem_base + (uint32_t)(acc * UNIT_COLS) + ((uint32_t)(quarter * 32) << 16);
							float em[NS];
#pragma unroll
							for (int hs = 0; hs < NS; hs++) em[hs] = subtile_min<LDW>(taddr + hs * TC_TRK);
							// the accumulator is in registers: hand it back to the MMA warp before the (rare) exact pass
							asm volatile("tcgen05.fence::before_thread_sync;");
							__syncwarp();
							if (lane == 0) mbar_arrive(&tempty_bar[acc]);
#pragma unroll
							for (int hs = 0; hs < NS; hs++) {
								const int h = un * NS + hs;                   // sub-tile of the tile
								const unsigned need = __ballot_sync(0xffffffffu, em[hs] <= tau[q]);
								n_tests += 1;
								if (need) {                                   // warp-uniform
									n_exact += 1;
									const int j0 = h * (TC_TRK / 4), j1 = j0 + TC_TRK / 4;
									const float sx = ox_s[sidx], sy = oy_s[sidx], sz = oz_s[sidx];
									const float th = thr_s[sidx];
									const u64 PX = pack2(sx, sx), PY = pack2(sy, sy), PZ = pack2(sz, sz);
									float mm = th;
#pragma unroll 4
									for (int jq = j0; jq < j1; jq++) {
										const float4 Xo = X4[jq], Yo = Y4[jq], Zo = Z4[jq];
										u64 dx = sub2(PX, pack2(Xo.x, Xo.y)), dy = sub2(PY, pack2(Yo.x, Yo.y)), dz = sub2(PZ, pack2(Zo.x, Zo.y));
										u64 d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
										float aa, bb;
										unpack2(d, aa, bb);
										mm = min3(mm, aa, bb);
										dx = sub2(PX, pack2(Xo.z, Xo.w)); dy = sub2(PY, pack2(Yo.z, Yo.w)); dz = sub2(PZ, pack2(Zo.z, Zo.w));
										d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
										unpack2(d, aa, bb);
										mm = min3(mm, aa, bb);
									}
									if (mm < th) {
										const float nt_ = lower_threshold<MODE>(mm);
										thr_s[sidx] = nt_; best_s[sidx] = t * SUBS + h;
										if (tau[q] > -inf) tau[q] = __fadd_ru(__fmul_ru(nt_, one8u), kk_s[sidx]);
									}
								}
							}
						}
					}
					// this warp is done with the tile's originals: release the ring stage (8 epilogue warps)
					__syncwarp();
					if (lane == 0) mbar_arrive(&empty_bar[st]);
				}
			}
			if (failed) s_fail = 1;
			it += t1 - t0;
			__syncthreads();
			if (s_fail) break;
			// ---- index recovery: the first target of the remembered sub-tile that attains the exact minimum ----
			if (is_epi) {
#pragma unroll 1
				for (int q = 0; q < SLABS / 2; q++) {
					const int a = grp + 2 * q;
					const int sidx = a * 128 + row;
					const int i = sb * SBN + sidx;
					const int bs = best_s[sidx];
					if (i < p.n && bs >= 0) {
						const float th = thr_s[sidx];
						const float sx = ox_s[sidx], sy = oy_s[sidx], sz = oz_s[sidx];
						const float* gx = p.tiles + (size_t)(bs / SUBS) * TC_TILE_FLOATS + TC_B_FLOATS + (size_t)(bs % SUBS) * TC_TRK;
						const float4* GX = reinterpret_cast<const float4*>(gx);
						const float4* GY = reinterpret_cast<const float4*>(gx + TC_TN);
						const float4* GZ = reinterpret_cast<const float4*>(gx + 2 * TC_TN);
						const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(th) : th;
						int found = -1;
						for (int jq = 0; jq < TC_TRK / 4 && found < 0; jq++) {
							const float4 X = __ldg(GX + jq), Y = __ldg(GY + jq), Z = __ldg(GZ + jq);
							float d0 = dist_chain(sx, sy, sz, X.x, Y.x, Z.x);
							float d1 = dist_chain(sx, sy, sz, X.y, Y.y, Z.y);
							float d2 = dist_chain(sx, sy, sz, X.z, Y.z, Z.z);
							float d3 = dist_chain(sx, sy, sz, X.w, Y.w, Z.w);
							if (MODE == ICPB_DIST_SQRT) { d0 = __fsqrt_rn(d0); d1 = __fsqrt_rn(d1); d2 = __fsqrt_rn(d2); d3 = __fsqrt_rn(d3); }
							if (d0 <= target) found = 4 * jq;
							else if (d1 <= target) found = 4 * jq + 1;
							else if (d2 <= target) found = 4 * jq + 2;
							else if (d3 <= target) found = 4 * jq + 3;
						}
						if (found >= 0) {
							const u64 key = ((u64)__float_as_uint(target) << 32) | (u64)(uint32_t)(bs * TC_TRK + found);
							atomicMin(p.keys + i, key);
						}
					}
				}
			}
		}
		if (s_fail) break;
	}
	__syncthreads();
	if (s_fail && tid == 0) *p.fail = 1;
	if (p.stats != nullptr && is_epi) {
		for (int o = 16; o > 0; o >>= 1) { n_tests += __shfl_xor_sync(0xffffffffu, n_tests, o); n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o); }
		if (lane == 0) { atomicAdd(p.stats, n_tests / 32); atomicAdd(p.stats + 1, n_exact / 32); }
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TMEM_COLS));
}

// ---------------------------------------------------------------------------------------------------------------------
// K1T, split form (the default). Two warps per TMEM lane quarter read one accumulator: warps that share a quarter share
// its 32 sources and take one 128-target sub-tile of the unit each, so a (slab, tile) unit is read by 8 warps at once and
// a handshake with the MMA warp buys half as much reading per warp — which is what lets 16 epilogue warps (GROUPS = 2,
// two groups of 8, group g on accumulator g and the slabs of parity g) keep the ALU pipe busy across each other's
// handshakes: the MMA's issue -> commit -> visible latency (389 cycles for the two K=8 instructions, `tools/
// ubench_tc_filter lat`) plus the barrier round trips cost more than reading a unit. GROUPS = 1: one group of 8 warps on
// two accumulators used alternately.
// Two warps now work on one source concurrently, so its state is one 64-bit word
//     key = (bits of the exact threshold) << 32 | remembered sub-tile            (0xffffffff = none yet)
// updated with atomicMin in shared memory: smallest threshold first, lowest sub-tile among equal thresholds — exactly what
// visiting the sub-tiles in ascending order with a strict `<` gives, whatever order the warps get there in. For that the
// exact pass always takes the sub-tile's true minimum and offers it whenever its threshold is <= the current one (an
// equal distance in an EARLIER sub-tile must still win; the key order decides). The starting threshold is the largest
// float below the sentinel's, so that "<=" keeps the reference's strict `d < sentinel`. A stale threshold read can only
// cost an unnecessary exact pass, never change the result.
// ---------------------------------------------------------------------------------------------------------------------
// TPC = targets per MMA column: 1, or 2 / 4 = the grouped forms (tiles of 256 TPC targets, see tc_pack_group_kernel).
// The filter test, the exact pass and the remembered unit are QUARTERS of a warp's sub-tile: 32 columns = 32 TPC targets
// (the running minimum costs the same split four ways; an exact pass then covers a quarter of the targets).
template <int MODE, int GROUPS, int SLABS, int STAGES, int LDW, int TPC, int QC, int NACC, bool SORTED>
__global__ void __launch_bounds__(64 + 256 * GROUPS, 1) k1_filter_tc_split(const KTParams p)
{
	constexpr int TILE_T = TC_TN * TPC;                  // targets per tile
	constexpr int TRK_T = TC_TRK * TPC;                  // targets per sub-tile (filter test / tracking unit)
	constexpr int TILE_FLOATS = tcg_tile_floats(TPC);
	constexpr int QT = QC * TPC;                         // targets per "quarter" (QC columns): the unit of the exact pass and of the key's index
	constexpr int QPT = TILE_T / QT;                     // quarters per tile: 256 / QC
	constexpr int TILE_BYTES = TILE_FLOATS * 4;
	// NACC = 1: one 256-column accumulator per group, a unit = (slab, tile). NACC = 2: two 128-column accumulators per group,
	// a unit = (slab, half tile): the MMA of the next half runs while the group reads this one.
	constexpr int UNIT_COLS = TC_TN / NACC;
	constexpr int WCOLS = UNIT_COLS / 2;                 // columns a warp reads per unit (two warps per TMEM lane quarter)
	constexpr int QPS = WCOLS / QC;                      // exact-pass units per warp and MMA unit
	constexpr int NBAR = (GROUPS == 2) ? 2 * NACC : 2;   // accumulators
	constexpr int TMEM_COLS = 512;
	static_assert(NACC == 1 || (NACC == 2 && GROUPS == 2), "double-buffered accumulators need the two-group form");
	constexpr int SBN = 128 * SLABS;                     // sources per source block
	constexpr int NEPI = 8 * GROUPS;                     // epilogue warps
	constexpr int SPT = SBN / (32 * NEPI);               // sources each epilogue thread sets up and flushes
	static_assert(GROUPS == 1 || GROUPS == 2, "one or two epilogue groups");
	static_assert(SLABS % (2 * GROUPS) == 0 && SLABS <= TC_SLABS && SPT >= 1, "slab count");
	if (p.done != nullptr && *p.done) return;
	extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
	unsigned char* tc_smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);      // 1 KB alignment by hand
	__shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[NBAR], tempty_bar[NBAR];
	__shared__ uint32_t tmem_base_s;
	__shared__ unsigned long long s_u0;
	__shared__ int s_len, s_fail;

	float* a_slabs = reinterpret_cast<float*>(tc_smem);
	float* ring    = reinterpret_cast<float*>(tc_smem + (size_t)SLABS * TC_A_BYTES);
	u64*   key_s   = reinterpret_cast<u64*>(tc_smem + (size_t)SLABS * TC_A_BYTES + (size_t)STAGES * TILE_BYTES);
	float* ox_s    = reinterpret_cast<float*>(key_s + SBN);
	float* oy_s    = ox_s + SBN;
	float* oz_s    = oy_s + SBN;
	float* kk_s    = oz_s + SBN;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	// the two single-thread roles get the HIGHEST warp indices: the scheduler favours high warp ids among eligible warps, and
	// every cycle the MMA issuer waits for an issue slot behind busy epilogue warps is added to each unit's handshake
	const bool is_epi = warp < NEPI;
	const int ew = warp;                      // epilogue warp 0 .. NEPI-1
	const int w_tma = NEPI, w_mma = NEPI + 1;
	const int quarter = warp & 3;             // TMEM lane quarter a warp may read: warp index mod 4
	const int wq = ew >> 2;                   // which of the quarter's warps this is: 0 .. 2*GROUPS-1
	const int half = wq & 1;                  // the 128-target sub-tile of a unit this warp reads
	const int grp = wq >> 1;                  // epilogue group (GROUPS = 2), else 0
	const int row = quarter * 32 + lane;      // source row inside a slab

	if (tid == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NEPI); }
		for (int a = 0; a < NBAR; a++) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
		s_fail = 0;
		fence_mbar_init();
	}
	if (warp == w_mma) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = tmem_base_s;
	// kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UNIT_COLS >> 3) << 17) | ((128u >> 4) << 24);

	int it = 0;                   // tiles this CTA has streamed so far: ring stage and parities follow it across segments
	unsigned long long n_tests = 0, n_exact = 0;
	unsigned n_sweeps = 0, n_small_s0 = 0;      // grouped form: source sweeps started, of which with s0 below the largest group radius
	const float inf = __int_as_float(0x7f800000);
	const float one8u = 1.0f + 8.0f * TC_U;
	// largest float below the starting threshold: with it, "threshold <= current" is the reference's strict d < sentinel
	const float thr_start = (p.thr0 > 0.0f) ? __uint_as_float(__float_as_uint(p.thr0) - 1u) : -1.0f;
	bool failed = false;

	while (true) {
		__syncthreads();
		if (tid == 0) {
			const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(p.work_counter);
			const long long left = p.total_units - (long long)seen;
			long long len = left / ((long long)p.gss_div * (long long)gridDim.x);
			if (len < p.min_chunk) len = p.min_chunk;
			if (len > p.max_chunk) len = p.max_chunk;
			s_len = (int)len;
			s_u0 = atomicAdd(p.work_counter, (unsigned long long)len);
		}
		__syncthreads();
		if ((long long)s_u0 >= p.total_units || s_fail) break;
		long long u = (long long)s_u0;
		const long long u_end = min(p.total_units, u + (long long)s_len);
		while (u < u_end) {
			const int sb = (int)(u / p.nt);
			const int t0 = (int)(u - (long long)sb * p.nt);
			const int t1 = (int)min((long long)p.nt, (long long)t0 + (u_end - u));
			u += t1 - t0;
			__syncthreads();                       // previous segment fully consumed (A slabs, per-source state, ring)
			// ---- per-source state and the A operand of the segment's sources (SPT sources per epilogue thread) ----
			if (is_epi) {
#pragma unroll
				for (int q = 0; q < SPT; q++) {
					const int a = wq + 2 * GROUPS * q;           // the slabs of this thread: wq, wq + 2 GROUPS, ...
					const int sidx = a * 128 + row;
					const int i = sb * SBN + sidx;
					// the warp's 32 rows should be neighbours in space (one exact pass serves them all): sources in arbitrary order are
					// taken in Morton order through src_perm; everything per source below is indexed by gi
					const int gi = (p.src_perm != nullptr && i < p.n) ? __ldg(p.src_perm + i) : i;
					const float x = p.px[gi], y = p.py[gi], z = p.pz[gi];    // arrays are padded beyond n
					ox_s[sidx] = x; oy_s[sidx] = y; oz_s[sidx] = z;
					const float pcx = __fsub_rn(x, p.cx), pcy = __fsub_rn(y, p.cy), pcz = __fsub_rn(z, p.cz);
					const float p2 = __fmaf_rn(pcz, pcz, __fmaf_rn(pcx, pcx, __fmul_rn(pcy, pcy)));
					const float p2lo = __fmul_rd(p2, 1.0f - 8.0f * TC_U);
					const float rp = __fmul_ru(__fsqrt_ru(p2), 1.0f + 8.0f * TC_U);
					// sources past the end of the cloud, and non-finite ones, never ask for an exact pass: kk = -inf makes tau = -inf
					const bool live = (i < p.n) && (p2 == p2) && (p2 < inf);
					float th = thr_start;
					if (p.seed_idx != nullptr && i < p.n) {
						const int j0 = p.seed_idx[gi];
						if (j0 >= 0 && j0 < p.m) {
							const float4 qq = __ldg(p.q4 + j0);
							const float us = dist_chain(x, y, z, qq.x, qq.y, qq.z);
							const float up = (MODE == ICPB_DIST_SQRT) ? __fmul_ru(us, one8u) : us;     // covers the floats sqrt.rn merges with us
							if (us == us && up < th) th = up;            // the seed itself (or an equal, lower-indexed target) is offered by the exact pass
						}
					}
					key_s[sidx] = ((u64)__float_as_uint(th) << 32) | 0xffffffffull;
					// grouped form: s0 >= sqrt of everything the threshold can be compared at during this sweep (it only falls), up to TF32
					float s0 = 0.0f;
					bool wide = false;                       // no usable starting threshold: every quarter goes to the exact pass
					if (TPC >= 2) {
						if (th > 0.0f && th < 1.0e30f) s0 = tf32_ru_pos(__fmul_ru(__fsqrt_ru(__fmul_ru(th, one8u)), 1.0f + 4.0f * TC_U));
						else if (!(th <= 0.0f)) wide = true;
					}
					float e = __fmul_ru(8.0f * p.rq, p.rq);
					e = __fmaf_ru(10.0f * rp, p.rq, e);
					e = __fmaf_ru(2.0f * rp, rp, e);
					if (TPC >= 2) { e = __fmaf_ru(8.0f * p.hmax, p.hmax, e); e = __fmaf_ru(4.0f * s0, p.hmax, e); }
					if (TPC >= 2 && live && !wide) { n_sweeps += 1; n_small_s0 += (s0 < p.hmax) ? 1u : 0u; }
					e = __fmul_ru(e, 1.05f * TC_EPS_SCALE * TC_U);
					kk_s[sidx] = live ? (wide ? inf : __fsub_ru(e, p2lo)) : -inf;
					// A row: a = -2 (p - c) as hi + lo; slots ax_hi ax_hi ax_lo | ay.. | az.. | 1 1 | 0...
					float* A = a_slabs + (size_t)a * (TC_A_BYTES / 4);
					float h, l;
					const float ax = live ? -2.0f * pcx : 0.0f, ay = live ? -2.0f * pcy : 0.0f, az = live ? -2.0f * pcz : 0.0f;
					tf32_split(ax, h, l); A[tc_elem(row, 0)] = h; A[tc_elem(row, 1)] = h; A[tc_elem(row, 2)] = l;
					tf32_split(ay, h, l); A[tc_elem(row, 3)] = h; A[tc_elem(row, 4)] = h; A[tc_elem(row, 5)] = l;
					tf32_split(az, h, l); A[tc_elem(row, 6)] = h; A[tc_elem(row, 7)] = h; A[tc_elem(row, 8)] = l;
					A[tc_elem(row, 9)] = 1.0f; A[tc_elem(row, 10)] = 1.0f;
					A[tc_elem(row, 11)] = live ? s0 : 0.0f;              // against -2 h of the column (0 for one target per column)
#pragma unroll
					for (int k = 12; k < TC_K; k++) A[tc_elem(row, k)] = 0.0f;
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes of A -> visible to the tensor core
			}
			__syncthreads();

			if (warp == w_tma) {
				// ---- TMA producer ----
				if (lane == 0) {
					for (int t = t0; t < t1 && !failed; t++) {
						const int k = it + (t - t0), st = k % STAGES, use = k / STAGES;
						if (use >= 1 && !mbar_wait_bounded(&empty_bar[st], (uint32_t)((use - 1) & 1))) { failed = true; break; }
						mbar_expect_tx(&full_bar[st], TILE_BYTES);
						tma_load_1d(ring + (size_t)st * TILE_FLOATS, p.tiles + (size_t)t * TILE_FLOATS, TILE_BYTES, &full_bar[st]);
					}
				}
			} else if (warp == w_mma) {
				// ---- MMA issuer ----
				if (lane == 0) {
					for (int t = t0; t < t1 && !failed; t++) {
						const int k = it + (t - t0), st = k % STAGES, use = k / STAGES;
						if (!mbar_wait_bounded(&full_bar[st], (uint32_t)(use & 1))) { failed = true; break; }
						const uint32_t sB = smem_u32(ring + (size_t)st * TILE_FLOATS);
#pragma unroll 1
						for (int a = 0; a < SLABS && !failed; a++) {
							// GROUPS = 2: slab a belongs to group a & 1; its NACC unit(s) go to the group's accumulator(s). GROUPS = 1: the units alternate
							const int v = (GROUPS == 2) ? (k * (SLABS / 2) + (a >> 1)) : (k * SLABS + a);
							const uint32_t sA = smem_u32(a_slabs + (size_t)a * (TC_A_BYTES / 4));
#pragma unroll 1
							for (int hs = 0; hs < NACC; hs++) {
								const int acc = (GROUPS == 2) ? ((a & 1) * NACC + hs) : (v & 1);
								const int j = (GROUPS == 2) ? v : (v >> 1);                       // use counter of accumulator `acc`
								if (j >= 1 && !mbar_wait_bounded(&tempty_bar[acc], (uint32_t)((j - 1) & 1))) { failed = true; break; }
								asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
								for (int kk2 = 0; kk2 < 2; kk2++) {
									// K advances by two 128 B core matrices; half hs starts at row hs * UNIT_COLS of the B block (row groups of 8 rows, 512 B apart)
									const uint64_t dA = umma_desc(sA + kk2 * 256), dB = umma_desc(sB + hs * (UNIT_COLS / 8) * 512 + kk2 * 256);
									asm volatile("{\n.reg .pred pacc;\nsetp.ne.b32 pacc, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pacc;\n}\n"
									             :: "r"(tmem_base + (uint32_t)(acc * UNIT_COLS)), "l"(dA), "l"(dB), "r"(idesc), "r"((uint32_t)kk2));
								}
								asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&tfull_bar[acc])) : "memory");
							}
						}
					}
				}
			} else {
				// ---- epilogue: minimum over this warp's sub-tile, test against tau, exact pass where it cannot be excluded ----
				for (int t = t0; t < t1 && !failed; t++) {
					const int k = it + (t - t0), st = k % STAGES, use = k / STAGES;
					if (!mbar_wait_bounded(&full_bar[st], (uint32_t)(use & 1))) { failed = true; break; }
					const float* tile = ring + (size_t)st * TILE_FLOATS;
					const float4* X4 = reinterpret_cast<const float4*>(tile + TC_B_FLOATS);
					const float4* Y4 = X4 + TILE_T / 4;
					const float4* Z4 = Y4 + TILE_T / 4;
#pragma unroll 1
					for (int un = 0; un < (SLABS / GROUPS) * NACC && !failed; un++) {
						// (a, hs): the group's slabs in ascending order, each slab's NACC units one after the other
						const int a = grp + GROUPS * (un / NACC), hs = un % NACC;
						const int sidx = a * 128 + row;
						const int v = (GROUPS == 2) ? (k * (SLABS / 2) + (a >> 1)) : (k * SLABS + a);
						const int acc = (GROUPS == 2) ? (grp * NACC + hs) : (v & 1);
						const int j = (GROUPS == 2) ? v : (v >> 1);
						if (!mbar_wait_bounded(&tfull_bar[acc], (uint32_t)(j & 1))) { failed = true; break; }
						asm volatile("tcgen05.fence::after_thread_sync;");
						const uint32_t taddr = tmem_base + (uint32_t)(acc * UNIT_COLS + half * WCOLS) + ((uint32_t)(quarter * 32) << 16);
						float mq[QPS];
						subtile_mins<QC, WCOLS>(taddr, mq);
						// the accumulator is in registers: hand it back to the MMA warp before the (rare) exact pass
						asm volatile("tcgen05.fence::before_thread_sync;");
						__syncwarp();
						if (lane == 0) mbar_arrive(&tempty_bar[acc]);
						const float kk = kk_s[sidx];                                    // -inf for dead rows
						float th = __uint_as_float((unsigned)(key_s[sidx] >> 32));
						float tau = __fadd_ru(__fmul_ru(th, one8u), kk);
						// most sub-tiles have nothing below tau: one vote on the minimum of the partial minima settles them
						// (a threshold that another warp lowers meanwhile only costs an unnecessary exact pass)
						float mall = mq[0];
#pragma unroll
						for (int qq = 1; qq + 1 < QPS; qq += 2) mall = min3(mall, mq[qq], mq[qq + 1]);
						mall = fminf(mall, mq[QPS - 1]);
						n_tests += QPS;
						if (__ballot_sync(0xffffffffu, mall <= tau) == 0u) continue;
#pragma unroll
						for (int qq = 0; qq < QPS; qq++) {
							const unsigned need = __ballot_sync(0xffffffffu, mq[qq] <= tau);
							if (need) {                                       // warp-uniform
								n_exact += 1;
								const int q_in_tile = (hs * UNIT_COLS + half * WCOLS) / QC + qq;      // exact-pass unit of the tile: slots [QT q, QT q + QT)
								const int j0 = q_in_tile * (QT / 4), j1 = j0 + QT / 4;
								const float sx = ox_s[sidx], sy = oy_s[sidx], sz = oz_s[sidx];
								const u64 PX = pack2(sx, sx), PY = pack2(sy, sy), PZ = pack2(sz, sz);
								float mm = inf;                               // the quarter's true minimum (see the tie rule above)
#pragma unroll 4
								for (int jq = j0; jq < j1; jq++) {
									const float4 Xo = X4[jq], Yo = Y4[jq], Zo = Z4[jq];
									u64 dx = sub2(PX, pack2(Xo.x, Xo.y)), dy = sub2(PY, pack2(Yo.x, Yo.y)), dz = sub2(PZ, pack2(Zo.x, Zo.y));
									u64 d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
									float aa, bb;
									unpack2(d, aa, bb);
									mm = min3(mm, aa, bb);
									dx = sub2(PX, pack2(Xo.z, Xo.w)); dy = sub2(PY, pack2(Yo.z, Yo.w)); dz = sub2(PZ, pack2(Zo.z, Zo.w));
									d  = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
									unpack2(d, aa, bb);
									mm = min3(mm, aa, bb);
								}
								// offered whenever its threshold is <= the one read above — compared AFTER lower_threshold: in sqrt mode a
								// distance of the current class above the class floor still ties with it (NaN / inf never pass)
								const float nt_ = (mm < inf) ? lower_threshold<MODE>(mm) : inf;
								if (nt_ <= th && kk > -inf) {
									if constexpr (SORTED) {
										// slots are in Morton order: the smallest ORIGINAL index among the slots that attain the minimum goes into the key
										const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(nt_) : nt_;
										const int* so = p.slot_index + (size_t)t * TILE_T;
										int bi = 0x7fffffff;
										for (int jq = j0; jq < j1; jq++) {
											const float4 X = X4[jq], Y = Y4[jq], Z = Z4[jq];
											float d0 = dist_chain(sx, sy, sz, X.x, Y.x, Z.x);
											float d1 = dist_chain(sx, sy, sz, X.y, Y.y, Z.y);
											float d2 = dist_chain(sx, sy, sz, X.z, Y.z, Z.z);
											float d3 = dist_chain(sx, sy, sz, X.w, Y.w, Z.w);
											if (MODE == ICPB_DIST_SQRT) { d0 = __fsqrt_rn(d0); d1 = __fsqrt_rn(d1); d2 = __fsqrt_rn(d2); d3 = __fsqrt_rn(d3); }
											if (d0 <= target) bi = min(bi, __ldg(so + 4 * jq));
											if (d1 <= target) bi = min(bi, __ldg(so + 4 * jq + 1));
											if (d2 <= target) bi = min(bi, __ldg(so + 4 * jq + 2));
											if (d3 <= target) bi = min(bi, __ldg(so + 4 * jq + 3));
										}
										if (bi != 0x7fffffff)
											atomicMin(reinterpret_cast<unsigned long long*>(key_s + sidx), ((u64)__float_as_uint(nt_) << 32) | (u64)(uint32_t)bi);
									} else {
										atomicMin(reinterpret_cast<unsigned long long*>(key_s + sidx), ((u64)__float_as_uint(nt_) << 32) | (u64)(uint32_t)(t * QPT + q_in_tile));
									}
								}
								th = __uint_as_float((unsigned)(key_s[sidx] >> 32));     // re-read: this pass (or the warp on the other half) may have lowered it
								tau = __fadd_ru(__fmul_ru(th, one8u), kk);
							}
						}
					}
					// this warp is done with the tile's originals: release the ring stage (all epilogue warps)
					__syncwarp();
					if (lane == 0) mbar_arrive(&empty_bar[st]);
				}
			}
			if (failed) s_fail = 1;
			it += t1 - t0;
			__syncthreads();
			if (s_fail) break;
			// ---- index recovery: the first target of the remembered quarter that attains the exact minimum ----
			// (done once per source and sweep here; finding the slot whenever the exact pass offers a quarter was measured
			// 15-60 % slower: it stalls the warp inside the pipeline, several times per source on cold passes)
			if (is_epi) {
#pragma unroll 1
				for (int q = 0; q < SPT; q++) {
					const int a = wq + 2 * GROUPS * q;
					const int sidx = a * 128 + row;
					const int i = sb * SBN + sidx;
					const u64 kv = key_s[sidx];
					const int bs = (int)(uint32_t)(kv & 0xffffffffull);          // -1: nothing below the starting threshold
					if (SORTED) {                                               // the key already holds the original index
						if (i < p.n && bs >= 0) {
							const float th = __uint_as_float((unsigned)(kv >> 32));
							const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(th) : th;
							const int gi = (p.src_perm != nullptr) ? __ldg(p.src_perm + i) : i;
							atomicMin(p.keys + gi, ((u64)__float_as_uint(target) << 32) | (u64)(uint32_t)bs);
						}
					} else if (i < p.n && bs >= 0) {
						const float th = __uint_as_float((unsigned)(kv >> 32));
						const float sx = ox_s[sidx], sy = oy_s[sidx], sz = oz_s[sidx];
						const float* gx = p.tiles + (size_t)(bs / QPT) * TILE_FLOATS + TC_B_FLOATS + (size_t)(bs % QPT) * QT;
						const float4* GX = reinterpret_cast<const float4*>(gx);
						const float4* GY = reinterpret_cast<const float4*>(gx + TILE_T);
						const float4* GZ = reinterpret_cast<const float4*>(gx + 2 * TILE_T);
						const float target = (MODE == ICPB_DIST_SQRT) ? __fsqrt_rn(th) : th;
						int found = -1;
						// the unit comes from L2: four float4 triples (16 slots) are requested together, one round trip per 16 slots
						static_assert((QT / 4) % 4 == 0, "whole groups of four float4 per unit");
						for (int jq = 0; jq < QT / 4 && found < 0; jq += 4) {
							float4 X[4], Y[4], Z[4];
#pragma unroll
							for (int u = 0; u < 4; u++) { X[u] = __ldg(GX + jq + u); Y[u] = __ldg(GY + jq + u); Z[u] = __ldg(GZ + jq + u); }
#pragma unroll
							for (int u = 3; u >= 0; u--) {               // descending, so that the lowest slot is what remains
								float d0 = dist_chain(sx, sy, sz, X[u].x, Y[u].x, Z[u].x);
								float d1 = dist_chain(sx, sy, sz, X[u].y, Y[u].y, Z[u].y);
								float d2 = dist_chain(sx, sy, sz, X[u].z, Y[u].z, Z[u].z);
								float d3 = dist_chain(sx, sy, sz, X[u].w, Y[u].w, Z[u].w);
								if (MODE == ICPB_DIST_SQRT) { d0 = __fsqrt_rn(d0); d1 = __fsqrt_rn(d1); d2 = __fsqrt_rn(d2); d3 = __fsqrt_rn(d3); }
								if (d3 <= target) found = 4 * (jq + u) + 3;
								if (d2 <= target) found = 4 * (jq + u) + 2;
								if (d1 <= target) found = 4 * (jq + u) + 1;
								if (d0 <= target) found = 4 * (jq + u);
							}
						}
						if (found >= 0) {
							// slot -> target index: one target per column, or column start + member (slots past a short column hold +inf)
							const int jidx = (TPC == 1) ? (bs * QT + found) : (__ldg(p.colstart + (size_t)bs * QC + found / TPC) + found % TPC);
							const u64 key = ((u64)__float_as_uint(target) << 32) | (u64)(uint32_t)jidx;
							const int gi = (p.src_perm != nullptr) ? __ldg(p.src_perm + i) : i;
							atomicMin(p.keys + gi, key);
						}
					}
				}
			}
		}
		if (s_fail) break;
	}
	__syncthreads();
	if (s_fail && tid == 0) *p.fail = 1;
	if (tid == 0) {
		// every CTA has taken its last grab before it gets here: the last one through re-arms the counters for the next pass
		__threadfence();
		if (atomicAdd(p.exit_counter, 1ull) == (unsigned long long)gridDim.x - 1ull) { *p.work_counter = 0ull; *p.exit_counter = 0ull; __threadfence(); }
	}
	if (p.stats != nullptr && is_epi) {      // one test = one (warp, slab, quarter of a sub-tile)
		for (int o = 16; o > 0; o >>= 1) { n_tests += __shfl_xor_sync(0xffffffffu, n_tests, o); n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o); }
		if (lane == 0) { atomicAdd(p.stats, n_tests / 32); atomicAdd(p.stats + 1, n_exact / 32); }
		if (TPC >= 2) {
			for (int o = 16; o > 0; o >>= 1) { n_sweeps += __shfl_xor_sync(0xffffffffu, n_sweeps, o); n_small_s0 += __shfl_xor_sync(0xffffffffu, n_small_s0, o); }
			if (lane == 0) { atomicAdd(p.stats + 2, (unsigned long long)n_sweeps); atomicAdd(p.stats + 3, (unsigned long long)n_small_s0); }
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == w_mma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TMEM_COLS));
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------

// Operand tiles of the current target (needs the centre chosen by build_filter_data); buffers kept across targets.
// kt_tpc >= 2: the grouped form — columns of up to kt_tpc consecutive targets that never span a jump of the scan.
template <int TPC> static void launch_pack_group(Ctx* c, int nt, int ncols, const float4* src, const int* perm, int* slot_index)
{
	tc_pack_group_kernel<TPC><<<nt * 2, 128, 0, c->stream>>>(src, c->m, c->kt_colstart, ncols, c->kf_center[0], c->kf_center[1], c->kf_center[2], c->kt_tiles, c->kt_hmax_d, perm, slot_index);
}
// targets per MMA column: forced by ICPB_KT_VAR (experiments), else what the exact-pass-rate policy currently holds
static int tc_current_tpc(const Ctx* c)
{
	if (c->kt_variant < 0) return c->kt_tpc_auto;
	const int v = c->kt_variant;
	if (v >= 21 && v <= 24) return 16 >> (v - 21);
	return (v == 19 || v == 20) ? 2 : (v == 17 || v == 18) ? 4 : (v == 15 || v == 16) ? 8 : (v >= 12) ? 16 : (v == 11) ? 8 : (v >= 10) ? 4 : (v >= 8 ? 2 : 1);
}
int build_filter_tc_data(Ctx* c)
{
	const int m = c->m;
	const int tpc = c->kt_tpc = tc_current_tpc(c);
	if (!c->kt_fail) { ICPB_CUDA(c, cudaMalloc((void**)&c->kt_fail, sizeof(int))); ICPB_CUDA(c, cudaMemsetAsync(c->kt_fail, 0, sizeof(int), c->stream)); }
	int nt, ncols = 0;
	bool sorted = false;
	const float4* src = c->q4;           // the cloud the columns are cut from: the targets in index order, or in Morton order
	const int* perm = nullptr;
	if (tpc >= 2) {
		if ((size_t)m + 1 > c->kt_cols_cap) {
			cudaFree(c->kt_colstart); cudaFree(c->kt_scan_a); cudaFree(c->kt_scan_b); c->kt_colstart = c->kt_scan_a = c->kt_scan_b = nullptr; c->kt_cols_cap = 0;
			const size_t cap = (size_t)m + 1 + (size_t)m / 8;
			ICPB_CUDA(c, cudaMalloc((void**)&c->kt_colstart, sizeof(int) * cap));
			ICPB_CUDA(c, cudaMalloc((void**)&c->kt_scan_a, sizeof(int) * cap));
			ICPB_CUDA(c, cudaMalloc((void**)&c->kt_scan_b, sizeof(int) * cap));
			c->kt_cols_cap = cap;
		}
		if (!c->kt_hmax_d) ICPB_CUDA(c, cudaMalloc((void**)&c->kt_hmax_d, 2 * sizeof(double)));      // [0] float hmax (+ pad), [1] double step sum
		double* step_sum = reinterpret_cast<double*>(c->kt_hmax_d) + 1;
		const int g = (m + 255) / 256;
		// Is the scan order coherent? Consecutive targets that are not neighbours in space (a cloud in arbitrary order) make
		// every group as wide as the cloud. The mean step of the scan against the spacing of m points spread over a surface
		// of the cloud's radius tells.
		double step_total = 0.0;
		ICPB_CUDA(c, cudaMemsetAsync(c->kt_hmax_d, 0, 2 * sizeof(double), c->stream));
		tcg_step_sum_kernel<<<g, 256, 0, c->stream>>>(c->q4, m, step_sum);
		ICPB_CUDA(c, cudaMemcpyAsync(&step_total, step_sum, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
		const bool scattered = m > 1 && step_total / (double)(m - 1) > 16.0 * (double)c->kf_rq * sqrt(3.14159265358979 / (double)m);
		// ICPB_KT_SORT: 1 always, 0 never; default: when the scan order is incoherent, and for large clouds whatever their order —
		// a group of 16 along the Z curve is a compact patch (h about 2 spacings), along a raster row a strip (7.5 spacings), and a
		// warp's 32 Morton-ordered sources share their units: measured on the raster saddle 6.7 ms against 8.4 ms per pass at 1000^2
		// points, equal at 317^2, 15 % slower at 128^2
		if (c->kt_sort_mode == 1 || (c->kt_sort_mode < 0 && (scattered || m >= TCG_SORT_MIN))) sorted = true;
		else if (scattered && c->kt_variant < 0) { c->kt_tpc_auto = 1; return build_filter_tc_data(c); }   // sorting switched off: one target per column
		size_t tmp_a = 0, tmp_b = 0, tmp_c = 0;
		cub::DeviceScan::InclusiveScan(nullptr, tmp_a, c->kt_scan_a, c->kt_scan_a, TcgMax(), m, c->stream);
		cub::DeviceScan::InclusiveSum(nullptr, tmp_b, c->kt_scan_b, c->kt_scan_b, m, c->stream);
		if (sorted) {
			if ((size_t)m > c->kt_sort_cap) {
				cudaFree(c->kt_mkeys); cudaFree(c->kt_perm2); cudaFree(c->kt_q4s); c->kt_mkeys = nullptr; c->kt_perm2 = nullptr; c->kt_q4s = nullptr; c->kt_sort_cap = 0;
				const size_t cap = (size_t)m + (size_t)m / 8 + 1;
				ICPB_CUDA(c, cudaMalloc((void**)&c->kt_mkeys, sizeof(unsigned long long) * 2 * cap));
				ICPB_CUDA(c, cudaMalloc((void**)&c->kt_perm2, sizeof(int) * 2 * cap));
				ICPB_CUDA(c, cudaMalloc((void**)&c->kt_q4s, sizeof(float4) * cap));
				c->kt_sort_cap = cap;
			}
			cub::DeviceRadixSort::SortPairs(nullptr, tmp_c, c->kt_mkeys, c->kt_mkeys + c->kt_sort_cap, c->kt_perm2, c->kt_perm2 + c->kt_sort_cap, m, 0, 63, c->stream);
		}
		const size_t tmp = std::max(tmp_a, std::max(tmp_b, tmp_c));
		if (tmp > c->kt_cub_cap) {
			cudaFree(c->kt_cub_tmp); c->kt_cub_tmp = nullptr; c->kt_cub_cap = 0;
			ICPB_CUDA(c, cudaMalloc(&c->kt_cub_tmp, tmp + 256));
			c->kt_cub_cap = tmp + 256;
		}
		if (sorted) {
			// Morton order over the cube [centre - Rq, centre + Rq]^3, then the same column pipeline over the sorted cloud
			const float inv = 0.5f / c->kf_rq;
			tcg_morton_kernel<<<g, 256, 0, c->stream>>>(c->q4, m, c->kf_center[0] - c->kf_rq, c->kf_center[1] - c->kf_rq, c->kf_center[2] - c->kf_rq, inv, c->kt_mkeys, c->kt_perm2);
			size_t t0 = c->kt_cub_cap;
			ICPB_CUDA(c, cub::DeviceRadixSort::SortPairs(c->kt_cub_tmp, t0, c->kt_mkeys, c->kt_mkeys + c->kt_sort_cap, c->kt_perm2, c->kt_perm2 + c->kt_sort_cap, m, 0, 63, c->stream));
			perm = c->kt_perm2 + c->kt_sort_cap;
			tcg_gather_kernel<<<g, 256, 0, c->stream>>>(c->q4, perm, m, c->kt_q4s);
			src = c->kt_q4s;
			ICPB_CUDA(c, cudaMemsetAsync(c->kt_hmax_d, 0, 2 * sizeof(double), c->stream));
			tcg_step_sum_kernel<<<g, 256, 0, c->stream>>>(src, m, step_sum);           // the jump threshold of the sorted scan
			c->launches += 4;
		}
		tcg_break_kernel<<<g, 256, 0, c->stream>>>(src, m, step_sum, c->kt_scan_a);
		size_t t1 = c->kt_cub_cap;
		ICPB_CUDA(c, cub::DeviceScan::InclusiveScan(c->kt_cub_tmp, t1, c->kt_scan_a, c->kt_scan_a, TcgMax(), m, c->stream));      // run start of every target
		tcg_colflag_kernel<<<g, 256, 0, c->stream>>>(c->kt_scan_a, m, tpc, c->kt_scan_b);
		t1 = c->kt_cub_cap;
		ICPB_CUDA(c, cub::DeviceScan::InclusiveSum(c->kt_cub_tmp, t1, c->kt_scan_b, c->kt_scan_b, m, c->stream));                 // 1-based column of every target
		tcg_colstart_kernel<<<g, 256, 0, c->stream>>>(c->kt_scan_a, c->kt_scan_b, m, tpc, c->kt_colstart);
		c->launches += 6;
		ICPB_CUDA(c, cudaMemcpyAsync(&ncols, c->kt_scan_b + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
		if (ncols < 1 || ncols > m) { snprintf(c->err, sizeof c->err, "k1_filter_tc: column scan returned %d columns for %d targets", ncols, m); return ICPB_ERR_CUDA; }
		nt = (ncols + TC_TN - 1) / TC_TN;
	} else {
		nt = (m + TC_TN - 1) / TC_TN;
	}
	const size_t need = (size_t)nt * tcg_tile_floats(tpc);
	if (need > c->kt_tiles_cap) {
		cudaFree(c->kt_tiles); c->kt_tiles = nullptr; c->kt_tiles_cap = 0;
		ICPB_CUDA(c, cudaMalloc((void**)&c->kt_tiles, sizeof(float) * need));
		c->kt_tiles_cap = need;
	}
	int* slot_index = nullptr;
	if (sorted) {
		const size_t slots = (size_t)nt * TC_TN * tpc;
		if (slots > c->kt_slot_cap) {
			cudaFree(c->kt_slot_index); c->kt_slot_index = nullptr; c->kt_slot_cap = 0;
			ICPB_CUDA(c, cudaMalloc((void**)&c->kt_slot_index, sizeof(int) * slots));
			c->kt_slot_cap = slots;
		}
		slot_index = c->kt_slot_index;
	}
	c->kt_hmax = 0.0f;
	if (tpc >= 2) ICPB_CUDA(c, cudaMemsetAsync(c->kt_hmax_d, 0, sizeof(float), c->stream));
	if (tpc == 16) launch_pack_group<16>(c, nt, ncols, src, perm, slot_index);
	else if (tpc == 8) launch_pack_group<8>(c, nt, ncols, src, perm, slot_index);
	else if (tpc == 4) launch_pack_group<4>(c, nt, ncols, src, perm, slot_index);
	else if (tpc == 2) launch_pack_group<2>(c, nt, ncols, src, perm, slot_index);
	else tc_pack_kernel<<<(nt * TC_TN + 255) / 256, 256, 0, c->stream>>>(c->q4, m, nt * TC_TN, c->kf_center[0], c->kf_center[1], c->kf_center[2], c->kt_tiles);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	if (tpc >= 2) {
		ICPB_CUDA(c, cudaMemcpyAsync(&c->kt_hmax, c->kt_hmax_d, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
		ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	c->kt_nt = nt;
	c->kt_ncols = ncols;
	c->kt_sorted = sorted;
	c->graph_gen++;              // a captured iteration graph holds the tile count, hmax and (when the group size changed) the kernel of the old tiles
	c->kt_built_tpc = tpc;
	c->kt_ready = true;
	return ICPB_OK;
}

// Once per source upload: are consecutive sources neighbours in space? The exact pass is warp-uniform — it runs for a unit
// as soon as one of a warp's 32 sources needs it — so a warp should hold 32 neighbours. Sources in arbitrary order are
// visited in Morton order (src_perm); the order is kept while the cloud moves rigidly from iteration to iteration.
int ensure_source_order(Ctx* c)
{
	if (c->kt_src_checked) return ICPB_OK;
	c->kt_src_checked = true;
	const bool was_sorted = c->kt_src_sorted;          // a captured iteration graph holds the src_perm pointer (or its absence)
	c->kt_src_sorted = false;
	const int n = c->n;
	if (n < 1024 || c->kt_sort_mode == 0 || !c->kf_ready) { if (was_sorted) c->graph_gen++; return ICPB_OK; }
	if (!c->kt_hmax_d) ICPB_CUDA(c, cudaMalloc((void**)&c->kt_hmax_d, 2 * sizeof(double)));
	double* step_sum = reinterpret_cast<double*>(c->kt_hmax_d) + 1;
	const int g = (n + 255) / 256;
	double step_total = 0.0;
	ICPB_CUDA(c, cudaMemsetAsync(step_sum, 0, sizeof(double), c->stream));
	tcg_step_sum_soa_kernel<<<g, 256, 0, c->stream>>>(c->px, c->py, c->pz, n, step_sum);
	c->launches++;
	ICPB_CUDA(c, cudaMemcpyAsync(&step_total, step_sum, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	const bool scattered = step_total / (double)(n - 1) > 16.0 * (double)c->kf_rq * sqrt(3.14159265358979 / (double)n);
	if (!(c->kt_sort_mode == 1 || scattered || c->m >= TCG_SORT_MIN)) { if (was_sorted) c->graph_gen++; return ICPB_OK; }
	bool moved = false;
	if ((size_t)n > c->kt_ssort_cap) {
		moved = true;
		cudaFree(c->kt_skeys); cudaFree(c->kt_sperm2); c->kt_skeys = nullptr; c->kt_sperm2 = nullptr; c->kt_ssort_cap = 0;
		const size_t cap = (size_t)n + (size_t)n / 8 + 1;
		ICPB_CUDA(c, cudaMalloc((void**)&c->kt_skeys, sizeof(unsigned long long) * 2 * cap));
		ICPB_CUDA(c, cudaMalloc((void**)&c->kt_sperm2, sizeof(int) * 2 * cap));
		c->kt_ssort_cap = cap;
	}
	size_t tmp = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->kt_skeys, c->kt_skeys + c->kt_ssort_cap, c->kt_sperm2, c->kt_sperm2 + c->kt_ssort_cap, n, 0, 63, c->stream);
	if (tmp > c->kt_cub_cap) {
		cudaFree(c->kt_cub_tmp); c->kt_cub_tmp = nullptr; c->kt_cub_cap = 0;
		ICPB_CUDA(c, cudaMalloc(&c->kt_cub_tmp, tmp + 256));
		c->kt_cub_cap = tmp + 256;
	}
	const float inv = 0.5f / c->kf_rq;            // the target's cube: the source lies on it (or is clamped to its faces)
	tcg_morton_soa_kernel<<<g, 256, 0, c->stream>>>(c->px, c->py, c->pz, n, c->kf_center[0] - c->kf_rq, c->kf_center[1] - c->kf_rq, c->kf_center[2] - c->kf_rq, inv, c->kt_skeys, c->kt_sperm2);
	size_t t0 = c->kt_cub_cap;
	ICPB_CUDA(c, cub::DeviceRadixSort::SortPairs(c->kt_cub_tmp, t0, c->kt_skeys, c->kt_skeys + c->kt_ssort_cap, c->kt_sperm2, c->kt_sperm2 + c->kt_ssort_cap, n, 0, 63, c->stream));
	c->launches += 2;
	c->kt_src_sorted = true;
	if (!was_sorted || moved) c->graph_gen++;
	return ICPB_OK;
}

// tiles present and laid out for the group size the next launch will use (called before the clock and before a capture)
int ensure_filter_tc_data(Ctx* c)
{
	int rc;
	if (!(c->kt_ready && c->kt_built_tpc == tc_current_tpc(c)) && (rc = build_filter_tc_data(c)) != ICPB_OK) return rc;
	return ensure_source_order(c);
}

static float sqrt_domain_threshold_tc(float sentinel)
{
	if (!(sentinel > 0.0f)) return 0.0f;
	float y = sentinel * sentinel;
	if (std::isinf(y)) return y;
	while (sqrtf(y) < sentinel) y = nextafterf(y, INFINITY);
	while (y > 0.0f && sqrtf(nextafterf(y, 0.0f)) >= sentinel) y = nextafterf(y, 0.0f);
	return y;
}

template <int NS, int NACC_G, int SLABS, int STAGES, int MINB, int LDW>
static int launch_tc_variant(Ctx* c, int dist_mode, KTParams& p, int variant)
{
	constexpr int SBN = 128 * SLABS;
	constexpr size_t SMEM = (size_t)SLABS * TC_A_BYTES + (size_t)STAGES * TC_TILE_BYTES + (size_t)6 * SBN * 4 + 1024;
	const int nb = (c->n + SBN - 1) / SBN;
	p.total_units = (long long)nb * p.nt;
	// guided self-scheduling as in K1F, in tiles of 256 targets x SBN sources
	p.min_chunk = 8; p.max_chunk = 128; p.gss_div = 4;
	if (c->kf_gss[0] > 0) { p.min_chunk = c->kf_gss[0]; p.max_chunk = c->kf_gss[1]; p.gss_div = c->kf_gss[2]; }
	if (c->kf_chunk_override > 0) p.min_chunk = p.max_chunk = c->kf_chunk_override;
	if (p.max_chunk > p.nt) p.max_chunk = p.nt;
	if (p.min_chunk > p.max_chunk) p.min_chunk = p.max_chunk;
	auto kern = (dist_mode == ICPB_DIST_SQRT) ? k1_filter_tc<ICPB_DIST_SQRT, NS, NACC_G, SLABS, STAGES, MINB, LDW> : k1_filter_tc<ICPB_DIST_SQ, NS, NACC_G, SLABS, STAGES, MINB, LDW>;
	static bool attr_set[2][8][64] = {};
	bool& done = attr_set[dist_mode == ICPB_DIST_SQRT ? 1 : 0][variant & 7][c->device & 63];
	if (!done) { ICPB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM)); done = true; }
	long long grid = (long long)c->sm_count * MINB;             // the CTAs of an SM share its 512 TMEM columns
	if (c->k1_grid_override > 0) grid = c->k1_grid_override;
	const long long max_ctas = (p.total_units + p.min_chunk - 1) / p.min_chunk;
	if (grid > max_ctas) grid = max_ctas;
	ICPB_CUDA(c, cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), c->stream));
	kern<<<(unsigned)grid, TC_THREADS, SMEM, c->stream>>>(p);
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

template <int GROUPS, int SLABS, int STAGES, int LDW, int TPC, int QC = 32, int NACC = 1, bool SORTED = false>
static int launch_tc_split(Ctx* c, int dist_mode, KTParams& p, int variant)
{
	constexpr int SBN = 128 * SLABS;
	constexpr size_t TILE_BYTES = (size_t)tcg_tile_floats(TPC) * 4;
	constexpr size_t SMEM = (size_t)SLABS * TC_A_BYTES + (size_t)STAGES * TILE_BYTES + (size_t)6 * SBN * 4 + 1024;
	const int nb = (c->n + SBN - 1) / SBN;
	p.total_units = (long long)nb * p.nt;
	// same number of targets per grab whatever TPC; up to 128k targets per segment: every segment costs a source set-up, three
	// block-wide barriers and an index recovery (measured at 1M x 1M: 8-tile segments 9.25 ms, 32-tile 8.35 ms per pass)
	p.min_chunk = (8 / TPC) > 0 ? 8 / TPC : 1; p.max_chunk = (512 / TPC) > 1 ? 512 / TPC : 2; p.gss_div = 4;
	if (c->kf_gss[0] > 0) { p.min_chunk = c->kf_gss[0]; p.max_chunk = c->kf_gss[1]; p.gss_div = c->kf_gss[2]; }
	if (c->kf_chunk_override > 0) p.min_chunk = p.max_chunk = c->kf_chunk_override;
	if (p.max_chunk > p.nt) p.max_chunk = p.nt;
	if (p.min_chunk > p.max_chunk) p.min_chunk = p.max_chunk;
	auto kern = (dist_mode == ICPB_DIST_SQRT) ? k1_filter_tc_split<ICPB_DIST_SQRT, GROUPS, SLABS, STAGES, LDW, TPC, QC, NACC, SORTED> : k1_filter_tc_split<ICPB_DIST_SQ, GROUPS, SLABS, STAGES, LDW, TPC, QC, NACC, SORTED>;
	static bool attr_set[2][64][64] = {};
	bool& done = attr_set[dist_mode == ICPB_DIST_SQRT ? 1 : 0][(variant & 31) + (SORTED ? 32 : 0)][c->device & 63];
	if (!done) { ICPB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM)); done = true; }
	long long grid = c->sm_count;
	if (c->k1_grid_override > 0) grid = c->k1_grid_override;
	const long long max_ctas = (p.total_units + p.min_chunk - 1) / p.min_chunk;
	if (grid > max_ctas) grid = max_ctas;
	kern<<<(unsigned)grid, 64 + 256 * GROUPS, SMEM, c->stream>>>(p);      // the counters re-arm themselves (exit_counter)
	c->launches++;
	ICPB_CUDA(c, cudaGetLastError());
	return ICPB_OK;
}

int launch_match_filter_tc(Ctx* c, int dist_mode, float sentinel)
{
	int rc;
	// targets per MMA column: forced by ICPB_KT_VAR (experiments), else chosen by the exact-pass-rate policy (kf_policy_update):
	// it starts from the target size and is halved when exact passes cost about as much as the tests and the groups are what asks for them
	const bool forced = c->kt_variant >= 0;
	c->kt_tpc = tc_current_tpc(c);
	if (!c->kt_ready || c->kt_built_tpc != c->kt_tpc) { if ((rc = build_filter_tc_data(c)) != ICPB_OK) return rc; }
	if ((rc = ensure_source_order(c)) != ICPB_OK) return rc;
	KTParams p;
	p.px = c->px; p.py = c->py; p.pz = c->pz;
	p.tiles = c->kt_tiles; p.q4 = c->q4; p.seed_idx = c->kf_use_seed ? c->seed : nullptr; p.keys = c->keys;
	p.n = c->n; p.m = c->m; p.nt = c->kt_nt;
	p.thr0 = (dist_mode == ICPB_DIST_SQRT) ? sqrt_domain_threshold_tc(sentinel) : sentinel;
	p.cx = c->kf_center[0]; p.cy = c->kf_center[1]; p.cz = c->kf_center[2]; p.rq = c->kf_rq;
	p.done = &c->st->done;
	p.stats = c->kf_stats;
	p.fail = c->kt_fail;
	p.colstart = c->kt_colstart; p.hmax = c->kt_hmax;
	c->kf_dims_last = 4;                       // reported as "4" = the 3-D bound on the tensor cores
	c->kf_seeded = true;
	c->pairs_acc += (double)c->n * (double)c->m;
	if (!c->kt_work) {
		ICPB_CUDA(c, cudaMalloc((void**)&c->kt_work, 2 * sizeof(unsigned long long)));
		ICPB_CUDA(c, cudaMemsetAsync(c->kt_work, 0, 2 * sizeof(unsigned long long), c->stream));
	}
	p.work_counter = c->kt_work; p.exit_counter = c->kt_work + 1;
	// ICPB_KT_VAR selects the pipeline shape (experiments; results are identical):
	//   0: one 256-column MMA per (slab, tile), one accumulator per epilogue group, 8 slabs, 1 CTA/SM
	//   1: two 128-column MMAs per (slab, tile), two accumulators per group (MMA of the next unit overlaps the read of this one)
	//   2: two CTAs per SM, each with half the TMEM: 128-column MMAs, one accumulator per group, 4 slabs, 2-stage ring
	//   3: as 2 with 16-column TMEM loads (fewer registers); 4: as 0 with 16-column loads
	//   5 / 6: split form, one group of 8 warps on two accumulators (per-source state merged with atomicMin), 32- / 16-column loads
	//   7: split form, two groups of 8 warps (16 epilogue warps), 16-column loads
	//   8 / 9: as 7 / 6 in the PAIRED form: one column per two consecutive targets (tc_pack_group_kernel); 10: as 7 with QUADS
	p.slot_index = c->kt_sorted ? c->kt_slot_index : nullptr;
	p.src_perm = c->kt_src_sorted ? c->kt_sperm2 + c->kt_ssort_cap : nullptr;
	if (c->kt_sorted) {            // tiles in Morton order (build_filter_tc_data decided, or ICPB_KT_SORT=1): the kernels that carry original indices
		switch (c->kt_tpc) {
		case 16: return launch_tc_split<2, 8, 2, 16, 16, 8, 1, true>(c, dist_mode, p, 14);
		case 8:  return launch_tc_split<2, 8, 3, 16, 8, 8, 1, true>(c, dist_mode, p, 16);
		case 4:  return launch_tc_split<2, 8, 3, 16, 4, 8, 1, true>(c, dist_mode, p, 18);
		default: return launch_tc_split<2, 8, 3, 16, 2, 8, 1, true>(c, dist_mode, p, 20);
		}
	}
	if (!forced) {
		switch (c->kt_tpc) {
		case 16: return launch_tc_split<2, 8, 2, 16, 16, 8>(c, dist_mode, p, 14);
		case 8:  return launch_tc_split<2, 8, 3, 16, 8, 8>(c, dist_mode, p, 16);
		case 4:  return launch_tc_split<2, 8, 3, 16, 4, 8>(c, dist_mode, p, 18);
		case 2:  return launch_tc_split<2, 8, 3, 16, 2, 8>(c, dist_mode, p, 20);
		default: return launch_tc_split<2, 8, 3, 16, 1>(c, dist_mode, p, 7);
		}
	}
	switch (c->kt_variant) {
	case 1:  return launch_tc_variant<1, 2, 8, 3, 1, 32>(c, dist_mode, p, 1);
	case 2:  return launch_tc_variant<1, 1, 4, 2, 2, 32>(c, dist_mode, p, 2);
	case 3:  return launch_tc_variant<1, 1, 4, 2, 2, 16>(c, dist_mode, p, 3);
	case 4:  return launch_tc_variant<2, 1, 8, 3, 1, 16>(c, dist_mode, p, 4);
	case 5:  return launch_tc_split<1, 8, 3, 32, 1>(c, dist_mode, p, 5);
	case 6:  return launch_tc_split<1, 8, 3, 16, 1>(c, dist_mode, p, 6);
	case 7:  return launch_tc_split<2, 8, 3, 16, 1>(c, dist_mode, p, 7);
	case 8:  return launch_tc_split<2, 8, 3, 16, 2>(c, dist_mode, p, 8);
	case 9:  return launch_tc_split<1, 8, 3, 16, 2>(c, dist_mode, p, 9);
	case 10: return launch_tc_split<2, 8, 3, 16, 4>(c, dist_mode, p, 10);
	case 11: return launch_tc_split<2, 8, 3, 16, 8>(c, dist_mode, p, 11);
	case 12: return launch_tc_split<2, 8, 2, 16, 16>(c, dist_mode, p, 12);
	case 13: return launch_tc_split<2, 8, 2, 16, 16, 16>(c, dist_mode, p, 13);      // 13-16: finer exact-pass units (16 / 8 columns)
	case 14: return launch_tc_split<2, 8, 2, 16, 16, 8>(c, dist_mode, p, 14);
	case 15: return launch_tc_split<2, 8, 3, 16, 8, 16>(c, dist_mode, p, 15);
	case 16: return launch_tc_split<2, 8, 3, 16, 8, 8>(c, dist_mode, p, 16);
	case 17: return launch_tc_split<2, 8, 3, 16, 4, 16>(c, dist_mode, p, 17);
	case 18: return launch_tc_split<2, 8, 3, 16, 4, 8>(c, dist_mode, p, 18);
	case 19: return launch_tc_split<2, 8, 3, 16, 2, 16>(c, dist_mode, p, 19);
	case 20: return launch_tc_split<2, 8, 3, 16, 2, 8>(c, dist_mode, p, 20);
	case 21: return launch_tc_split<2, 8, 2, 16, 16, 8, 2>(c, dist_mode, p, 21);     // 21-24: two 128-column accumulators per group
	case 22: return launch_tc_split<2, 8, 3, 16, 8, 8, 2>(c, dist_mode, p, 22);
	case 23: return launch_tc_split<2, 8, 3, 16, 4, 8, 2>(c, dist_mode, p, 23);
	case 24: return launch_tc_split<2, 8, 3, 16, 2, 8, 2>(c, dist_mode, p, 24);
	default: return launch_tc_variant<2, 1, 8, 3, 1, 32>(c, dist_mode, p, 0);
	}
}

// after a synchronisation point: did any launch report a protocol time-out?
int filter_tc_check(Ctx* c)
{
	if (!c->kt_fail) return ICPB_OK;
	int h = 0;
	ICPB_CUDA(c, cudaMemcpyAsync(&h, c->kt_fail, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (h) { snprintf(c->err, sizeof c->err, "k1_filter_tc: an mbarrier wait timed out (tensor-core pipeline protocol error)"); return ICPB_ERR_CUDA; }
	return ICPB_OK;
}

} // namespace icpb
