// solve_device.cuh — K3/K8: the per-iteration small solves, one thread, FP64 registers. Shared by the
// streaming path (reduce_kernels.cu) and the batched persistent kernel (batched.cu).
#pragma once
#include "common.cuh"

namespace icpb {

// ------------------------------------------------------------------------------------------------
// K3: R = U V^T from W by one-sided Jacobi in FP64 registers, T = qbar - R pbar.
// ------------------------------------------------------------------------------------------------
__device__ inline void polar_rotation(const double W[9] /*col-major*/, double R[9], bool fix_reflection = false)
{
	double a[3][3], v[3][3];
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) { a[i][j] = W[i + 3 * j]; v[i][j] = (i == j) ? 1.0 : 0.0; }
	// One thread, so the latency of every FP64 square root and division is paid in full: a rotation costs one sqrt, one
	// division and one rsqrt (t from the half-angle form with the division by gamma folded in, the convergence test on
	// gamma^2 against alpha beta instead of their quotient's root).
	for (int sweep = 0; sweep < 60; sweep++) {
		bool moved = false;
#pragma unroll
		for (int pq = 0; pq < 3; pq++) {
			const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
			double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
			for (int i = 0; i < 3; i++) { alpha += a[i][p] * a[i][p]; beta += a[i][q] * a[i][q]; gamma += a[i][p] * a[i][q]; }
			const double g2 = gamma * gamma, ab = alpha * beta;
			if (gamma == 0.0 || g2 < 1e-34 * ab) continue;           // |gamma| / sqrt(alpha beta) < 1e-17: columns orthogonal to working precision
			if (g2 >= 1e-32 * ab) moved = true;                      // ... >= 1e-16: another sweep is needed
			// t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (beta - alpha) / (2 gamma), numerator and denominator times 2 |gamma|
			const double d = beta - alpha;
			const double sgn = (d == 0.0 || ((d > 0.0) == (gamma > 0.0))) ? 1.0 : -1.0;
			const double t = sgn * 2.0 * fabs(gamma) / (fabs(d) + sqrt(d * d + 4.0 * g2));
			const double c = rsqrt(1.0 + t * t), s = c * t;
#pragma unroll
			for (int i = 0; i < 3; i++) {
				double x = a[i][p], y = a[i][q]; a[i][p] = c * x - s * y; a[i][q] = s * x + c * y;
				x = v[i][p]; y = v[i][q]; v[i][p] = c * x - s * y; v[i][q] = s * x + c * y;
			}
		}
		if (!moved) break;
	}
	double u[3][3], sv[3];
#pragma unroll
	for (int j = 0; j < 3; j++) {
		const double n2 = a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j];
		sv[j] = sqrt(n2);
		const double inv = (n2 > 0) ? 1.0 / sv[j] : 0.0;
#pragma unroll
		for (int i = 0; i < 3; i++) u[i][j] = a[i][j] * inv;
	}
	int small = 0;
	if (sv[1] < sv[small]) small = 1;
	if (sv[2] < sv[small]) small = 2;
	const double big = fmax(sv[0], fmax(sv[1], sv[2]));
	if (sv[small] <= 1e-14 * big) {   // rank-deficient W: complete the basis (never on the reference's inputs)
		const int j1 = (small + 1) % 3, j2 = (small + 2) % 3;
		u[0][small] = u[1][j1] * u[2][j2] - u[2][j1] * u[1][j2];
		u[1][small] = u[2][j1] * u[0][j2] - u[0][j1] * u[2][j2];
		u[2][small] = u[0][j1] * u[1][j2] - u[1][j1] * u[0][j2];
	}
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) R[i + 3 * j] = u[i][0] * v[j][0] + u[i][1] * v[j][1] + u[i][2] * v[j][2];
	if (fix_reflection) {
		const double det = R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) + R[6] * (R[1] * R[5] - R[4] * R[2]);
		if (det < 0.0) {      // Kabsch: flip the pair (u_k, v_k) of the smallest singular value: R -= 2 u_k v_k^T
#pragma unroll
			for (int i = 0; i < 3; i++)
#pragma unroll
				for (int j = 0; j < 3; j++) R[i + 3 * j] -= 2.0 * u[i][small] * v[j][small];
		}
	}
}

__device__ inline void compose_total(IterState* st)
{
	double Rn[9], tn[3];
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++)
			Rn[r + 3 * c] = (double)st->R[r] * st->Rtot[3 * c] + (double)st->R[r + 3] * st->Rtot[1 + 3 * c] + (double)st->R[r + 6] * st->Rtot[2 + 3 * c];
	for (int r = 0; r < 3; r++)
		tn[r] = (double)st->R[r] * st->ttot[0] + (double)st->R[r + 3] * st->ttot[1] + (double)st->R[r + 6] * st->ttot[2] + (double)st->T[r];
	for (int k = 0; k < 9; k++) st->Rtot[k] = Rn[k];
	for (int k = 0; k < 3; k++) st->ttot[k] = tn[k];
}

// point-to-point: moments[0..15] -> R, T                                   (one thread)
__device__ inline void solve_p2p(IterState* st)
{
	const double* mom = st->moments;
	const double N = mom[15];
	double pb[3], qb[3], W[9], R[9];
	const double invN = 1.0 / N;
	for (int c = 0; c < 3; c++) { pb[c] = mom[c] * invN; qb[c] = mom[3 + c] * invN; }
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++) W[r + 3 * c] = mom[6 + r + 3 * c] - N * qb[r] * pb[c];
	polar_rotation(W, R, (st->flags & ICPB_FLAG_FIX_REFLECTION) != 0);
	for (int k = 0; k < 9; k++) st->R[k] = (float)R[k];
	for (int r = 0; r < 3; r++) st->T[r] = (float)(qb[r] - (R[r] * pb[0] + R[r + 3] * pb[1] + R[r + 6] * pb[2]));
	compose_total(st);
}

// point-to-plane: moments[0..20] = upper triangle of C (row by row), [21..26] = b, [27] = N.
// Cholesky C = U^T U in FP64 on the float-rounded sums, two triangular solves, then the reference's
// host code: float cos/sin of the solution, R = Rz(g) Ry(b) Rx(a) in float (src/ICP_point_to_plane.cu:585-593).
__device__ inline void solve_p2plane(IterState* st)
{
	double C[6][6], U[6][6], b[6], y[6], x[6];
	int k = 0;
	for (int r = 0; r < 6; r++)
		for (int c = r; c < 6; c++) C[r][c] = (double)(float)st->moments[k++];
	for (int r = 0; r < 6; r++) b[r] = (double)(float)st->moments[21 + r];
	for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) U[r][c] = 0.0;
	for (int j = 0; j < 6; j++) {
		double s = C[j][j];
		for (int q = 0; q < j; q++) s -= U[q][j] * U[q][j];
		if (!(s > 0.0)) { st->numeric_error = j + 1; st->done = 1; return; }
		U[j][j] = sqrt(s);
		for (int i = j + 1; i < 6; i++) {
			double t = C[j][i];
			for (int q = 0; q < j; q++) t -= U[q][j] * U[q][i];
			U[j][i] = t / U[j][j];
		}
	}
	for (int i = 0; i < 6; i++) { double t = b[i]; for (int q = 0; q < i; q++) t -= U[q][i] * y[q]; y[i] = t / U[i][i]; }
	for (int i = 5; i >= 0; i--) { double t = y[i]; for (int q = i + 1; q < 6; q++) t -= U[i][q] * x[q]; x[i] = t / U[i][i]; }
	float hb[6];
	for (int i = 0; i < 6; i++) hb[i] = (float)x[i];
	const float cx = (float)cos((double)hb[0]), cy = (float)cos((double)hb[1]), cz = (float)cos((double)hb[2]);
	const float sx = (float)sin((double)hb[0]), sy = (float)sin((double)hb[1]), sz = (float)sin((double)hb[2]);
	float* R = st->R;
	R[0] = __fmul_rn(cy, cz);
	R[3] = __fsub_rn(__fmul_rn(__fmul_rn(cz, sx), sy), __fmul_rn(cx, sz));
	R[6] = __fadd_rn(__fmul_rn(__fmul_rn(cx, cz), sy), __fmul_rn(sx, sz));
	R[1] = __fmul_rn(cy, sz);
	R[4] = __fadd_rn(__fmul_rn(cx, cz), __fmul_rn(__fmul_rn(sx, sy), sz));
	R[7] = __fsub_rn(__fmul_rn(__fmul_rn(cx, sy), sz), __fmul_rn(cz, sx));
	R[2] = -sy;
	R[5] = __fmul_rn(cy, sx);
	R[8] = __fmul_rn(cx, cy);
	st->T[0] = hb[3]; st->T[1] = hb[4]; st->T[2] = hb[5];
	compose_total(st);
}


} // namespace icpb
