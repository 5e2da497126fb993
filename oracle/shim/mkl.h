/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Minimal stand-in for the Intel MKL entry points the reference calls, so that the
 * reference sources can be compiled *unmodified, from where they lie* into oracle/_ref/
 * (Intel MKL is not installed in this image; the reference pins no MKL version — its build
 * files never mention it, only `#include "mkl.h"`: /root/reference/src/ICP_point_to_plane.cu:15-16,
 * and /root/reference/src/ICP_CPU.c uses the symbols without including any header).
 *
 * Call sites covered:
 *   ICP_CPU.c:132-133,141,239,246,247,251  cblas_dgemm (row-major)
 *   ICP_CPU.c:203-205,225,254              cblas_dcopy (incx = 0 broadcast at :225)
 *   ICP_CPU.c:227,228,230,231,248,253,264  vdSub / vdSqr / vdAdd (element-wise, one rounding each)
 *   ICP_CPU.c:232                          cblas_idamin (FIRST index of the minimum |x|)
 *   ICP_CPU.c:240                          LAPACKE_dgesvd (3x3, jobu = jobvt = 'A')
 *   ICP_CPU.c:266                          cblas_dnrm2
 *   ICP_CPU.c:215,270,272                  dsecnd
 *   ICP_point_to_plane.cu:435-436          LAPACKE_ssyev (3x3, 'V','U') + cblas_isamin
 *
 * Semantics that decide results: vd* round once per element with no FMA
 * (build with -ffp-contract=off); idamin/isamin return the first minimum; the aliased
 * cblas_dgemm(r, rz -> r) at ICP_CPU.c:133 computes the full product before storing
 * (MKL's behaviour there is formally undefined: PARITY UNPINNED for that one line; the
 * mathematically intended r = rx*ry*rz is what this shim yields).
 * dgesvd / ssyev are cyclic Jacobi iterations in double precision; singular/eigen-vectors are
 * unique only up to sign, which is irrelevant downstream (R = U*Vt and the point-to-plane
 * system are invariant to those signs).
 */
#ifndef ORACLE_MKL_SHIM_H
#define ORACLE_MKL_SHIM_H
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int MKL_INT;
typedef int lapack_int;
enum { CblasRowMajor = 101, CblasColMajor = 102 };
enum { CblasNoTrans = 111, CblasTrans = 112 };
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

static inline double dsecnd(void)
{
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline void cblas_dcopy(int n, const double* x, int incx, double* y, int incy)
{
	for (int i = 0; i < n; i++) y[(size_t)i * incy] = x[(size_t)i * incx];
}
static inline void cblas_scopy(int n, const float* x, int incx, float* y, int incy)
{
	for (int i = 0; i < n; i++) y[(size_t)i * incy] = x[(size_t)i * incx];
}
static inline void vdSub(int n, const double* a, const double* b, double* y) { for (int i = 0; i < n; i++) y[i] = a[i] - b[i]; }
static inline void vdAdd(int n, const double* a, const double* b, double* y) { for (int i = 0; i < n; i++) y[i] = a[i] + b[i]; }
static inline void vdSqr(int n, const double* a, double* y) { for (int i = 0; i < n; i++) y[i] = a[i] * a[i]; }

static inline size_t cblas_idamin(int n, const double* x, int incx)
{
	size_t best = 0; double bv = fabs(x[0]);
	for (int i = 1; i < n; i++) { double v = fabs(x[(size_t)i * incx]); if (v < bv) { bv = v; best = (size_t)i; } }
	return best;
}
static inline size_t cblas_isamin(int n, const float* x, int incx)
{
	size_t best = 0; float bv = fabsf(x[0]);
	for (int i = 1; i < n; i++) { float v = fabsf(x[(size_t)i * incx]); if (v < bv) { bv = v; best = (size_t)i; } }
	return best;
}
static inline double cblas_dnrm2(int n, const double* x, int incx)
{
	/* scaled sum of squares, as the reference BLAS does, to stay overflow-safe */
	double scale = 0.0, ssq = 1.0;
	for (int i = 0; i < n; i++) {
		double v = fabs(x[(size_t)i * incx]);
		if (v != 0.0) {
			if (scale < v) { ssq = 1.0 + ssq * (scale / v) * (scale / v); scale = v; }
			else ssq += (v / scale) * (v / scale);
		}
	}
	return scale * sqrt(ssq);
}

/* C(m x n) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C ; row- or column-major. The product is
 * formed in a temporary first so that an output aliasing an input (ICP_CPU.c:133) is well defined. */
static inline void cblas_dgemm(int layout, int transa, int transb, int m, int n, int k, double alpha,
	const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc)
{
	double* tmp = (double*)malloc(sizeof(double) * (size_t)m * (size_t)n);
	for (int i = 0; i < m; i++)
		for (int j = 0; j < n; j++) {
			double acc = 0.0;
			for (int q = 0; q < k; q++) {
				double a, b;
				if (layout == CblasRowMajor) {
					a = (transa == CblasNoTrans) ? A[(size_t)i * lda + q] : A[(size_t)q * lda + i];
					b = (transb == CblasNoTrans) ? B[(size_t)q * ldb + j] : B[(size_t)j * ldb + q];
				} else {
					a = (transa == CblasNoTrans) ? A[(size_t)q * lda + i] : A[(size_t)i * lda + q];
					b = (transb == CblasNoTrans) ? B[(size_t)j * ldb + q] : B[(size_t)q * ldb + j];
				}
				acc += a * b;
			}
			tmp[(size_t)i * n + j] = acc;
		}
	for (int i = 0; i < m; i++)
		for (int j = 0; j < n; j++) {
			double* c = (layout == CblasRowMajor) ? &C[(size_t)i * ldc + j] : &C[(size_t)j * ldc + i];
			*c = (beta == 0.0) ? alpha * tmp[(size_t)i * n + j] : alpha * tmp[(size_t)i * n + j] + beta * (*c);
		}
	free(tmp);
}

/* ---- 3x3 kernels: one-sided (Hestenes) Jacobi SVD and cyclic Jacobi symmetric eigen-solver ---- */
static inline void shim_svd3(const double Ain[9] /*row-major*/, double U[9], double S[3], double Vt[9])
{
	double a[3][3], v[3][3];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] = Ain[3 * i + j]; v[i][j] = (i == j); }
	for (int sweep = 0; sweep < 60; sweep++) {
		double off = 0.0;
		for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
			double alpha = 0, beta = 0, gamma = 0;
			for (int i = 0; i < 3; i++) { alpha += a[i][p] * a[i][p]; beta += a[i][q] * a[i][q]; gamma += a[i][p] * a[i][q]; }
			if (gamma == 0.0) continue;
			double lim = fabs(gamma) / sqrt(alpha * beta);
			if (lim > off) off = lim;
			if (lim < 1e-17) continue;
			double zeta = (beta - alpha) / (2.0 * gamma);
			double t = ((zeta >= 0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
			double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
			for (int i = 0; i < 3; i++) {
				double x = a[i][p], y = a[i][q]; a[i][p] = c * x - s * y; a[i][q] = s * x + c * y;
				x = v[i][p]; y = v[i][q]; v[i][p] = c * x - s * y; v[i][q] = s * x + c * y;
			}
		}
		if (off < 1e-16) break;
	}
	double sv[3]; int ord[3] = { 0, 1, 2 };
	for (int j = 0; j < 3; j++) sv[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
	for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) if (sv[ord[j]] > sv[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
	double u[3][3];
	for (int jj = 0; jj < 3; jj++) {
		int j = ord[jj]; S[jj] = sv[j];
		for (int i = 0; i < 3; i++) { u[i][jj] = (sv[j] > 0) ? a[i][j] / sv[j] : 0.0; Vt[3 * jj + i] = v[i][j]; }
	}
	/* rank-deficient input: complete U to an orthonormal basis (never hit on the reference's inputs) */
	if (S[2] <= 1e-300 * (S[0] > 0 ? 1 : 0) || S[2] == 0.0) {
		if (S[1] == 0.0) {
			if (S[0] == 0.0) { u[0][0] = 1; u[1][0] = 0; u[2][0] = 0; }
			int k = (fabs(u[0][0]) < 0.9) ? 0 : 1; double e[3] = { 0, 0, 0 }; e[k] = 1;
			double d = e[0] * u[0][0] + e[1] * u[1][0] + e[2] * u[2][0], nn = 0;
			for (int i = 0; i < 3; i++) { u[i][1] = e[i] - d * u[i][0]; nn += u[i][1] * u[i][1]; }
			nn = sqrt(nn); for (int i = 0; i < 3; i++) u[i][1] /= nn;
		}
		u[0][2] = u[1][0] * u[2][1] - u[2][0] * u[1][1];
		u[1][2] = u[2][0] * u[0][1] - u[0][0] * u[2][1];
		u[2][2] = u[0][0] * u[1][1] - u[1][0] * u[0][1];
	}
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) U[3 * i + j] = u[i][j];
}

static inline lapack_int LAPACKE_dgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, double* a, lapack_int lda,
	double* s, double* u, lapack_int ldu, double* vt, lapack_int ldvt, double* superb)
{
	(void)jobu; (void)jobvt; (void)superb;
	if (m != 3 || n != 3) return -4;
	double A[9], U[9], S[3], Vt[9];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[3 * i + j] = (layout == LAPACK_ROW_MAJOR) ? a[i * lda + j] : a[j * lda + i];
	shim_svd3(A, U, S, Vt);
	for (int i = 0; i < 3; i++) {
		s[i] = S[i];
		for (int j = 0; j < 3; j++) {
			if (layout == LAPACK_ROW_MAJOR) { u[i * ldu + j] = U[3 * i + j]; vt[i * ldvt + j] = Vt[3 * i + j]; }
			else { u[j * ldu + i] = U[3 * i + j]; vt[j * ldvt + i] = Vt[3 * i + j]; }
		}
	}
	return 0;
}

/* symmetric 3x3 eigen-decomposition, eigenvalues ascending in w, eigenvectors in the COLUMNS of the output */
static inline void shim_eig3(const double Ain[9] /*full symmetric, row-major*/, double w[3], double V[9])
{
	double a[3][3], v[3][3];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] = Ain[3 * i + j]; v[i][j] = (i == j); }
	for (int sweep = 0; sweep < 60; sweep++) {
		double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
		double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
		if (off <= 1e-18 * diag || off == 0.0) break;
		for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
			if (a[p][q] == 0.0) continue;
			double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
			double t = ((theta >= 0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
			double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
			for (int k = 0; k < 3; k++) { double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
			for (int k = 0; k < 3; k++) { double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
			for (int k = 0; k < 3; k++) { double x = v[k][p], y = v[k][q]; v[k][p] = c * x - s * y; v[k][q] = s * x + c * y; }
		}
	}
	int ord[3] = { 0, 1, 2 };
	for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) if (a[ord[j]][ord[j]] < a[ord[i]][ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
	for (int jj = 0; jj < 3; jj++) { w[jj] = a[ord[jj]][ord[jj]]; for (int i = 0; i < 3; i++) V[3 * i + jj] = v[i][ord[jj]]; }
}

static inline lapack_int LAPACKE_ssyev(int layout, char jobz, char uplo, lapack_int n, float* a, lapack_int lda, float* w)
{
	(void)jobz;
	if (n != 3) return -4;
	double A[9], W[3], V[9];
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		int r = i, c = j;
		/* only the `uplo` triangle is referenced */
		if ((uplo == 'U' || uplo == 'u') ? (r > c) : (r < c)) { int t = r; r = c; c = t; }
		A[3 * i + j] = (layout == LAPACK_ROW_MAJOR) ? a[r * lda + c] : a[c * lda + r];
	}
	shim_eig3(A, W, V);
	for (int i = 0; i < 3; i++) {
		w[i] = (float)W[i];
		for (int j = 0; j < 3; j++) {
			if (layout == LAPACK_ROW_MAJOR) a[i * lda + j] = (float)V[3 * i + j]; else a[j * lda + i] = (float)V[3 * i + j];
		}
	}
	return 0;
}

/* ---- MSVC-only CRT calls used by the reference's dataset programs (src/CUDA/GPU_point_to_point_bunny.cu:468,480) ---- */
#ifndef _MSC_VER
#include <stdio.h>
static inline int oracle_fopen_s(FILE** pf, const char* name, const char* mode) { *pf = fopen(name, mode); return *pf ? 0 : 1; }
#define fopen_s(pf, name, mode) oracle_fopen_s((pf), (name), (mode))
#define strtok_s(str, delim, ctx) strtok_r((str), (delim), (ctx))
#endif
#endif
