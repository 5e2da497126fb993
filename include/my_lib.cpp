/* my_lib.cpp — implementation of include/my_lib.h (see there). Written from the behaviour of the
 * reference's src/my_lib.cpp, not from its text: one generic column-major product and one generic
 * row printer, instantiated for the nine public names. Build with -ffp-contract=off: the reference
 * accumulates `temp += a*b` without fused multiply-adds on the host. */
#ifndef MY_LIB_IMPL_INCLUDED
#define MY_LIB_IMPL_INCLUDED
#include <stdio.h>

namespace my_lib_detail {

/* C[row + col*m] = sum_q A[row + q*m] * B[q + col*k], q ascending, accumulator of type T starting at 0 */
template <typename T>
static inline void colmajor_product(const T* A, const T* B, T* C, int m, int n, int k)
{
	for (int col = 0; col < n; ++col) {
		const T* bcol = B + (long)col * k;
		T* ccol = C + (long)col * m;
		for (int row = 0; row < m; ++row) {
			T acc = T(0);
			for (int q = 0; q < k; ++q) acc = acc + A[row + (long)q * m] * bcol[q];
			ccol[row] = acc;
		}
	}
}

/* "x\ty\tz\n" header, then one "%.4f\t%.4f\t%.4f\t\n" line per point (AoS xyz) */
template <typename T>
static inline void show_cloud(const T* cloud, int num_points, int points2show)
{
	printf("x\ty\tz\n");
	if (points2show > num_points) { printf("The cloud can't be printed\n\n"); return; }
	for (int p = 0; p < points2show; ++p) {
		const T* v = cloud + 3L * p;
		printf("%.4f\t%.4f\t%.4f\t\n", (double)v[0], (double)v[1], (double)v[2]);
	}
}

template <typename T>
static inline void show_row(const T* a, int count, const char* fmt)
{
	for (int i = 0; i < count; ++i) printf(fmt, a[i]);
	printf("\n");
}

} /* namespace my_lib_detail */

void fmatrixMul(float* A, float* B, float* C, int m, int n, int k) { my_lib_detail::colmajor_product<float>(A, B, C, m, n, k); }
void dmatrixMul(double* A, double* B, double* C, int m, int n, int k) { my_lib_detail::colmajor_product<double>(A, B, C, m, n, k); }
void SmatrixMul(float* A, float* B, float* C, int m, int n, int k) { my_lib_detail::colmajor_product<float>(A, B, C, m, n, k); }

void print_cloud(double* cloud, int num_points, int points2show) { my_lib_detail::show_cloud<double>(cloud, num_points, points2show); }
void printScloud(float* cloud, int num_points, int points2show) { my_lib_detail::show_cloud<float>(cloud, num_points, points2show); }

void print_darray(double* array, int points2show) { my_lib_detail::show_row<double>(array, points2show, "%.3f "); }
void printSarray(float* array, int points2show) { my_lib_detail::show_row<float>(array, points2show, "%.4f "); }
void print_iarray(int* array, int points2show) { my_lib_detail::show_row<int>(array, points2show, "%d "); }
void printIarray(int* array, int points2show) { my_lib_detail::show_row<int>(array, points2show, "%d "); }

#endif
