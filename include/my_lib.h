/* my_lib.h — drop-in for the reference's host helper library (src/my_lib.h:4-12).
 *
 * Same nine entry points, same signatures, same C++ linkage, same behaviour:
 * three naive column-major GEMMs (C(m x n) = A(m x k) * B(k x n), accumulated left to right in the
 * element type — src/my_lib.cpp:6-35,80-93) and six printers with the reference's exact formats
 * (src/my_lib.cpp:38-77,96-134). Like the reference header, this one pulls in its implementation
 * so that `#include "my_lib.h"` alone is enough (src/my_lib.h:14); define MY_LIB_NO_IMPL to get
 * declarations only and link libicp_lib.a instead.
 */
#ifndef _MYLIB
#define _MYLIB

void fmatrixMul(float *A, float *B, float *C, int m, int n, int k);
void dmatrixMul(double *A, double *B, double *C, int m, int n, int k);
void print_cloud(double* cloud, int num_points, int points2show);
void print_darray(double* array, int points2show);
void print_iarray(int* array, int points2show);
void SmatrixMul(float* A, float* B, float* C, int m, int n, int k);
void printScloud(float* cloud, int num_points, int points2show);
void printSarray(float* array, int points2show);
void printIarray(int* array, int points2show);

#ifndef MY_LIB_NO_IMPL
#include "my_lib.cpp"
#endif
#endif
