/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Harness around the UNMODIFIED reference CUDA sources. Nothing is copied: the three reference
 * programs are pulled in by #include from /root/reference/src (REF_SRC_DIR), each inside its own
 * namespace with `main` renamed, so that
 *   (a) a whole reference program can be run (`run std|p2p|p2l`) to capture its stdout, and
 *   (b) a single reference kernel can be launched on caller-supplied inputs, giving golden
 *       vectors from the reference's own device code:
 *         match <std|p2p|p2l> n m P.bin Q.bin idx.bin     Matching   (ICP_standard.cu:21-39, ICP_point_to_point.cu:31-57, ICP_point_to_plane.cu:163-181)
 *         ryt n R.bin T.bin P.bin out.bin                 RyT        (ICP_point_to_point.cu:81-88)
 *         knn m k1 Q.bin nbr.bin                          knn        (ICP_point_to_plane.cu:48-70)
 *         normalsA m k Q.bin nbr.bin A.bin                Normals    (ICP_point_to_plane.cu:72-102; the 9-float covariance slots)
 *         cxb n m P.bin Q.bin idx.bin normals.bin C.bin b.bin   Q_index + Cxb + cublasSgemv x2 (ICP_point_to_plane.cu:183-234, 546-556)
 * The reference kernels have no bounds checks and hard-wire n = grid*block, so n (and m for knn)
 * must be a multiple of 128; ICP_standard's Matching only handles n <= 1024 (its NUM_POINTS).
 * Built by oracle/Makefile into oracle/_ref/ref_kernels; runs only on a GPU box.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <time.h>
#define _USE_MATH_DEFINES
#include <math.h>
#include <cuda_runtime.h>
#include <device_launch_parameters.h>
#include <cublas_v2.h>
#include <curand.h>
#include <cusolverDn.h>
#include "mkl.h"
#include "mkl_lapacke.h"
#include "my_lib.h"      /* the reference's own helper library (src/my_lib.h includes my_lib.cpp) */

#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#define REF_FILE(name) REF_STR(REF_SRC_DIR/name)

namespace ref_std {
#define main ref_main
#include REF_FILE(ICP_standard.cu)
#undef main
}
#undef WIDTH
#undef NUM_POINTS
#undef XY_min
#undef XY_max
#undef MAX_ITER
namespace ref_p2p {
#define main ref_main
#include REF_FILE(ICP_point_to_point.cu)
#undef main
}
#undef WIDTH
#undef NUM_POINTS
#undef XY_min
#undef XY_max
#undef MAX_ITER
namespace ref_p2l {
#define main ref_main
#include REF_FILE(ICP_point_to_plane.cu)
#undef main
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static void* slurp(const char* path, size_t bytes)
{
	FILE* f = fopen(path, "rb"); if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
	void* p = malloc(bytes); if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read %s\n", path); exit(2); }
	fclose(f); return p;
}
static void spit(const char* path, const void* p, size_t bytes)
{
	FILE* f = fopen(path, "wb"); if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
	fwrite(p, 1, bytes, f); fclose(f);
}
template <typename T> static T* to_dev(const void* h, size_t count)
{
	T* d; CK(cudaMalloc(&d, count * sizeof(T))); if (h) CK(cudaMemcpy(d, h, count * sizeof(T), cudaMemcpyHostToDevice)); else CK(cudaMemset(d, 0, count * sizeof(T))); return d;
}
static int pick_block(int n) { int b = 1024; while (b > 128 && n % b) b >>= 1; if (n % b) { fprintf(stderr, "n=%d must be a multiple of 128\n", n); exit(2); } return b; }

int main(int argc, char** argv)
{
	if (argc < 2) { fprintf(stderr, "usage: ref_kernels run|match|ryt|knn|normalsA|cxb ...\n"); return 2; }
	if (!strcmp(argv[1], "run")) {
		if (!strcmp(argv[2], "std")) return ref_std::ref_main();
		if (!strcmp(argv[2], "p2p")) return ref_p2p::ref_main();
		if (!strcmp(argv[2], "p2l")) return ref_p2l::ref_main();
		return 2;
	}
	if (!strcmp(argv[1], "match")) {
		const char* which = argv[2]; int n = atoi(argv[3]), m = atoi(argv[4]);
		float* hP = (float*)slurp(argv[5], 12ull * n); float* hQ = (float*)slurp(argv[6], 12ull * m);
		float* dP = to_dev<float>(hP, 3ull * n); float* dQ = to_dev<float>(hQ, 3ull * m);
		int* dI = to_dev<int>(NULL, n);
		int block = pick_block(n);
		if (!strcmp(which, "std")) { if (n > 1024) { fprintf(stderr, "std Matching handles n<=1024\n"); return 2; } ref_std::Matching<<<n / block, block>>>(dP, dQ, m, dI); }
		else if (!strcmp(which, "p2p")) ref_p2p::Matching<<<n / block, block>>>(dP, dQ, m, dI);
		else ref_p2l::Matching<<<n / block, block>>>(dP, dQ, m, dI);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		int* hI = (int*)malloc(4ull * n); CK(cudaMemcpy(hI, dI, 4ull * n, cudaMemcpyDeviceToHost));
		spit(argv[7], hI, 4ull * n); return 0;
	}
	if (!strcmp(argv[1], "ryt")) {
		int n = atoi(argv[2]);
		float* dR = to_dev<float>(slurp(argv[3], 36), 9); float* dT = to_dev<float>(slurp(argv[4], 12), 3);
		float* dP = to_dev<float>(slurp(argv[5], 12ull * n), 3ull * n); float* dO = to_dev<float>(NULL, 3ull * n);
		int block = pick_block(n);
		ref_p2p::RyT<<<n / block, block>>>(dR, dT, dP, dO);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		float* hO = (float*)malloc(12ull * n); CK(cudaMemcpy(hO, dO, 12ull * n, cudaMemcpyDeviceToHost));
		spit(argv[6], hO, 12ull * n); return 0;
	}
	if (!strcmp(argv[1], "knn") || !strcmp(argv[1], "normalsA")) {
		int isA = !strcmp(argv[1], "normalsA");
		int m = atoi(argv[2]), k1 = isA ? atoi(argv[3]) + 1 : atoi(argv[3]);
		float* dQ = to_dev<float>(slurp(argv[4], 12ull * m), 3ull * m);
		int* dN = to_dev<int>(NULL, (size_t)m * k1);
		float* dD = to_dev<float>(NULL, (size_t)m * m);
		int block = pick_block(m);
		ref_p2l::knn<<<m / block, block>>>(dQ, m, dQ, m, dN, k1, dD);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		int* hN = (int*)malloc(4ull * m * k1); CK(cudaMemcpy(hN, dN, 4ull * m * k1, cudaMemcpyDeviceToHost));
		spit(argv[5], hN, 4ull * m * k1);
		if (isA) {
			float* dBar = to_dev<float>(NULL, 3ull * m); float* dA = to_dev<float>(NULL, 9ull * m); float* dNrm = to_dev<float>(NULL, 3ull * m);
			ref_p2l::Normals<<<m / block, block>>>(dQ, dN, m, m, k1 - 1, dBar, dA, dNrm);
			CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
			float* hA = (float*)malloc(36ull * m); CK(cudaMemcpy(hA, dA, 36ull * m, cudaMemcpyDeviceToHost));
			spit(argv[6], hA, 36ull * m);
		}
		return 0;
	}
	if (!strcmp(argv[1], "cxb")) {
		int n = atoi(argv[2]), m = atoi(argv[3]);
		float* dP = to_dev<float>(slurp(argv[4], 12ull * n), 3ull * n); float* dQ = to_dev<float>(slurp(argv[5], 12ull * m), 3ull * m);
		int* dI = to_dev<int>(slurp(argv[6], 4ull * n), n); float* dNrm = to_dev<float>(slurp(argv[7], 12ull * m), 3ull * m);
		float* dQi = to_dev<float>(NULL, 3ull * n); float* dcn = to_dev<float>(NULL, 6ull * n);
		float* dCt = to_dev<float>(NULL, 36ull * n); float* dbt = to_dev<float>(NULL, 6ull * n);
		float* dC = to_dev<float>(NULL, 36); float* db = to_dev<float>(NULL, 6);
		float* hU = (float*)malloc(4ull * n); for (int i = 0; i < n; i++) hU[i] = 1.0f; float* dU = to_dev<float>(hU, n);
		int block = pick_block(n);
		ref_p2l::Q_index<<<n / block, block>>>(dQ, dI, dQi);
		ref_p2l::Cxb<<<n / block, block>>>(dP, dQi, dI, dNrm, dcn, dCt, dbt);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		cublasHandle_t h; cublasCreate(&h); float alpha = 1, beta = 0;
		cublasSgemv(h, CUBLAS_OP_N, 36, n, &alpha, dCt, 36, dU, 1, &beta, dC, 1);
		cublasSgemv(h, CUBLAS_OP_N, 6, n, &alpha, dbt, 6, dU, 1, &beta, db, 1);
		CK(cudaDeviceSynchronize());
		float hC[36], hb[6]; CK(cudaMemcpy(hC, dC, 144, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb, db, 24, cudaMemcpyDeviceToHost));
		spit(argv[8], hC, 144); spit(argv[9], hb, 24); return 0;
	}
	fprintf(stderr, "unknown command %s\n", argv[1]);
	return 2;
}
