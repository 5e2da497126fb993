/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Harness around the UNMODIFIED reference dataset programs. Nothing is copied: the four
 * programs are pulled in by #include from /root/reference/src/CUDA (REF_SRC_DIR), each inside its own namespace with
 * `main` renamed, so that
 *   (a) a whole reference program can be run to capture its stdout (it opens its dataset in the working directory):
 *         run bunny_p2p|bunny_p2l|lidar_p2p|lidar_p2l
 *   (b) single reference functions / kernels can be run on caller-supplied inputs:
 *         readdata file nfloats out.bin            readData            GPU_point_to_point_bunny.cu:463-497   (host only)
 *         lidar P.bin Q.bin                        Read_data           GPU_point_to_point_real.cu:432-620    (parser, Conversion, RyT;
 *                                                                      16384 points each, before the mm -> m scaling)
 *         knn_sq m k1 Q.bin nbr.bin                knn on squared distances   GPU_point_to_plane_bunny.cu:47-82
 *         match n m P.bin Q.bin idx.bin            Matching, sentinel 1e6     GPU_point_to_point_real.cu:38-79
 * Intel MKL (dsecnd, LAPACKE_ssyev, cblas_isamin) and the two MSVC CRT calls come from oracle/shim/mkl.h.
 * Built by oracle/Makefile into oracle/_ref/ref_datasets; (a), lidar, knn_sq and match need a GPU.
 */
#define _USE_MATH_DEFINES
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <cuda_runtime.h>
#include <device_launch_parameters.h>
#include <cublas_v2.h>
#include <curand.h>
#include <cusolverDn.h>
#include "mkl.h"
#include "mkl_lapacke.h"

#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#define REF_FILE(name) REF_STR(REF_SRC_DIR/name)

namespace bunny_p2p {
#define main ref_main
#include REF_FILE(GPU_point_to_point_bunny.cu)
#undef main
}
#undef MAX_ITER
namespace bunny_p2l {
#define main ref_main
#include REF_FILE(GPU_point_to_plane_bunny.cu)
#undef main
}
#undef MAX_ITER
namespace lidar_p2p {
#define main ref_main
#include REF_FILE(GPU_point_to_point_real.cu)
#undef main
}
#undef MAX_ITER
namespace lidar_p2l {
#define main ref_main
#include REF_FILE(GPU_point_to_plane_real.cu)
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

int main(int argc, char** argv)
{
	if (argc < 3) { fprintf(stderr, "usage: ref_datasets run|readdata|lidar|knn_sq|match ...\n"); return 2; }
	if (!strcmp(argv[1], "run")) {
		if (!strcmp(argv[2], "bunny_p2p")) return bunny_p2p::ref_main();
		if (!strcmp(argv[2], "bunny_p2l")) return bunny_p2l::ref_main();
		if (!strcmp(argv[2], "lidar_p2p")) return lidar_p2p::ref_main();
		if (!strcmp(argv[2], "lidar_p2l")) return lidar_p2l::ref_main();
		return 2;
	}
	if (!strcmp(argv[1], "readdata")) {
		int nf = atoi(argv[3]);
		float* out = (float*)calloc((size_t)nf + 16, sizeof(float));
		if (bunny_p2p::readData(argv[2], out) != 0) return 1;
		spit(argv[4], out, 4ull * nf); return 0;
	}
	if (!strcmp(argv[1], "lidar")) {
		const int n = 16384;
		float* dP = to_dev<float>(NULL, 3ull * n); float* dQ = to_dev<float>(NULL, 3ull * n);
		if (lidar_p2p::Read_data(dP, dQ, n) != 0) return 1;
		float* h = (float*)malloc(12ull * n);
		CK(cudaMemcpy(h, dP, 12ull * n, cudaMemcpyDeviceToHost)); spit(argv[2], h, 12ull * n);
		CK(cudaMemcpy(h, dQ, 12ull * n, cudaMemcpyDeviceToHost)); spit(argv[3], h, 12ull * n);
		return 0;
	}
	if (!strcmp(argv[1], "knn_sq")) {
		int m = atoi(argv[2]), k1 = atoi(argv[3]);
		float* dQ = to_dev<float>(slurp(argv[4], 12ull * m), 3ull * m);
		int* dN = to_dev<int>(NULL, (size_t)m * k1);
		float* dD = to_dev<float>(NULL, (size_t)m * m);
		bunny_p2l::knn<<<60, m / 60 + 1>>>(dQ, m, dQ, m, dN, k1, dD);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		int* hN = (int*)malloc(4ull * m * k1); CK(cudaMemcpy(hN, dN, 4ull * m * k1, cudaMemcpyDeviceToHost));
		spit(argv[5], hN, 4ull * m * k1); return 0;
	}
	if (!strcmp(argv[1], "match")) {
		int n = atoi(argv[2]), m = atoi(argv[3]);
		float* dP = to_dev<float>(slurp(argv[4], 12ull * n), 3ull * n); float* dQ = to_dev<float>(slurp(argv[5], 12ull * m), 3ull * m);
		int* dI = to_dev<int>(NULL, n);
		lidar_p2p::Matching<<<60, n / 60 + 1>>>(n, dP, dQ, m, dI);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		int* hI = (int*)malloc(4ull * n); CK(cudaMemcpy(hI, dI, 4ull * n, cudaMemcpyDeviceToHost));
		spit(argv[6], hI, 4ull * n); return 0;
	}
	fprintf(stderr, "unknown command %s\n", argv[1]);
	return 2;
}
