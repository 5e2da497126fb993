// lidar.cu — dataset front-end kernels: the LiDAR polar -> Cartesian conversion, a stand-alone rigid
// transform and a scale, on caller buffers.
//
// Replaces (reference file:line):
//   Conversion<<<32, n/32>>>            src/CUDA/GPU_point_to_point_real.cu:20-36, launched at :546
//   RyT<<<>>> outside the loop          src/CUDA/GPU_point_to_point_real.cu:113-123, launched at :604 (target = R*scan + T)
//   cublasSscal(3n, 1/1000)             src/CUDA/GPU_point_to_point_real.cu:169-171
// All three are streaming kernels: 4 B in / 12 B out per point (conversion), 12 B in / 12 B out (transform, scale) — HBM
// bound by construction and microseconds at the 16 384 points of one capture; they exist so that a capture goes from raw
// ranges to the two registered clouds without leaving the device. One thread per point, bounds-checked (the reference's
// Conversion has no bounds check and needs n to be a multiple of its grid).
#include "common.cuh"

namespace icpb {

constexpr int LIDAR_MAX_BEAMS = 128;

struct LidarBeams { float altitude[LIDAR_MAX_BEAMS]; float azimuth[LIDAR_MAX_BEAMS]; };

// The arithmetic is the reference kernel's, spelled out: theta and phi are formed in double and rounded to float; the
// three products are float, left to right, with the float cosf/sinf the device overloads of cos/sin resolve to.
__global__ void __launch_bounds__(256) lidar_convert_kernel(const float* __restrict__ r, int n, unsigned long long encoder_count,
                                                            const __grid_constant__ LidarBeams beams, int nbeams, int ticks_per_block,
                                                            unsigned long long ticks_per_rev, float* __restrict__ xyz)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int block = i / nbeams, channel = i - block * nbeams;
	const unsigned long long counter = (encoder_count + (unsigned long long)(block * ticks_per_block)) % ticks_per_rev;
	const double two_pi = 2 * 3.14159265358979323846;
	const float theta = (float)(two_pi * ((double)counter / (double)ticks_per_rev + (double)beams.azimuth[channel] / 360.0));
	const float phi = (float)(two_pi * (double)beams.altitude[channel] / 360.0);
	const float ri = r[i];
	const float ct = cosf(theta), st = sinf(theta), cp = cosf(phi), sp = sinf(phi);
	xyz[3 * (size_t)i + 0] = __fmul_rn(__fmul_rn(ri, ct), cp);
	xyz[3 * (size_t)i + 1] = __fmul_rn(__fmul_rn(-ri, st), cp);
	xyz[3 * (size_t)i + 2] = __fmul_rn(ri, sp);
}

// out_r = fma(R[r+6], z, fma(R[r], x, R[r+3]*y)) + T[r]: the contraction nvcc gives RyT (arithmetic contract)
struct RigidRT { float R[9]; float T[3]; };
__global__ void __launch_bounds__(256) apply_transform_kernel(const float* __restrict__ in, int n, const __grid_constant__ RigidRT rt, float* __restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float x = in[3 * (size_t)i], y = in[3 * (size_t)i + 1], z = in[3 * (size_t)i + 2];
#pragma unroll
	for (int r = 0; r < 3; r++)
		out[3 * (size_t)i + r] = __fadd_rn(__fmaf_rn(rt.R[r + 6], z, __fmaf_rn(rt.R[r], x, __fmul_rn(rt.R[r + 3], y))), rt.T[r]);
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ v, size_t count, float alpha)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) v[i] = __fmul_rn(alpha, v[i]);
}

} // namespace icpb

using namespace icpb;
static inline Ctx* C(icpb_ctx* p) { return reinterpret_cast<Ctx*>(p); }

// scratch device buffer for host callers (separate from the context's cloud staging, which set_source/set_target own)
static int scratch(Ctx* c, float** p, size_t floats)
{
	cudaError_t e = cudaMalloc((void**)p, sizeof(float) * floats);
	if (e != cudaSuccess) { fail_cuda(c, e, "cudaMalloc", __FILE__, __LINE__); return ICPB_ERR_NOMEM; }
	return ICPB_OK;
}
#define LIDAR_CUDA(c, call, cleanup) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup; return fail_cuda((c), e__, #call, __FILE__, __LINE__); } } while (0)

extern "C" {

int icpb_lidar_convert(icpb_ctx* ctx, const float* range, int n, unsigned long long encoder_count, const float* altitude_deg,
                       const float* azimuth_deg, int beams, int ticks_per_block, int ticks_per_rev, float* xyz_out, int on_device,
                       float* elapsed_ms)
{
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	if (!range || !altitude_deg || !azimuth_deg || !xyz_out || n <= 0 || beams < 1 || beams > LIDAR_MAX_BEAMS || ticks_per_rev < 1 || ticks_per_block < 0) {
		snprintf(c->err, sizeof c->err, "icpb_lidar_convert: bad arguments (1 <= beams <= %d)", LIDAR_MAX_BEAMS);
		return ICPB_ERR_BADARG;
	}
	ICPB_CUDA(c, cudaSetDevice(c->device));
	LidarBeams b;
	memset(&b, 0, sizeof b);
	memcpy(b.altitude, altitude_deg, sizeof(float) * (size_t)beams);
	memcpy(b.azimuth, azimuth_deg, sizeof(float) * (size_t)beams);
	float *d_r = nullptr, *d_xyz = nullptr;
	int rc;
	if (!on_device) {
		if ((rc = scratch(c, &d_r, (size_t)n)) != ICPB_OK) return rc;
		if ((rc = scratch(c, &d_xyz, 3 * (size_t)n)) != ICPB_OK) { cudaFree(d_r); return rc; }
		LIDAR_CUDA(c, cudaMemcpyAsync(d_r, range, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream), (cudaFree(d_r), cudaFree(d_xyz)));
	}
	const float* src = on_device ? range : d_r;
	float* dst = on_device ? xyz_out : d_xyz;
	cudaEventRecord(c->ev[2], c->stream);
	lidar_convert_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(src, n, encoder_count, b, beams, ticks_per_block, (unsigned long long)ticks_per_rev, dst);
	c->launches++;
	cudaEventRecord(c->ev[3], c->stream);
	LIDAR_CUDA(c, cudaGetLastError(), (cudaFree(d_r), cudaFree(d_xyz)));
	if (!on_device) LIDAR_CUDA(c, cudaMemcpyAsync(xyz_out, d_xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream), (cudaFree(d_r), cudaFree(d_xyz)));
	LIDAR_CUDA(c, cudaStreamSynchronize(c->stream), (cudaFree(d_r), cudaFree(d_xyz)));
	if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, c->ev[2], c->ev[3]);
	cudaFree(d_r); cudaFree(d_xyz);
	return ICPB_OK;
}

int icpb_apply_transform(icpb_ctx* ctx, const float R[9], const float T[3], const float* xyz_in, int n, float* xyz_out, int on_device,
                         float* elapsed_ms)
{
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	if (!R || !T || !xyz_in || !xyz_out || n <= 0) { snprintf(c->err, sizeof c->err, "icpb_apply_transform: bad arguments"); return ICPB_ERR_BADARG; }
	ICPB_CUDA(c, cudaSetDevice(c->device));
	RigidRT rt;
	memcpy(rt.R, R, sizeof rt.R); memcpy(rt.T, T, sizeof rt.T);
	float* d = nullptr;
	int rc;
	if (!on_device) {
		if ((rc = scratch(c, &d, 3 * (size_t)n)) != ICPB_OK) return rc;
		LIDAR_CUDA(c, cudaMemcpyAsync(d, xyz_in, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream), cudaFree(d));
	}
	cudaEventRecord(c->ev[2], c->stream);
	apply_transform_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(on_device ? xyz_in : d, n, rt, on_device ? xyz_out : d);
	c->launches++;
	cudaEventRecord(c->ev[3], c->stream);
	LIDAR_CUDA(c, cudaGetLastError(), cudaFree(d));
	if (!on_device) LIDAR_CUDA(c, cudaMemcpyAsync(xyz_out, d, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream), cudaFree(d));
	LIDAR_CUDA(c, cudaStreamSynchronize(c->stream), cudaFree(d));
	if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, c->ev[2], c->ev[3]);
	cudaFree(d);
	return ICPB_OK;
}

int icpb_scale_cloud(icpb_ctx* ctx, float alpha, float* xyz, int n, int on_device)
{
	if (!ctx) return ICPB_ERR_BADARG;
	Ctx* c = C(ctx);
	if (!xyz || n <= 0) { snprintf(c->err, sizeof c->err, "icpb_scale_cloud: bad arguments"); return ICPB_ERR_BADARG; }
	ICPB_CUDA(c, cudaSetDevice(c->device));
	const size_t count = 3 * (size_t)n;
	float* d = nullptr;
	int rc;
	if (!on_device) {
		if ((rc = scratch(c, &d, count)) != ICPB_OK) return rc;
		LIDAR_CUDA(c, cudaMemcpyAsync(d, xyz, sizeof(float) * count, cudaMemcpyHostToDevice, c->stream), cudaFree(d));
	}
	scale_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(on_device ? xyz : d, count, alpha);
	c->launches++;
	LIDAR_CUDA(c, cudaGetLastError(), cudaFree(d));
	if (!on_device) LIDAR_CUDA(c, cudaMemcpyAsync(xyz, d, sizeof(float) * count, cudaMemcpyDeviceToHost, c->stream), cudaFree(d));
	LIDAR_CUDA(c, cudaStreamSynchronize(c->stream), cudaFree(d));
	cudaFree(d);
	return ICPB_OK;
}

} // extern "C"
