// k1_device.cuh — device helpers shared by the matching kernels (K1, K5, K9): packed FP32 PTX wrappers,
// TMA/mbarrier wrappers, the reference's distance chain and the sqrt-class threshold.
#pragma once
#include "common.cuh"

namespace icpb {

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 bcast2v(float a) { u64 r; asm volatile("mov.b64 %0, {%1,%1};" : "=l"(r) : "f"(a)); return r; }
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// The reference's distance chain, scalar, with the contraction nvcc applies to
// src/ICP_point_to_point.cu:48-50 spelled out (SASS: FADD, FADD, FMUL dy*dy, FFMA dx, FADD, FFMA dz).
__device__ __forceinline__ float dist_chain(float xp, float yp, float zp, float xq, float yq, float zq)
{
	float dx = __fsub_rn(xp, xq), dy = __fsub_rn(yp, yq), dz = __fsub_rn(zp, zq);
	return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Smallest float y with sqrt.rn(y) == sqrt.rn(m): comparing squared distances against it is the
// same as comparing their correctly-rounded square roots with strict `<` against sqrt(m).
__device__ __forceinline__ float sqrt_class_floor(float m)
{
	float s = __fsqrt_rn(m), y = m;
	for (int k = 0; k < 8; k++) {
		if (!(y > 0.0f)) break;
		float yd = __uint_as_float(__float_as_uint(y) - 1u);
		if (__fsqrt_rn(yd) != s) break;
		y = yd;
	}
	return y;
}
template <int MODE> __device__ __forceinline__ float lower_threshold(float m)
{
	if (MODE == ICPB_DIST_SQRT) return sqrt_class_floor(m);
	return m;
}


} // namespace icpb
