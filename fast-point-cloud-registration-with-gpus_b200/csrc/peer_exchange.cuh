// peer_exchange.cuh — C1 fused: the per-iteration exchange of moment sums done INSIDE the reduction kernels over
// NVLink/NVSwitch peer memory, instead of an ncclAllReduce launched between them.
//
// The payload is 16-28 doubles per rank per exchange, so this is a latency problem, not a bandwidth one: with NCCL an
// iteration on G GPUs is 7 launches (match, moments, allreduce, solve, transform, allreduce, finish); with the exchange
// fused into the last block of K2/K7 and of K4 it is the same 3 kernels as on one GPU.
//
// Every rank owns a mailbox in its own HBM, mapped into every peer process with CUDA IPC (dist.cpp):
//     mailbox[slot 0..1][source rank 0..PEER_MAX-1][PEER_ROW doubles]      row = payload[0..31] | flag (u64) | pad
// One exchange = the last block of the kernel
//   1. stores its row into slot (seq & 1), row `rank`, of EVERY rank's mailbox (remote stores over NVLink, st.relaxed.sys),
//   2. fences (fence.sc.sys via __threadfence_system) and then publishes flag = seq in each of those rows (st.release.sys),
//   3. spins on the `world` flags of its OWN mailbox (local HBM: polling costs no NVLink traffic) until all equal seq,
//   4. adds the rows in RANK ORDER — every rank computes bit-identical sums, so every rank then runs the same 3x3 / 6x6
//      solve on identical bits and no broadcast of R,T is needed (same argument as the NCCL path).
// `seq` counts exchanges since the mailboxes were zeroed; all ranks execute the same sequence of exchanges because the
// loop state they branch on (IterState::done, ...) is itself a function of the exchanged sums. Two slots suffice: a rank
// can start exchange k+1 while a peer still reads slot k & 1, but cannot reach k+2 before that peer has published k+1.
// A peer that never arrives (crashed process) trips a ~20 s clock64() timeout that raises IterState::numeric_error = 100
// and IterState::done instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace icpb {

__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys_u64(u64* p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys_u64(const u64* p) { u64 v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// In-place sum over ranks of vals[0..nv) (nv <= PEER_PAYLOAD). Called by ALL threads of ONE block (blockDim >= 32).
// Returns false on timeout.
__device__ __forceinline__ bool peer_allreduce_block(const PeerXchg& px, double* vals, int nv)
{
	__shared__ u64 s_seq;
	__shared__ int s_ok;
	const int t = threadIdx.x;
	if (t == 0) { s_seq = *px.seq + 1; s_ok = 1; }
	__syncthreads();
	const u64 seq = s_seq;
	const int slot = (int)(seq & 1);
	const size_t my_row = ((size_t)slot * PEER_MAX + (size_t)px.rank) * PEER_ROW;
	// 1. my row -> every mailbox (including my own)
	for (int k = t; k < nv * px.world; k += blockDim.x) {
		const int r = k / nv, e = k - r * nv;
		st_relaxed_sys_f64(px.mailbox[r] + my_row + e, vals[e]);
	}
	__threadfence_system();
	__syncthreads();
	// 2. publish
	if (t < px.world) st_release_sys_u64(reinterpret_cast<u64*>(px.mailbox[t] + my_row + PEER_PAYLOAD), seq);
	// 3. wait for every rank's row in my own mailbox
	if (t < px.world) {
		const u64* flag = reinterpret_cast<const u64*>(px.mailbox[px.rank] + ((size_t)slot * PEER_MAX + (size_t)t) * PEER_ROW + PEER_PAYLOAD);
		const long long t0 = clock64();
		while (ld_acquire_sys_u64(flag) != seq) {
			if (clock64() - t0 > 40000000000ll) { s_ok = 0; break; }     // ~20 s at 2 GHz
			__nanosleep(100);
		}
		__threadfence_system();
	}
	__syncthreads();
	const bool ok = s_ok != 0;
	// 4. fixed-order sum
	if (ok && t < nv) {
		double s = 0.0;
		for (int r = 0; r < px.world; r++) s += ld_relaxed_sys_f64(px.mailbox[px.rank] + ((size_t)slot * PEER_MAX + (size_t)r) * PEER_ROW + t);
		vals[t] = s;
	}
	if (t == 0) *px.seq = seq;
	__syncthreads();
	return ok;
}

} // namespace icpb
