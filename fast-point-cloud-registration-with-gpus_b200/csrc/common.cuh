// common.cuh — shared declarations of the B200 ICP engine (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "icp_b200.h"
#include <nvtx3/nvToolsExt.h>

namespace icpb {

typedef unsigned long long u64;

// ---- K1 (brute-force matching) geometry ------------------------------------------------------
// Targets live in HBM re-tiled as [tile][X|Y|Z][K1_TT] floats so that one 1-D TMA bulk copy brings
// one tile (12 KB) into shared memory. Correspondence tracking is done per sub-tile of K1_TRK
// targets: the inner loop only keeps a running minimum; the winning index is recovered afterwards
// by re-scanning one sub-tile per source.
constexpr int K1_TT     = 1024;   // targets per shared-memory tile
constexpr int K1_TRK    = 128;    // targets per tracking sub-tile
constexpr int K1_STAGES = 3;      // TMA ring depth
constexpr int K1_TILE_BYTES = 3 * K1_TT * 4;
constexpr u64 KEY_UNMATCHED = ~0ull;

// ---- device-resident control block of a registration ----------------------------------------
struct IterState {
	int    done;          // set on the device when the reference's loop would `break` / hit MAX_ITER
	int    iteration;     // the reference's `iteration` counter
	int    iters_run;     // loop bodies executed
	int    numeric_error; // 6x6 system not SPD
	int    ticket_a;      // last-block tickets of the two reduction kernels
	int    ticket_b;
	int    max_iter;
	int    stop_early;
	int    flags;         // ICPB_FLAG_*
	int    count_slot;    // moments[count_slot] = number of points that entered the sums (15 point-to-point, 27 point-to-plane)
	double tol;
	double n_total;       // global number of source points (all ranks)
	double moments[32];   // reduced moment sums (16 point-to-point, 28 point-to-plane)
	double err_sum;       // sum of squared residuals
	float  R[9];          // this iteration's rotation (column-major) and translation
	float  T[3];
	double Rtot[9];       // accumulated transform
	double ttot[3];
	float  last_err;
};

struct K1Params {
	const float* px; const float* py; const float* pz;   // SoA sources, padded to nb * block sources
	const float* qtiles;                                  // [nt][3][K1_TT]
	u64*         keys;                                    // per source: (distance bits << 32) | target index
	int          n;                                       // valid sources
	int          nt;                                      // target tiles
	long long    units;                                   // nb * nt (source block x target tile work units)
	float        thr0;                                    // initial threshold in the squared-distance domain
	float        sentinel;
	const int*   done;
	const int*   remap;                                   // optional: process sources remap[0..*count_dev) only (grid fallback)
	const int*   count_dev;
};

// ---- peer-memory exchange (peer_exchange.cuh, dist.cpp) ---------------------------------------
constexpr int PEER_MAX     = 8;     // ranks of one NVSwitch domain served by the fused exchange (more: NCCL path)
constexpr int PEER_PAYLOAD = 32;    // doubles per row
constexpr int PEER_ROW     = 40;    // payload + flag (u64) + pad: 320 B, a row never shares a 32 B sector with another rank's row
struct PeerXchg {
	int     rank, world;            // world == 0: exchange disabled (single GPU, or NCCL fallback)
	u64*    seq;                    // device counter of completed exchanges (this rank)
	double* mailbox[PEER_MAX];      // mailbox[r] = rank r's mailbox as mapped in this process (mailbox[rank] is local)
};

struct ReduceParams {
	const float* px; const float* py; const float* pz;
	float*       ox; float* oy; float* oz;               // transform output (in place allowed)
	const float4* q4;                                     // targets AoS padded to float4
	const float4* nrm4;                                   // normals (point-to-plane)
	u64*         keys;
	int*         idx;
	int*         seed;
	int          n;
	double*      partials;                                // [grid][32]
	IterState*   st;
	float*       errors;                                  // [max_iter + 1]
	int          fuse_tail;                               // the last block also runs the solve / bookkeeping (single GPU, or
	                                                      // several GPUs with the peer-memory exchange fused in)
	int          metric;
	int          flags;                                   // ICPB_FLAG_* of the run
	PeerXchg     peer;                                    // peer.world > 1: exchange the sums inside the last block
};

// ---- uniform grid over the target (grid_nn.cu) and its occupancy pyramid (grid_tree.cuh) -----------------
struct GridGeom { float ox, oy, oz, inv_h, h; int nx, ny, nz; };
constexpr int GP_MAX_L = 18;       // pyramid levels: dims up to 2^17 per axis
struct GridPyramid {
	int levels;                    // level 0 = the grid's cells, level levels-1 = one root node
	int nx[GP_MAX_L], ny[GP_MAX_L], nz[GP_MAX_L];
	long long off[GP_MAX_L];       // offset of each level's occupancy bytes in `occ`
	const unsigned char* occ;
	float slack;                   // absolute per-axis slack of the box distance (grid_tree.cuh)
};

// ---- tracing: NVTX ranges per API call and per phase of an iteration; header-only NVTX v3 costs a
// null-pointer test when no tool is attached. nsys / ncu --nvtx show them as icpb:<name>.
struct NvtxRange {
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
	NvtxRange(const NvtxRange&) = delete;
	NvtxRange& operator=(const NvtxRange&) = delete;
};
#define ICPB_NVTX(name) ::icpb::NvtxRange nvtx_range__(name)

// ---- error handling -----------------------------------------------------------------------------
struct Ctx;
int  fail_cuda(Ctx* c, cudaError_t e, const char* what, const char* file, int line);
#define ICPB_CUDA(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::icpb::fail_cuda((c), e__, #call, __FILE__, __LINE__); } while (0)

// ---- kernel launchers (defined next to their kernels) -------------------------------------------
int  k1_max_block_sources();
int  launch_match_brute(Ctx* c, int dist_mode, float sentinel);
int  launch_key_reset(Ctx* c);
int  launch_resolve(Ctx* c, float sentinel);
int  launch_match(Ctx* c, int dist_mode, int nn_method, float sentinel);
int  launch_match_grid(Ctx* c, int dist_mode, float sentinel);
int  launch_knn_tree(Ctx* c, int k1, int knn_dist_mode, int* nbr);   // K5 through the grid's occupancy pyramid (grid_nn.cu)
int  launch_match_filter(Ctx* c, int dist_mode, float sentinel);
int  kf_policy_update(Ctx* c);
int  prepare_match_filter(Ctx* c);
int  prepare_match_grid(Ctx* c);
int  build_filter_tc_data(Ctx* c);
int  ensure_filter_tc_data(Ctx* c);
int  ensure_source_order(Ctx* c);
int  launch_match_filter_tc(Ctx* c, int dist_mode, float sentinel);
int  filter_tc_check(Ctx* c);
int  launch_moments(Ctx* c, int metric);
int  launch_solve(Ctx* c, int metric);
int  launch_transform(Ctx* c);
int  launch_finish(Ctx* c);
int  launch_pack_source(Ctx* c, const float* d_xyz, int n, bool reset_seed);
int  launch_unpack_source(Ctx* c, float* d_xyz);
int  launch_pack_target(Ctx* c, const float* d_xyz, int m);
int  launch_target_same(Ctx* c, const float* d_xyz, int m, int* d_differ);
int  launch_fp32_peak(Ctx* c, float* d_out, int iters, int blocks);

// ---- NCCL (loaded lazily with dlopen; see dist.cpp) ------------------------------------------------
struct Dist;
int  dist_unique_id(void* id128);
int  dist_init(Dist** out, int rank, int world, const void* id128, char* err, size_t errlen);
int  dist_allreduce_f64(Dist* d, double* dev_buf, int count, cudaStream_t s, char* err, size_t errlen);
// Maps every rank's mailbox into this process (CUDA IPC, handles exchanged with ncclAllGather); all ranks agree
// (ncclAllReduce min) on whether the fused exchange can be used. out->world stays 0 when it cannot.
int  dist_peer_init(Dist* d, int device, cudaStream_t s, PeerXchg* out, char* err, size_t errlen);
void dist_destroy(Dist* d);
int  dist_check_async(Dist* d, char* err, size_t errlen);       // ncclCommGetAsyncError (NCCL path only)
bool dist_has_comm(const Dist* d);
int  dist_init_local(Dist** outs, PeerXchg* peers, const int* devices, int world, char* err, size_t errlen);   // in-process group
int  create_context(Ctx** out, int device);                     // engine.cu

// device / pinned buffers of icpb_run_batched, kept across calls (grow-only)
struct K9Buffers {
	float *s = nullptr, *t = nullptr, *e = nullptr; int* ints = nullptr; double* dbl = nullptr;
	size_t s_cap = 0, t_cap = 0, e_cap = 0, ints_cap = 0, dbl_cap = 0;
	cudaStream_t copy_stream = nullptr; cudaEvent_t ev = nullptr;
	int* one = nullptr;              // pinned {1, 0}: source of the "chunk landed" flag copies
};

// ---- the context ---------------------------------------------------------------------------------
struct Ctx {
	int device = 0, sm_count = 0, sm_clock_khz = 0;
	char name[64] = {0};
	cudaStream_t stream = nullptr;
	char err[512] = {0};
	long long launches = 0;

	// distributed
	int rank = 0, world = 1;
	Dist* dist = nullptr;
	double n_total = 0.0; bool n_total_valid = false;   // global source count (all ranks), refreshed after icpb_set_source
	PeerXchg peer = {};           // peer.world > 1: moments / residual sums are exchanged inside K2/K7/K4 over peer memory

	// target
	int m = 0, nt = 0;
	float4* q4 = nullptr;        // [m] x,y,z,0
	float*  qtiles = nullptr;    // [nt][3][K1_TT], padded with +inf
	float4* nrm4 = nullptr;      // [m] normals
	int*    nbr = nullptr;       // [m][k+1]
	int     knn_k = 0;
	bool    have_normals = false;

	// exact uniform-grid index of the target (ICPB_NN_GRID)
	bool    grid_ready = false;
	int*    grid_cell_start = nullptr;
	float4* grid_sorted4 = nullptr;
	int     grid_dim[3] = {0, 0, 0};
	float   grid_origin[3] = {0, 0, 0};
	float   grid_cell = 0.f;
	int*    grid_open_list = nullptr;   // sources the grid search left open (finished by brute force)
	int     grid_open_cap = 0;
	unsigned long long* grid_counters = nullptr;   // [0] open sources of the last pass, [1] candidates visited
	bool    grid_pyramid = true;        // best-first descent of an occupancy pyramid (grid_tree.cuh); ICPB_GRID_PYRAMID=0: rings + brute-force fallback
	unsigned char* grid_occ = nullptr;  // pyramid occupancy bytes, all levels
	bool    knn_pyramid = true;         // neighbour lists through the pyramid (0.33 ms at 100k points on B200 against 6.5 ms for the tiled
	                                    // brute-force scan, identical lists); ICPB_KNN_PYRAMID=0 selects the tiled scan
	// build scratch of the grid, kept across targets (grow-only: a host-driven loop re-uploads the target every step)
	int*    grid_counts = nullptr; int* grid_cell_of = nullptr; int* grid_sums = nullptr; unsigned* grid_mm = nullptr;
	size_t  grid_cells_cap = 0, grid_pts_cap = 0, grid_sums_cap = 0, grid_occ_cap = 0;
	GridPyramid grid_py = {};           // level dimensions / offsets (host copy; `occ` points at grid_occ)

	// source (this rank's shard)
	int n = 0, n_cap = 0;        // n_cap: padded capacity
	int n_cap_at_graph = 0;      // capacity when graph_gen was last checked (a reallocation invalidates captured graphs)
	float *px = nullptr, *py = nullptr, *pz = nullptr;
	u64*  keys = nullptr;
	int*  idx = nullptr;
	int*  seed = nullptr;        // last resolved correspondences: warm start of K1F; survives set_source of the same size
	int   seed_n = -1;
	float* dmin = nullptr;       // winning distance per source (icpb_match / icpb_time_match)
	float* stage_xyz = nullptr;  // device AoS staging for H2D/D2H
	size_t stage_cap = 0;

	// K1F: lower-bound filter data of the target (nn_filter.cu)
	bool    kf_ready = false;
	float*  kf_tiles7 = nullptr;        // [nt][X Y Z | Xc Yc Zc | W3 W2][512]
	int     kf_tiles_cap = 0;           // tiles allocated (kept across targets of the same size)
	bool    kf_tiles_built = false;     // kf_tiles7 holds the current target (false when only centre + radius were computed for K1T)
	unsigned long long* kf_scratch = nullptr;   // bbox / radius / axis-score scratch of build_filter_data
	float   kf_center[3] = {0, 0, 0};
	float   kf_rq = 0.f;                // >= max |q - centre|
	int     kf_nt = 0;
	unsigned long long* kf_stats = nullptr;
	unsigned long long* kf_work_counter = nullptr;
	int     kf_gss[3] = {0, 0, 0};      // ICPB_KF_GSS=min,max,div: guided self-scheduling parameters (experiments)
	int     kf_chunk_override = 0;      // ICPB_KF_CHUNK: tiles per work chunk
	int     kf_s = 8;                   // sources per thread of the filter kernel (ICPB_KF_S=8|16)
	int     kf_drop = 2;                // axis left out of the planar (2-FMA) bound, chosen per target (kf_score_kernel)
	int     kf_drop_forced = -1;        // ICPB_KF_DROP=0|1|2
	double  kf_score[3] = {0, 0, 0};
	int     kf_dims = 2;                // bound the next warm launch uses: 2 = planar, 3 = full (kf_policy_update)
	int     kf_dims_last = 0;           // bound of the last launch
	int     kf_dims_forced = 0;         // ICPB_KF_DIMS=2|3
	int     kf_bounces = 0, kf_hold = 0; // planar -> full switches so far; full-bound launches left before planar is retried
	bool    kf_seeded = false;          // the seeds hold real correspondences (not the reset value)
	unsigned long long kf_stats_seen[4] = {0, 0, 0, 0};   // tests, exact passes; K1T: source sweeps, of which with s0 < the largest group radius
	double  kf_last_frac = 0.0;
	bool    kf_use_seed = true;         // warm start from the previous correspondences
	double  kf_min_pairs = 2.5e8;       // below this many pairs per pass the direct kernel is used (ICPB_K1_FILTER_MIN_PAIRS); K1T pays off
	                                    // from 16 384 x 16 384 (62 us against 85 us per pass), K1F (ICPB_K1_TC=0) from ~3e4 x 3e4
	bool    k1_use_filter = true;       // ICPB_NN_BRUTE goes through the filter kernel (ICPB_K1_FILTER=0 disables)
	// K1T: the filter on the tensor cores (nn_filter_tc.cu)
	bool    k1_use_tc = true;           // ICPB_NN_BRUTE goes through K1T (ICPB_K1_TC=0: the FP32 filter kernel K1F)
	int     kt_variant = -1;            // ICPB_KT_VAR: forces a pipeline shape of K1T (nn_filter_tc.cu); -1 = automatic
	int     kt_tpc_start = 16;          // ICPB_KT_TPC: where the policy starts for a new target (1, 2, 4, 8, 16)
	int     kt_policy_m = -1;           // target size the policy state belongs to
	float   kt_policy_fp[4] = {0, 0, 0, -1};   // ... and its centre and radius: the same cloud uploaded again keeps the state
	bool    kt_tpc_forced_start = false;       // ICPB_KT_TPC given: start there whatever the target size
	int     kt_tpc_auto = 16;           // targets per MMA column the policy currently uses (halved when exact passes pile up)
	bool    kt_ready = false;
	float*  kt_tiles = nullptr;         // [nt][B operand block 16 KB | X Y Z originals 3 KB], 256 targets per tile
	size_t  kt_tiles_cap = 0;           // floats allocated
	int     kt_nt = 0;
	int     kt_tpc = 1, kt_built_tpc = 0;   // targets per MMA column: 1, or 2 / 4 = the grouped forms (consecutive targets share a column)
	int*    kt_fail = nullptr;          // device flag: a bounded mbarrier wait of the pipeline timed out
	int*    tgt_differ = nullptr;       // device flag of icpb_set_target's "same cloud again?" comparison
	bool    tgt_packed = false;         // q4 holds a packed target
	unsigned long long* kt_work = nullptr;   // [0] work counter of the K1T pass, [1] CTAs through (the last one re-arms both)
	int*    kt_colstart = nullptr;      // grouped form: first target of every column; kt_scan_a / kt_scan_b: scan scratch (run starts, column ids)
	int*    kt_scan_a = nullptr;
	int*    kt_scan_b = nullptr;
	size_t  kt_cols_cap = 0;
	void*   kt_cub_tmp = nullptr;
	size_t  kt_cub_cap = 0;
	float*  kt_hmax_d = nullptr;        // device: [0] largest group radius (float), then the step sum (double at byte 8)
	float   kt_hmax = 0.0f;
	int     kt_ncols = 0;
	int     kt_sort_mode = -1;          // ICPB_KT_SORT: 1 = always build the tiles over the targets in Morton order, 0 = never, -1 = when the scan order is incoherent
	bool    kt_sorted = false;          // what the current tiles are
	unsigned long long* kt_mkeys = nullptr;   // [2][cap] Morton keys (in / out of the radix sort)
	int*    kt_perm2 = nullptr;         // [2][cap] indices (in / out): the second half is the sorted order's original indices
	float4* kt_q4s = nullptr;           // the targets in Morton order
	size_t  kt_sort_cap = 0;
	int*    kt_slot_index = nullptr;    // original index of the target in every slot of every tile
	bool    kt_src_checked = false;     // the order of the current source has been looked at (once per upload)
	bool    kt_src_sorted = false;      // ... and it is visited in Morton order through kt_sperm2's second half
	unsigned long long* kt_skeys = nullptr;
	int*    kt_sperm2 = nullptr;
	size_t  kt_ssort_cap = 0;
	size_t  kt_slot_cap = 0;

	// iteration state
	IterState* st = nullptr;     // device
	IterState* st_host = nullptr;// pinned
	double* partials = nullptr;  // [reduce_grid][32]
	int reduce_grid = 0;
	float* errors = nullptr;     // device [err_cap]
	int err_cap = 0;
	float* errors_host = nullptr;// pinned
	bool   step_state_ready = false;
	int    run_flags = 0;        // ICPB_FLAG_* of the current run (step-wise API: 0)

	// K1 configuration
	int k1_cfg = 6;              // index into the K1 tuning table (nn_bruteforce.cu); 6 = best measured on B200
	bool k1_cfg_forced = false;  // ICPB_K1_CFG given: no size-based choice
	int k1_grid_override = 0;
	double pairs_acc = 0;

	// CUDA-graph replay of a batch of iterations (launch-bound small problems, engine.cu)
	cudaGraphExec_t graph_exec = nullptr;
	long long graph_gen = 0, graph_built_gen = -1;   // bumped whenever clouds / buffers change
	int  graph_batch = 0, graph_launches = 0;
	int  graph_key[4] = {-1, -1, -1, -1};             // metric, dist_mode, nn_method, flags
	float graph_sentinel = 0.f;
	bool graphs_enabled = false;                     // ICPB_GRAPHS=1 (or ICPB_FLAG_GRAPH) enables: capture + instantiate cost
	                                                 // milliseconds, which only repeated registrations of one size amortise

	K9Buffers k9;

	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t* ev_match = nullptr; int ev_match_cap = 0;
};

} // namespace icpb
