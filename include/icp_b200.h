/* icp_b200.h — C ABI of the B200-native ICP registration engine (libicp_b200.so).
 *
 * This is the drop-in boundary for the hot path of
 * Carlos310197/Fast-Point-Cloud-Registration-with-GPUs. The reference has no library API: its
 * hot path is the body of `while (iteration < MAX_ITER)` in three single-file programs. Each entry
 * point below names the reference code it replaces (paths relative to the reference repository).
 *
 * Conventions kept from the reference
 *   - clouds are AoS float xyzxyz... (a column-major 3xN matrix, ld = 3)  src/ICP_point_to_point.cu:142-152
 *   - R is column-major, R[row + 3*col]                                    src/ICP_point_to_point.cu:85-87
 *   - the error array has slot 0 = 0 and iteration k writes slot k+1       src/ICP_point_to_point.cu:416
 *   - matching scans targets in ascending order and updates on strict `<`: the LOWEST index wins
 *     ties; the running minimum starts at `sentinel` (100000) and a source whose every distance is
 *     >= sentinel keeps its previous correspondence                        src/ICP_point_to_point.cu:36,51-55
 *
 * Plain C: pointers and sizes only. All functions return ICPB_OK (0) or a negative ICPB_ERR_*;
 * they never print. One host thread drives one context; calls are synchronous at return.
 * There is NO CPU fallback: every entry point needs a CUDA device of compute capability 10.0.
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ICPB_VERSION 100

#define ICPB_OK            0
#define ICPB_ERR_CUDA     -1   /* a CUDA runtime call or kernel failed (see icpb_last_error) */
#define ICPB_ERR_BADARG   -2
#define ICPB_ERR_NCCL     -3
#define ICPB_ERR_STATE    -4   /* e.g. icpb_run before icpb_set_target, normals missing for point-to-plane */
#define ICPB_ERR_NOMEM    -5
#define ICPB_ERR_NUMERIC  -6   /* 6x6 system not positive definite (cusolver's devInfo > 0 in the reference) */
#define ICPB_ERR_NODEVICE -7

/* Distance formula of the matching step — one per reference executable (SURVEY.md §7 hard part 2). */
#define ICPB_DIST_SQ    0   /* d = dx*dx + dy*dy + dz*dz, fused as nvcc fuses it     src/ICP_point_to_point.cu:48-50 */
#define ICPB_DIST_SQRT  1   /* (float)sqrt of the same chain                         src/ICP_point_to_plane.cu:172-174 */
#define ICPB_DIST_STD   2   /* (float)sqrt(pow(dx,2)+pow(dy,2)+pow(dz,2)) in double  src/ICP_standard.cu:31 */

#define ICPB_POINT_TO_POINT 0   /* centroids, 3x3 cross-covariance, SVD, R = U*V^T   src/ICP_point_to_point.cu:313-398 */
#define ICPB_POINT_TO_PLANE 1   /* 6x6 normal equations, Cholesky, Euler -> R        src/ICP_point_to_plane.cu:537-601 */

#define ICPB_NN_BRUTE 0         /* exact brute force over every target (the reference's method): a rigorous lower bound is
                                   evaluated for every pair — on the tensor cores (tcgen05.mma kind::tf32, K1T) or with 2-3 FMAs
                                   (K1F) — and the reference's chain wherever the bound cannot exclude the pair's sub-tile;
                                   identical indices (icpb_get_filter_config, icpb_get_filter_tc_config) */
#define ICPB_NN_GRID  1         /* exact uniform-grid search, identical indices */
#define ICPB_NN_BRUTE_DIRECT 2  /* exact brute force, the reference's chain evaluated for every pair (no pruning) */

/* Robustness options the reference lacks; all off by default so parity is untouched. */
#define ICPB_FLAG_FIX_REFLECTION 1   /* if det(U*V^T) < 0 flip the singular vector of the smallest singular value (Kabsch);
                                        the reference keeps the reflection (src/ICP_point_to_point.cu:379-381) */

#define ICPB_FLAG_PROFILE 2          /* per-iteration phase times wanted (match_ms, minimize_ms, transform_ms): never replay the
                                        loop from a CUDA graph */
#define ICPB_FLAG_GRAPH   4          /* replay batches of `sync_every` iterations from a CUDA graph (problems below 4e9 pairs
                                        per pass, one GPU). Identical results; match_ms is then reported as 0. Pays off only
                                        when many registrations of one size reuse the instantiated graph. */

#define ICPB_FLAG_REJECT_UNMATCHED 8   /* max-distance rejection: a source with no target below `sentinel` is dropped from the
                                        moment sums and from the RMS of that iteration and reports correspondence -1, instead of
                                        keeping its previous correspondence as the reference does (its LiDAR programs raise the
                                        sentinel to 1e6 precisely because real scans have outliers,
                                        src/CUDA/GPU_point_to_point_real.cu:18,44). The RMS is then over the matched points. */

typedef struct icpb_ctx icpb_ctx;

typedef struct icpb_params {
	int    metric;       /* ICPB_POINT_TO_POINT | ICPB_POINT_TO_PLANE */
	int    dist_mode;    /* ICPB_DIST_* */
	int    nn_method;    /* ICPB_NN_* */
	int    max_iter;     /* MAX_ITER: 40 (src/ICP_point_to_point.cu:24), 50 (ICP_point_to_plane.cu:23) */
	int    stop_early;   /* 1: break when e < tol or |e_k+1 - e_k| < tol (src/ICP_point_to_point.cu:420-421);
	                        0: always run max_iter iterations (src/ICP_standard.cu:369) */
	int    sync_every;   /* iterations enqueued between host reads of the device-side stop flag; 0 = adaptive
	                        (4 below 4e9 pairs per pass, else 1). Never changes results. */
	float  sentinel;     /* 100000 (src/ICP_point_to_point.cu:36); 1e6 in the LiDAR programs */
	double tol;          /* 0.000001 (GPU programs), 0.00001 (src/ICP_CPU.c:267) */
	int    flags;        /* ICPB_FLAG_*; 0 = the reference's behaviour */
} icpb_params;

typedef struct icpb_result {
	int    iterations;      /* the reference's `iteration` counter when its loop exits */
	int    iterations_run;  /* loop bodies executed (= iterations + 1 after a break) */
	double R[9];            /* accumulated rotation, column-major: R_tot <- R_k * R_tot */
	double t[3];            /* accumulated translation:           t_tot <- R_k * t_tot + T_k */
	float  last_R[9];       /* transform of the last iteration (d_temp_r / d_temp_T) */
	float  last_T[3];
	float  elapsed_ms;      /* cudaEvent time around the loop, as src/ICP_point_to_point.cu:294,424-426 */
	float  match_ms;        /* part of elapsed_ms spent in the matching kernels (0 when the loop was replayed from a
	                           CUDA graph, see ICPB_FLAG_GRAPH) */
	double nn_pairs;        /* source x target pairs evaluated by brute-force matching (this rank) */
	float  minimize_ms;     /* with ICPB_FLAG_PROFILE: time of the minimisation step (moments, allreduce, solve) and of the */
	float  transform_ms;    /* fused transformation + error + stop test — the phases the reference's instrumented programs
	                           report (src/CUDA/ICP_point_to_point_clean.cu:464-481); 0 without the flag */
} icpb_result;

/* ---- library ------------------------------------------------------------------------------ */
int         icpb_version(void);
const char* icpb_status_string(int status);
void        icpb_default_params(icpb_params* p);   /* the constants of src/ICP_point_to_point.cu */
int         icpb_device_count(int* count);

/* ---- context ------------------------------------------------------------------------------ */
/* Replaces the cudaMalloc / cublasCreate / cusolverDnCreate block of each main
 * (src/ICP_point_to_point.cu:203-285). All device memory is owned by the context and allocated
 * once per cloud size — nothing is allocated inside the loop (cf. src/ICP_point_to_plane.cu:577). */
int  icpb_create(icpb_ctx** out, int device);
/* Multi-GPU: one context per process/GPU. Source points are sharded (each rank passes its own
 * shard to icpb_set_source), the target is replicated, and the per-iteration moment partials are
 * combined across ranks (inside the kernels over peer memory, or with ncclAllReduce). `nccl_unique_id` is the 128-byte ncclUniqueId made by rank 0 with
 * icpb_nccl_unique_id and handed to every rank by the caller (MPI, torch.distributed, a file...). */
int  icpb_nccl_unique_id(void* id128);
int  icpb_create_dist(icpb_ctx** out, int device, int rank, int world, const void* nccl_unique_id);
/* `peer_exchange` = 1 when the per-iteration sums are exchanged INSIDE the reduction kernels through peer memory
 * (NVLink/NVSwitch; every rank's mailbox mapped into every process with CUDA IPC at icpb_create_dist time), 0 when
 * they go through ncclAllReduce launches between the kernels (more than 8 ranks, ranks that cannot map each other's
 * memory, or ICPB_PEER=0 in the environment). All ranks always agree on the choice. */
int  icpb_dist_info(const icpb_ctx* ctx, int* rank, int* world, int* peer_exchange);
int  icpb_destroy(icpb_ctx* ctx);
const char* icpb_last_error(const icpb_ctx* ctx);
int  icpb_device_info(const icpb_ctx* ctx, int* sm_count, int* sm_clock_khz, char* name64);

/* ---- several GPUs, one process, plain C/C++ (north star: "host code stays C/C++ ... source points sharded across the
 * 8 GPUs of one box with the target replicated") ------------------------------------------------------------------
 * A group is `ndev` contexts (rank r on devices[r]; devices == NULL means 0..ndev-1) created and driven from ONE
 * process: no launcher, no MPI, no unique id. The per-iteration moment sums are exchanged inside the reduction kernels
 * over NVLink peer memory (cudaDeviceEnablePeerAccess), or with ncclAllReduce launches on communicators made by
 * ncclCommInitAll when peer access is missing, ndev > 8 or ICPB_PEER=0. Every call below drives all devices at once
 * (one host thread per device inside the call) and returns when all of them are done. Results — correspondences,
 * trajectories, transforms — are those of one GPU: indices bit for bit, sums to FP64 summation-order noise.
 * Replaces nothing in the reference (it is single-GPU, src/ICP_point_to_point.cu:90-460); it is that main()'s shape. */
typedef struct icpb_group icpb_group;
int  icpb_group_create(icpb_group** out, const int* devices, int ndev);
int  icpb_group_destroy(icpb_group* g);
int  icpb_group_size(const icpb_group* g);
icpb_ctx* icpb_group_ctx(icpb_group* g, int rank);        /* rank's context, e.g. for icpb_get_filter_stats; owned by the group */
const char* icpb_group_last_error(const icpb_group* g);
int  icpb_group_info(const icpb_group* g, int* ndev, int* peer_exchange, int* devices /* [ndev] or NULL */);
int  icpb_group_set_target(icpb_group* g, const float* xyz_host, int m);                  /* replicated on every device */
/* Shards the source: block > 0 deals blocks of `block` consecutive points round-robin (2048 = the matching kernels'
 * source block: every rank sees the same mix of regions of the cloud, which balances the matching step); block == 0
 * gives contiguous shards. */
int  icpb_group_set_source(icpb_group* g, const float* xyz_host, int n, int block);
int  icpb_group_estimate_normals(icpb_group* g, int k, int knn_dist_mode, float* elapsed_ms);
/* icpb_run on every rank; `result` is rank 0's (all ranks hold identical bits; the call checks it) with the times taken
 * as the maximum over the ranks and nn_pairs summed. */
int  icpb_group_run(icpb_group* g, const icpb_params* params, float* errors, icpb_result* result);
int  icpb_group_get_source(icpb_group* g, float* xyz_host);           /* in the original point order */
int  icpb_group_get_correspondences(icpb_group* g, int* idx_host);    /* in the original point order */

/* ---- clouds ------------------------------------------------------------------------------- */
/* `xyz` is AoS float[3*count]; `on_device` != 0 means it already is a device pointer on the
 * context's GPU. Replaces the cudaMemcpy H2D of src/ICP_point_to_point.cu:207-208. The target is
 * re-tiled once here (it never moves during a registration). Uploading a cloud that is bit for bit the
 * current target again (a host-driven loop does, at every step) keeps everything derived from it: normals
 * (they stay valid and in place), the matching tiles and grid, captured graphs. */
int  icpb_set_target(icpb_ctx* ctx, const float* xyz, int m, int on_device);
int  icpb_set_source(icpb_ctx* ctx, const float* xyz, int n, int on_device);
int  icpb_get_source(icpb_ctx* ctx, float* xyz, int on_device);          /* current (transformed) source */
int  icpb_get_correspondences(icpb_ctx* ctx, int* idx, int on_device);   /* d_idx */
int  icpb_get_min_distances(icpb_ctx* ctx, float* d, int on_device);     /* the winning distance per source, in dist_mode units */

/* ---- the steps of one iteration ------------------------------------------------------------ */
/* Matching — `Matching<<<>>>` (src/ICP_point_to_point.cu:31-57,298; variants src/ICP_standard.cu:21-39,
 * src/ICP_point_to_plane.cu:163-181). Exact; bit-identical indices. */
int  icpb_match(icpb_ctx* ctx, int dist_mode, int nn_method, float sentinel);
/* Minimisation — Q_index + cublasSgemv x2 + deviation + cublasSgemm + cusolverDnSgesvd + 2 cublasSgemm
 * (src/ICP_point_to_point.cu:308-397) or Cxb + cublasSgemv x2 + Spotrf/Spotrs + Euler
 * (src/ICP_point_to_plane.cu:532-601). Outputs this iteration's R (column-major) and T. */
int  icpb_minimize(icpb_ctx* ctx, int metric, float R[9], float T[3]);
/* Transformation + error — RyT + cublasScopy + Scopy/Saxpy/Snrm2 (src/ICP_point_to_point.cu:403-416).
 * Applies the R,T of the last icpb_minimize; returns the RMS  ||P' - Q_idx|| / sqrt(N). */
int  icpb_transform(icpb_ctx* ctx, float* rms);
/* Overrides the R (column-major), T that the next icpb_transform applies — the reference's d_temp_r / d_temp_T
 * (src/ICP_point_to_point.cu:379-397), for callers that minimise elsewhere and for testing K4 against `RyT` alone. */
int  icpb_set_transform(icpb_ctx* ctx, const float R[9], const float T[3]);
/* Moments of the last icpb_minimize (after the allreduce when distributed): 16 doubles for
 * point-to-point {sum p (3), sum q (3), sum q p^T (9, column-major, rows = q), N},
 * 28 for point-to-plane {upper triangle of C row by row (21), b (6), N}. */
int  icpb_get_moments(icpb_ctx* ctx, double* mom, int count);

/* ---- normals (point-to-plane) --------------------------------------------------------------- */
/* knn<<<>>> + Normals<<<>>> + host LAPACKE_ssyev/cblas_isamin (src/ICP_point_to_plane.cu:378-447)
 * on the target: exact k+1 nearest (self first, lowest index first on ties, over sqrt'ed float
 * distances), PCA normal = eigenvector of the eigenvalue of smallest magnitude. No MxM matrix. */
int  icpb_estimate_normals(icpb_ctx* ctx, int k, float* elapsed_ms);
/* Same with the distance the k-NN ranks by: ICPB_DIST_SQRT = the canonical program (above); ICPB_DIST_SQ = the squared
 * chain without sqrt, as the dataset programs and the "clean" variant do (src/CUDA/GPU_point_to_plane_bunny.cu:47-82,
 * src/CUDA/ICP_point_to_plane_clean.cu) — the two orders differ where sqrt.rn merges neighbouring squares. */
int  icpb_estimate_normals_ex(icpb_ctx* ctx, int k, int knn_dist_mode, float* elapsed_ms);
int  icpb_get_neighbors(icpb_ctx* ctx, int* nbr /* m*(k+1), row-major */, int on_device);
int  icpb_get_normals(icpb_ctx* ctx, float* normals /* 3*m AoS */, int on_device);
int  icpb_set_normals(icpb_ctx* ctx, const float* normals, int on_device);

/* ---- whole loop ------------------------------------------------------------------------------ */
/* The `while (iteration < MAX_ITER)` loop of src/ICP_point_to_point.cu:295-423,
 * src/ICP_standard.cu:369-463 (stop_early = 0, dist_mode = STD) and src/ICP_point_to_plane.cu:517-631,
 * entirely on the device; `errors` receives max_iter+1 floats (slot 0 = 0). */
int  icpb_run(icpb_ctx* ctx, const icpb_params* params, float* errors, icpb_result* result);
/* One iteration from HOST buffers: uploads both clouds, matches, minimises, transforms, downloads
 * the correspondences and the transform. The end-to-end entry point bench.py times. */
int  icpb_iterate_host(icpb_ctx* ctx, const icpb_params* params, const float* source_xyz, int n,
                       const float* target_xyz, int m, int* idx_out, float R[9], float T[3], float* rms);

/* ---- batched registration (BASELINE.json config 5) ------------------------------------------- */
/* `batch` independent registrations, one per thread block (CTA), the whole loop of each inside one persistent kernel
 * (registrations are drawn from an atomic counter: their iteration counts differ).
 * sources: batch*n*3 floats, targets: batch*m*3 floats — host memory (pinned: streamed in chunks of 32 registrations
 * behind the running kernel; pageable: uploaded first) or device memory (used in place). errors: batch*(max_iter+1);
 * iterations: batch; R: batch*9 (accumulated, column-major, double); t: batch*3 — host. Device buffers are kept in
 * the context between calls. */
int  icpb_run_batched(icpb_ctx* ctx, const icpb_params* params, int batch, const float* sources, int n,
                      const float* targets, int m, float* errors, int* iterations, double* R, double* t,
                      float* elapsed_ms);

/* Page-locked host memory for callers that have no CUDA of their own (the drop-in programs are plain C++): clouds handed
 * to icpb_run_batched from such a buffer are uploaded in chunks BEHIND the running kernel; from ordinary memory they are
 * uploaded first. Also what icpb_set_source / icpb_get_source transfer fastest from. */
int  icpb_host_alloc(void** ptr, unsigned long long bytes);
int  icpb_host_free(void* ptr);

/* ---- dataset front ends ----------------------------------------------- */
/* These work on caller buffers (host, or device when `on_device` != 0), not on the context's clouds; the context only
 * supplies the device and the stream. `elapsed_ms` (optional) is the CUDA-event time of the kernel, which the
 * reference prints ("Conversion kernel's elapsed time", src/CUDA/GPU_point_to_point_real.cu:557). */
/* `Conversion<<<>>>` (src/CUDA/GPU_point_to_point_real.cu:20-36): polar -> Cartesian for an Ouster OS1 capture.
 * Point i is beam i % beams of azimuth block i / beams; the encoder advances `ticks_per_block` (88) per block modulo
 * `ticks_per_rev` (90112); theta = 2*pi*(count/ticks_per_rev + azimuth_deg[beam]/360), phi = 2*pi*altitude_deg[beam]/360
 * in double rounded to float, then x = r*cos(theta)*cos(phi), y = -r*sin(theta)*cos(phi), z = r*sin(phi) in float.
 * range: n floats (millimetres in the reference's capture); xyz_out: 3*n floats AoS. */
int  icpb_lidar_convert(icpb_ctx* ctx, const float* range, int n, unsigned long long encoder_count, const float* altitude_deg,
                        const float* azimuth_deg, int beams, int ticks_per_block, int ticks_per_rev, float* xyz_out, int on_device,
                        float* elapsed_ms);
/* `RyT<<<>>>` on an arbitrary cloud (src/CUDA/GPU_point_to_point_real.cu:113-123, used at :604 to synthesise the target
 * from the scan): out = R*p + T with R column-major; same arithmetic as the loop's transformation step. in == out allowed. */
int  icpb_apply_transform(icpb_ctx* ctx, const float R[9], const float T[3], const float* xyz_in, int n, float* xyz_out, int on_device,
                          float* elapsed_ms);
/* `cublasSscal(3*n, alpha)` (src/CUDA/GPU_point_to_point_real.cu:169-171, millimetres -> metres): xyz *= alpha in place. */
int  icpb_scale_cloud(icpb_ctx* ctx, float alpha, float* xyz, int n, int on_device);

/* ---- measurement helpers --------------------------------------------------------------------- */
/* Register-resident FFMA loop on every SM: the measured FP32 roofline denominator (TFLOP/s). */
int  icpb_measure_fp32_peak(icpb_ctx* ctx, double* tflops);
/* Times `reps` launches of the matching step alone with CUDA events on the context's stream
 * (the protocol of src/CUDA/Matching_opt.cu:213-226); returns the mean and the minimum in ms. */
int  icpb_time_match(icpb_ctx* ctx, int dist_mode, int nn_method, float sentinel, int reps, float* mean_ms, float* min_ms);
/* Uniform-grid matching statistics: candidates actually visited so far (vs the N*M a brute-force pass
 * evaluates), sources the last pass had to hand to the brute-force kernel, grid dimensions and cell size. */
int  icpb_get_grid_stats(icpb_ctx* ctx, double* candidates_visited, int* last_open_sources, int dims[3], float* cell);
/* ICPB_NN_BRUTE statistics since the target was set: (warp x source x 128-target sub-tile) bound tests, and
 * how many of them had to be evaluated with the exact chain. */
int  icpb_get_filter_stats(icpb_ctx* ctx, double* subtile_tests, double* subtile_exact);
/* The lower bound ICPB_NN_BRUTE evaluates per pair is either the full 3-FMA one or its planar 2-FMA restriction (one
 * coordinate axis left out; still a rigorous lower bound, weaker, cheaper). dims_next / dims_last: 2 or 3 for the next /
 * the last launch (chosen from the measured exact-pass rate; ICPB_KF_DIMS forces it); drop_axis: the axis the planar
 * bound leaves out for the current target (-1 before the first launch); last_exact_fraction: share of sub-tile tests
 * that needed the exact chain, as last sampled. Results never depend on any of this. */
int  icpb_get_filter_config(icpb_ctx* ctx, int* dims_next, int* dims_last, int* drop_axis, double* last_exact_fraction);
/* ICPB_NN_BRUTE on the tensor cores (K1T, the default; ICPB_K1_TC=0 selects the FP32 filter): `enabled`, and how many
 * consecutive targets share one MMA column in the launches to come (1 ... 16; chosen from the target size and the measured
 * exact-pass rate).
 * icpb_get_filter_config reports dims_last == 4 after a K1T launch. Results never depend on any of this. */
int  icpb_get_filter_tc_config(icpb_ctx* ctx, int* enabled, int* targets_per_column);
/* Whether the tiles of the last K1T build hold the targets in Morton order (1) or in index order (0). Morton order is
 * chosen when consecutive targets are not neighbours in space (a cloud in arbitrary order: mean step of the scan far
 * above the point spacing), so that one MMA column can still stand for a group of targets; ICPB_KT_SORT=1 / 0 forces /
 * forbids it. Indices returned are always the caller's, the lowest one among equal distances. */
int  icpb_get_filter_tc_order(icpb_ctx* ctx, int* morton_order);
/* Number of kernels this context has launched since creation. */
long long icpb_launch_count(const icpb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ICP_B200_H */
