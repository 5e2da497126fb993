# BASELINE.json configs 1-3 and 5 on one B200 (config 4 is bench.py): whole registrations through the C ABI
import os, sys, json, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
out = {}
ctx = ib.Context(0)
# config 1: src/ICP_CPU.c's own clouds (100x100, its pose), tol 1e-5, MAX_ITER 200 — on the GPU engine
D, M = icp_synth.cpu_clouds(100)
ctx.set_target(M); ctx.set_source(D); ctx.run(ib.default_params(max_iter=200, tol=1e-5))
ctx.set_source(D); e, r = ctx.run(ib.default_params(max_iter=200, tol=1e-5))
out["config1_icp_cpu_clouds_10k"] = {"iterations_run": r.iterations_run, "elapsed_ms": r.elapsed_ms, "final_rms": float(e[r.iterations + 1]),
                                     "reference_ICP_CPU.c_ms_same_box_class": 18678.0, "reference_iterations": 61}
# config 2: 100k x 100k point-to-point, brute force
D, M = icp_synth.p2p_clouds(317, 100000)
ctx.set_target(M); ctx.set_source(D); ctx.run(ib.default_params(max_iter=64))
ctx.set_source(D); e, r = ctx.run(ib.default_params(max_iter=64, flags=ib.FLAG_PROFILE))
out["config2_p2p_100k"] = {"iterations_run": r.iterations_run, "elapsed_ms": r.elapsed_ms, "match_ms": r.match_ms, "icp_iters_per_sec": r.iterations_run / (r.elapsed_ms * 1e-3),
                           "nn_pairs_per_sec": r.nn_pairs / (r.match_ms * 1e-3), "final_rms": float(e[r.iterations + 1]), "idx_identity": bool(np.array_equal(ctx.correspondences(), np.arange(100000)))}
# config 3: 100k point-to-plane with k-NN PCA normals
ctx.set_source(D); nms = ctx.estimate_normals(4); nms = ctx.estimate_normals(4)
e, r = ctx.run(ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT, max_iter=50, flags=ib.FLAG_PROFILE))
out["config3_p2plane_100k"] = {"normals_ms": nms, "knn_pairs_per_sec": 1e10 / (nms * 1e-3), "iterations_run": r.iterations_run, "elapsed_ms": r.elapsed_ms, "match_ms": r.match_ms,
                               "nn_pairs_per_sec": r.nn_pairs / (r.match_ms * 1e-3), "final_rms": float(e[r.iterations + 1])}
# config 5: 4096 pairs of 2048 points (one GPU; 8 GPUs = replicas, tools/bench_batched.py under torchrun)
S, T, rr, tt = icp_synth.batched_pairs(4096)
ctx.run_batched(ib.default_params(max_iter=40), S[:64], T[:64])
errors, iters, R, t, ms = ctx.run_batched(ib.default_params(max_iter=40), S, T)
its = int((iters + 1).sum())
out["config5_batched_4096x2048_one_gpu"] = {"kernel_ms": ms, "registrations_per_sec": 4096 / (ms * 1e-3), "icp_iters_per_sec": its / (ms * 1e-3), "nn_pairs_per_sec": its * 2048.0 ** 2 / (ms * 1e-3),
                                            "all_poses_recovered": bool(all(np.abs(R[b] - icp_synth.euler_matrix(rr[b])).max() < 2e-5 for b in range(4096)))}
print(json.dumps(out, indent=1))
