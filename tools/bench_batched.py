# BASELINE.json config 5 on one GPU: `batch` independent pairs of 2048-point clouds, one ICP per CTA.
import os, sys, time, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
total = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
S, T, r, t = icp_synth.batched_pairs(total)
lo, hi = rank * total // world, (rank + 1) * total // world          # replicas only: each GPU takes its share of the pairs
S, T, r, t = S[lo:hi], T[lo:hi], r[lo:hi], t[lo:hi]
batch = hi - lo
ctx = ib.Context(local)
p = ib.default_params(max_iter=40)
ctx.run_batched(p, S[:8], T[:8])
best = None
for rep in range(3):
    t0 = time.perf_counter()
    errors, iters, R, tt, ms = ctx.run_batched(p, S, T)
    wall = time.perf_counter() - t0
    best = ms if best is None else min(best, ms)
its = int((iters + 1).sum())
ok = all(np.abs(R[b] - icp_synth.euler_matrix(r[b])).max() < 2e-5 for b in range(batch))
print(json.dumps({"rank": rank, "world": world, "batch": batch, "points": 2048, "kernel_ms": best, "wall_ms_incl_copies": wall * 1e3, "registrations_per_sec": batch / (best * 1e-3),
                  "icp_iterations": its, "icp_iters_per_sec": its / (best * 1e-3), "nn_pairs_per_sec": its * 2048.0 * 2048.0 / (best * 1e-3),
                  "all_poses_recovered": bool(ok), "mean_iterations": float((iters + 1).mean())}))
