#!/usr/bin/env python
"""Device times of the phases that are not brute-force matching: per-iteration K2 / K4 times at
1M and 100k points (ICPB_FLAG_PROFILE events), whole small registrations (1 024 / 16 384 points, the reference's own
sizes), the grid variant's per-iteration matching time along a 1M registration, and the k-NN + normals step with and
without the pyramid."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import numpy as np
import icp_b200 as ib
import icp_synth

out = {}
with ib.Context(0) as ctx:
    for name, (W, npts) in {"1M": (1000, None), "100k": (317, 100000), "16384": (128, None), "1024": (32, None)}.items():
        D, M = icp_synth.p2p_clouds(W, npts)
        ctx.set_target(M); ctx.set_source(D)
        ctx.run(ib.default_params(max_iter=3, stop_early=0, nn_method=ib.NN_GRID, sync_every=1))      # warm (grid build)
        ctx.set_source(D)
        e, r = ctx.run(ib.default_params(max_iter=64, nn_method=ib.NN_GRID, sync_every=1, flags=ib.FLAG_PROFILE))
        out["p2p_grid_" + name] = {"iterations_run": r.iterations_run, "elapsed_ms": r.elapsed_ms, "match_ms_per_it": r.match_ms / r.iterations_run,
                                   "minimize_us_per_it": 1e3 * r.minimize_ms / r.iterations_run, "transform_us_per_it": 1e3 * r.transform_ms / r.iterations_run}
        ctx.set_source(D)
        e, r = ctx.run(ib.default_params(max_iter=64, nn_method=ib.NN_GRID))
        out["p2p_grid_" + name]["elapsed_ms_default_sync"] = r.elapsed_ms
        if D.shape[0] <= 100000:
            for nn, tag in ((ib.NN_BRUTE, "brute"),):
                ctx.set_source(D)
                ctx.run(ib.default_params(max_iter=64, nn_method=nn))
                ctx.set_source(D)
                e, r = ctx.run(ib.default_params(max_iter=64, nn_method=nn))
                out["p2p_%s_%s" % (tag, name)] = {"iterations_run": r.iterations_run, "elapsed_ms": r.elapsed_ms, "us_per_iteration": 1e3 * r.elapsed_ms / r.iterations_run}
                ctx.set_source(D)
                e, r = ctx.run(ib.default_params(max_iter=64, nn_method=nn, sync_every=1, flags=ib.FLAG_PROFILE))
                out["p2p_%s_%s" % (tag, name)].update({"match_us_per_it": 1e3 * r.match_ms / r.iterations_run, "minimize_us_per_it": 1e3 * r.minimize_ms / r.iterations_run,
                                                       "transform_us_per_it": 1e3 * r.transform_ms / r.iterations_run, "elapsed_ms_profiled": r.elapsed_ms})
    D, M = icp_synth.p2p_clouds(317, 100000)
    ctx.set_target(M); ctx.set_source(D)
    ms = [ctx.estimate_normals(4) for _ in range(3)]
    nb = ctx.neighbors(4)
    out["normals_100k_bruteforce_knn_ms"] = min(ms)
os.environ["ICPB_KNN_PYRAMID"] = "1"
with ib.Context(0) as ctx:
    ctx.set_target(M); ctx.set_source(D)
    ms = [ctx.estimate_normals(4) for _ in range(3)]
    out["normals_100k_pyramid_knn_ms"] = min(ms)
    out["knn_lists_equal"] = bool(np.array_equal(nb, ctx.neighbors(4)))
    e, r = ctx.run(ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT, max_iter=50, flags=ib.FLAG_PROFILE, sync_every=1))
    out["p2plane_100k"] = {"iterations_run": r.iterations_run, "elapsed_ms": r.elapsed_ms, "minimize_us_per_it": 1e3 * r.minimize_ms / r.iterations_run,
                           "transform_us_per_it": 1e3 * r.transform_ms / r.iterations_run, "match_ms_per_it": r.match_ms / r.iterations_run}
print(json.dumps(out, indent=1))
