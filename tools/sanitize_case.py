# small ragged inputs through every kernel; meant to run under compute-sanitizer (memcheck / racecheck)
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
os.environ["ICPB_K1_FILTER_MIN_PAIRS"] = "0"          # force the filter kernel on these small clouds
import icp_b200 as ib, icp_synth
rng = np.random.default_rng(0)
Q = (rng.normal(size=(2777, 3))).astype(np.float32)
P = (rng.normal(size=(1333, 3)) * 1.1).astype(np.float32)
c = ib.Context(0)
c.set_target(Q); c.set_source(P)
res = {}
for mode in (0, 1, 2):
    res[("direct", mode)] = c.match(mode, ib.NN_BRUTE_DIRECT)
for mode in (0, 1):
    res[("filter", mode)] = c.match(mode, ib.NN_BRUTE)
    res[("filter2", mode)] = c.match(mode, ib.NN_BRUTE)
    res[("grid", mode)] = c.match(mode, ib.NN_GRID)
    assert np.array_equal(res[("filter", mode)], res[("direct", mode)]) and np.array_equal(res[("grid", mode)], res[("direct", mode)])
D, M = icp_synth.p2p_clouds(40)
c.set_target(M); c.set_source(D)
e, r = c.run(ib.default_params(max_iter=30)); print("p2p iterations", r.iterations_run)
c.set_source(D); e, r = c.run(ib.default_params(max_iter=30, nn_method=ib.NN_GRID)); print("p2p grid iterations", r.iterations_run)
c.set_source(D); e, r = c.run(ib.default_params(max_iter=6, flags=ib.FLAG_GRAPH)); print("p2p graph iterations", r.iterations_run)
c.set_source(D); c.estimate_normals(4)
e, r = c.run(ib.default_params(metric=ib.POINT_TO_PLANE, dist_mode=ib.DIST_SQRT, max_iter=30)); print("p2plane iterations", r.iterations_run)
c.set_source(D); e, r = c.run(ib.default_params(dist_mode=ib.DIST_STD, max_iter=5, stop_early=0)); print("std iterations", r.iterations_run)
S, T, rr, tt = icp_synth.batched_pairs(5, n=700)
out = c.run_batched(ib.default_params(max_iter=30), S, T); print("batched iterations", out[1])
idx, R, Tt, rms = c.iterate_host(ib.default_params(), D, M); print("iterate_host rms", rms)
print("fp32 peak", c.fp32_peak_tflops() > 1)
c.close()
print("sanitize_case done")
