# K1F (lower-bound filter + exact refine) vs K1 direct: full registration, bitwise-equal trajectory, timing per iteration
import os, sys, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else W * W
D, M = icp_synth.p2p_clouds(W, n)
ctx = ib.Context(0); ctx.set_target(M)
out = {}
for name, nn in (("filter", ib.NN_BRUTE), ("direct", ib.NN_BRUTE_DIRECT)):
    ctx.set_source(D)
    per_it = []
    errs = []
    for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 50):
        s0 = ctx.filter_stats()
        e, res = ctx.run(ib.default_params(max_iter=1, stop_early=0, nn_method=nn))
        s1 = ctx.filter_stats()
        frac = (s1["subtile_exact"] - s0["subtile_exact"]) / max(1.0, s1["subtile_tests"] - s0["subtile_tests"])
        per_it.append((round(res.match_ms, 3), round(frac, 4)))
        errs.append(float(e[1]))
    out[name] = {"match_ms_per_iteration(ms, exact fraction)": per_it, "total_match_ms": sum(p[0] for p in per_it)}
    out[name + "_errs"] = errs
    out[name + "_idx_identity"] = bool(np.array_equal(ctx.correspondences(), np.arange(n)))
out["identical_trajectory"] = out["filter_errs"] == out["direct_errs"]
del out["filter_errs"], out["direct_errs"]
print(json.dumps(out))
