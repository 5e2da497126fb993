# exact uniform-grid NN vs brute force on the 1M x 1M registration (BASELINE config 4 size), one GPU
import os, sys, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
D, M = icp_synth.p2p_clouds(W)
ctx = ib.Context(0); ctx.set_target(M)
out = {}
for name, nn in (("grid", ib.NN_GRID), ("brute", ib.NN_BRUTE)):
    ctx.set_source(D)
    err, res = ctx.run(ib.default_params(max_iter=64, nn_method=nn))
    out[name] = {"iterations_run": res.iterations_run, "elapsed_ms": res.elapsed_ms, "match_ms": res.match_ms,
                 "equiv_pairs_per_sec": res.nn_pairs / (res.match_ms * 1e-3), "final_rms": float(err[res.iterations + 1]), "errors": err[:res.iterations + 2].tolist()}
    if nn == ib.NN_GRID: out["grid_stats"] = ctx.grid_stats()
out["identical_trajectory"] = out["grid"]["errors"] == out["brute"]["errors"]
for k in ("grid", "brute"): del out[k]["errors"]
print(json.dumps(out))
