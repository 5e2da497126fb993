# whole 1M x 1M registration with the uniform-grid variant: rings + brute-force fallback vs occupancy pyramid (one GPU)
import os, sys, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
D, M = icp_synth.p2p_clouds(W)
out = {}
for name, env in (("pyramid", "1"), ("rings", "0")):
    os.environ["ICPB_GRID_PYRAMID"] = env
    ctx = ib.Context(0); ctx.set_target(M); ctx.set_source(D)
    err, res = ctx.run(ib.default_params(max_iter=64, nn_method=ib.NN_GRID, flags=ib.FLAG_PROFILE))
    out[name] = {"iterations_run": res.iterations_run, "elapsed_ms": res.elapsed_ms, "match_ms": res.match_ms, "icp_iters_per_sec": res.iterations_run / (res.elapsed_ms * 1e-3),
                 "final_rms": float(err[res.iterations + 1]), "grid_stats": ctx.grid_stats(), "errors": err[:res.iterations + 2].tolist()}
    ctx.close()
    if name == "pyramid" and len(sys.argv) > 2 and sys.argv[2] == "only":
        break
if "rings" in out:
    out["identical_trajectory"] = out["pyramid"]["errors"] == out["rings"]["errors"]
for k in out:
    if isinstance(out[k], dict): out[k].pop("errors", None)
print(json.dumps(out))
