"""Small K1T / small-kernel registrations for compute-sanitizer (memcheck, racecheck, synccheck): every brute-force pass goes through the tensor-core filter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ["ICPB_K1_FILTER_MIN_PAIRS"] = "0"
import numpy as np
import icp_b200 as ib
import icp_synth
with ib.Context(0) as ctx:
    for W in (37, 128):
        D, M = icp_synth.p2p_clouds(W)
        ctx.set_target(M); ctx.set_source(D)
        e, r = ctx.run(ib.default_params(max_iter=6, stop_early=0))
        print(W, r.iterations_run, float(e[r.iterations_run]), ctx.filter_tc_config(), ctx.filter_config()["dims_last"])
    D, M = icp_synth.p2p_clouds(32)
    ctx.set_target(M); ctx.set_source(D)
    e, r = ctx.run(ib.default_params(max_iter=6, stop_early=0, nn_method=ib.NN_BRUTE_DIRECT))
    print("small", r.iterations_run, float(e[r.iterations_run]))
