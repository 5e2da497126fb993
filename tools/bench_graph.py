# repeated registrations of one size: plain loop vs CUDA-graph replay (ICPB_FLAG_GRAPH)
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
for W in (32, 128):
    D, M = icp_synth.p2p_clouds(W)
    ctx = ib.Context(0); ctx.set_target(M)
    for name, flags in (("plain", 0), ("graph", ib.FLAG_GRAPH)):
        ts = []
        for rep in range(12):
            ctx.set_source(D)
            t0 = time.perf_counter(); e, r = ctx.run(ib.default_params(max_iter=40, flags=flags)); ts.append((time.perf_counter() - t0) * 1e3)
        print("W %3d %-5s: first run %.3f ms, then median %.3f ms (device elapsed %.3f ms), iterations %d" % (W, name, ts[0], float(np.median(ts[2:])), r.elapsed_ms, r.iterations_run))
    ctx.close()
