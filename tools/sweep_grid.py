import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
n = 400000
W = int(np.ceil(np.sqrt(n)))
D, M = icp_synth.p2p_clouds(W, n)
for grid in (74, 148, 222, 296, 370, 444, 592, 1184, 2368):
    os.environ["ICPB_K1_GRID"] = str(grid)
    c = ib.Context(0); c.set_target(M); c.set_source(D)
    c.match(0, ib.NN_BRUTE)
    _, f = c.time_match(0, ib.NN_BRUTE, reps=3)
    _, d = c.time_match(0, ib.NN_BRUTE_DIRECT, reps=3)
    print("grid %5d : filter %.2f ms (%.3e pairs/s)   direct %.2f ms (%.3e pairs/s)" % (grid, f, float(n) * n / (f * 1e-3), d, float(n) * n / (d * 1e-3)), flush=True)
    c.close()
