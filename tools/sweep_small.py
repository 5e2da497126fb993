import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
for n in (1024, 4096, 16384, 65536, 262144):
    W = int(np.ceil(np.sqrt(n)))
    D, M = icp_synth.p2p_clouds(W, n)
    row = []
    c = ib.Context(0); c.set_target(M); c.set_source(D); c.match(0, ib.NN_BRUTE)
    _, f = c.time_match(0, ib.NN_BRUTE, reps=20); row.append("filter %.4f" % f); c.close()
    for cfg in (6, 2, 3, 8, 9, 11, 14):
        os.environ["ICPB_K1_CFG"] = str(cfg)
        c = ib.Context(0); c.set_target(M); c.set_source(D); c.match(0, ib.NN_BRUTE_DIRECT)
        _, d = c.time_match(0, ib.NN_BRUTE_DIRECT, reps=20); row.append("d%d %.4f" % (cfg, d)); c.close()
    print("n %7d ms: " % n + "  ".join(row), flush=True)
