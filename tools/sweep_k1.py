# sweep the K1 tuning table (ICPB_K1_CFG) on the GPU: time + index equality against config 0
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib, icp_synth
cfgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(16))
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [100000, 1000000]
for n in sizes:
    W = int(np.ceil(np.sqrt(n)))
    D, M = icp_synth.p2p_clouds(W, n)
    ref = {}
    for cfg in cfgs:
        os.environ["ICPB_K1_CFG"] = str(cfg)
        c = ib.Context(0); c.set_target(M); c.set_source(D)
        out = []
        for mode in (0, 1):
            mean, mn = c.time_match(mode, reps=3 if n > 300000 else 10)
            idx = c.correspondences()
            if mode not in ref: ref[mode] = idx
            ok = bool(np.array_equal(idx, ref[mode]))
            out.append("mode%d min %.3f ms %.3e pairs/s %s" % (mode, mn, float(n) * n / (mn * 1e-3), "OK" if ok else "MISMATCH"))
        print("n %8d cfg %2d : %s" % (n, cfg, " | ".join(out)), flush=True)
        c.close()
