"""Guided self-scheduling parameters of the filter kernel (ICPB_KF_GSS=min,max,div): matching time of the full 1M x 1M
pass and of a 1/8 shard, iterations 3-6 of a registration (warm planar bound)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib      # noqa: E402
import icp_dist            # noqa: E402
import icp_synth           # noqa: E402

D, M = icp_synth.p2p_clouds(1000)
for gss in ("2,64,4", "1,64,4", "4,64,4", "2,32,4", "2,128,4", "2,64,2", "2,64,8", "1,32,8", "2,256,2"):
    os.environ["ICPB_KF_GSS"] = gss
    row = []
    for G in (1, 8):
        ctx = ib.Context(0)
        ctx.set_target(M)
        ctx.set_source(np.ascontiguousarray(D[icp_dist.shard_indices(D.shape[0], 0, G)]))
        ms = [ctx.run(ib.default_params(max_iter=1, stop_early=0))[1].match_ms for _ in range(6)]
        ctx.close()
        row.append(sum(ms[2:]) / 4)
    print("GSS %-9s  full %.3f ms   1/8 shard %.3f ms" % (gss, row[0], row[1]))
