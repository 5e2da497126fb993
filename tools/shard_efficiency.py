"""How efficient is one rank's share of an N-GPU step? Times the matching step of a 1/G shard of the 1M source cloud
against the full 1M target on ONE GPU (warm planar filter) and compares with 1/G of the full-cloud time; sweeps the
work-chunk size (ICPB_KF_CHUNK is read at context creation, so one context per value)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
import icp_b200 as ib      # noqa: E402
import icp_dist            # noqa: E402
import icp_synth           # noqa: E402

D, M = icp_synth.p2p_clouds(1000)


def run(G, chunk, interleaved=True, iters=6):
    if chunk:
        os.environ["ICPB_KF_CHUNK"] = str(chunk)
    else:
        os.environ.pop("ICPB_KF_CHUNK", None)
    ctx = ib.Context(0)
    ctx.set_target(M)
    shard = D[icp_dist.shard_indices(D.shape[0], 0, G)] if interleaved else D[:D.shape[0] // G]
    ctx.set_source(np.ascontiguousarray(shard))
    ms = []
    for it in range(iters):
        err, res = ctx.run(ib.default_params(max_iter=1, stop_early=0))
        ms.append(res.match_ms)
    cfg = ctx.filter_config()
    ctx.close()
    return ms, cfg


full, _ = run(1, 0)
print("G=1 full cloud: match ms per iteration", ["%.2f" % v for v in full])
for G in (2, 8):
    for chunk in (0, 4, 8, 16, 32, 64):
        ms, cfg = run(G, chunk)
        eff = [f / G / v for f, v in zip(full, ms)]
        print("G=%d chunk=%2d: %s  efficiency vs full/G: %s  (dims %d, exact %.2f%%)" % (
            G, chunk, ["%.2f" % v for v in ms[2:]], ["%.3f" % e for e in eff[2:]], cfg["dims_last"], 100 * cfg["last_exact_fraction"]))
