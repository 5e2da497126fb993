#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native ICP engine (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--width 1000]

Workload (BASELINE.json configs[3], the configuration its metric and target are quoted on; it fits one
GPU): synthetic z = x^2 - y^2 saddle, 1 000 000 source x 1 000 000 target points, point-to-point ICP with
exact brute-force nearest neighbours. A step = one ICP iteration (matching -> moments -> 3x3 SVD ->
transform -> error). With N GPUs the SOURCE is sharded over the ranks, the target replicated, and the 16
moment sums are exchanged inside the library (in the reduction kernels over NVLink peer memory; ICPB_PEER=0
selects ncclAllReduce launches instead): total work is fixed => "scaling": "strong".

One JSON line on stdout (rank 0):
  value      NN pairs/s, whole job, inputs resident in HBM, device time (CUDA events on the engine's
             stream around each step, max over ranks, summed over the K steps)
  e2e        same metric through the C-ABI call icpb_iterate_host with HOST buffers (pinned): H2D of both
             clouds, one iteration, D2H of the correspondences + transform, wall-clock
  roofline   the matching kernel: 8 FLOP per (source,target) pair / its CUDA-event duration, against the
             FP32 FFMA peak measured in the same process (icpb_measure_fp32_peak); nominal peak beside it
  cpu_baseline  the oracle's brute-force matching (OpenMP, all host cores) on a bounded sample (N=1 only)
--impl reference: the reference's CPU algorithm (oracle port of src/ICP_CPU.c / CPU_ICP_point_to_point.cpp
matching, all host threads) on bounded samples of the same workload; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))

METRIC = "nn_pairs_per_sec"
UNIT = "pairs/s"
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 * 1e-12    # SMs x lanes x 2 x max SM clock


def clocks_start(device_index):
    try:
        f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        p = subprocess.Popen(["nvidia-smi", "-i", str(device_index),
                              "--query-gpu=timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                              "--format=csv,noheader,nounits", "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        return p, f
    except Exception:
        return None, None


def _smi_time(text):
    """nvidia-smi's timestamp ("2026/10/18 11:19:44.901", local time) -> seconds since the epoch, or None."""
    import datetime
    try:
        return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
    except ValueError:
        return None


def clocks_stop(p, f, t_begin=None, t_end=None):
    """Median SM clock under load and the throttle reasons seen. The sampler is started BEFORE the warm-up (starting
    nvidia-smi initialises NVML, which stalls kernel launches on the box for tens of milliseconds — inside a timed
    region of 10 x 12 ms steps that is a 25 % error); only the samples taken between t_begin and t_end (wall clock of
    the timed region, +- one sampling period) are used when the timestamps parse."""
    if p is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    f.flush(); f.seek(0)
    rows = []
    for line in f.read().splitlines():
        c = [x.strip() for x in line.split(",")]
        if len(c) < 9:
            continue
        try:
            rows.append((_smi_time(c[0]), float(c[1]), float(c[2]), c[5:9]))
        except ValueError:
            continue
    inside = [r for r in rows if r[0] is not None and t_begin is not None and t_begin - 0.25 <= r[0] <= t_end + 0.25]
    window = "timed region" if inside else "warm-up + timed region"
    sm, mx, reasons = [], [], set()
    for _, a, b, flags in (inside or rows):
        sm.append(a); mx.append(b)
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
            if v.lower().startswith("active"):
                reasons.add(name)
    f.close()
    try:
        os.unlink(f.name)
    except OSError:
        pass
    busy = [s for s in sm if s > 500] or sm
    return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_match_rate(orc, D, M, seconds_target, mode=0):
    """Oracle brute-force matching on a bounded slice of the sources against the FULL target."""
    n = D.shape[0]
    probe = min(n, 256)
    t0 = time.perf_counter(); orc.match(D[:probe], M, mode); dt = time.perf_counter() - t0
    rate = probe * M.shape[0] / max(dt, 1e-9)
    want = int(min(n, max(probe, rate * seconds_target / M.shape[0])))
    want = max(256, (want // 256) * 256)
    sel = np.ascontiguousarray(D[:: max(1, n // want)][:want])
    t0 = time.perf_counter(); orc.match(sel, M, mode); dt = time.perf_counter() - t0
    return sel.shape[0] * M.shape[0] / dt, sel.shape[0], dt


def run_reference_arm(args, rank, emit):
    """The reference's CPU implementation of the path, on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    import icp_synth
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm would then be timed on ONE thread (r1: SCALE ratios at
    # N > 1 were against a single-threaded baseline). Use every processor OpenMP sees, and record the count in effect.
    cores = orc.set_num_threads(0)
    D, M = icp_synth.p2p_clouds(args.width)
    n = D.shape[0]
    # each step = one bounded sample: S sources x all targets, S sized for ~3 s of CPU work
    rate0, _, _ = cpu_match_rate(orc, D, M, 1.0)
    S = int(max(256, min(n, (rate0 * 3.0 / M.shape[0]) // 256 * 256)))
    sel = np.ascontiguousarray(D[:: max(1, n // S)][:S])
    idx_prev = None
    for _ in range(args.warmup):
        orc.match(sel, M, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        idx = orc.match(sel, M, 0)
        mom = orc.moments(sel, M, idx)
        R, T = orc.rt_from_moments(mom)
        idx_prev = idx
    dt = time.perf_counter() - t0
    pairs = float(S) * M.shape[0] * args.steps
    value = pairs / dt
    cores = orc.num_threads()
    if cores == 1 and (os.cpu_count() or 1) > 1:
        sys.stderr.write("bench.py: WARNING: the CPU arm ran on 1 thread of %d processors\n" % os.cpu_count())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic z=x^2-y^2, %dx%d points, point-to-point ICP, exact brute-force NN" % (n, M.shape[0]),
                   "width": args.width, "step": "one ICP iteration (CPU: matching on a bounded source sample + moments + SVD)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of %d sources x all %d targets per step, oracle/icp_oracle.c orc_match_f32 "
                                   "(restates src/ICP_point_to_point.cu:31-57 = the float twin of src/ICP_CPU.c:220-234), OpenMP %d threads"
                                   % (S, n, M.shape[0], cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "icp_iters_per_sec_extrapolated": value / (float(n) * M.shape[0]),
    }
    # the unmodified reference program (config 0: 100x100 points, double, 1 thread) for context
    exe = os.path.join(ROOT, "oracle", "_ref", "icp_cpu")
    if os.path.exists(exe) and not args.skip_ref_binary:
        try:
            t0 = time.perf_counter()
            out = subprocess.run([exe], capture_output=True, text=True, timeout=300).stdout
            wall = time.perf_counter() - t0
            import re
            m = re.search(r"computed in ([\d.]+) ms with (\d+) iterations", out)
            if m:
                ms, its = float(m.group(1)), int(m.group(2))
                line["reference_binary"] = {"program": "src/ICP_CPU.c (unmodified, MKL shim, 1 thread, double)", "points": 10000,
                                            "iterations": its, "ms": ms, "nn_pairs_per_sec": (its + 1) * 1e8 / (ms * 1e-3), "wall_s": wall}
        except Exception as e:      # noqa: BLE001
            line["reference_binary"] = {"error": str(e)}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1000, help="grid width W (W*W points per cloud); 1000 = BASELINE config 4")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-ref-binary", action="store_true")
    ap.add_argument("--skip-grid", action="store_true", help="do not run the extra whole-registration measurement with the exact grid variant (N=1)")
    ap.add_argument("--balance", type=int, default=0, help="N > 1: 1 = deal source blocks in proportion to each GPU's measured matching rate (two extra untimed steps); 0 = even deal (B200s of one box measured within +-1.2 %, so this is off by default)")
    ap.add_argument("--shard", default="interleaved", choices=["interleaved", "contiguous"],
                    help="how the source is dealt to the ranks (N > 1): blocks of 2048 points round-robin, or contiguous ranges")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # Libraries (NCCL's "NCCL version ..." banner, torchrun chatter) write to fd 1; the contract is ONE JSON line on
    # stdout, so everything else is sent to stderr and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, emit)
        return

    import torch
    import icp_b200 as ib
    import icp_dist
    import icp_synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        nccl_id = icp_dist.broadcast_bytes(ib.nccl_unique_id() if rank == 0 else None, 128)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = ib.Context(local_rank, rank, world, nccl_id)
    clk_p, clk_f = clocks_start(local_rank) if rank == 0 else (None, None)     # seconds before the timed region, see clocks_stop
    D, M = icp_synth.p2p_clouds(args.width)
    n_total, m = D.shape[0], M.shape[0]
    if args.shard == "contiguous":
        lo, hi = icp_dist.shard_bounds(n_total, rank, world)
        shard_index = np.arange(lo, hi, dtype=np.int64)
        shard = np.ascontiguousarray(D[lo:hi])
    else:                                   # blocks of 2048 sources dealt round-robin: balances the ranks' matching cost
        shard_index = icp_dist.shard_indices(n_total, rank, world)
        shard = np.ascontiguousarray(D[shard_index])
    n_rank = shard.shape[0]

    fp32_peak = ctx.fp32_peak_tflops()
    ctx.set_target(M)
    ctx.set_source(shard)
    # the direct kernel (reference chain on every pair, no pruning) for comparison, outside the timed region
    _, direct_ms = ctx.time_match(ib.DIST_SQ, ib.NN_BRUTE_DIRECT, reps=2)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    step_errs_all = []                      # RMS after every step since the last source upload (warm-up included)

    def one_step():
        flush_buf.fill_(1)
        torch.cuda.synchronize()
        err, res = ctx.run(ib.default_params(max_iter=1, stop_early=0))
        step_errs_all.append(float(err[1]))
        return res

    # N > 1: the step ends when the slowest rank ends, and B200s of one box differ by a few per cent in sustained speed.
    # Two untimed calibration steps measure every rank's matching rate; the blocks are then dealt in proportion to it
    # (icp_dist.deal_blocks) and the registration restarts from the original cloud. --balance 0 keeps the even deal.
    balance_weights = None
    if world > 1 and args.balance and args.shard != "contiguous":
        one_step(); r1 = one_step(); r2 = one_step()
        speed = torch.zeros(world, dtype=torch.float64, device="cuda")
        speed[rank] = float(n_rank) / max(1e-6, r1.match_ms + r2.match_ms)
        dist.all_reduce(speed)
        balance_weights = [float(v) for v in speed.cpu().tolist()]
        shard_index = icp_dist.shard_indices_weighted(n_total, rank, balance_weights)
        shard = np.ascontiguousarray(D[shard_index])
        n_rank = shard.shape[0]
        ctx.set_source(shard)
        step_errs_all.clear()
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = ctx.launch_count()
    step_ms, match_ms = [], []
    t_epoch0 = time.time()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        res = one_step()
        step_ms.append(res.elapsed_ms); match_ms.append(res.match_ms)
    last_res = res
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    clocks = clocks_stop(clk_p, clk_f, t_epoch0, time.time()) if rank == 0 else None

    # ---- parity carried by the bench line: the state after the W + K steps — every source point's
    # correspondence, the last transform, the error trajectory — is checksummed, and at N > 1 rank 0 replays the same
    # steps on ONE GPU (its own, full source) and the run FAILS unless the sharded run gave the same correspondences bit
    # for bit and the same errors / transform to FP64 summation-order noise.
    import zlib
    idx_rank = ctx.correspondences()
    step_errs = np.array(step_errs_all, dtype=np.float32)
    parity = {"steps_checked": len(step_errs_all)}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (shard_index.astype(np.int64), idx_rank))
        if rank == 0:
            idx_all = np.full(n_total, -1, np.int32)
            for sidx, sval in gathered:
                idx_all[sidx] = sval
            with ib.Context(local_rank) as one:
                one.set_target(M); one.set_source(D)
                errs1 = []
                for _ in range(len(step_errs_all)):
                    e1, r1 = one.run(ib.default_params(max_iter=1, stop_early=0))
                    errs1.append(float(e1[1]))
                idx1 = one.correspondences()
                last_R1 = np.array(r1.last_R[:], np.float64)
            errs1 = np.array(errs1, np.float32)
            parity.update({"one_gpu_replay": True, "idx_equal": bool(np.array_equal(idx_all, idx1)),
                           "max_rel_err_diff": float(np.max(np.abs(step_errs - errs1) / np.maximum(np.abs(errs1), 1e-12))),
                           "max_R_diff": float(np.max(np.abs(np.array(last_res.last_R[:], np.float64) - last_R1))),
                           "idx_crc32": int(zlib.crc32(idx_all.tobytes())), "idx_crc32_one_gpu": int(zlib.crc32(idx1.tobytes()))})
            if not (parity["idx_equal"] and parity["max_rel_err_diff"] <= 1e-5 and parity["max_R_diff"] <= 1e-6):
                raise SystemExit("bench.py: the %d-GPU run does not reproduce the 1-GPU run: %s" % (world, json.dumps(parity)))
    else:
        parity.update({"one_gpu_replay": False, "idx_crc32": int(zlib.crc32(idx_rank.tobytes()))})
    parity["errors_crc32"] = int(zlib.crc32(step_errs.tobytes()))
    parity["final_rms"] = float(step_errs[-1])

    total_ms = allmax(sum(step_ms))
    match_total_ms = allmax(sum(match_ms))
    match_fastest_rank_ms = -allmax(-sum(match_ms))
    launches_all = int(allsum(launches))
    pairs_total = float(n_total) * m * args.steps
    value = pairs_total / (total_ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI entry point ------------------------------------------
    hD = torch.from_numpy(shard).pin_memory()
    hM = torch.from_numpy(M).pin_memory()
    hD_np, hM_np = hD.numpy(), hM.numpy()
    # every host buffer of the loop is pinned: the clouds that go up and the results that come down
    h_cur = [torch.empty_like(hD).pin_memory(), torch.empty_like(hD).pin_memory()]
    h_idx = torch.empty(shard.shape[0], dtype=torch.int32).pin_memory()
    p = ib.default_params()
    # a genuine host-driven loop: every step uploads the CURRENT source and the target from pinned host memory, runs one
    # iteration, and downloads the correspondences, the transform and the transformed source (which feeds the next step)
    idx, R, T, rms = ctx.iterate_host(p, hD_np, hM_np, idx_out=h_idx.numpy())            # warm
    cur = ctx.get_source(out=h_cur[0].numpy())
    barrier()
    t0 = time.perf_counter()
    for k in range(args.e2e_steps):
        idx, R, T, rms = ctx.iterate_host(p, cur, hM_np, idx_out=h_idx.numpy())
        cur = ctx.get_source(out=h_cur[(k + 1) & 1].numpy())
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = float(n_total) * m * args.e2e_steps / e2e_s
    h2d = allsum(float(shard.nbytes + M.nbytes))
    d2h = allsum(float(idx.nbytes + R.nbytes + T.nbytes + 4 + cur.nbytes))

    # per-GPU roofline of the dominant kernel (brute-force matching)
    pairs_rank = float(n_rank) * m * args.steps
    achieved = 8.0 * pairs_rank / (sum(match_ms) * 1e-3) * 1e-12
    achieved = allmax(-achieved) * -1.0 if world > 1 else achieved      # the slowest rank's figure

    fcfg = ctx.filter_config()
    # Extra, outside every timed region and never allowed to disturb the contract line: the same registration from its
    # initial pose to convergence with the exact uniform-grid variant (ICPB_NN_GRID: occupancy pyramid over the cells,
    # identical correspondences), device time from the engine's events. One GPU only.
    grid_extra = None
    if world == 1 and not args.skip_grid:
        try:
            ctx.set_source(shard)
            ctx.run(ib.default_params(max_iter=64, nn_method=ib.NN_GRID))        # builds the grid, warms up
            ctx.set_source(shard)
            g_err, g_res = ctx.run(ib.default_params(max_iter=64, nn_method=ib.NN_GRID))
            grid_extra = {"what": "whole registration of the same clouds with the exact uniform-grid variant (ICPB_NN_GRID), initial pose to convergence",
                          "iterations_run": int(g_res.iterations_run), "elapsed_ms": float(g_res.elapsed_ms),
                          "icp_iters_per_sec": float(g_res.iterations_run) / (float(g_res.elapsed_ms) * 1e-3),
                          "final_rms": float(g_err[g_res.iterations + 1])}
        except Exception as exc:      # noqa: BLE001
            grid_extra = {"error": str(exc)[:200]}
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json")))
        traffic = tj.get("k1_filter", {}).get(str(args.width))
    except Exception:      # noqa: BLE001
        pass
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic z=x^2-y^2, %dx%d points, point-to-point ICP, exact brute-force NN (BASELINE configs[3])" % (n_total, m),
                       "width": args.width, "step": "one ICP iteration: matching + moments + 3x3 SVD + transform + error",
                       "parallelism": "source sharded x%d (%s), target replicated, 16 FP64 moment sums %s" % (
                           world, "contiguous shards" if args.shard == "contiguous" else "blocks of 2048 points dealt round-robin", "exchanged inside the reduction kernels over NVLink peer memory (no collective launch)" if ctx.dist_info()["peer_exchange"]
                           else ("combined with ncclAllReduce" if world > 1 else "(single GPU: no exchange)")),
                       "l2": "256 MiB device write between timed steps (untimed); timing = CUDA events per step on the engine stream"},
            "icp_iters_per_sec": args.steps / (total_ms * 1e-3),
            "exact_grid_registration": grid_extra,
            "match_ms_per_step": match_total_ms / args.steps,
            "match_ms_per_step_fastest_rank": match_fastest_rank_ms / args.steps,
            "rank_balance": ({"weights": [w / (sum(balance_weights) / world) for w in balance_weights],
                              "how": "2 untimed calibration steps; blocks of 2048 sources dealt in proportion to each rank's measured matching rate"}
                             if balance_weights else None),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "icpb_iterate_host (pinned host clouds -> idx, R, T, rms) + icpb_get_source; host-driven loop, each step uploads the previous step's transformed source", "steps": args.e2e_steps},
            "gpu_launches": launches_all,
            "parity": parity,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "peak_source": "FFMA microbenchmark measured in this process (icpb_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry",
                         "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_nominal": achieved / NOMINAL_FP32_TFLOPS,
                         "kernel": "k1_filter (brute-force NN: a %d-FMA lower bound on every pair + the reference's chain on the sub-tiles it cannot exclude)" % fcfg["dims_last"],
                         "flop_per_pair": 8, "filter": fcfg,
                         "executed": {"fma_per_pair": fcfg["dims_last"], "min3_per_pair": 0.5,
                                      "fma_pipe_frac": (achieved / 8.0) * fcfg["dims_last"] * 2.0 / fp32_peak,
                                      "note": "FP32 work the filter actually executes per pair (exact re-evaluations of the few unexcluded sub-tiles not counted): "
                                              "FMAs on the FMA pipe, half a 3-input min on the 16-lane ALU pipe; ncu shows the two pipes never overlapping in this loop"},
                         "traffic": traffic,
                         "traffic_source": "profiles/r01_k1_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                         "note": "compute-bound (FP32 issue slots); `achieved` counts the ALGORITHMIC 8 FLOP per pair SURVEY.md 8(d) defines, so frac can exceed 1: "
                                 "the direct form executes 6 FP32 ops per pair (ceiling 66.7% of FFMA peak), the filter 3 (full bound) or 2 (planar bound) FMAs per pair"},
            "roofline_direct_kernel": {"kernel": "k1_match (reference chain on every pair)", "ms_per_launch": direct_ms,
                                       "achieved": 8.0 * float(n_rank) * m / (direct_ms * 1e-3) * 1e-12, "unit": "TFLOP/s",
                                       "frac": 8.0 * float(n_rank) * m / (direct_ms * 1e-3) * 1e-12 / fp32_peak},
        }
        if world == 1 and not args.skip_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as orc
            orc.set_num_threads(0)             # all host cores, whatever OMP_NUM_THREADS says
            rate, S, dt = cpu_match_rate(orc, D, M, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                    "sample": "%d of %d sources x all %d targets, %.1f s, oracle orc_match_f32 (OpenMP)" % (S, n_total, m, dt)}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
