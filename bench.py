#!/usr/bin/env python
"""bench.py -- ICP registrations/s on synthetic KITTI-shaped 120k-point scans.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--no-extras] [--pairs-per-gpu P] [--map-points-per-gpu M]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over
one scan: index build of the incoming scan + one whole registration against
the previous scan (BASELINE.json config C2, raw 120,000-point scans, the three
shipped YAML files).  `value` times that with the scans already resident in
HBM; `e2e` times the same steps through the LidarOdometry module with pinned
HOST buffers (H2D of the scan and D2H of the result inside the timed region).
N > 1 runs one independent sequence per GPU (weak scaling, no data-path
collective: SURVEY 8e "independent sequences: replicas").
The same line carries, outside the timed region of `value` (each with its own
CUDA-event timing, max over ranks):
  `knn`          kNN queries/s of the search kernel (BASELINE metric, part 2),
  `batch_lc`     config C4's shape: loop-closure candidate pairs x 10
                 Monte-Carlo guesses, pairs sharded i -> rank i mod N,
  `sharded_knn`  config C5's shape: a 128-beam scan against a map split by
                 spatial cell over the N GPUs, partial arg-min keys merged over
                 NCCL (all-reduce MIN for k=1; all-gather + k-way merge kernel
                 for k=6).
`--impl reference` times the CPU oracle port (the reference's own ICP cannot
be built here) on all host threads, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SCANS = 10          # synthetic scans generated per rank (set from --steps/--warmup in main(): a forward-only
                      # sequence when it fits, else walked back and forth)
# search kernel of the matcher, per query: 16 B query + 6 x 16 B neighbour coordinates read
# + 6 x 4 B neighbour indices written (SURVEY 8d's kNN figure without the d2 output)
ALGO_BYTES_PER_QUERY = 136
KNN_ALGO_BYTES = {1: 16 + 16 + 8, 6: 16 + 6 * 16 + 6 * 8}  # packed 8-byte keys out


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def scan_index(step):
    """0,1,..,n-1,n-2,..,0,1,..: consecutive steps always use adjacent scans."""
    period = 2 * (N_SCANS - 1)
    k = step % period
    return k if k < N_SCANS else period - k


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a
    thread (5 ms period), `nvidia-smi` as the fallback."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
           "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index, period_s=0.005):
        self.idx = gpu_index
        self.period = period_s
        self.rows, self.stop_flag, self.t = [], False, None
        self.nv, self.h, self.max_mhz = None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    mhz = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    try:
                        rs = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        rs = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((time.time(), mhz, rs))
                else:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=clocks.sm,clocks.max.sm",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    c = [x.strip() for x in o.stdout.strip().split(",")]
                    self.max_mhz = float(c[1])
                    self.rows.append((time.time(), float(c[0]), 0))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=2)
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        scope = "timed region"
        if not rows:  # a timed region shorter than one sampling period
            rows, scope = self.rows, "whole run (timed region shorter than the sampling period)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["clock sampling unavailable"], "samples": 0}
        reasons = sorted({n for n, bit in self.BAD.items() for r in rows if r[2] & bit})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(rows), "scope": scope,
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def next_guess(step, pose):
    """Constant-velocity initial guess for step+1 from the result of `step`
    with equal time steps: x, y, z, yaw only (LidarOdometry.cpp:272-275,
    305-308); identity when the back-and-forth walk over the scans turns round."""
    back = scan_index(step + 2) < scan_index(step + 1)
    was_back = scan_index(step + 1) < scan_index(step)
    if back != was_back:
        return np.zeros(6)
    return np.array([pose[0], pose[1], pose[2], pose[3], 0.0, 0.0])


def workload_config(n_pts):
    """The SAME dict in both arms (`--impl b200` and `--impl reference`)."""
    return {"workload": "c2_kitti64_120k_scan_to_scan_odometry", "points_per_scan": int(n_pts),
            "icp_settings": "icp-settings-regular.yaml via kitti-default.yaml",
            "guess": "constant-velocity from the previous registration of the same sequence "
                     "(LidarOdometry.cpp:272-275); the first registration of a sequence (identity guess) is warm-up",
            "sequence": "synthetic 64-beam, 1.0 m / 0.29 deg per scan, seed 1 + rank; forward-only when "
                        "warmup + steps + 3 <= 48 scans, else walked back and forth",
            "l2": "inputs (2 x 1.9 MB float4 + index) are L2-resident by nature; each step "
                  "registers a different scan pair, no flush"}


def make_scans(seed):
    from mola_fe_lidar_b200 import scene
    t = time.time()
    scans, poses = scene.make_sequence(N_SCANS, seed=seed)
    log(f"[bench] generated {N_SCANS} synthetic 64-beam scans of {len(scans[0])} pts in {time.time() - t:.1f}s")
    return scans, poses


# --------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """CPU oracle port on all host threads.  Every thread walks the SAME kind of
    sequence as the GPU arm -- consecutive scan pairs, each registration started
    from the constant-velocity guess of its own previous result -- staggered by
    one scan per thread; a step = one registration per thread."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_api as O
    O.build()
    scans, _ = make_scans(1)
    threads = max(1, os.cpu_count() or 1)
    clouds = [O.Cloud(s) for s in scans]
    prm = O.default_params()
    for c in clouds:  # kd-trees are built lazily on first query (as in MRPT): warm them once
        O.knn(c, scans[0][:8], 1, 0.49, kdtree=True)
    guesses = [np.zeros(6) for _ in range(threads)]
    iters = []

    def one_step(step, record):
        def work(t):
            s_t = step + t  # this thread's position in its own walk over the scans
            r = O.icp_align(clouds[scan_index(s_t)], clouds[scan_index(s_t + 1)], guesses[t], prm, kdtree=True)
            guesses[t] = next_guess(s_t, r["pose"])
            if record:
                iters.append(r["n_iterations"] + 1)
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        [x.start() for x in th]
        [x.join() for x in th]

    warm = max(args.warmup, 1)  # the first registration of every thread has no velocity guess yet
    for w in range(warm):
        one_step(w, False)
    t0 = time.time()
    for s in range(args.steps):
        one_step(warm + s, True)
    dt = time.time() - t0
    regs = args.steps * threads
    value = regs / dt
    out = {
        "impl": "reference", "metric": "icp_registrations_per_sec", "value": value,
        "unit": "registrations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 search / f64 solve", "data": "synthetic",
        "config": workload_config(len(scans[0])),
        "mean_outer_iterations": float(np.mean(iters)) if iters else None,
        "timing": "host clock, CPU only",
        "cpu_baseline": {"value": value, "unit": "registrations/s", "cores": threads, "kind": "port",
                         "sample": f"{regs} registrations of consecutive 120k-pt scans from constant-velocity guesses, "
                                   f"one per thread per step (oracle/icp_oracle.c, kd-tree leaf 10); the "
                                   f"reference's own ICP (mp2p_icp/MRPT) is not buildable here"},
        "e2e": {"value": value, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- GPU arm
def timed(torch, fn):
    """fn() between two CUDA events on the legacy default stream (the library's
    blocking streams order against it); returns milliseconds."""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def world_points(scans, poses):
    """The scans in the world frame (float32), for the map workloads."""
    out = []
    for s, T in zip(scans, poses):
        out.append((s.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    return out


def knn_microbench(torch, icp, scans, poses, dev, peak):
    """kNN queries/s of the search kernel on device-resident clouds: the
    BASELINE kNN metric.  Kernel time = the library's CUDA events around the
    launch; inputs larger than L2 are not possible for 120k-pt clouds, so each
    repetition queries a different scan pair (no flush)."""
    out = []
    wp = world_points(scans, poses)
    clouds = [icp.upload(w) for w in wp[:10]]
    # C3's local map: every scan of this run in one frame, merged by voxel -- the coarsest of these
    # resolutions that leaves at least 1,000,000 points (BASELINE config 3: "1M-pt local map")
    big_raw = icp.upload(np.concatenate(wp))
    big, map_res = None, None
    for res in (0.1, 0.07, 0.05, 0.035, 0.025, 0.015):
        cand = icp.voxel_decimate(big_raw, res)
        if big is not None:
            big.free()
        big, map_res = cand, res
        if len(big) >= 1_000_000:
            break
    big_raw.free()
    log(f"[bench] C3 map: {len(wp)} scans merged at {map_res} m -> {len(big)} points")
    cases = [("120k_vs_120k", 1, clouds, None), ("120k_vs_120k", 6, clouds, None),
             ("120k_vs_1M_map", 6, clouds, big)]
    icp.profile_enable(True)
    for name, k, cl, ref in cases:
        nq = len(cl[0])
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        reps = 6
        for w in range(2):
            icp.knn_keys_device(ref or cl[w], cl[w + 1], k, 0.7, keys.data_ptr())
        icp.profile_reset()
        for r in range(reps):
            icp.knn_keys_device(ref or cl[r % 8], cl[r % 8 + 1], k, 0.7, keys.data_ptr())
        pr = icp.profile()
        ms = pr["knn_ms"] / max(pr["knn_launches"], 1)
        qps = nq / (ms * 1e-3)
        gbs = qps * KNN_ALGO_BYTES[k] / 1e9
        out.append({"case": name, "n_queries": nq, "n_ref": len(ref) if ref else len(cl[0]), "k": k,
                    "radius_m": 0.7, "kernel_ms": ms, "queries_per_s": qps,
                    "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak})
    # uncapped (max_dist = +inf): the capped pass at the index radius plus the completion of the incomplete rows
    # over the whole cloud (b200icp_knn, results to the host; the kernel time is the library's events around both)
    try:
        for k in (1, 6):
            icp.knn(clouds[0], clouds[1], k, float("inf"))
            icp.profile_reset()
            for r in range(3):
                idx_u, _ = icp.knn(clouds[r], clouds[r + 1], k, float("inf"))
            pr = icp.profile()
            ms = pr["knn_ms"] / max(pr["knn_launches"], 1)
            qps = len(clouds[1]) / (ms * 1e-3)
            out.append({"case": "120k_vs_120k_uncapped", "n_queries": len(clouds[1]), "n_ref": len(clouds[0]), "k": k,
                        "radius_m": None, "kernel_ms": ms, "queries_per_s": qps,
                        "algorithmic_GBps": qps * KNN_ALGO_BYTES[k] / 1e9,
                        "frac_of_hbm_peak": qps * KNN_ALGO_BYTES[k] / 1e9 / peak,
                        "rows_without_k_neighbours": int((idx_u[:, k - 1] == 0xFFFFFFFF).sum())})
    except Exception as e:
        out.append({"case": "120k_vs_120k_uncapped", "error": f"{type(e).__name__}: {e}"[:200]})
    # the same kernel on points uniform in a box (SURVEY 8d asks for both distributions): 120k x 120k in
    # 60 x 60 x 6 m, about 8 points within the 0.7 m radius of a query
    try:
        rngu = np.random.default_rng(3)
        box = np.float32([60.0, 60.0, 6.0])
        ua = icp.upload((rngu.random((120000, 3), dtype=np.float32) * box))
        ub = icp.upload((rngu.random((120000, 3), dtype=np.float32) * box))
        for k in (1, 6):
            keys = torch.empty((120000, k), dtype=torch.int64, device=dev)
            for w in range(2):
                icp.knn_keys_device(ua, ub, k, 0.7, keys.data_ptr())
            icp.profile_reset()
            for r in range(6):
                icp.knn_keys_device(ua, ub, k, 0.7, keys.data_ptr())
            pr = icp.profile()
            ms = pr["knn_ms"] / max(pr["knn_launches"], 1)
            qps = 120000 / (ms * 1e-3)
            gbs = qps * KNN_ALGO_BYTES[k] / 1e9
            out.append({"case": "uniform_box_120k_vs_120k", "n_queries": 120000, "n_ref": 120000, "k": k,
                        "radius_m": 0.7, "kernel_ms": ms, "queries_per_s": qps, "algorithmic_GBps": gbs,
                        "frac_of_hbm_peak": gbs / peak})
        ua.free(), ub.free()
    except Exception as e:
        out.append({"case": "uniform_box_120k_vs_120k", "error": f"{type(e).__name__}: {e}"[:200]})
    icp.profile_enable(False)
    # config C3's shape: scan-to-map registration, raw 120k-pt scans (sensor frame) against the merged map
    from mola_fe_lidar_b200 import scene
    rng = np.random.default_rng(17)
    sens = [icp.upload(s) for s in scans[:6]]
    errs, iters = [], 0

    def run_c3():
        nonlocal iters
        for i, c in enumerate(sens):
            truth = scene.matrix_to_pose6(poses[i])
            guess = truth + np.r_[rng.normal(0, 0.15, 3), rng.normal(0, np.deg2rad(0.3), 1), 0.0, 0.0]
            r = icp.align(big, c, guess)
            errs.append(float(np.abs(r["pose"][:3] - truth[:3]).max()))
            iters += r["n_iterations"] + 1

    run_c3()
    errs, iters = [], 0
    ms, _ = timed(torch, run_c3)
    c3 = {"workload": "c3_scan_to_map", "map_points": len(big), "scan_points": len(sens[0]), "registrations": len(sens),
          "ms_per_registration": ms / len(sens), "registrations_per_s": len(sens) / (ms * 1e-3),
          "mean_matcher_runs": iters / len(sens), "max_abs_translation_error_m": max(errs),
          "map": f"the {len(wp)} scans of this run in one frame, merged at {map_res} m (b200icp_voxel_decimate)",
          "guess": "true pose + N(0, 0.15 m) / N(0, 0.3 deg yaw)"}
    for c in clouds + sens:
        c.free()
    big.free()
    return out, c3


def batch_lc(torch, dist, capi, lidar_odometry, scans, rank, world, local_rank, dev, n_pairs, mc=10):
    """Config C4: `n_pairs` candidate pairs (1024) of 20k-pt clouds x `mc`
    Monte-Carlo guesses (sigma 3 m / 2 deg, LidarOdometry.cpp:768-769) with
    icp-settings-loop-closure.yaml; the SAME pairs at every N, pair i -> rank
    i mod world (strong scaling), results gathered with one all-gather."""
    from mola_fe_lidar_b200 import multi_gpu as M
    yaml_txt = open(os.path.join(lidar_odometry.PARAMS_DIR, "icp-settings-loop-closure.yaml")).read()
    icp = capi.ICP(yaml_text=yaml_txt, device=local_rank)
    rng = np.random.default_rng(99)  # same on every rank
    sub = []
    for s in scans:
        sel = np.sort(rng.choice(len(s), size=20000, replace=False))
        sub.append(s[sel])
    pair_scans = [(int(rng.integers(0, len(sub) - 1)),) for _ in range(n_pairs)]
    guesses = np.zeros((n_pairs, mc, 6))
    guesses[:, :, :3] = rng.normal(0.0, 3.0, size=(n_pairs, mc, 3))
    guesses[:, :, 3] = rng.normal(0.0, np.deg2rad(2.0), size=(n_pairs, mc))
    guesses[:, :, 0] += 1.0  # consecutive scans are ~1 m apart
    mine = M.pairs_of_rank(n_pairs, rank, world)
    clouds = [icp.upload(x) for x in sub]  # each rank indexes the clouds once
    fr, to, gs = [], [], []
    for p in mine:
        i = pair_scans[p][0]
        fr += [clouds[i]] * mc
        to += [clouds[i + 1]] * mc
        gs.append(guesses[p])
    gs = np.concatenate(gs) if gs else np.zeros((0, 6))

    def run():
        res = icp.align_batch(fr, to, gs) if fr else []
        rec = np.zeros((len(mine), 9))
        for j in range(len(mine)):
            best = max(res[j * mc:(j + 1) * mc], key=lambda r: r["quality"])  # cpp:785-786
            rec[j, :6], rec[j, 6], rec[j, 7], rec[j, 8] = best["pose"], best["quality"], best["n_iterations"], \
                sum(r["n_iterations"] + 1 for r in res[j * mc:(j + 1) * mc])
        return M.gather_pair_results(rec, n_pairs, rank, world, dist, device=dev)

    run()  # warm-up (workspace growth)
    if dist is not None:
        dist.barrier()
    ms, allrec = timed(torch, run)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    for c in clouds:
        c.free()
    icp.close()
    return {"workload": "c4_loop_closure_candidates", "pairs": n_pairs, "montecarlo_samples": mc,
            "points_per_cloud": 20000, "registrations": n_pairs * mc, "ms": ms,
            "registrations_per_s": n_pairs * mc / (ms * 1e-3),
            "outer_iterations_total": float(allrec[:, 8].sum()), "best_quality_mean": float(allrec[:, 6].mean()),
            "sharding": "pair i -> rank i mod N; MC samples of a pair on one rank; one all-gather of results",
            "scaling": "strong"}


def sharded_knn(torch, dist, capi, scans, poses, rank, world, local_rank, dev, map_per_gpu):
    """Config C5's shape: one 128-beam scan (260,096 queries) against a map of
    map_per_gpu x N points split by spatial cell over the N GPUs."""
    from mola_fe_lidar_b200 import multi_gpu as M, scene
    icp = capi.ICP(capi.default_params(), device=local_rank)
    wp = np.concatenate(world_points(scans, poses))
    rng = np.random.default_rng(5)  # same map on every rank
    total = map_per_gpu * world
    # the mapped stretch = the scans in one frame, cropped to 60 m around the vehicle, jittered copies up to
    # map_per_gpu points.  It is repeated `world` times on a 130 m lattice: the map grows with the GPUs at
    # constant density, as a larger mapped area does, and stays inside the 700 m the 0.7 m block grid spans
    ctr = wp.mean(axis=0)
    near = wp[(np.abs(wp[:, 0] - ctr[0]) < 60.0) & (np.abs(wp[:, 1] - ctr[1]) < 60.0)]
    reps = (map_per_gpu + len(near) - 1) // len(near)
    stretch = np.concatenate([near + rng.normal(0, 0.03, size=near.shape).astype(np.float32)
                              for _ in range(reps)])[:map_per_gpu]
    themap = np.concatenate([stretch + np.float32([130.0 * (j % 3), 130.0 * (j // 3), 0.0]) for j in range(world)])
    owner = M.partition_by_cell(themap, world, cell=4.0, mode="interleaved")
    mine = M.shard_indices(owner, rank)
    # one 128-beam scan (260,096 points) taken in the middle of the mapped stretch
    q_pose = scene.trajectory(6)[5]
    q_scan = scene.make_scan(scene.World(1), q_pose, np.random.default_rng(1005), n_beams=128, n_azimuth=2032,
                             elev=(15.0, -25.0))
    queries = world_points([q_scan], [q_pose])[0]
    search = M.CudaShardSearch(icp, themap[mine], mine, 0.7, dev)
    sm = M.ShardedMap(search, rank, world, dist)
    qc = search.upload_queries(queries, 0.7)
    out = {"workload": "c5_sharded_map_knn", "n_queries": len(queries), "map_points": int(total),
           "map_points_this_rank": int(len(mine)), "partition": "4 m (x,y) cells interleaved over the ranks",
           "scaling": "map size grows with N at constant density (the mapped stretch repeated on a 130 m lattice); "
                      "the one scan overlaps one stretch, so each rank searches 1/N of its neighbourhood",
           "timing": "best of 4 single queries, max over ranks each", "cases": []}
    for k in (1, 6):
        sm.query(qc, k, 0.7)
        if dist is not None:
            dist.barrier()
        icp.profile_enable(True)
        icp.profile_reset()
        ms = 1e30
        for _ in range(4):  # best of 4: single calls of a few hundred microseconds are sensitive to host hiccups
            m1, keys = timed(torch, lambda: sm.query(qc, k, 0.7))
            if dist is not None:
                tt = torch.tensor([m1], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                m1 = float(tt[0])
            ms = min(ms, m1)
        pr = icp.profile()
        icp.profile_enable(False)
        search_ms = pr["knn_ms"] / max(pr["knn_launches"], 1)
        found = int((keys != M.NO_KEY).sum().item())
        case = {"k": k, "radius_m": 0.7, "ms": ms, "queries_per_s": len(queries) / (ms * 1e-3),
                "neighbours_found": found, "exchange_bytes_per_rank": int(sm.last_exchange_bytes),
                "merge": "all_reduce(MIN, int64)" if k == 1 else "all_gather + k-way merge kernel",
                "search_kernel_ms_rank0": search_ms}
        # the fused variant: the search kernel stores into every rank's buffer over NVLink (peer memory),
        # only barriers go through NCCL
        try:
            fk = sm.query_fused(qc, k, 0.7)
            same = bool(torch.equal(fk, keys))
            fms = 1e30
            for _ in range(4):
                m1, fk = timed(torch, lambda: sm.query_fused(qc, k, 0.7))
                if dist is not None:
                    tt = torch.tensor([m1], device=dev, dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    m1 = float(tt[0])
                fms = min(fms, m1)
            case["fused_peer_memory"] = {"ms": fms, "queries_per_s": len(queries) / (fms * 1e-3),
                                         "identical_to_nccl_path": same,
                                         "how": "search + atomicMin_system into the owner's slot of each query, owner stores its slice to every rank" if k == 1
                                         else "search + P2P row stores to the owner of each query, owner merges and stores the merged rows to every rank"}
        except Exception as e:
            case["fused_peer_memory"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        out["cases"].append(case)
    sm.close()
    qc.free()
    search.close()
    icp.close()
    return out


def sharded_native(torch, dist, capi, scans, poses, rank, world, local_rank, dev, map_per_gpu):
    """Config C5 through the library's own multi-GPU host (b200icp_comm_* / b200icp_sharded_*, csrc/sharded.inl):
    the NCCL communicator is the library's, nothing of torch.distributed is in the data path (it only hands the
    communicator id around).  kNN against the sharded map, and a whole REGISTRATION of the 128-beam scan against
    it -- bit-identical to the unsharded registration, which rank 0 also runs (the whole map fits one B200) and
    compares."""
    from mola_fe_lidar_b200 import scene
    icp = capi.ICP(capi.default_params(), device=local_rank)
    ids = [capi.comm_unique_id() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(ids, src=0)
    comm = capi.Comm(icp, ids[0], world, rank)
    wp = np.concatenate(world_points(scans, poses))
    rng = np.random.default_rng(5)  # the same map on every rank (as in sharded_knn)
    ctr = wp.mean(axis=0)
    near = wp[(np.abs(wp[:, 0] - ctr[0]) < 60.0) & (np.abs(wp[:, 1] - ctr[1]) < 60.0)]
    reps = (map_per_gpu + len(near) - 1) // len(near)
    stretch = np.concatenate([near + rng.normal(0, 0.03, size=near.shape).astype(np.float32)
                              for _ in range(reps)])[:map_per_gpu]
    themap = np.concatenate([stretch + np.float32([130.0 * (j % 3), 130.0 * (j // 3), 0.0]) for j in range(world)])
    q_pose = scene.trajectory(6)[5]
    q_scan = scene.make_scan(scene.World(1), q_pose, np.random.default_rng(1005), n_beams=128, n_azimuth=2032,
                             elev=(15.0, -25.0))
    truth = scene.matrix_to_pose6(q_pose)
    guess = truth + np.array([0.10, -0.05, 0.02, 0.003, 0.0, 0.0])
    smap = capi.NativeShardedMap(comm, themap, cell=4.0, interleaved=True, search_radius=0.7)
    local = icp.upload(q_scan)
    out = {"workload": "c5_sharded_map_native", "map_points": int(len(themap)), "map_points_this_rank": smap.local_size(),
           "scan_points": int(len(q_scan)), "host": "C++ (libb200icp.so owns the NCCL communicator)", "knn": []}

    def maxed(ms):
        if dist is None:
            return ms
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0])

    keep = {}
    for k in (1, 6):
        keys = torch.empty((len(q_scan), k), dtype=torch.int64, device=dev)
        smap.knn_keys(local, k, 0.7, keys.data_ptr(), pose6=truth)
        ms = 1e30
        for _ in range(4):
            m1, _ = timed(torch, lambda: smap.knn_keys(local, k, 0.7, keys.data_ptr(), pose6=truth))
            ms = min(ms, maxed(m1))
        keep[k] = keys
        out["knn"].append({"k": k, "ms": ms, "queries_per_s": len(q_scan) / (ms * 1e-3),
                           "merge": "ncclAllReduce(MIN, uint64)" if k == 1 else "ncclAllGather + k-way merge kernel"})
    r = smap.align(local, guess)  # warm-up
    ms = 1e30
    for _ in range(3):
        m1, r = timed(torch, lambda: smap.align(local, guess))
        ms = min(ms, maxed(m1))
    out["registration"] = {"ms": ms, "registrations_per_s": 1e3 / ms, "outer_iterations": int(r["n_iterations"]) + 1,
                           "quality": float(r["quality"]), "n_pairings": int(r["n_pairings"]),
                           "max_abs_translation_error_m": float(np.abs(r["pose"][:3] - truth[:3]).max()),
                           "per_iteration": "search on every shard, reduce-scatter of the lists by slice "
                                            "(ncclSend/ncclRecv), k-way merge + plane fit + moments on the owner, "
                                            "all-gather of the group partials, the same solve on every rank"}
    if rank == 0:  # the unsharded map on ONE GPU: same keys, same registration, bit for bit
        try:
            whole = icp.upload(themap, search_radius=0.7)
            same_keys = {}
            for k in (1, 6):
                kp = torch.empty((len(q_scan), k), dtype=torch.int64, device=dev)
                icp.knn_keys_device(whole, local, k, 0.7, kp.data_ptr(), pose6=truth)
                torch.cuda.synchronize()
                same_keys[k] = bool(torch.equal(kp, keep[k]))
            ms1, r1 = timed(torch, lambda: icp.align(whole, local, guess))
            out["unsharded_on_one_gpu"] = {
                "keys_identical": same_keys, "registration_ms": ms1,
                "registration_identical": bool(np.array_equal(r1["pose"], r["pose"]) and np.array_equal(r1["cov"], r["cov"])
                                               and r1["n_iterations"] == r["n_iterations"] and r1["quality"] == r["quality"])}
            whole.free()
        except Exception as e:
            out["unsharded_on_one_gpu"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if dist is not None:
        dist.barrier()
    local.free()
    smap.close()
    comm.close()
    icp.close()
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    from mola_fe_lidar_b200 import capi, lidar_odometry

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    scans, poses = make_scans(1 + rank)
    n_pts = len(scans[0])
    yaml_txt = open(os.path.join(lidar_odometry.PARAMS_DIR, "icp-settings-regular.yaml")).read()
    icp = capi.ICP(yaml_text=yaml_txt, device=local_rank)
    # inputs resident in HBM (SoA float, as the reference's CPointsMap buffers)
    dscans = [torch.from_numpy(np.ascontiguousarray(s.T)).to(dev) for s in scans]  # (3, n)
    # pinned host copies for the end-to-end leg
    hscans = [torch.from_numpy(np.ascontiguousarray(s.T)).pin_memory() for s in scans]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: scans resident in HBM, C ABI calls
    def make_cloud(i):
        t = dscans[i]
        return icp.from_device(t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), n_pts)

    state = {"prev": make_cloud(scan_index(0)), "next": make_cloud(scan_index(1)), "guess": np.zeros(6), "iters": 0,
             "pairs": 0, "rel": []}

    def step_value(step):
        # one step = the index build of one scan + one registration.  The scan of the NEXT step is handed to the
        # library before this step's registration is started: its index is built on the upload stream while the
        # registration runs (scans arrive independently of the registration results).
        cur = state["next"]
        state["next"] = make_cloud(scan_index(step + 2))
        r = icp.align(state["prev"], cur, state["guess"])
        state["prev"].free()
        state["prev"] = cur
        state["guess"] = next_guess(step, r["pose"])
        state["iters"] += r["n_iterations"] + 1
        state["pairs"] += r["n_pairings"]
        state["rel"].append((scan_index(step), scan_index(step + 1), np.array(r["pose"]), r["quality"]))
        return r

    total = args.warmup + args.steps
    # rank 0 samples its GPU every 5 ms; the other ranks every 50 ms (NVML calls of many processes
    # at a high rate get in the way of each other's kernel launches)
    sampler = ClockSampler(local_rank, 0.005 if rank == 0 else 0.05)
    sampler.start()
    for s in range(args.warmup):
        step_value(s)
    # ---- the timed region: K steps, no per-launch events (the library replays the first batch of a predicted
    # registration as one CUDA graph when it is not asked to time its kernels)
    icp.profile_enable(False)
    icp.profile_reset()
    state["iters"] = state["pairs"] = 0
    state["rel"] = []
    barrier()
    t_mark0 = sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.warmup, total):
        step_value(s)
    ev1.record()
    barrier()
    t_mark1 = sampler.mark()
    ms_value = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_mark0, t_mark1)
    launches_timed = int(icp.profile()["total_kernel_launches"])
    graph_replays_timed = int(icp.profile()["graph_replays"])
    timed_iters, timed_rel = state["iters"], list(state["rel"])
    # ---- K more steps of the same sequence with the library's CUDA events around every kernel launch: the
    # per-kernel times behind `roofline` and `kernel_ms` (events between the launches keep the graph replay off, so
    # this pass is a little slower than the timed one; its own time is reported as profiled_pass_ms_per_step)
    icp.profile_enable(True)
    icp.profile_reset()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for s in range(total, total + args.steps):
        step_value(s)
    ev3.record()
    barrier()
    ms_profiled = ev2.elapsed_time(ev3)
    prof = icp.profile()
    icp.profile_enable(False)
    state["iters"], state["rel"] = timed_iters, timed_rel
    state["prev"].free()
    state["next"].free()
    mean_iters = state["iters"] / max(args.steps, 1)
    # accuracy against the generator's ground truth: per-step relative pose error and the absolute
    # trajectory error of the chained estimates over the timed steps
    from mola_fe_lidar_b200 import scene as _scene
    terr, rerr, T_est, T_gt, ate = [], [], np.eye(4), np.eye(4), []
    for (i, j, p6, q) in state["rel"]:
        gt = _scene.relative_pose6(poses[i], poses[j])
        terr.append(float(np.linalg.norm(p6[:3] - gt[:3])))
        rerr.append(float(np.abs(p6[3:] - gt[3:]).max()))
        T_est = T_est @ _scene.pose_matrix(*p6)
        T_gt = T_gt @ _scene.pose_matrix(*gt)
        ate.append(float(np.linalg.norm(T_est[:3, 3] - T_gt[:3, 3])))
    accuracy = {"steps": len(terr), "rel_translation_error_m_mean": float(np.mean(terr)) if terr else None,
                "rel_translation_error_m_max": float(np.max(terr)) if terr else None,
                "rel_rotation_error_rad_max": float(np.max(rerr)) if rerr else None,
                "ate_rmse_m": float(np.sqrt(np.mean(np.square(ate)))) if ate else None,
                "mean_quality": float(np.mean([q for *_, q in state["rel"]])) if terr else None,
                "note": "vs the synthetic generator's ground truth (range noise sigma 0.02 m), rank 0's sequence"}

    # ---------------- e2e: the LidarOdometry module, pinned host buffers.
    # Extra-edge / loop-closure checks (checkForNearbyKFs) are switched off with
    # the additive key b200_extra_edge_checks so that every timed scan costs
    # exactly one consecutive-scan registration -- the unit `value` and the
    # reference arm count; `e2e_full_module` below runs the module as shipped.
    replays = {"n": 0}

    def run_module(extra_yaml, steps, warm, voxel=None, feed="async"):
        """feed "async": scans go in through onNewObservation as the reference's data source delivers them
        (LidarOdometry.cpp:162-187: enqueue on the 1-thread pool), as fast as the module takes them (at most 4
        waiting, so the >10-queued drop rule never fires); "sync": each scan is processed on the calling thread
        before the next one is handed over.  Returns (ms, registrations, scans, key-frames) of the timed region."""
        # additive key b200_device: this rank's GPU
        lo = lidar_odometry.LidarOdometry(
            yaml_text=lidar_odometry.system_yaml(voxel_resolution=voxel,
                                                 extra=f"  b200_device: {local_rank}\n" + extra_yaml))
        stamp = 0.0

        def hand_over(s, stamp):
            h = hscans[scan_index(s)]
            if feed == "async":
                lo.enqueueObservationSoA(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), n_pts, stamp)
            else:
                lo.onNewObservationSoA(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), n_pts, stamp, sync=True)

        for s in range(warm + 1):  # +1: the first scan only creates a keyframe; warm-up in the same feed mode
            hand_over(s, stamp)
            stamp += 0.1
        lo.wait_idle()
        st0 = lo.state()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for s in range(warm + 1, warm + 1 + steps):
            hand_over(s, stamp)
            stamp += 0.1
        lo.wait_idle()  # every queued scan and extra-edge registration belongs to the timed region
        e3.record()
        barrier()
        ms = e2.elapsed_time(e3)
        st = lo.state()
        prof_mod = lo.profile()
        replays["n"] = sum(capi.ICP.profile_of_handle(lo.icp_handle(kind))["graph_replays"] for kind in (0, 1, 2))
        lo.close()
        n_all = max(st["n_processed"], 1)
        sections = {k.replace("doProcessNewObservation.", ""): [round(v[1] / max(v[0], 1) * 1e3, 3), round(v[2] * 1e3, 3)]
                    for k, v in prof_mod.items() if v[0] > 0 and not k.startswith("exception") and v[1] > 1e-5}
        log(f"[bench] module sections, [mean, max] ms per call over {n_all} scans:", json.dumps(sections))
        return ms, int(st["n_icp"] - st0["n_icp"]), int(st["n_processed"] - st0["n_processed"]), \
            int(st["n_keyframes"])

    # the headline leg: the same K scans three times, each time through a fresh module, the median run reported (a
    # rare stall of some tens of milliseconds -- all the module's threads at once -- otherwise decides a
    # 20-millisecond timed region)
    e2e_runs = sorted(run_module("  b200_extra_edge_checks: false\n", args.steps, args.warmup) for _ in range(3))
    e2e_windows_ms = [r[0] for r in e2e_runs]
    ms_e2e, e2e_regs, e2e_scans, _ = e2e_runs[1]
    e2e_graph_replays = int(replays["n"])
    ms_sync, sync_regs, _, _ = run_module("  b200_extra_edge_checks: false\n", args.steps, args.warmup, feed="sync")
    ms_full, full_regs, full_scans, full_kfs = run_module("", args.steps, max(args.warmup, 12))
    # C2's decimated variant: the voxel filter kitti-default.yaml hints at (1.0 m) in front of the same ICP
    ms_dec, dec_regs, _, _ = run_module("  b200_extra_edge_checks: false\n", args.steps, args.warmup, voxel=1.0)

    # ---------------- aggregate over ranks (max time, sum of units)
    t = torch.tensor([ms_value, ms_e2e, ms_full, ms_dec, ms_sync], device=dev, dtype=torch.float64)
    u = torch.tensor([float(args.steps), float(e2e_regs), float(full_regs), float(full_scans), float(dec_regs),
                      float(sync_regs)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    ms_value_max, ms_e2e_max, ms_full_max = float(t[0]), float(t[1]), float(t[2])
    value = float(u[0]) / (ms_value_max * 1e-3)
    e2e_value = float(u[1]) / (ms_e2e_max * 1e-3)

    peak, peak_src = load_peaks()
    extras = {}
    if not args.no_extras:
        def section(name, fn):
            try:
                return fn()
            except Exception as e:  # the contract line must still be printed, the other sections still run
                extras.setdefault("extras_error", {})[name] = f"{type(e).__name__}: {e}"[:300]
                log(f"[bench] section {name} failed:", extras["extras_error"][name])
                return None

        if rank == 0:
            r = section("knn", lambda: knn_microbench(torch, icp, scans, poses, dev, peak))
            if r is not None:
                extras["knn"], extras["scan_to_map"] = r
        shared_scans, shared_poses = (scans, poses) if world == 1 else make_scans(1)
        r = section("batch_lc", lambda: batch_lc(torch, dist, capi, lidar_odometry, shared_scans, rank, world,
                                                 local_rank, dev, args.lc_pairs))
        if r is not None:
            extras["batch_lc"] = r
        r = section("sharded_knn", lambda: sharded_knn(torch, dist, capi, shared_scans, shared_poses, rank, world,
                                                       local_rank, dev, args.map_points_per_gpu))
        if r is not None:
            extras["sharded_knn"] = r
        r = section("sharded_native", lambda: sharded_native(torch, dist, capi, shared_scans, shared_poses, rank, world,
                                                             local_rank, dev, args.map_points_per_gpu))
        if r is not None:
            extras["sharded_native"] = r

    if rank == 0:
        launches = max(prof["match_launches"], 1)
        avg_ms = prof["match_ms"] / launches
        achieved = ALGO_BYTES_PER_QUERY * n_pts / (avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("search_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": "icp_registrations_per_sec", "value": value, "unit": "registrations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 search / f64 solve", "data": "synthetic",
            "config": workload_config(n_pts),
            "mean_outer_iterations": mean_iters, "sequences": world,
            "timing": "CUDA events on the legacy default stream bracketing the library's blocking streams; "
                      "max over ranks; `value` = K steps without per-launch events, `roofline` / `kernel_ms` = the "
                      "next K steps of the same sequence with the library's CUDA events around every launch",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "search_tile_kernel<6> (the matcher's kNN search)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_QUERY * n_pts,
                         "avg_launch_ms": avg_ms, "launches_timed": int(prof["match_launches"]),
                         "note": "working set is L2-resident; the kernel is issue/latency bound, see DESIGN.md"},
            "e2e": {"value": e2e_value, "unit": "registrations/s", "h2d_bytes_per_step": n_pts * 12,
                    "d2h_bytes_per_step": 1128, "ms_per_step": ms_e2e_max / max(float(u[1]) / world, 1.0),
                    "cuda_graph_replays_rank0": e2e_graph_replays,
                    "windows_ms_rank0": [round(w, 3) for w in e2e_windows_ms],
                    "windows_note": "the same K scans three times, each through a fresh module; value = the median run",
                    "api": "LidarOdometry.onNewObservation (b200lo_enqueue_observation: the reference's asynchronous "
                           "entry, LidarOdometry.cpp:162-187), pinned host SoA, at most 4 scans waiting; "
                           "b200_extra_edge_checks: false (one consecutive-scan registration per scan); the module "
                           "uploads and indexes scan i+1 on its own stream while scan i is being registered",
                    "sync_feed_value": float(u[5]) / (float(t[4]) * 1e-3),
                    "sync_feed_note": "b200lo_process_observation: each scan processed on the calling thread before "
                                      "the next is handed over (no overlap of upload / index build and registration)"},
            "e2e_full_module": {"registrations_per_s": float(u[2]) / (ms_full_max * 1e-3),
                                "scans_per_s": float(u[3]) / (ms_full_max * 1e-3),
                                "registrations": int(u[2]), "scans": int(u[3]), "keyframes_rank0": full_kfs,
                                "note": "module as shipped: keyframes + extra-edge registrations between "
                                        "keyframes 5-20 m apart run on the pool threads inside the timed region"},
            "e2e_decimated_1m": {"registrations_per_s": float(u[4]) / (float(t[3]) * 1e-3),
                                 "note": "same module and host buffers with pointcloud_filter = FilterDecimateVoxels "
                                         "(voxel_filter_resolution 1.0 m): 120k-pt scans -> ~3k points per cloud"},
            "accuracy": accuracy,
            "gpu_launches": launches_timed,
            "cuda_graph_replays_timed_region": graph_replays_timed,
            "profiled_pass_ms_per_step": ms_profiled / args.steps,
            "kernel_ms": {"search": prof["match_ms"], "fit": prof["fit_ms"], "solve": prof["solve_ms"],
                          "index": prof["index_ms"], "index_builds": int(prof["index_builds"])},
            "clocks": clocks,
        }
        out.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(scans)
        print(json.dumps(out), flush=True)
    icp.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(scans, budget_s=20.0):
    """The oracle port on ONE host thread (how the reference runs one ICP,
    LidarOdometry.h:167-168): a bounded sample of the same workload -- the first
    scan pairs of the sequence, each from the constant-velocity guess."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_api as O
    O.build()
    prm = O.default_params()
    clouds = {}

    def cloud(i):
        if i not in clouds:
            clouds[i] = O.Cloud(scans[i])
            O.knn(clouds[i], scans[0][:8], 1, 0.49, kdtree=True)  # lazy kd-tree build, as on first query in MRPT
        return clouds[i]

    cloud(0), cloud(1), cloud(2)
    r = O.icp_align(cloud(0), cloud(1), np.zeros(6), prm, kdtree=True)  # untimed: gives the first velocity guess
    guess = next_guess(0, r["pose"])
    t_all, n, iters, step = 0.0, 0, [], 1
    while n < 3 or (t_all < budget_s and n < 24):
        a, b = cloud(scan_index(step)), cloud(scan_index(step + 1))
        t0 = time.time()
        r = O.icp_align(a, b, guess, prm, kdtree=True)
        t_all += time.time() - t0
        guess = next_guess(step, r["pose"])
        iters.append(r["n_iterations"] + 1)
        n, step = n + 1, step + 1
    return {"value": n / t_all, "unit": "registrations/s", "cores": 1, "kind": "port",
            "mean_outer_iterations": float(np.mean(iters)),
            "sample": f"{n} registrations of consecutive 120k-pt scans from constant-velocity guesses in "
                      f"{t_all:.1f}s, 1 thread (oracle/icp_oracle.c with its kd-tree)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the knn / batch_lc / sharded_knn sections")
    ap.add_argument("--lc-pairs", type=int, default=1024, help="C4 section: candidate pairs in total (sharded over the GPUs)")
    ap.add_argument("--map-points-per-gpu", type=int, default=2_500_000, help="C5-shaped section")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # one distinct scan per processed step when that stays affordable (0.45 s of ray casting per scan):
    # the module's constant-velocity guess is then right at every step, as on a real sequence
    global N_SCANS
    if args.impl == "b200":
        N_SCANS = min(max(10, args.warmup + args.steps + 3), 48)
    else:  # the same sequence; the threads of the reference arm walk it staggered by one scan each
        N_SCANS = min(max(10, max(args.warmup, 1) + args.steps + 3 + (os.cpu_count() or 1)), 48)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: anything native libraries print there (NCCL's
    # version banner) is sent to stderr, and the line is written to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
