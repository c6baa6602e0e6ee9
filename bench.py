#!/usr/bin/env python
"""bench.py -- LiDAR points registered / s (scan registration + scan-to-map registration, whole box).

Default workload = BASELINE.json configs[1]: HDL-64E-shaped sweeps (64 x 2048 = 131,072 points) registered against a
>= 1,000,000-point local map per stream (asserted on the RESIDENT map, after the 0.4 m map voxel merge).  One step = one
sweep of EVERY stream through the full hot path (scan registration -> frame voxel filters -> <= 10 Gauss-Newton
iterations of exact-5-NN correspondence + solve -> map insertion).  Streams are independent LiDARs (own pose chain, own
map) that walk a fresh trajectory: no sweep is registered twice by a stream inside the run, so map growth (new voxels,
cell growth) happens inside the timed region.  Ranks shard streams (weak scaling, no data-path collective).
`value` = sweeps resident in HBM, the product path exactly as a caller runs it (Gauss-Newton loop as a WHILE graph);
`e2e` = pinned host sweeps through the C ABI (cm_pipeline_prefetch_host + cm_pipeline_step_host), H2D of the sweeps and
D2H of the poses inside the timed region.  Per-kernel times come from a separate event-timed pass after the timed arms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5] [--streams S]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "the-cooper-mapper_b200"
ROWS, COLS = 64, 2048
NPTS = ROWS * COLS
SEED = 0x5EED0002
CFG = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
for _k in ("cell_surf", "cell_corner"):                  # development sweeps of the search-cell size (results do not depend on it)
    if os.environ.get("BENCH_" + _k.upper()):
        CFG[_k] = float(os.environ["BENCH_" + _k.upper()])
ORACLE_MAP = dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4)
MAP_SPACING = 0.3          # sampling lattice of the prebuilt map: 1.8 M samples -> 1.02 M resident points at the 0.4 m map leaf
METRIC = "lidar_points_registered_per_s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line, the JSON result: libraries that write to file descriptor 1 on their own (NCCL prints its version
# when torch.distributed creates the communicator) are sent to stderr for the whole run, the result goes to the saved descriptor
def claim_stdout():
    # (kept on the sys module: bench_configs.py imports this file a second time under the name `bench`)
    if getattr(sys, "_coopermap_result_out", None) is None:
        sys.stdout.flush()
        sys._coopermap_result_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = getattr(sys, "_coopermap_result_out", None) or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------------------------
def _sim_one(args):
    synth = importlib.import_module(PKG + ".synth")
    sc, R, t, model, seed = args
    return synth.simulate_scan(sc, R, t, model, seed=seed)


def simulate_pool(synth, sc, traj, model, seed0, procs=None):
    """One sweep per trajectory pose, simulated on `procs` host processes (fork: call before CUDA is initialised)."""
    jobs = [(sc, R, t, model, seed0 + k) for k, (R, t) in enumerate(traj)]
    procs = procs or max(1, min(len(jobs), (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if procs <= 1 or len(jobs) <= 2:
        return [_sim_one(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_sim_one, jobs)


def make_workload(n_pool, synth):
    """Scene, map samples (corner, surf; >= 1 M points once voxel-merged at 0.4 m) and a trajectory of n_pool HDL-64E sweeps."""
    sc = synth.make_scene(seed=SEED & 0xFFFF, extent=125.0, n_boxes=44, n_poles=40)
    mc, ms = synth.sample_map(sc, MAP_SPACING, seed=2)
    traj = synth.trajectory(n_pool, speed=2.0)
    frames = np.stack(simulate_pool(synth, sc, traj, "HDL-64E", 1000)).astype(np.float32)
    poses = [(R.astype(np.float32), t.astype(np.float32)) for R, t in traj]
    return mc, ms, frames, poses


def noisy_odom(poses, idx, rng, synth):
    """Odometry prediction = truth + U(+-0.1 m, +-0.5 deg)."""
    R, t = poses[idx]
    d = np.deg2rad(rng.uniform(-0.5, 0.5, 3))
    dR, _ = synth.pose_matrix(d[2], d[1], d[0])
    return (R @ dR.astype(np.float32)).astype(np.float32), (t + rng.uniform(-0.1, 0.1, 3)).astype(np.float32)


def pack_isos(isos):
    a = np.empty((len(isos), 12), np.float32)
    for i, (R, t) in enumerate(isos):
        a[i, :9] = np.asarray(R, np.float32).ravel(); a[i, 9:] = t
    return a


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (the reference restated line by line + the reference's own nanoflann) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wid, mc, ms, frames, poses, order, warm = args
    from oracle import oracle_py as O
    synth = importlib.import_module(PKG + ".synth")
    rng = np.random.default_rng(500 + wid)
    m = O.Mapping(map_params=ORACLE_MAP, nanoflann=True, fast=True)
    m.map_update(np.zeros(3, np.float32))
    m.map_add(mc, ms, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    times = []
    for k, fi in enumerate(order):
        oR, ot = noisy_odom(poses, fi, rng, synth)
        t0 = time.perf_counter()
        f = O.scanreg_organised(frames[fi], fast=True)
        t1 = time.perf_counter()
        m.process(oR, ot, f["lessSharp"], f["lessFlat"])
        t2 = time.perf_counter()
        if k >= warm:
            times.append((t2 - t0, t1 - t0, t2 - t1))
    return times


def cpu_arm(mc, ms, frames, poses, frames_per_worker, warm=1, workers=None):
    """Runs `workers` independent streams, one per process / core.  Returns (points/s, cores, per-frame seconds [total, stage 1, stage 3], wall)."""
    import multiprocessing as mp
    workers = workers or min(os.cpu_count() or 1, 32)
    ctx = mp.get_context("fork")
    jobs = []
    for w in range(workers):
        order = [(3 * w + k) % len(frames) for k in range(warm + frames_per_worker)]
        jobs.append((w, mc, ms, frames, poses, order, warm))
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    per_frame = np.array([x for r in res for x in r], np.float64).reshape(-1, 3)
    # throughput from the timed frames only: every worker runs concurrently, so the box rate is workers / mean frame time
    rate = workers * NPTS / float(np.mean(per_frame[:, 0]))
    return rate, workers, per_frame, wall


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """The profiling recipe's clocks line: ONE `nvidia-smi -lms` process, started before the warm-up (its start-up takes the
    driver's lock for tens of ms and must not land in a timed region) and killed after the last timed step; the summary uses
    the samples that fall inside the timed regions (mark_begin() / mark_end())."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index, period_ms=50):
        self.gpu = gpu_index; self.period_ms = period_ms; self.rows = []; self.proc = None; self.thread = None
        self.windows = []

    def start(self):
        try:
            import shutil
            pre = ["stdbuf", "-oL"] if shutil.which("stdbuf") else []   # line-buffered pipe: samples arrive as they are taken
            self.proc = subprocess.Popen(pre + ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                                "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()

    def _read(self):
        try:
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                try:
                    try:   # nvidia-smi's own sample time (local time, ms resolution); arrival time if it does not parse
                        ts = time.mktime(time.strptime(f[0][:19], "%Y/%m/%d %H:%M:%S")) + float("0" + f[0][19:])
                    except ValueError:
                        ts = time.time()
                    self.rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(self.NAMES, f[3:]) if v.lower().startswith("active")]))
                except (ValueError, IndexError):
                    pass
        except Exception:
            pass

    def mark_begin(self):
        self.windows.append([time.time(), None])

    def mark_end(self):
        self.windows[-1][1] = time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows if any(w[1] is not None and w[0] - 0.03 <= r[0] <= w[1] + 0.03 for w in self.windows)]
        where = "timed regions"
        if not rows:
            rows = self.rows; where = "whole run (no sample fell inside the timed regions)"
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": rows[-1][2] if rows else None,
                "reasons": reasons, "samples": len(rows), "window": where}


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture, if any (profiles/kernel_traffic.json)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        for k, v in d.items():
            if k in kernel:
                return v
    except Exception:
        pass
    return None


def bind_to_gpu_numa(local_rank):
    """Pin this rank (and the pinned buffers it allocates from now on) to the CPUs of its GPU's NUMA node: the sweeps of 8 ranks
    then cross 8 different PCIe root ports from node-local memory instead of funnelling through one socket."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        dev = "/sys/bus/pci/devices/" + bus
        cpus = open(dev + "/local_cpulist").read().strip()
        node = int(open(dev + "/numa_node").read().strip())
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"pci": bus, "numa_node": node, "cpus": cpus}
    except Exception as e:   # no sysfs entry (virtualised box): leave the affinity alone
        return {"error": str(e)[:80]}


# per-kernel ALGORITHMIC bytes: what a launch must move at least, from the run-time counters of the
# step it belongs to.  c = {raw: raw sweep points, feat: feature points written by scan registration, qi: query-iterations,
# q: filtered queries, ins: inserted points, fin: points entering the frame voxel filters}
KERNEL_BYTES = [
    ("sr_ring_kernel", lambda c: 16.0 * c["raw"] + 16.0 * c["feat"], "16 B per raw point read + 16 B per feature point written"),
    ("sr_assemble_kernel", lambda c: 32.0 * c["feat"], "16 B read + 16 B written per feature point"),
    ("search_kernel", lambda c: 96.0 * c["qi"], "16 B query + 5 x 16 B neighbours per query-iteration (SURVEY 8d)"),
    ("search_hard_kernel", None, "the sparse-surroundings tail of search_kernel's queries (its bytes are counted there)"),
    ("fit_solve_kernel", lambda c: 128.0 * c["qi"], "16 B query + 5 x 16 B neighbours read, 32 B row written per query-iteration"),
    ("vox_", lambda c: None, None),
    ("map_merge_kernel", lambda c: 32.0 * c["ins"], "16 B new point read, 16 B voxel point written / merged"),
]


def kernel_bytes(name, counters):
    for key, fn, _ in KERNEL_BYTES:
        if key in name:
            try:
                return fn(counters) if fn else None
            except Exception:
                return None
    return None


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, synth, rank):
    if rank != 0:
        return
    K = max(args.steps, 1)
    workload = "HDL-64E 64x2048 sweeps, scan registration + scan-to-map vs >=1M-point local map per stream"
    mc, ms, frames, poses = make_workload(min(args.pool or 6, 6), synth)
    rate, cores, per_frame, wall = cpu_arm(mc, ms, frames, poses, frames_per_worker=K, warm=max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "points/s",
            "n_gpus": args.gpus, "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(per_frame[:, 0])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "map_points_sampled": int(len(mc) + len(ms)), "streams": cores,
                       "note": "CPU oracle (line-by-line restatement + the reference's vendored nanoflann), one stream per core"},
            "cpu_baseline": {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": "%d sweeps per core after %d warm-up, wall %.1f s" % (K, max(args.warmup, 1), wall)},
            "p50_latency_ms": 1e3 * float(np.median(per_frame[:, 0])),
            "stage1_alone_p50_ms": 1e3 * float(np.median(per_frame[:, 1])), "stage3_alone_p50_ms": 1e3 * float(np.median(per_frame[:, 2])),
            "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_config2(args, synth, rank, world, local_rank):
    K, W, S = args.steps, max(args.warmup, 3), args.streams
    n_steps = W + K
    workload = "HDL-64E 64x2048 sweeps, scan registration + scan-to-map vs >=1M-point local map per stream"
    # ---- workload on the host (before CUDA is touched: the simulation and the CPU baseline fork) -------------------
    t_setup = time.time()
    P = args.pool or (2 * n_steps + 8)                       # one fresh sweep per stream and step over BOTH timed arms
    mc, ms, frames, poses = make_workload(P, synth)
    log("[bench] workload: map samples %d corner + %d surf, %d sweeps (%.1f s)" % (len(mc), len(ms), len(frames), time.time() - t_setup))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, per_frame, wall = cpu_arm(mc, ms, frames[:6], poses[:6], frames_per_worker=12, warm=1)
        cpu = {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
               "sample": "12 sweeps per core after 1 warm-up (oracle -O3 + reference nanoflann, one stream per core), wall %.1f s, p50 %.0f ms/sweep"
                         % (wall, 1e3 * float(np.median(per_frame[:, 0]))),
               "stage1_alone_points_per_s": cores * NPTS / float(np.mean(per_frame[:, 1])),
               "stage3_alone_points_per_s": cores * NPTS / float(np.mean(per_frame[:, 2])),
               "p50_ms": {"end_to_end": 1e3 * float(np.median(per_frame[:, 0])), "stage1": 1e3 * float(np.median(per_frame[:, 1])),
                          "stage3": 1e3 * float(np.median(per_frame[:, 2]))}}
        log("[bench] cpu baseline: %.3e points/s on %d cores" % (rate, cores))

    import torch
    import torch.distributed as dist
    numa = bind_to_gpu_numa(local_rank) if not args.no_numa_bind else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cmb = importlib.import_module(PKG)
    ctx = cmb.Context(device=local_rank, **CFG)
    cap_c, cap_s = max(8 * len(mc), 200000), int(1.2 * len(ms)) + 400000
    ctx.mapping_create(S, max_corner_points=cap_c, max_surf_points=cap_s)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    # prebuilt map in every stream (inserted through the product's own insert kernels, in chunks)
    chunk = 1 << 18
    for o in range(0, len(ms), chunk):
        c_part = mc if o == 0 else mc[:0]
        ctx.map_insert([c_part] * S, [ms[o:o + chunk]] * S, [eye] * S)
    map_pts = [len(ctx.map_export(0, 0)[0]), len(ctx.map_export(0, 1)[0])]
    log("[bench] rank %d: %d streams, resident map per stream: %d corner + %d surf points" % (rank, S, map_pts[0], map_pts[1]))
    assert sum(map_pts) >= 1000000, "BASELINE config 2 asks for a 1M-point local map: resident %d" % sum(map_pts)

    rng = np.random.default_rng(77 + rank)
    # stream s registers sweep (7 s + 5 rank + g) mod P at global step g (device arm: g = k, end-to-end arm: g = n_steps + k):
    # every stream walks the trajectory forward, 2 m per step, and never meets a sweep twice
    def sweep_of(s, g):
        return (7 * s + 5 * rank + g) % P
    G_total = 2 * n_steps + 16
    idx = np.array([[sweep_of(s, g) for s in range(S)] for g in range(G_total)])
    odom = [pack_isos([noisy_odom(poses, idx[g][s], rng, synth) for s in range(S)]) for g in range(G_total)]
    pool_dev = torch.from_numpy(frames).to(dev)
    step_dev = [pool_dev[torch.from_numpy(idx[g]).to(dev)].contiguous() for g in range(n_steps)]
    torch.cuda.synchronize()
    mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled by rank 0 only (its own GPU): the other ranks' GPUs run the same work at the same time
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident arm (the product path: WHILE-graph Gauss-Newton loop, filter / insert chains as graphs) ------
    ctx.pipeline_prefetch_dev(step_dev[0].data_ptr(), ROWS, COLS)
    for k in range(W):                                       # warm-up through the same (pipelined) path as the timed steps
        ctx.pipeline_prefetch_dev(step_dev[k + 1].data_ptr(), ROWS, COLS)
        ctx.pipeline_step_dev(step_dev[k].data_ptr(), ROWS, COLS, odom[k], mapped, stats)
    barrier()
    sampler.mark_begin()
    launches0 = ctx.launch_count()
    gb0 = ctx.graph_builds()
    cnt = dict(qi=0, q=0, ins=0, feat=0)
    iters = []
    ctx.timer_record(0)
    # software pipeline across steps: scan registration of step k+1 (cm_pipeline_prefetch_dev, side stream) is issued before
    # step k's matching, so the issue-bound feature extraction overlaps the latency-bound Gauss-Newton loop
    # (the sweep of step W was prefetched by the last warm-up step)
    for k in range(W, W + K):
        if k + 1 < W + K:
            ctx.pipeline_prefetch_dev(step_dev[k + 1].data_ptr(), ROWS, COLS)
        ctx.pipeline_step_dev(step_dev[k].data_ptr(), ROWS, COLS, odom[k], mapped, stats)
        c = ctx.last_step_counters()
        cnt["qi"] += c["query_iters"]; cnt["q"] += c["queries"]; cnt["ins"] += c["inserted"]; cnt["feat"] += c["features"]
        iters += [st.iterations for st in stats]
    ctx.timer_record(1)
    ms_total = ctx.timer_elapsed_ms()
    sampler.mark_end()
    barrier()
    launches = ctx.launch_count() - launches0
    gb1 = ctx.graph_builds()
    conv = float(np.mean([st.converged for st in stats]))
    map_after = [len(ctx.map_export(0, 0)[0]), len(ctx.map_export(0, 1)[0])]
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * S * K * NPTS / (ms_max * 1e-3)
    del step_dev
    torch.cuda.empty_cache()

    # ---- end-to-end arm: pinned host sweeps through the C ABI --------------------------------------------------------
    # every step has its own pinned buffer (fresh sweeps); beyond --pinned-gb the buffers are reused round-robin
    per_step = S * NPTS * 16
    nbuf = max(4, min(n_steps, int(args.pinned_gb * 1e9 // per_step)))
    host_t = [torch.empty((S, ROWS, COLS, 4), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    host_np = [h.numpy() for h in host_t]

    def fill(b, g):
        np.take(frames, idx[g], axis=0, out=host_np[b])

    for b in range(min(nbuf, n_steps)):
        fill(b, n_steps + b)
    A = min(max(args.ahead, 1), 3)
    ge = lambda k: n_steps + (k % nbuf)                      # the global step whose sweeps buffer k % nbuf holds

    def e2e_steps(first, count):
        # `A` sweeps ahead: upload of step k+A (copy streams) | scan registration of the steps before it (side stream) | matching
        # of step k (main stream); every step's sweeps cross PCIe, the poses of step k are read back before step k+1 is issued
        for j in range(first, min(first + A, first + count)):
            ctx.pipeline_prefetch(host_np[j % nbuf])
        for k in range(first, first + count):
            if k + A < first + count:
                ctx.pipeline_prefetch(host_np[(k + A) % nbuf])
            ctx.pipeline_step_packed(host_np[k % nbuf], odom[ge(k)], mapped, stats)

    e2e_steps(0, W)
    barrier()
    sampler.mark_begin()
    ctx.timer_record(0)
    e2e_steps(W, K)
    ctx.timer_record(1)
    e2e_ms = ctx.timer_elapsed_ms()
    barrier()
    sampler.mark_end()
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * S * K * NPTS / (float(t.item()) * 1e-3)

    # ---- end-to-end with the xyz-only entry: pinned packed coordinates, 12 B per point over PCIe (the pipeline never reads the
    #      intensity of a raw point: scan registration replaces it with ring + relTime, OrganizedScanRegistration.cpp:109-110) ---------
    e2e16_value = e2e_value
    e2e_bytes = int(S * NPTS * 16 + S * 48)
    e2e_input = "pinned packed cm_point sweeps (x, y, z, intensity: 16 B per point)"
    if not args.no_xyz12_arm:
        x_t = [torch.empty((S, NPTS, 3), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
        x_np = [h.numpy() for h in x_t]
        for b in range(min(nbuf, n_steps)):
            np.copyto(x_np[b], frames[idx[n_steps + b]].reshape(S, NPTS, 4)[:, :, :3])
        x_ptr = [ctx.cloud_ptrs([x_np[b][si] for si in range(S)]) for b in range(nbuf)]

        def xyz_steps(first, count):
            for j in range(first, min(first + A, first + count)):
                ctx.pipeline_prefetch_strided_ptrs(x_ptr[j % nbuf][0], 12, ROWS, COLS)
            for k in range(first, first + count):
                if k + A < first + count:
                    ctx.pipeline_prefetch_strided_ptrs(x_ptr[(k + A) % nbuf][0], 12, ROWS, COLS)
                ctx.pipeline_step_strided_ptrs(x_ptr[k % nbuf][0], 12, ROWS, COLS, odom[ge(k)], mapped, stats)

        xyz_steps(0, W)
        barrier()
        sampler.mark_begin()
        ctx.timer_record(0)
        xyz_steps(W, K)
        ctx.timer_record(1)
        x_ms = ctx.timer_elapsed_ms()
        barrier()
        sampler.mark_end()
        t = torch.tensor([x_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * S * K * NPTS / (float(t.item()) * 1e-3)
        e2e_bytes = int(S * NPTS * 12 + S * 48)
        e2e_input = "pinned packed coordinates through the xyz-only entry (cm_pipeline_*_strided_host, stride 12: 12 B per point)"
        del x_t, x_np, x_ptr

    sampler.stop()

    # ---- end-to-end from the layout a nodelet holds: one PAGEABLE pcl::PointXYZI cloud (32-byte stride) per stream ------------------
    # cm_pipeline_prefetch_strided_host packs x, y, z into library-owned pinned staging with worker threads and uploads 12 B / point;
    # the packing is inside the timed region
    e2e_pcl = None
    if not args.no_pcl_arm:
        nbp = max(4, min(n_steps, 6))
        pcl_bufs = []
        for b in range(nbp):
            a = np.empty((S, NPTS, 8), np.float32)
            src = frames[idx[n_steps + b]].reshape(S, NPTS, 4)
            a[:, :, :3] = src[:, :, :3]; a[:, :, 3] = 1.0; a[:, :, 4] = src[:, :, 3]; a[:, :, 5:] = 0.0
            pcl_bufs.append([a[si] for si in range(S)])
        gp = lambda k: n_steps + (k % nbp)

        def pcl_steps(first, count):
            for j in range(first, min(first + A, first + count)):
                ctx.pipeline_prefetch_strided(pcl_bufs[j % nbp], ROWS, COLS)
            for k in range(first, first + count):
                if k + A < first + count:
                    ctx.pipeline_prefetch_strided(pcl_bufs[(k + A) % nbp], ROWS, COLS)
                ctx.pipeline_step_strided(pcl_bufs[k % nbp], ROWS, COLS, odom[gp(k)], mapped, stats)

        pcl_steps(0, W)
        barrier()
        ctx.timer_record(0)
        pcl_steps(W, K)
        ctx.timer_record(1)
        pcl_ms = ctx.timer_elapsed_ms()
        barrier()
        t = torch.tensor([pcl_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pcl = {"value": world * S * K * NPTS / (float(t.item()) * 1e-3), "unit": "points/s",
                   "input": "one pageable pcl::PointXYZI cloud per stream (32-byte stride), packed to 12 B / point by the library's worker "
                            "threads into pinned staging inside the timed region",
                   "h2d_bytes_per_step": int(S * NPTS * 12 + S * 48), "host_bytes_read_per_step": int(S * NPTS * 32),
                   "stage_threads": int(os.environ.get("COOPERMAP_STAGE_THREADS", "0")) or min(len(os.sched_getaffinity(0)), 16)}
        del pcl_bufs

    # ---- per-kernel times: a separate pass with every launch bracketed by CUDA events (launch-by-launch, no graphs) -------
    kernels = []
    stage = {}
    if rank == 0 and not args.no_kernel_pass:
        T = 3
        g0 = 2 * n_steps
        bufs = [pool_dev[torch.from_numpy(idx[g0 + j]).to(dev)].contiguous() for j in range(T + 9)]
        # warm all four prefetch slots (the timed arms use two): their buffers are allocated on first use
        for j in range(4):
            ctx.pipeline_prefetch_dev(bufs[j].data_ptr(), ROWS, COLS)
        ctx.pipeline_wait()
        for j in range(4):
            ctx.pipeline_discard(bufs[j].data_ptr())
        # stage 3 alone: scan registration finished before the clock starts (no overlap), mapping stage timed on its own stream
        s3 = []
        for j in range(4):
            ctx.pipeline_prefetch_dev(bufs[j].data_ptr(), ROWS, COLS); ctx.pipeline_wait()
            ctx.timer_record(0)
            ctx.pipeline_step_dev(bufs[j].data_ptr(), ROWS, COLS, odom[g0 + j], mapped, stats)
            ctx.timer_record(1)
            s3.append(ctx.timer_elapsed_ms())
        # stage 1 alone: four sweeps' scan registration back to back on the side stream
        ctx.timer_record_side(0)
        for j in range(4, 8):
            ctx.pipeline_prefetch_dev(bufs[j].data_ptr(), ROWS, COLS)
        ctx.timer_record_side(1)
        ctx.pipeline_wait()
        s1 = ctx.timer_elapsed_ms() / 4.0
        for j in range(4, 8):
            ctx.pipeline_discard(bufs[j].data_ptr())
        stage = {"stage1_alone_points_per_s": S * NPTS / (s1 * 1e-3), "stage1_alone_ms_per_step": s1,
                 "stage3_alone_points_per_s": S * NPTS / (float(np.mean(s3[1:])) * 1e-3), "stage3_alone_ms_per_step": float(np.mean(s3[1:]))}
        # per-kernel pass: every launch between two events on its stream; the two stages run one after the other here (no
        # prefetch in flight during a step), so a kernel's interval holds no time spent waiting for the other stage's CTAs
        ctx.timeline_enable(True)
        tc = dict(qi=0, q=0, ins=0, feat=0)
        for j in range(T):
            ctx.pipeline_prefetch_dev(bufs[8 + j].data_ptr(), ROWS, COLS); ctx.pipeline_wait()
            ctx.pipeline_step_dev(bufs[8 + j].data_ptr(), ROWS, COLS, odom[g0 + 8 + j], mapped, stats)
            c = ctx.last_step_counters()
            tc["qi"] += c["query_iters"]; tc["q"] += c["queries"]; tc["ins"] += c["inserted"]; tc["feat"] += c["features"]
        rep = ctx.timeline_report(); ctx.timeline_enable(False)
        peak, _ = measured_peak()
        counters = dict(raw=float(S * NPTS), feat=tc["feat"] / T, qi=tc["qi"] / T, q=tc["q"] / T, ins=tc["ins"] / T)
        rows = []
        for l in rep.strip().splitlines():
            name, us, n = l.rsplit(" ", 2)
            rows.append((name, float(us) / T / 1e3, int(n) / T))
        tot = sum(r[1] for r in rows)
        for name, ms_k, n in rows:
            b = kernel_bytes(name, counters)
            e = {"kernel": name, "ms_per_step": ms_k, "launches_per_step": n, "share_of_kernel_time": ms_k / tot if tot else None,
                 "algorithmic_bytes_per_step": b, "achieved_gbs": (b / (ms_k * 1e-3) / 1e9) if (b and ms_k > 0) else None}
            e["frac"] = e["achieved_gbs"] / peak if e["achieved_gbs"] else None
            kernels.append(e)
        kernels.sort(key=lambda e: -e["ms_per_step"])
        log("[kernels] %d steps launch by launch, %.3f ms of kernel time per step" % (T, tot))
        for e in kernels:
            log("  %-44s %8.1f us/step x%-5.1f %5.1f%%  %s" % (e["kernel"][:44], 1e3 * e["ms_per_step"], e["launches_per_step"], 100 * e["share_of_kernel_time"],
                                                            ("%.0f GB/s = %.4f of peak" % (e["achieved_gbs"], e["frac"])) if e["frac"] else ""))

    # ---- single-stream latency (one LiDAR, host sweep in -> host pose out) ---------------------------------------------
    lat = {}
    if rank == 0 and not args.no_latency:
        c1 = cmb.Context(device=local_rank, **CFG)
        c1.mapping_create(1, max_corner_points=cap_c, max_surf_points=cap_s)
        for o in range(0, len(ms), chunk):
            c1.map_insert([mc if o == 0 else mc[:0]], [ms[o:o + chunk]], [eye])
        m1 = np.empty((1, 12), np.float32); s1_ = (cmb.MatchStats * 1)()
        l_all, l_s1, l_s3 = [], [], []
        for k in range(4 + 20):
            fi = k % len(frames)
            fr = torch.from_numpy(np.ascontiguousarray(frames[fi:fi + 1])).pin_memory()
            od = pack_isos([noisy_odom(poses, fi, rng, synth)])
            t0 = time.perf_counter()
            c1.pipeline_step_packed(fr.numpy(), od, m1, s1_)
            t1 = time.perf_counter()
            # the two stages on their own: upload + scan registration until the features are on the device, then the mapping stage
            fi2 = (k + 7) % len(frames)
            fr2 = torch.from_numpy(np.ascontiguousarray(frames[fi2:fi2 + 1])).pin_memory()
            od2 = pack_isos([noisy_odom(poses, fi2, rng, synth)])
            t2 = time.perf_counter()
            c1.pipeline_prefetch(fr2.numpy()); c1.pipeline_wait()
            t3 = time.perf_counter()
            c1.pipeline_step_packed(fr2.numpy(), od2, m1, s1_)
            t4 = time.perf_counter()
            if k >= 4:
                l_all.append(1e3 * (t1 - t0)); l_s1.append(1e3 * (t3 - t2)); l_s3.append(1e3 * (t4 - t3))
        lat = {"end_to_end": float(np.median(l_all)), "stage1": float(np.median(l_s1)), "stage3": float(np.median(l_s3))}
        c1.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        step_bytes = 16.0 * S * K * NPTS + 16.0 * cnt["feat"] + 96.0 * cnt["qi"] + 32.0 * cnt["ins"]
        step_gbs = step_bytes / (ms_total * 1e-3) / 1e9
        # every kernel >= 5 % of the step's kernel time, then the next longest ones until the list covers >= 92 % of it
        big, cum = [], 0.0
        for e in kernels:   # sorted, longest first
            sh = e["share_of_kernel_time"] or 0.0
            if sh >= 0.05 or cum < 0.92:
                big.append(e); cum += sh
        dom = next((e for e in kernels if e["frac"]), None)          # longest kernel with an algorithmic-bytes figure
        if kernels and kernels[0]["frac"]:
            dom = kernels[0]
        roof = {"bound": "hbm", "kernel": dom["kernel"] if dom else None, "achieved": dom["achieved_gbs"] if dom else step_gbs,
                "peak": peak, "unit": "GB/s", "frac": (dom["frac"] if dom else step_gbs / peak),
                "traffic": ncu_traffic(dom["kernel"]) if dom else None, "peak_source": peak_src,
                "kernel_ms_per_step": dom["ms_per_step"] if dom else None,
                "kernel_share_of_step": (dom["ms_per_step"] / (ms_max / K)) if dom else None,
                "how": "per-kernel CUDA events on the launching stream in a separate launch-by-launch pass of 3 steps after the timed arms "
                       "(the timed arms replay CUDA graphs); algorithmic bytes per kernel: DESIGN.md section 5",
                "step_algorithmic_gbs": step_gbs, "step_frac": step_gbs / peak,
                "kernels": big, "kernels_listed_share": sum(e["share_of_kernel_time"] for e in big) if big else None,
                "kernel_time_ms_per_step": sum(e["ms_per_step"] for e in kernels) if kernels else None}
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "config": 2, "streams_per_gpu": S, "points_per_sweep": NPTS,
                       "map_points_per_stream": int(sum(map_pts)), "map_points_per_stream_after": int(sum(map_after)),
                       "map_samples_inserted": int(len(mc) + len(ms)),
                       "frame_leaf": [CFG["filter_corner"], CFG["filter_surf"]], "map_leaf": [CFG["map_filter_corner"], CFG["map_filter_surf"]],
                       "mean_gn_iterations": float(np.mean(iters)), "converged_frac": conv,
                       "queries_per_sweep": cnt["q"] / float(S * K), "sweeps": "fresh trajectory: no stream registers a sweep twice",
                       "l2": "working set (S maps + S sweeps) > 126 MB L2, no explicit flush", "e2e_sweeps_ahead": A,
                       "e2e_pinned_buffers": nbuf, "graphs_built_in_timed_region": [gb1[0] - gb0[0], gb1[1] - gb0[1]],
                       "numa_bind": numa, "parallelism": "streams sharded over ranks, no collective"},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": e2e_bytes,
                    "d2h_bytes_per_step": int(S * 48 + S * C.sizeof(cmb.MatchStats)), "input": e2e_input},
            "e2e_xyzi16": {"value": e2e16_value, "unit": "points/s", "h2d_bytes_per_step": int(S * NPTS * 16 + S * 48),
                           "input": "pinned packed cm_point sweeps (16 B per point) through cm_pipeline_prefetch_host / cm_pipeline_step_host"},
            "e2e_pcl_layout": e2e_pcl,
            "stage_alone": stage,
            "p50_latency_ms": lat.get("end_to_end"), "p50_latency_ms_by_stage": lat,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs (1-based); 2 = the headline")
    ap.add_argument("--streams", type=int, default=128, help="LiDAR streams per GPU (config 2); r01 and the first half of r02 used 64")
    ap.add_argument("--pool", type=int, default=0, help="distinct synthetic sweeps generated on the host (0: one per stream-step)")
    ap.add_argument("--ahead", type=int, default=2, help="end-to-end arm: sweeps uploaded ahead of the one being registered (1..3)")
    ap.add_argument("--pinned-gb", type=float, default=6.0, help="end-to-end arm: pinned host memory for sweep buffers per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-kernel-pass", action="store_true")
    ap.add_argument("--no-xyz12-arm", action="store_true", help="report the 16-byte end-to-end arm as e2e instead of the xyz-only 12-byte entry")
    ap.add_argument("--no-pcl-arm", action="store_true", help="skip the end-to-end arm that starts from pageable 32-byte-stride clouds")
    ap.add_argument("--no-numa-bind", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    synth = importlib.import_module(PKG + ".synth")
    if args.impl == "reference":
        return run_reference(args, synth, rank)
    if args.config == 2:
        return run_config2(args, synth, rank, world, local_rank)
    extra = importlib.import_module("bench_configs")
    return extra.run(args, synth, rank, world, local_rank)


if __name__ == "__main__":
    main()
