"""bench.py --config {1,3,4,5}: the other BASELINE.json configurations (config 2, the headline, lives in bench.py).

  1  VLP-16 16x1800 sequence, ONE stream, scan registration -> laserOdometry -> laserMapping, every stage on the GPU through the
     host-buffer C ABI, next to the same chain of the CPU oracle on one host core (the reference drives each stage from one thread)
  3  256 independent VLP-16 streams, stream i on rank i mod G (strong scaling: the 256 streams are the job), no cross-GPU traffic
  4  ONE 50 M-point map sharded over the ranks by the FeatureMap cube lattice (cm_dist_init), HDL-64E sweeps, the partial normal
     equations exchanged once per Gauss-Newton iteration by the library's own kernel over NVLink peer memory
  5  sparse tilted-RPLidar sweeps (12 revolutions x 800 points): the low-feature / degenerate paths
Every line keeps bench.py's contract keys; `config.workload` names the configuration.
"""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

import bench as B

PKG = B.PKG


def _barrier(world, dist, torch):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world, dist, torch, dev):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _line(args, world, value, ms_per_step, scaling, workload, extra_cfg, e2e=None, **more):
    line = {"metric": B.METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(workload=workload, config=args.config, **extra_cfg),
            "roofline": None, "cpu_baseline": None,
            "e2e": e2e or {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(more)
    return line


# ---------------------------------------------------------------------------------------------------------------------------------
def run_config3(args, synth, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    TOTAL = 256
    ROWS, COLS = 16, 1800
    NP = ROWS * COLS
    K, W = args.steps, max(args.warmup, 3)
    n_steps = W + K
    mine = list(range(rank, TOTAL, world))                      # stream i -> rank i mod G
    S = len(mine)
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    P = 2 * n_steps + 8
    traj = synth.trajectory(P, speed=0.1, yaw_amp=0.02)         # 1 m/s at 10 Hz
    frames = np.stack(B.simulate_pool(synth, sc, traj, "VLP-16", 0x100)).astype(np.float32)
    poses = [(R.astype(np.float32), t.astype(np.float32)) for R, t in traj]
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cmb = importlib.import_module(PKG)
    ctx = cmb.Context(device=local_rank, **B.CFG)
    ctx.mapping_create(S, max_corner_points=60000, max_surf_points=400000)
    rng = np.random.default_rng(7 + rank)
    idx = np.array([[(5 * i + g) % P for i in mine] for g in range(2 * n_steps)])
    odom = [B.pack_isos([B.noisy_odom(poses, idx[g][j], rng, synth) for j in range(S)]) for g in range(2 * n_steps)]
    pool_dev = torch.from_numpy(frames).to(dev)
    step_dev = [pool_dev[torch.from_numpy(idx[g]).to(dev)].contiguous() for g in range(n_steps)]
    mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()
    ctx.pipeline_prefetch_dev(step_dev[0].data_ptr(), ROWS, COLS)
    for k in range(W):
        ctx.pipeline_prefetch_dev(step_dev[k + 1].data_ptr(), ROWS, COLS)
        ctx.pipeline_step_dev(step_dev[k].data_ptr(), ROWS, COLS, odom[k], mapped, stats)
    _barrier(world, dist, torch)
    l0 = ctx.launch_count()
    ctx.timer_record(0)
    for k in range(W, W + K):
        if k + 1 < W + K:
            ctx.pipeline_prefetch_dev(step_dev[k + 1].data_ptr(), ROWS, COLS)
        ctx.pipeline_step_dev(step_dev[k].data_ptr(), ROWS, COLS, odom[k], mapped, stats)
    ctx.timer_record(1)
    ms = _max_over_ranks(ctx.timer_elapsed_ms(), world, dist, torch, dev)
    launches = ctx.launch_count() - l0
    _barrier(world, dist, torch)
    value = TOTAL * K * NP / (ms * 1e-3)
    # end to end: pinned host sweeps
    host = [torch.from_numpy(np.ascontiguousarray(frames[idx[n_steps + g]])).pin_memory() for g in range(n_steps)]
    hn = [h.numpy() for h in host]

    def e2e(first, count):
        for j in range(first, min(first + 2, first + count)):
            ctx.pipeline_prefetch(hn[j])
        for k in range(first, first + count):
            if k + 2 < first + count:
                ctx.pipeline_prefetch(hn[k + 2])
            ctx.pipeline_step_packed(hn[k], odom[n_steps + k], mapped, stats)
    e2e(0, W)
    _barrier(world, dist, torch)
    ctx.timer_record(0); e2e(W, K); ctx.timer_record(1)
    ms_e = _max_over_ranks(ctx.timer_elapsed_ms(), world, dist, torch, dev)
    conv = float(np.mean([st.converged for st in stats]))
    if rank == 0:
        line = _line(args, world, value, ms / K, "strong",
                     "config 3: 256 independent VLP-16 16x1800 streams, stream i on rank i mod G, scan registration + scan-to-map (maps grow from empty)",
                     dict(streams_total=TOTAL, streams_per_gpu=S, points_per_sweep=NP, converged_frac=conv, parallelism="streams sharded over ranks, no collective"),
                     e2e={"value": TOTAL * K * NP / (ms_e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": int(S * NP * 16 + S * 48),
                          "d2h_bytes_per_step": int(S * 48 + S * C.sizeof(cmb.MatchStats))}, gpu_launches=int(launches))
        B.emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------------------
def run_config4(args, synth, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    ROWS, COLS, NP = B.ROWS, B.COLS, B.NPTS
    K, W = args.steps, max(args.warmup, 3)
    n_steps = W + K
    tiles = int(os.environ.get("BENCH_C4_TILES", "7"))          # tiles x tiles copies of the 1.02 M-point base map: 7 -> 50 M points
    PITCH = 400.0                                               # a multiple of the map leaf (0.4) and of the cube size (50): identical voxelisation per tile
    sc = synth.make_scene(seed=B.SEED & 0xFFFF, extent=125.0, n_boxes=44, n_poles=40)
    mc, ms = synth.sample_map(sc, B.MAP_SPACING, seed=2)
    P = n_steps * 2 + 4
    traj = synth.trajectory(P, speed=2.0)
    frames = np.stack(B.simulate_pool(synth, sc, traj, "HDL-64E", 1000)).astype(np.float32)
    poses = [(R.astype(np.float32), t.astype(np.float32)) for R, t in traj]
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cmb = importlib.import_module(PKG)
    ctx = cmb.Context(device=local_rank, **B.CFG)
    info = dict(rank=0, nranks=1, p2p=False)
    if world > 1:
        idb = [cmb.Context.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idb, src=0)
        ctx.dist_init(idb[0], rank, world)
        info = ctx.dist_info()
    total_pts = tiles * tiles * 1021434
    share = 1.0 / world
    halo = 1.0 if world == 1 else 1.35
    ctx.mapping_create(1, max_corner_points=int(tiles * tiles * 12000 * share * halo) + 100000,
                       max_surf_points=int(tiles * tiles * 1030000 * share * halo) + 1000000)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    t0 = time.time()
    half = (tiles - 1) / 2.0
    chunk = 1 << 20
    for ti in range(tiles):
        for tj in range(tiles):
            off = np.array([(ti - half) * PITCH, (tj - half) * PITCH, 0.0, 0.0], np.float32)
            c = mc + off; s_ = ms + off
            for o in range(0, len(s_), chunk):
                ctx.map_insert([c if o == 0 else c[:0]], [s_[o:o + chunk]], [eye])
    resident = len(ctx.map_export(0, 0)[0]) + len(ctx.map_export(0, 1)[0])
    B.log("[bench c4] rank %d: %d resident map points (%.1f s to build)" % (rank, resident, time.time() - t0))
    rng = np.random.default_rng(77)                              # the SAME odometry on every rank: the ranks work on one sweep together
    odom = [B.pack_isos([B.noisy_odom(poses, g % P, rng, synth)]) for g in range(2 * n_steps)]
    fr_dev = [torch.from_numpy(frames[g % P][None]).to(dev).contiguous() for g in range(n_steps)]
    mapped = np.empty((1, 12), np.float32); stats = (cmb.MatchStats * 1)()
    for k in range(W):
        ctx.pipeline_step_dev(fr_dev[k].data_ptr(), ROWS, COLS, odom[k], mapped, stats)
    _barrier(world, dist, torch)
    l0 = ctx.launch_count()
    its = []
    ctx.timer_record(0)
    for k in range(W, W + K):
        ctx.pipeline_step_dev(fr_dev[k].data_ptr(), ROWS, COLS, odom[k], mapped, stats)
        its.append(stats[0].iterations)
    ctx.timer_record(1)
    ms_t = _max_over_ranks(ctx.timer_elapsed_ms(), world, dist, torch, dev)
    launches = ctx.launch_count() - l0
    value = K * NP / (ms_t * 1e-3)
    host = [torch.from_numpy(np.ascontiguousarray(frames[(n_steps + g) % P][None])).pin_memory() for g in range(n_steps)]
    for k in range(W):
        ctx.pipeline_step_packed(host[k].numpy(), odom[n_steps + k], mapped, stats)
    _barrier(world, dist, torch)
    ctx.timer_record(0)
    for k in range(W, W + K):
        ctx.pipeline_step_packed(host[k].numpy(), odom[n_steps + k], mapped, stats)
    ctx.timer_record(1)
    ms_e = _max_over_ranks(ctx.timer_elapsed_ms(), world, dist, torch, dev)
    ex_ms = None
    if world > 1:
        _, ex_ms = ctx.dist_allreduce(np.zeros(32), repeat=400)
    pose = mapped[0].copy()
    allp = [None] * world
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allp, pose.tolist()); dist.all_gather_object(allr, resident)
    else:
        allp = [pose.tolist()]; allr = [resident]
    if rank == 0:
        line = _line(args, world, value, ms_t / K, "strong",
                     "config 4: HDL-64E 64x2048 sweeps, scan registration + scan-to-map against ONE %.1f M-point map sharded over the ranks" % (total_pts / 1e6),
                     dict(map_points_total=int(total_pts), map_points_per_rank=allr, halo_overhead=sum(allr) / float(total_pts) - 1.0,
                          tiles=tiles * tiles, area_km2=(tiles * PITCH / 1000.0) ** 2, ownership="50 m cube (i, j, k) -> rank (i + 3 j + 5 k) mod G, sqrt(5) m halo",
                          exchange="library kernel over CUDA-IPC peer mailboxes (NVLink), ranks added in rank order" if info["p2p"] else
                                   ("ncclAllGather + ordered sum" if world > 1 else "none (one rank)"),
                          exchange_us_per_call_32_doubles=(1e3 * ex_ms) if ex_ms is not None else None, exchanges_per_sweep=11 if world > 1 else 0,
                          mean_gn_iterations=float(np.mean(its)), poses_identical_on_all_ranks=all(p == allp[0] for p in allp),
                          parallelism="one map over all ranks; every rank evaluates the queries in its cubes"),
                     e2e={"value": K * NP / (ms_e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": int(NP * 16 + 48) * world,
                          "d2h_bytes_per_step": int(48 + C.sizeof(cmb.MatchStats)) * world}, gpu_launches=int(launches))
        B.emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------------------
def run_config1(args, synth, rank, world, local_rank):
    if rank != 0:
        return
    from oracle import oracle_py as O
    cmb = importlib.import_module(PKG)
    NF = int(os.environ.get("BENCH_C1_FRAMES", "100"))
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    traj = synth.trajectory(NF, speed=0.1, yaw_amp=0.02)        # 1 m/s at 10 Hz
    frames = B.simulate_pool(synth, sc, traj, "VLP-16", 0x1000)
    cfg = B.CFG
    # CPU: the oracle chain on one core (the reference runs each stage on one thread)
    ncpu = min(NF, int(os.environ.get("BENCH_C1_CPU_FRAMES", "40")))
    oo = O.Odometry(fast=True); om = O.Mapping(map_params=B.ORACLE_MAP, fast=True)
    tc = []
    for k in range(ncpu):
        t0 = time.perf_counter()
        f = O.scanreg_organised(frames[k], fast=True)
        od = oo.process(f["sharp"], f["lessSharp"], f["flat"], f["lessFlat"])
        om.process(od["R"], od["t"], od["corner_last"], od["surf_last"])
        tc.append(time.perf_counter() - t0)
    cpu_rate = 28800 / float(np.mean(tc[2:]))
    c_sr = cmb.Context(device=local_rank)
    c_od = cmb.Context(device=local_rank); c_od.odometry_reset()
    c_mp = cmb.Context(device=local_rank, **cfg); c_mp.mapping_create(1, 100000, 800000)
    tg = []
    l0 = c_sr.launch_count()
    for k in range(NF):
        t0 = time.perf_counter()
        g = c_sr.scanreg_organised(frames[k])
        go = c_od.odometry_process(g["sharp"], g["lessSharp"], g["flat"], g["lessFlat"])
        c_mp.mapping_process([(go["R"], go["t"])], [go["corner_last"]], [go["surf_last"]])
        tg.append(time.perf_counter() - t0)
    launches = c_sr.launch_count() - l0
    warm = max(3, args.warmup)
    rate3 = 28800 / float(np.mean(tg[warm:]))
    three_calls = {"value": rate3, "unit": "points/s", "p50_ms": 1e3 * float(np.median(tg[warm:])),
                   "note": "cm_scanreg_organised_host + cm_odometry_process_host + cm_mapping_process_host: every feature cloud crosses PCIe twice"}
    # the call a user makes: the three stages in one (cm_pipeline_chain_step_host), the clouds stay on the device
    import torch
    c1 = cmb.Context(device=local_rank, **cfg)
    c1.mapping_create(1, 100000, 800000)
    c1.pipeline_chain_create(16, 1800)
    od1 = np.empty((1, 12), np.float32); mp1 = np.empty((1, 12), np.float32)
    ost1 = (cmb.OdomStats * 1)(); mst1 = (cmb.MatchStats * 1)()
    pin1 = [torch.from_numpy(np.ascontiguousarray(frames[k][None].astype(np.float32))).pin_memory().numpy() for k in range(NF)]
    tg = []
    l0 = c1.launch_count()
    for k in range(NF):
        t0 = time.perf_counter()
        c1.pipeline_chain_step_packed(pin1[k], od1, mp1, ost1, mst1)
        tg.append(time.perf_counter() - t0)
    c1.mapping_sync()
    launches = c1.launch_count() - l0
    c1.close()
    rate = 28800 / float(np.mean(tg[warm:]))
    # the same three-stage chain BATCHED over S independent streams (stream s starts s frames into the sequence): one launch set
    # per stage for all streams (cm_scanreg_organised_host, cm_odometry_batch_process_host, cm_mapping_process_host)
    SB = int(os.environ.get("BENCH_C1_STREAMS", "32"))
    NB = min(NF - 1, int(os.environ.get("BENCH_C1_BATCH_FRAMES", "24")))
    cb = cmb.Context(device=local_rank, **cfg)
    cb.mapping_create(SB, 100000, 800000)
    cb.odometry_batch_create(SB, 4000, 28800, 8000, 28800)
    fstack = np.stack(frames).astype(np.float32)
    tb = []
    for k in range(NB):
        fr = fstack[[(k + s) % NF if (k + s) < NF else NF - 1 for s in range(SB)]]
        t0 = time.perf_counter()
        gs = cb.scanreg_organised(fr)
        go = cb.odometry_batch_process([g["sharp"] for g in gs], [g["lessSharp"] for g in gs], [g["flat"] for g in gs], [g["lessFlat"] for g in gs])
        cb.mapping_process([(o["R"], o["t"]) for o in go], [o["corner_last"] for o in go], [o["surf_last"] for o in go])
        tb.append(time.perf_counter() - t0)
    batched = {"streams": SB, "frames": NB, "value": SB * 28800 / float(np.mean(tb[warm:])), "unit": "points/s",
               "ms_per_step": 1e3 * float(np.mean(tb[warm:])),
               "note": "scan registration -> batched odometry -> mapping for all streams per call, through host buffers and the ctypes layer "
                       "(feature clouds cross PCIe between the stages; the Python marshalling of the clouds is inside the time)"}
    cb.close()
    # the same batch through cm_pipeline_chain_step_host: ONE call per sweep set, the clouds stay on the device between the stages
    import torch
    cc = cmb.Context(device=local_rank, **cfg)
    cc.mapping_create(SB, 100000, 800000)
    cc.pipeline_chain_create(16, 1800)
    od_o = np.empty((SB, 12), np.float32); mp_o = np.empty((SB, 12), np.float32)
    ost = (cmb.OdomStats * SB)(); mst = (cmb.MatchStats * SB)()
    pinned = [torch.from_numpy(np.ascontiguousarray(fstack[[(k + s) % NF if (k + s) < NF else NF - 1 for s in range(SB)]])).pin_memory().numpy()
              for k in range(NB)]
    tcn = []
    cc.pipeline_prefetch(pinned[0])
    for k in range(NB):
        t0 = time.perf_counter()
        if k + 1 < NB:
            cc.pipeline_prefetch(pinned[k + 1])     # the next sweep set uploads and is scan-registered beside this one's odometry + mapping
        cc.pipeline_chain_step_packed(pinned[k], od_o, mp_o, ost, mst)
        tcn.append(time.perf_counter() - t0)
    cc.mapping_sync()
    batched["device_chain"] = {"value": SB * 28800 / float(np.mean(tcn[warm:])), "unit": "points/s", "ms_per_step": 1e3 * float(np.mean(tcn[warm:])),
                               "note": "cm_pipeline_prefetch_host (next sweeps) + cm_pipeline_chain_step_host: pinned sweeps in, two poses per stream out, "
                                       "feature clouds never leave the device"}
    cc.close()
    line = _line(args, 1, rate, 1e3 * float(np.mean(tg[warm:])), "weak",
                 "config 1: VLP-16 16x1800 sequence, ONE stream, scan registration -> laserOdometry -> laserMapping in one call per sweep (cm_pipeline_chain_step_host: host sweep in, two poses out)",
                 dict(frames=NF, points_per_sweep=28800, p50_ms=1e3 * float(np.median(tg[warm:])), streams=1, three_calls=three_calls, batched=batched,
                      note="single stream: latency-bound by construction (three stages, ~35 dependent kernel rounds per sweep); throughput configs are 2 and 3"),
                 gpu_launches=int(launches))
    line["steps"] = NF - warm; line["warmup"] = warm
    line["cpu_baseline"] = {"value": cpu_rate, "unit": "points/s", "cores": 1, "kind": "port",
                            "sample": "%d sweeps through the oracle chain (scan registration, odometry, mapping; -O3, reference nanoflann) on one core, p50 %.1f ms"
                                      % (ncpu, 1e3 * float(np.median(tc[2:])))}
    line["p50_latency_ms"] = 1e3 * float(np.median(tg[warm:]))
    B.emit(line)
    for c in (c_sr, c_od, c_mp):
        c.close()


# ---------------------------------------------------------------------------------------------------------------------------------
def run_config5(args, synth, rank, world, local_rank):
    if rank != 0:
        return
    cmb = importlib.import_module(PKG)
    tilts = np.linspace(-30.0, 30.0, 12)
    cfg = dict(filter_corner=0.2, filter_surf=0.4, map_filter_corner=0.2, map_filter_surf=0.4, blind_radius=0.3)
    scenes = {"corridor": synth.make_scene(seed=1, extent=40.0, corridor=True),
              "wall": synth.Scene([[6.0, 0.0, 0.0, 0.5, 30.0, 0.0, 8.0]], [], 40.0), "field": synth.Scene([], [], 40.0)}
    NF = max(args.steps + args.warmup, 8)
    S = len(scenes)
    traj = synth.trajectory(NF, speed=0.15)
    fr = np.stack([np.stack([synth.simulate_scan(sc, R, t, tilt_deg=tilts, cols=800, seed=500 + k, dropout=0.05) for sc in scenes.values()])
                   for k, (R, t) in enumerate(traj)]).astype(np.float32)            # [frame][scene]
    ctx = cmb.Context(device=local_rank, **cfg)
    ctx.mapping_create(S, 50000, 200000)
    seen = {}
    tg = []
    l0 = ctx.launch_count()
    for k, (R, t) in enumerate(traj):
        od = [(R.astype(np.float32), t.astype(np.float32))] * S
        t0 = time.perf_counter()
        isos, stats = ctx.pipeline_step(fr[k], od)
        tg.append(time.perf_counter() - t0)
        for name, st in zip(scenes, stats):
            key = "%s:%s%s" % (name, {0: "ok", 1: "too_few_ref", 2: "too_few_matches", 3: "not_converged", 4: "low_score"}[st["status"]], "+degenerate" if st["degenerate"] else "")
            seen[key] = seen.get(key, 0) + 1
    warm = max(3, args.warmup)
    npts = 12 * 800 * S
    rate = npts / float(np.mean(tg[warm:]))
    line = _line(args, 1, rate, 1e3 * float(np.mean(tg[warm:])), "weak",
                 "config 5: tilted RPLidar-A2-like sweeps (12 revolutions x 800 points, +-30 deg nod), sparse scenes (corridor, single wall, open field), scan registration + scan-to-map through the host-buffer C ABI",
                 dict(streams=S, points_per_sweep=12 * 800, outcomes=seen, p50_ms=1e3 * float(np.median(tg[warm:]))), gpu_launches=int(ctx.launch_count() - l0))
    B.emit(line)
    ctx.close()


def run(args, synth, rank, world, local_rank):
    return {1: run_config1, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](args, synth, rank, world, local_rank)
