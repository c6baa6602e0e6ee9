#!/usr/bin/env python
"""BASELINE config 4 on real GPUs: one map split into x-slabs over the ranks of a torchrun job (one process per GPU), the 32
partial sums of the normal equations all-reduced over NCCL every Gauss-Newton iteration, redundant solve on every rank.
Checks that every rank's pose is bit-identical to the unsharded single-GPU solve and reports the time per registration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_nccl_check.py [map_points]
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    target = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cmb = importlib.import_module("the-cooper-mapper_b200")
    dmod = importlib.import_module("the-cooper-mapper_b200.dist")
    synth = importlib.import_module("the-cooper-mapper_b200.synth")
    # same scene / map / sweep on every rank (seeded); the map spacing is chosen to reach ~`target` points
    sc = synth.make_scene(seed=0x5EED0004 & 0xFFFF, extent=200.0, n_boxes=120, n_poles=80)
    spacing = 0.4 * np.sqrt(2.6e6 / target)            # ~2.6M points at 0.4 m for this scene
    mc, ms = synth.sample_map(sc, float(spacing), seed=4)
    R, t = synth.pose_matrix(0.03, 0.0, 0.0, (1.0, 0.2, 0.0))
    fr = synth.simulate_scan(sc, R, t, "HDL-64E", seed=23)
    ctx = cmb.Context(device=local, map_filter_corner=float(spacing), map_filter_surf=float(spacing))
    out = ctx.scanreg_organised(fr)
    corner = ctx.voxel_filter(out["lessSharp"], 0.4); surf = ctx.voxel_filter(out["lessFlat"], 0.8)
    truth = np.array([0.0, 0.0, 0.03, 1.0, 0.2, 0.0], np.float32)
    init = truth + np.array([0.004, -0.003, 0.006, 0.06, -0.05, 0.04], np.float32)
    bounds = dmod.slab_bounds(ms, world)
    lo, hi = dmod.own_box(bounds, rank)
    sm = dmod.ShardedScanMatch(ctx, dmod.shard_cloud(mc, bounds, rank), dmod.shard_cloud(ms, bounds, rank), len(mc), len(ms), lo, hi)
    ok, pose, stats = sm.scanMatchScan(corner, surf, init)          # warm-up + result
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ok, pose, stats = sm.scanMatchScan(corner, surf, init)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    poses = [torch.zeros(6, device="cuda") for _ in range(world)]
    mine = torch.from_numpy(pose.copy()).cuda()
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_gather(poses, mine)
    else:
        poses = [mine]
    if rank == 0:
        ref_ctx = cmb.Context(device=local, map_filter_corner=float(spacing), map_filter_surf=float(spacing))
        ref_pose, ref_stats, _ = ref_ctx.match_stateless(mc, ms, corner, surf, init)
        same = all(np.array_equal(p.cpu().numpy().view(np.uint32), ref_pose.view(np.uint32)) for p in poses)
        print(json.dumps({"check": "sharded map == single map (bit-identical pose on every rank)", "ok": bool(same), "world": world,
                          "map_points": int(len(mc) + len(ms)), "shard_points_rank0": int(len(dmod.shard_cloud(ms, bounds, 0))),
                          "queries": int(len(corner) + len(surf)), "iterations": stats["iterations"], "ref_iterations": ref_stats["iterations"],
                          "ms_per_registration_max_over_ranks": 1e3 * float(tt.item()),
                          "pose_error_m": float(np.linalg.norm(ref_pose[3:] - truth[3:]))}), flush=True)
        assert same and stats["iterations"] == ref_stats["iterations"]
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
