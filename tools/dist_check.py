#!/usr/bin/env python
"""BASELINE config 4 check (run under torchrun, one rank per GPU): the scan-to-map pipeline on ONE map sharded over the ranks
(cm_dist_init: cube-lattice ownership + sqrt(5) m halo, partial normal equations exchanged by the library's own kernel once per
Gauss-Newton iteration) against the same pipeline on one GPU holding the whole map.  Prints one JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    cmb = importlib.import_module("the-cooper-mapper_b200"); synth = importlib.import_module("the-cooper-mapper_b200.synth")
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    extent = float(os.environ.get("DIST_EXTENT", "125"))
    sc = synth.make_scene(seed=0x5EED, extent=extent, n_boxes=44, n_poles=40)
    mc, ms = synth.sample_map(sc, float(os.environ.get("DIST_SPACING", "0.18")), seed=2)      # ~4-5 M samples at the default
    NF = int(os.environ.get("DIST_FRAMES", "6"))
    traj = synth.trajectory(NF, speed=2.0)
    frames = [synth.simulate_scan(sc, R, t, "HDL-64E", seed=1000 + k) for k, (R, t) in enumerate(traj)]
    rng = np.random.default_rng(5)
    odoms = []
    for R, t in traj:
        d = np.deg2rad(rng.uniform(-0.5, 0.5, 3)); dR, _ = synth.pose_matrix(d[2], d[1], d[0])
        odoms.append(((R @ dR).astype(np.float32), (t + rng.uniform(-0.1, 0.1, 3)).astype(np.float32)))
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    cap_c, cap_s = max(8 * len(mc), 200000), int(1.3 * len(ms)) + 400000

    def run(ctx):
        chunk = 1 << 20
        for o in range(0, len(ms), chunk):
            ctx.map_insert([mc if o == 0 else mc[:0]], [ms[o:o + chunk]], [eye])
        npts = len(ctx.map_export(0, 0)[0]) + len(ctx.map_export(0, 1)[0])
        out = []; times = []
        for k in range(NF):
            t0 = time.perf_counter()
            isos, stats = ctx.pipeline_step(frames[k][None], [odoms[k]])
            times.append(time.perf_counter() - t0)
            out.append((isos[0], stats[0]))
        return npts, out, times

    # the sharded map
    idb = [cmb.Context.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(idb, src=0)
    ctx = cmb.Context(device=local, **cfg)
    ctx.dist_init(idb[0], rank, world)
    info = ctx.dist_info()
    ctx.mapping_create(1, cap_c, cap_s)
    npts, out, times = run(ctx)
    tot, ms_call = ctx.dist_allreduce(np.arange(32, dtype=np.float64) * (rank + 1), repeat=200)
    ok_sum = bool(np.array_equal(tot, np.arange(32, dtype=np.float64) * (world * (world + 1) // 2)))
    ctx.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(npts=npts, poses=[(o[0][0].tolist(), o[0][1].tolist()) for o in out], its=[o[1]["iterations"] for o in out], rows=[o[1]["rows"] for o in out]))
    if rank == 0:
        ref = cmb.Context(device=local, **cfg)
        ref.mapping_create(1, cap_c, cap_s)
        n_ref, out_ref, times_ref = run(ref)
        ref.close()
        same_ranks = all(g["poses"] == gathered[0]["poses"] for g in gathered)
        dt = [float(np.max(np.abs(np.array(gathered[0]["poses"][k][1]) - out_ref[k][0][1]))) for k in range(NF)]
        dR = [float(np.max(np.abs(np.array(gathered[0]["poses"][k][0]) - out_ref[k][0][0]))) for k in range(NF)]
        line = dict(check="sharded map == one GPU", world=world, p2p=info["p2p"], map_points_one_gpu=n_ref, map_points_per_rank=[g["npts"] for g in gathered],
                    halo_overhead=sum(g["npts"] for g in gathered) / float(n_ref) - 1.0, poses_identical_on_all_ranks=same_ranks,
                    max_abs_dt=max(dt), max_abs_dR=max(dR), bit_identical_frames=int(sum(1 for a, b in zip(dt, dR) if a == 0 and b == 0)), frames=NF,
                    iterations=gathered[0]["its"], iterations_one_gpu=[o[1]["iterations"] for o in out_ref], rows=gathered[0]["rows"], rows_one_gpu=[o[1]["rows"] for o in out_ref],
                    exchange_ok=ok_sum, exchange_ms_per_call_32_doubles=ms_call, ms_per_frame_sharded=1e3 * float(np.median(times)), ms_per_frame_one_gpu=1e3 * float(np.median(times_ref)))
        print(json.dumps(line), flush=True)
        assert same_ranks and ok_sum and max(dt) <= 1e-4 and max(dR) <= 1e-5, line
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
