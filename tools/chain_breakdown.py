#!/usr/bin/env python
"""Where the time of one cm_pipeline_chain_step_host goes (ONE VLP-16 stream, scan registration -> odometry -> mapping): wall time per
sweep on the graph path and the device time of every kernel / copy, event-timed launch by launch.
usage: python tools/chain_breakdown.py            (CHAIN_STREAMS=32: the same for a batch of streams, stream s starting s sweeps in)"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
synth = importlib.import_module(bench.PKG + ".synth"); cmb = importlib.import_module(bench.PKG)
NF = 40
S = int(os.environ.get("CHAIN_STREAMS", "1"))
sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
traj = synth.trajectory(NF, speed=0.1, yaw_amp=0.02)
frames = bench.simulate_pool(synth, sc, traj, "VLP-16", 0x1000)
ctx = cmb.Context(device=0, **bench.CFG)
ctx.mapping_create(S, 100000, 800000)
ctx.pipeline_chain_create(16, 1800)
pin = [torch.from_numpy(np.ascontiguousarray(np.stack([frames[(k + s) % NF] for s in range(S)]).astype(np.float32))).pin_memory().numpy() for k in range(NF)]
od = np.empty((S, 12), np.float32); mp = np.empty((S, 12), np.float32); ost = (cmb.OdomStats * S)(); mst = (cmb.MatchStats * S)()
t = []; its = []
for k in range(30):
    t0 = time.perf_counter(); ctx.pipeline_chain_step_packed(pin[k], od, mp, ost, mst); t.append(1e3 * (time.perf_counter() - t0)); its.append((ost[0].iterations, mst[0].iterations))
print("%d stream(s); graph path: p50 %.3f ms, min %.3f ms per step; (odometry, mapping) iterations of the last sweeps: %s" % (S, np.median(t[5:]), min(t[5:]), its[-5:]))
ctx.timeline_enable(True)
n = 8
for k in range(30, 30 + n):
    ctx.pipeline_chain_step_packed(pin[k], od, mp, ost, mst)
rep = ctx.timeline_report(); ctx.timeline_enable(False)
tot = 0.0
for l in rep.strip().splitlines():
    name, us, cnt = l.rsplit(" ", 2)
    tot += float(us) / n
    print("  %-40s %8.1f us/step  x%-4.1f" % (name[:40], float(us) / n, int(cnt) / n))
print("  sum of device time %.1f us/step" % tot)
ctx.close()
