#!/usr/bin/env python
"""Where the single-sweep latency goes: ONE HDL-64E stream, one synchronous cm_pipeline_step_host per sweep (host sweep in, host pose
out), wall time per step and the device time of every kernel / copy of the step (event-timed launch by launch, no graphs).
usage: python tools/latency_breakdown.py"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
synth = importlib.import_module(bench.PKG + ".synth"); cmb = importlib.import_module(bench.PKG)
mc, ms, frames, poses = bench.make_workload(12, synth)
ctx = cmb.Context(device=0, **bench.CFG)
ctx.mapping_create(1, max_corner_points=max(4 * len(mc), 100000), max_surf_points=int(1.6 * len(ms)) + 200000)
eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
for o in range(0, len(ms), 1 << 20):
    ctx.map_insert([mc if o == 0 else mc[:0]], [ms[o:o + (1 << 20)]], [eye])
rng = np.random.default_rng(7)
host = [torch.from_numpy(np.ascontiguousarray(frames[k:k + 1])).pin_memory().numpy() for k in range(len(frames))]
mapped = np.empty((1, 12), np.float32); stats = (cmb.MatchStats * 1)()
def run(n, first=0):
    t = []
    for k in range(n):
        f = (first + k) % len(frames)
        od = bench.pack_isos([bench.noisy_odom(poses, f, rng, synth)])
        t0 = time.perf_counter()
        ctx.pipeline_step_packed(host[f], od, mapped, stats)
        t.append(1e3 * (time.perf_counter() - t0))
    return t
run(6)
t = run(12, 3)
print("graph path: p50 %.3f ms, min %.3f ms per sweep (wall, host in -> pose out), %d Gauss-Newton iterations" % (np.median(t), min(t), stats[0].iterations))
ctx.timeline_enable(True)
n = 6; t = run(n, 5)
rep = ctx.timeline_report(); ctx.timeline_enable(False)
print("launch by launch with events: p50 %.3f ms" % np.median(t))
tot = 0.0
for l in rep.strip().splitlines():
    name, us, cnt = l.rsplit(" ", 2)
    tot += float(us) / n
    print("  %-40s %8.1f us/step  x%-4.1f" % (name[:40], float(us) / n, int(cnt) / n))
print("  sum of device time %.1f us/step" % tot)
ctx.close()
