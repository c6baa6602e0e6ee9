#!/usr/bin/env python
"""Per-kernel / per-copy device time of the end-to-end arm (pinned host sweeps, 2-ahead prefetch), event-timed launch by launch.
usage: python tools/e2e_timeline.py [streams]"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
synth = importlib.import_module(bench.PKG + ".synth"); cmb = importlib.import_module(bench.PKG)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mc, ms, frames, poses = bench.make_workload(6, synth)
ctx = cmb.Context(device=0, **bench.CFG)
ctx.mapping_create(S, max_corner_points=max(4 * len(mc), 100000), max_surf_points=int(1.6 * len(ms)) + 200000)
eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
for o in range(0, len(ms), 1 << 18):
    ctx.map_insert([mc if o == 0 else mc[:0]] * S, [ms[o:o + (1 << 18)]] * S, [eye] * S)
rng = np.random.default_rng(7)
NB = 6
order = [[(3 * s + b) % len(frames) for s in range(S)] for b in range(NB)]
host = [torch.from_numpy(np.ascontiguousarray(frames[order[b]])).pin_memory().numpy() for b in range(NB)]
mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()
def run(n):
    od = [bench.pack_isos([bench.noisy_odom(poses, order[k % NB][s], rng, synth) for s in range(S)]) for k in range(n)]
    ctx.pipeline_prefetch(host[0]); ctx.pipeline_prefetch(host[1])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if k + 2 < n: ctx.pipeline_prefetch(host[(k + 2) % NB])
        ctx.pipeline_step_packed(host[k % NB], od[k], mapped, stats)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
run(5)
print("e2e: %.2f ms/step (wall)" % run(10))
ctx.timeline_enable(True)
n = 6; ms_step = run(n)
rep = ctx.timeline_report(); ctx.timeline_enable(False)
print("with per-launch events: %.2f ms/step" % ms_step)
for l in rep.strip().splitlines():
    name, us, cnt = l.rsplit(" ", 2)
    print("  %-40s %9.1f us/step  x%-4d" % (name[:40], float(us) / n, int(cnt) // n))
ctx.close()
