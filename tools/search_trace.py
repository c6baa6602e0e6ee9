#!/usr/bin/env python
"""Per-warp timeline of search_kernel (development aid): where does one launch's time go?
usage: python tools/search_trace.py [streams] [iter]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import torch
    synth = importlib.import_module(bench.PKG + ".synth")
    cmb = importlib.import_module(bench.PKG)
    mc, ms, frames, poses = bench.make_workload(6, synth)
    ctx = cmb.Context(device=0, **bench.CFG)
    ctx.mapping_create(S, max_corner_points=max(4 * len(mc), 100000), max_surf_points=int(1.6 * len(ms)) + 200000)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    chunk = 1 << 18
    for o in range(0, len(ms), chunk):
        ctx.map_insert([mc if o == 0 else mc[:0]] * S, [ms[o:o + chunk]] * S, [eye] * S)
    rng = np.random.default_rng(7)
    dev = torch.device("cuda", 0)
    pool = torch.from_numpy(frames).to(dev)
    mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()
    for k in range(5):
        order = [(3 * s + k) % len(frames) for s in range(S)]
        fr = pool[torch.tensor(order, device=dev)].contiguous()
        od = bench.pack_isos([bench.noisy_odom(poses, order[s], rng, synth) for s in range(S)])
        if k == 4:
            ctx._check(ctx.L.cm_debug_search_trace_enable(ctx.h, C.c_int(it)))
        ctx.pipeline_step_dev(fr.data_ptr(), bench.ROWS, bench.COLS, od, mapped, stats)
    n = C.c_size_t(0)
    ctx._check(ctx.L.cm_debug_search_trace_read(ctx.h, None, C.c_size_t(0), C.byref(n)))
    buf = np.zeros(n.value, np.uint64)
    ctx._check(ctx.L.cm_debug_search_trace_read(ctx.h, buf.ctypes.data_as(C.c_void_p), C.c_size_t(len(buf)), C.byref(n)))
    w = buf.reshape(-1, 4)
    w = w[w[:, 0] > 0]
    t0 = w[:, 0].astype(np.int64); t1 = w[:, 1].astype(np.int64)
    mx = (w[:, 2] >> np.uint64(32)).astype(np.int64); sm = (w[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
    hard = (w[:, 3] >> np.uint64(32)).astype(np.int64); smid = ((w[:, 3] >> np.uint64(16)) & np.uint64(0xffff)).astype(np.int64)
    corner = (w[:, 3] & np.uint64(1)).astype(np.int64)
    base = t0.min()
    dur = (t1 - t0) / 1e3
    print("warps %d, span %.1f us, start spread %.1f us; duration us: mean %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f" % (
        len(w), (t1.max() - base) / 1e3, (t0.max() - base) / 1e3, dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
    print("candidates/lane mean %.1f; max-over-lanes per warp: mean %.1f p90 %d max %d; hard queries %d of %d (corner warps: %d hard, surf warps: %d hard)" % (
        sm.sum() / (32.0 * len(w)), mx.mean(), np.percentile(mx, 90), mx.max(), hard.sum(), 32 * len(w), hard[corner == 1].sum(), hard[corner == 0].sum()))
    order = np.argsort(-dur)[:12]
    for i in order:
        print("  warp: start %.1f us dur %.1f us, max cand %d, sum cand %d, hard %d, corner %d, sm %d" % ((t0[i] - base) / 1e3, dur[i], mx[i], sm[i], hard[i], corner[i], smid[i]))
    c = np.corrcoef(mx, dur)[0, 1]
    print("corr(max candidates, duration) = %.2f;  us per 4 candidates (fit) = %.3f" % (c, 4 * np.polyfit(mx, dur, 1)[0]))
    ends = np.sort((t1 - base) / 1e3)
    print("fraction of warps finished by 25/50/75%% of the span: %.2f %.2f %.2f" % tuple(np.searchsorted(ends, q * ends[-1]) / len(ends) for q in (0.25, 0.5, 0.75)))
    ctx.close()


if __name__ == "__main__":
    main()
