"""GPU parity: exact 5-NN and the scan-to-map solver against the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from conftest import frame_features

pytestmark = pytest.mark.gpu


def _rand_cloud(rng, n, extent=20.0):
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(-extent, extent, (n, 3))
    p[:, 2] *= 0.2
    p[:, 3] = np.arange(n)
    return p


@pytest.mark.parametrize("n,cell", [(5000, 1.2), (20000, 0.7), (300, 2.4), (7, 1.0), (3, 1.0)])
def test_knn5_exact(ctx, oracle, n, cell):
    rng = np.random.default_rng(n)
    pts = _rand_cloud(rng, n, 20.0 if n > 10 else 1.0)
    q = pts[rng.integers(0, n, 800), :3] + rng.normal(0, 0.3, (800, 3)).astype(np.float32)
    idx, d2 = ctx.knn5(pts, q, cell=cell, gate=5.0)
    oi, od = oracle.knn(pts, q, 5, nanoflann=False)
    ok = od[:, 4] < 5.0           # the reference only uses queries whose 5th neighbour passes the gate
    assert ok.sum() > 0 or n < 5
    if n < 5:
        assert np.all(idx[:, n:] == -1) and np.all(d2[:, 4] >= 5.0)
    assert np.array_equal(idx[ok], oi[ok])
    assert np.array_equal(d2[ok], od[ok])          # bit-exact float distances
    assert np.all(d2[~ok][:, 4] >= 5.0)            # rejected queries stay rejected
    if oracle.lib().has_nanoflann:                 # the reference's own KD-tree agrees on the sets
        ni, nd = oracle.knn(pts, q, 5, nanoflann=True)
        assert np.array_equal(np.sort(ni[ok], 1), np.sort(idx[ok], 1))


def test_knn5_ties_by_index(ctx, oracle):
    # duplicated points: equal distances must be ordered by index
    rng = np.random.default_rng(5)
    base = _rand_cloud(rng, 400, 5.0)
    pts = np.concatenate([base, base, base])
    q = base[:200, :3] + 0.01
    idx, d2 = ctx.knn5(pts, q, cell=1.0, gate=5.0)
    oi, od = oracle.knn(pts, q, 5, nanoflann=False)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def _compare_match(ctx, oracle, mc, ms, c, s, init):
    pg, sg, lg = ctx.match_stateless(mc, ms, c, s, init, trace=True)
    po, so, lo = oracle.scan_match(mc, ms, c, s, init, params=dict(deltaTAbort=ctx.cfg.delta_t_abort, deltaRAbort=ctx.cfg.delta_r_abort,
                                                                maxIterations=ctx.cfg.max_iterations), keep_log=True)
    assert sg["iterations"] == so["iterations"]
    assert sg["converged"] == so["converged"] and sg["degenerate"] == so["degenerate"]
    assert len(lg) == len(lo)
    for it, (a, b) in enumerate(zip(lg, lo)):
        assert np.array_equal(a["pose_in"], b["pose_in"]), it
        assert np.array_equal(a["nnCorner"], b["nnCorner"]), it      # neighbour index sets, in (d2, index) order
        assert np.array_equal(a["nnSurf"], b["nnSurf"]), it
        assert np.array_equal(a["counts"][:3], b["counts"][:3]), it
        assert np.array_equal(a["AtA"], b["AtA"]), it
        assert np.array_equal(a["AtB"], b["AtB"]), it
        assert np.array_equal(a["x"], b["x"]), it
    # north-star tolerance (1e-4 m, 1e-5 rad); in practice the poses are bit-identical
    assert np.all(np.abs(pg[:3] - po[:3]) <= 1e-5) and np.all(np.abs(pg[3:] - po[3:]) <= 1e-4)
    assert np.array_equal(pg, po)
    return pg, sg


@pytest.mark.parametrize("model", ["VLP-16", "HDL-64E"])
def test_match_stateless_bit_exact(ctx, oracle, synth, scene_small, model):
    sc, mc, ms = scene_small
    c, s, truth = frame_features(synth, oracle, sc, (0.05, 0.01, -0.008, (3.0, 0.4, 0.1)), model)
    init = truth + np.array([0.005, -0.004, 0.008, 0.08, -0.06, 0.05], np.float32)
    pg, sg = _compare_match(ctx, oracle, mc, ms, c, s, init)
    assert sg["converged"]
    assert np.all(np.abs(pg[3:] - truth[3:]) < 0.03) and np.all(np.abs(pg[:3] - truth[:3]) < 2e-3)


def test_match_identity_converges_immediately(ctx, oracle, synth, scene_small):
    sc, mc, ms = scene_small
    c, s, truth = frame_features(synth, oracle, sc, (0.0, 0.0, 0.0, (0.0, 0.0, 0.0)), "VLP-16", seed=5)
    pg, sg = _compare_match(ctx, oracle, mc, ms, c, s, truth)
    assert sg["converged"] and sg["iterations"] <= 3


def test_match_too_few_reference(ctx, oracle, scene_small):
    sc, mc, ms = scene_small
    init = np.array([0.01, 0.02, 0.03, 1, 2, 3], np.float32)
    pg, sg, _ = ctx.match_stateless(mc[:40], ms, mc[:100], ms[:500], init)
    assert sg["status"] == 1 and np.array_equal(pg, init) and sg["iterations"] == 0
    pg, sg, _ = ctx.match_stateless(mc, ms[:99], mc[:100], ms[:500], init)
    assert sg["status"] == 1 and np.array_equal(pg, init)


def test_match_too_few_matches(ctx, oracle, scene_small):
    sc, mc, ms = scene_small
    far = ms[:600].copy(); far[:, :3] += 500.0     # queries nowhere near the map
    init = np.zeros(6, np.float32)
    pg, sg, lg = ctx.match_stateless(mc, ms, mc[:0], far, init, trace=True)
    po, so, lo = oracle.scan_match(mc, ms, mc[:0], far, init, keep_log=True)
    assert sg["status"] == 2 and so["tooFewMatches"]
    assert np.array_equal(pg, po) and sg["iterations"] == so["iterations"] == 0
    assert sg["rows"] == so["rows"]


def test_match_corridor_degenerate(ctx, oracle, synth):
    sc = synth.make_scene(seed=1, extent=60.0, corridor=True)
    mc, ms = synth.sample_map(sc, 0.4, seed=2)
    # the corridor has no edges: add a synthetic pole line so the corner gate (>= 50 reference corners) passes
    pole = np.zeros((80, 4), np.float32); pole[:, 0] = 200.0; pole[:, 2] = np.linspace(-1.8, 6, 80)
    mc = np.concatenate([mc, pole])
    R, t = synth.pose_matrix(0.0, 0.0, 0.0, (0.0, 0.0, 0.0))
    fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=9)
    r = oracle.scanreg_organised(fr)
    c = oracle.voxel_filter(r["lessSharp"], 0.4); s = oracle.voxel_filter(r["lessFlat"], 0.8)
    init = np.array([0.002, -0.003, 0.004, 0.05, 0.04, -0.03], np.float32)
    pg, sg = _compare_match(ctx, oracle, mc, ms, c, s, init)
    assert sg["degenerate"]        # along-corridor direction is unobservable (ScanMatch.cpp:211-240)


def test_scanmatch_mirror_class(cmb, oracle, synth, scene_small):
    sc, mc, ms = scene_small
    c, s, truth = frame_features(synth, oracle, sc, (0.02, 0.0, 0.0, (1.0, 0.2, 0.0)), "VLP-16", seed=8)
    sm = cmb.ScanMatch(10)
    sm.setConvergeThreshold(0.1, 0.1)
    sm.setUseCore(False)
    ok, pose = sm.scanMatchScan(mc, ms, c, s, truth + np.float32(0.01))
    assert ok is False          # quirk 5: mapping returns false even on convergence (useScore = false)
    assert sm.last_stats["converged"]
    sm2 = cmb.ScanMatch(10)     # class defaults: useScore = true, thresholds 0.05
    ok2, pose2 = sm2.scanMatchScan(mc, ms, c, s, truth + np.float32(0.01))
    po, so, _ = oracle.scan_match(mc, ms, c, s, truth + np.float32(0.01), params=dict(deltaTAbort=0.05, deltaRAbort=0.05, useScore=1))
    assert ok2 == so["ok"] and np.array_equal(pose2, po)
    assert abs(sm2.last_stats["score"] - so["score"]) <= 1e-9 * max(1.0, so["score"]) or not so["converged"]


def test_sharded_map_equals_unsharded(cmb, ctx, oracle, synth, scene_small):
    """BASELINE config 4 in miniature: the map split into 3 x-slabs (+ halo), per-iteration sum of the partial normal
    equations, redundant solve -> the pose is bit-identical to the single-map solve."""
    import importlib
    dmod = importlib.import_module("the-cooper-mapper_b200.dist")
    sc, mc, ms = scene_small
    c, s, truth = frame_features(synth, oracle, sc, (0.05, 0.01, -0.008, (3.0, 0.4, 0.1)), "HDL-64E")
    init = truth + np.array([0.005, -0.004, 0.008, 0.08, -0.06, 0.05], np.float32)
    ref_pose, ref_stats, _ = ctx.match_stateless(mc, ms, c, s, init)
    world = 3
    bounds = dmod.slab_bounds(ms, world)
    ranks = []
    pending = []

    def make_reduce(r):
        def red(sums):            # emulate the all-reduce in one process: every rank contributes before anyone solves
            pending.append(sums)
            return sums
        return red

    ctxs = [cmb.Context() for _ in range(world)]
    sms = []
    for r in range(world):
        lo, hi = dmod.own_box(bounds, r)
        sms.append(dmod.ShardedScanMatch(ctxs[r], dmod.shard_cloud(mc, bounds, r), dmod.shard_cloud(ms, bounds, r), len(mc), len(ms), lo, hi))
    # lock-step iteration over the "ranks"
    import ctypes as C
    from conftest import frame_features as _ff  # noqa: F401
    poses = [init.copy() for _ in range(world)]
    cc = np.ascontiguousarray(c, np.float32); ss = np.ascontiguousarray(s, np.float32)
    for r in range(world):
        sm = sms[r]
        sm.ctx._check(sm.L.cm_shard_begin_host(sm.ctx.h, cc.ctypes.data_as(C.c_void_p), C.c_size_t(len(cc)), ss.ctypes.data_as(C.c_void_p),
                                               C.c_size_t(len(ss)), poses[r].ctypes.data_as(C.c_void_p), C.c_size_t(len(mc)), C.c_size_t(len(ms))))
    done = False; iters = 0
    while not done and iters < ctx.cfg.max_iterations:
        parts = []
        for r in range(world):
            sums = np.zeros(32, np.float64); sm = sms[r]
            sm.ctx._check(sm.L.cm_shard_partial_host(sm.ctx.h, C.c_int(iters), sm.lo.ctypes.data_as(C.c_void_p), sm.hi.ctypes.data_as(C.c_void_p),
                                                     sums.ctypes.data_as(C.c_void_p)))
            parts.append(sums)
        total = np.sum(parts, axis=0)
        assert total[27] == sum(p[27] for p in parts)
        flags = []
        for r in range(world):
            d = C.c_int(0); st = cmb.MatchStats(); sm = sms[r]
            sm.ctx._check(sm.L.cm_shard_solve_host(sm.ctx.h, C.c_int(iters), total.ctypes.data_as(C.c_void_p), poses[r].ctypes.data_as(C.c_void_p),
                                                   C.byref(d), C.byref(st)))
            flags.append(d.value)
        assert len(set(flags)) == 1                       # every rank takes the same decision
        done = bool(flags[0]); iters += 1
    for r in range(world):
        assert np.array_equal(poses[r], ref_pose)         # bit-identical to the unsharded solve, on every rank
    assert iters == ref_stats["iterations"]
    for cx in ctxs:
        cx.close()


def test_scan_match_local_with_score_gate(cmb, oracle, synth, scene_small):
    """ScanMatch::scanMatchLocal as the pose-graph consumers call it: pre-voxelised clouds, score / percentage gate."""
    sc, mc, ms = scene_small
    R, t = synth.pose_matrix(0.03, 0.0, 0.0, (2.0, -0.3, 0.05))
    f = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, "HDL-64E", seed=31))
    truth = np.array([0.0, 0.0, 0.03, 2.0, -0.3, 0.05], np.float32)
    sm = cmb.ScanMatch(10)                                   # class defaults: useScore, thresholds 0.05
    for delta in (np.float32(0.02), np.array([0.01, -0.01, 0.02, 0.3, -0.2, 0.1], np.float32)):
        ok, pose = sm.scanMatchLocal(mc, ms, f["lessSharp"], f["lessFlat"], truth + delta)
        po, so, _ = oracle.scan_match_local(mc, ms, f["lessSharp"], f["lessFlat"], truth + delta)
        assert ok == so["ok"] and np.array_equal(pose, po)
        assert sm.last_stats["iterations"] == so["iterations"]
        if so["converged"]:
            assert abs(sm.last_stats["score"] - so["score"]) <= 1e-9 * max(1.0, so["score"])
