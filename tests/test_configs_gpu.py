"""GPU parity at the BASELINE.json configurations: config 2 at full size (HDL-64E 64 x 2048 sweeps against a ~1M-point map),
config 3 (many independent VLP-16 streams in one batch), config 5 (sparse tilted-RPLidar sweeps: degenerate and
low-feature frames).  Full-size checks go through the oracle on the same inputs and through size-independent properties:
stream independence, determinism, one point per map voxel."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_config2_hdl64_full_size(cmb, oracle, synth):
    bench = importlib.import_module("bench")
    mc, ms, frames, poses = bench.make_workload(3, synth)                  # the bench workload itself: ~1M-point map
    assert len(mc) + len(ms) > 900000 and frames.shape[1:] == (64, 2048, 4)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    rng = np.random.default_rng(3)
    odoms = [bench.noisy_odom(poses, k, rng, synth) for k in range(3)]

    def run(S):
        ctx = cmb.Context(**bench.CFG)
        ctx.mapping_create(S, max_corner_points=4 * len(mc), max_surf_points=int(1.6 * len(ms)) + 200000)
        for s_ in range(S):
            ctx.map_update(s_, np.zeros(3, np.float32))                    # FeatureMap::update at the origin, like the oracle below
        ctx.map_insert([mc] * S, [ms] * S, [eye] * S)
        out = []
        for k in range(3):
            isos, stats = ctx.pipeline_step(np.stack([frames[k]] * S), [odoms[k]] * S)
            out.append((isos, stats))
        maps = [ctx.map_export_sorted(0, cls)[0] for cls in (0, 1)]
        ctx.close()
        return out, maps

    out2, maps2 = run(2)
    out1, maps1 = run(1)
    om = oracle.Mapping(map_params=bench.ORACLE_MAP)
    om.map_update(np.zeros(3, np.float32))
    om.map_add(mc, ms, eye[0], eye[1])
    for k in range(3):
        (isos, stats), (isos1, stats1) = out2[k], out1[k]
        # stream independence + determinism: both streams of the batch and the single-stream run agree bit for bit
        assert np.array_equal(isos[0][0], isos[1][0]) and np.array_equal(isos[0][1], isos[1][1])
        assert np.array_equal(isos[0][0], isos1[0][0]) and np.array_equal(isos[0][1], isos1[0][1])
        assert stats[0]["iterations"] == stats[1]["iterations"] == stats1[0]["iterations"]
        # the oracle on the same inputs
        f = oracle.scanreg_organised(frames[k])
        oR, ot, ost = om.process(odoms[k][0], odoms[k][1], f["lessSharp"], f["lessFlat"])
        assert stats[0]["iterations"] == ost["iterations"] and stats[0]["rows"] == ost["rows"]
        assert np.max(np.abs(isos[0][1] - ot)) <= 1e-4 and np.max(np.abs(isos[0][0] - oR)) <= 1e-5     # north-star tolerance
        assert np.array_equal(isos[0][0], oR) and np.array_equal(isos[0][1], ot)
        # the registration recovers the true pose from a +-0.1 m / +-0.5 deg prediction
        assert np.linalg.norm(isos[0][1] - poses[k][1]) < 0.05 and stats[0]["converged"]
    for cls, which in ((0, 4), (1, 5)):
        assert _same(maps2[cls], maps1[cls])
        # cubes inside the 150 m validity window are filtered, farther cubes hold the raw samples until they become valid -- in the
        # reference and here alike; inside an unfiltered cube the order is push order there and cell order here: compare per cube
        # as sorted multisets, and the filtered part (|x|, |y| < 75 m) in the reference's own order
        near = lambda a: a[(np.abs(a[:, 0]) < 75.0) & (np.abs(a[:, 1]) < 75.0)]
        assert _same(near(maps2[cls]), near(om.cloud(which)))
        key = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
        assert _same(key(maps2[cls]), key(om.cloud(which)))
        # one point per (cube, voxel) -- up to centroids that rounding put exactly on a voxel face: such a point shares its new
        # voxel with the resident one until the next filter pass (also in the reference), a handful per million
        def dups(a):
            v = np.floor(a[:, :3] * np.float32(1.0 / 0.4)).astype(np.int64)
            cube = np.round(a[:, :3] / 50.0).astype(np.int64)
            keys = np.concatenate([v, cube], 1)
            return len(keys) - len(np.unique(keys, axis=0))
        assert dups(near(maps2[cls])) <= 4 and dups(near(maps2[cls])) == dups(near(om.cloud(which))) and dups(maps2[cls]) == dups(om.cloud(which))


def test_config3_batched_streams_equal_single_stream_runs(cmb, oracle, synth):
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    S, NF = 24, 3
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    seqs = []
    for s in range(S):
        seq = []
        for k, (R, t) in enumerate(synth.trajectory(NF, seed=s, speed=0.3 + 0.05 * s)):
            seq.append((R.astype(np.float32), t.astype(np.float32), synth.simulate_scan(sc, R, t, "VLP-16", seed=0x100 + 16 * s + k, cols=600)))
        seqs.append(seq)
    ctx = cmb.Context(**cfg)
    ctx.mapping_create(S, 60000, 300000)
    batched = []
    for k in range(NF):
        isos, stats = ctx.pipeline_step(np.stack([seqs[s][k][2] for s in range(S)]), [(seqs[s][k][0], seqs[s][k][1]) for s in range(S)])
        batched.append((isos, stats))
    for s in (0, 7, 23):                                                    # the same streams alone, and the oracle
        c1 = cmb.Context(**cfg); c1.mapping_create(1, 60000, 300000)
        om = oracle.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4))
        for k in range(NF):
            isos, stats = c1.pipeline_step(seqs[s][k][2][None], [(seqs[s][k][0], seqs[s][k][1])])
            assert np.array_equal(isos[0][0], batched[k][0][s][0]) and np.array_equal(isos[0][1], batched[k][0][s][1])
            f = oracle.scanreg_organised(seqs[s][k][2])
            oR, ot, _ = om.process(seqs[s][k][0], seqs[s][k][1], f["lessSharp"], f["lessFlat"])
            assert np.array_equal(isos[0][0], oR) and np.array_equal(isos[0][1], ot)
        for cls, which in ((0, 4), (1, 5)):
            assert _same(ctx.map_export_sorted(s, cls)[0], om.cloud(which))
        c1.close()
    ctx.close()


@pytest.mark.parametrize("scene_kind", ["corridor", "wall", "field"])
def test_config5_sparse_tilted_rplidar(cmb, oracle, synth, scene_kind):
    """One planar 360 deg scanner nodding +-30 deg, 12 revolutions per sweep, 800 points per revolution, 12 m range."""
    if scene_kind == "corridor":
        sc = synth.make_scene(seed=1, extent=40.0, corridor=True)
    elif scene_kind == "wall":
        sc = synth.Scene([[6.0, 0.0, 0.0, 0.5, 30.0, 0.0, 8.0]], [], 40.0)
    else:
        sc = synth.Scene([], [], 40.0)                                     # open field: only the ground
    tilts = np.linspace(-30.0, 30.0, 12)
    cfg = dict(filter_corner=0.2, filter_surf=0.4, map_filter_corner=0.2, map_filter_surf=0.4, blind_radius=0.3)
    ctx = cmb.Context(**cfg)
    ctx.mapping_create(1, 50000, 200000)
    om = oracle.Mapping(map_params=dict(filterCorner=0.2, filterSurf=0.4, mapFilterCorner=0.2, mapFilterSurf=0.4))
    seen = set()
    for k, (R, t) in enumerate(synth.trajectory(5, speed=0.15)):
        fr = synth.simulate_scan(sc, R, t, tilt_deg=tilts, cols=800, seed=500 + k, dropout=0.05)
        g = ctx.scanreg_organised(fr, debug=True)
        o = oracle.scanreg_organised(fr, params=dict(blindRadius=0.3))
        for name in ("sharpIdx", "lessSharpIdx", "flatIdx", "lessFlatRawIdx"):
            assert np.array_equal(g[name], o[name]), (scene_kind, k, name)
        for name in ("sharp", "lessSharp", "flat", "lessFlat"):
            assert _same(g[name], o[name]), (scene_kind, k, name)
        odom = (R.astype(np.float32), t.astype(np.float32))
        isos, stats = ctx.mapping_process([odom], [g["lessSharp"]], [g["lessFlat"]])
        oR, ot, ost = om.process(odom[0], odom[1], o["lessSharp"], o["lessFlat"])
        assert stats[0]["iterations"] == ost["iterations"] and stats[0]["rows"] == ost["rows"], (scene_kind, k, stats[0], ost)
        assert (stats[0]["status"] == cmb.CM_TOO_FEW_REF) == bool(ost["tooFewRef"])
        assert (stats[0]["status"] == cmb.CM_TOO_FEW_MATCHES) == bool(ost["tooFewMatches"])
        assert bool(stats[0]["degenerate"]) == bool(ost["degenerate"])
        assert np.array_equal(isos[0][0], oR) and np.array_equal(isos[0][1], ot), (scene_kind, k)
        seen.add((stats[0]["status"], bool(stats[0]["degenerate"])))
    assert (cmb.CM_TOO_FEW_REF, False) in seen                              # first frame: empty map
    ctx.close()


def test_config1_vlp16_full_chain(cmb, oracle, synth):
    """BASELINE config 1 (the reference's CPU-runnable case): a VLP-16 16 x 1800 sequence through scan registration ->
    laserOdometry -> laserMapping, every stage on the GPU, against the same chain of the oracle."""
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    c_sr = cmb.Context()                       # stage 1
    c_od = cmb.Context(); lo = cmb.LaserOdometry(ctx=c_od)          # stage 2
    c_mp = cmb.Context(**cfg); c_mp.mapping_create(1, 100000, 800000)   # stage 3
    oo = oracle.Odometry()
    om = oracle.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4))
    NF = 25
    path = []
    for k, (R, t) in enumerate(synth.trajectory(NF, speed=0.1, yaw_amp=0.02)):      # 1 m/s at 10 Hz
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=0x1000 + k)
        g = c_sr.scanreg_organised(fr)
        o = oracle.scanreg_organised(fr)
        for name in ("sharp", "lessSharp", "flat", "lessFlat"):
            assert _same(g[name], o[name]), (k, name)
        go = c_od.odometry_process(g["sharp"], g["lessSharp"], g["flat"], g["lessFlat"])
        oo_ = oo.process(o["sharp"], o["lessSharp"], o["flat"], o["lessFlat"])
        assert np.array_equal(go["R"], oo_["R"]) and np.array_equal(go["t"], oo_["t"]), k
        assert _same(go["corner_last"], oo_["corner_last"]) and _same(go["surf_last"], oo_["surf_last"]), k
        isos, stats = c_mp.mapping_process([(go["R"], go["t"])], [go["corner_last"]], [go["surf_last"]])
        oR, ot, ost = om.process(oo_["R"], oo_["t"], oo_["corner_last"], oo_["surf_last"])
        assert stats[0]["iterations"] == ost["iterations"], (k, stats[0], ost)
        assert np.max(np.abs(isos[0][1] - ot)) <= 1e-4 and np.max(np.abs(isos[0][0] - oR)) <= 1e-5      # north-star tolerance
        assert np.array_equal(isos[0][0], oR) and np.array_equal(isos[0][1], ot), k
        path.append(isos[0][1].copy())
    for cls, which in ((0, 4), (1, 5)):
        assert _same(c_mp.map_export_sorted(0, cls)[0], om.cloud(which))
    assert np.linalg.norm(path[-1] - path[0]) > 1.0          # the chain did track the motion
    for c in (c_sr, c_od, c_mp):
        c.close()


def test_chain_step_equals_the_three_stages(cmb, oracle, synth):
    """cm_pipeline_chain_step_host: the three stages in one call, feature clouds and projected clouds staying on the device, for two
    streams at once -- the poses of every sweep and the final maps are those of the oracle's chain (scan registration -> odometry ->
    mapping), stream by stream."""
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    S, NF = 2, 12
    ctx = cmb.Context(**cfg)
    ctx.mapping_create(S, 100000, 800000)
    ctx.pipeline_chain_create(16, 1800)
    oos = [oracle.Odometry() for _ in range(S)]
    oms = [oracle.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4)) for _ in range(S)]
    trajs = [synth.trajectory(NF, speed=0.1, yaw_amp=0.02), synth.trajectory(NF, speed=0.07, yaw_amp=-0.03)]
    for k in range(NF):
        fr = np.stack([synth.simulate_scan(sc, trajs[s][k][0], trajs[s][k][1], "VLP-16", seed=0x2000 + 50 * s + k) for s in range(S)])
        odoms, isos, ostats, stats = ctx.pipeline_chain_step(fr)
        for s in range(S):
            o = oracle.scanreg_organised(fr[s])
            oo_ = oos[s].process(o["sharp"], o["lessSharp"], o["flat"], o["lessFlat"])
            assert np.array_equal(odoms[s][0], oo_["R"]) and np.array_equal(odoms[s][1], oo_["t"]), (k, s)
            oR, ot, ost = oms[s].process(oo_["R"], oo_["t"], oo_["corner_last"], oo_["surf_last"])
            assert stats[s]["iterations"] == ost["iterations"], (k, s)
            assert np.array_equal(isos[s][0], oR) and np.array_equal(isos[s][1], ot), (k, s)
    for s in range(S):
        for cls, which in ((0, 4), (1, 5)):
            assert _same(ctx.map_export_sorted(s, cls)[0], oms[s].cloud(which))
    ctx.close()


def test_chain_step_device_sweeps_and_prefetch_equal_host_sweeps(cmb, synth):
    """cm_pipeline_chain_step_dev (sweeps already in device memory) and a chain step whose sweep was announced with
    cm_pipeline_prefetch_host give the poses of the plain host call, bit for bit."""
    import torch
    sc = synth.make_scene(seed=77, extent=50.0, n_boxes=18, n_poles=14)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    S, NF, rows, cols = 2, 6, 16, 900
    trajs = [synth.trajectory(NF, speed=0.1, yaw_amp=0.02), synth.trajectory(NF, speed=0.08, yaw_amp=-0.02)]
    frames = [np.stack([synth.simulate_scan(sc, trajs[s][k][0], trajs[s][k][1], "VLP-16", seed=0x3000 + 50 * s + k, cols=cols) for s in range(S)]).astype(np.float32)
              for k in range(NF)]
    results = {}
    for mode in ("host", "dev", "prefetch"):
        ctx = cmb.Context(**cfg)
        ctx.mapping_create(S, 100000, 600000)
        ctx.pipeline_chain_create(rows, cols)
        od = np.empty((S, 12), np.float32); mp = np.empty((S, 12), np.float32); ost = (cmb.OdomStats * S)(); mst = (cmb.MatchStats * S)()
        pinned = [torch.from_numpy(f).pin_memory().numpy() for f in frames]
        out = []
        for k in range(NF):
            if mode == "host":
                ctx.pipeline_chain_step_packed(frames[k], od, mp, ost, mst)
            elif mode == "dev":
                d = torch.from_numpy(frames[k]).cuda()
                torch.cuda.synchronize()
                ctx.pipeline_chain_step_dev(d.data_ptr(), rows, cols, od, mp, ost, mst)
            else:
                if k == 0:
                    ctx.pipeline_prefetch(pinned[0])
                if k + 1 < NF:
                    ctx.pipeline_prefetch(pinned[k + 1])                     # the next sweep uploads while this one is registered
                ctx.pipeline_chain_step_packed(pinned[k], od, mp, ost, mst)
            out.append((od.copy(), mp.copy(), [ost[s].iterations for s in range(S)], [mst[s].iterations for s in range(S)]))
        ctx.mapping_sync()
        results[mode] = (out, [ctx.map_export_sorted(s, cls)[0] for s in range(S) for cls in (0, 1)])
        ctx.close()
    for mode in ("dev", "prefetch"):
        for a, b in zip(results["host"][0], results[mode][0]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3], mode
        for a, b in zip(results["host"][1], results[mode][1]):
            assert _same(a, b), mode


def test_raw_sweep_chain_equals_oracle_chain(cmb, oracle, synth):
    """cm_pipeline_chain_step_sweep_host: the unorganised sweep a driver publishes -> MultiScanRegistration front end -> feature
    extraction -> odometry -> mapping in one call; poses and the final map are those of the oracle's chain started from
    oracle.scanreg_sweep (the raw-sweep restatement, MultiScanRegistration.cpp:95-200)."""
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    NF = 8
    ctx = cmb.Context(**cfg)
    ctx.mapping_create(1, 100000, 800000)
    ctx.pipeline_chain_sweep_create(16 * 1800)
    oo = oracle.Odometry()
    om = oracle.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4))
    for k, (R, t) in enumerate(synth.trajectory(NF, speed=0.1, yaw_amp=0.02)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=0x4000 + k, cols=1200)
        sweep = synth.organised_to_sweep(fr)
        ok = np.where(np.isfinite(sweep[:, 0]))[0]
        sweep = sweep[ok[0]:ok[-1] + 1]
        (gR, gt), (mR, mt), ost, mst = ctx.pipeline_chain_step_sweep(sweep, 0)
        o = oracle.scanreg_sweep(sweep, 0)
        oo_ = oo.process(o["sharp"], o["lessSharp"], o["flat"], o["lessFlat"])
        assert np.array_equal(gR, oo_["R"]) and np.array_equal(gt, oo_["t"]), k
        oR, ot, ostm = om.process(oo_["R"], oo_["t"], oo_["corner_last"], oo_["surf_last"])
        assert mst["iterations"] == ostm["iterations"], k
        assert np.array_equal(mR, oR) and np.array_equal(mt, ot), k
    for cls, which in ((0, 4), (1, 5)):
        assert _same(ctx.map_export_sorted(0, cls)[0], om.cloud(which))
    ctx.close()


def test_raw_sweep_chain_with_imu_deskew(cmb, oracle, synth):
    """The raw-sweep chain with scan_time >= 0: every sweep is de-skewed with the IMU states pushed before (ScanRegistration.cpp:89-188)
    ahead of odometry and mapping; poses equal the oracle chain started from oracle.scanreg_sweep_imu."""
    sc = synth.make_scene(seed=23, extent=50.0, n_boxes=16, n_poles=12)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    NF = 4
    stamps = np.arange(99.90, 100.60, 0.01)
    rng = np.random.default_rng(4)
    imu = np.zeros((len(stamps), 7))
    imu[:, 0] = stamps
    imu[:, 1] = 0.01 * np.sin(3 * stamps); imu[:, 2] = 0.015 * np.cos(2 * stamps); imu[:, 3] = 0.3 * (stamps - 100.0)
    imu[:, 4:] = rng.normal(0, 0.3, (len(stamps), 3)) + np.array([0.0, 0.0, 9.81])
    ctx = cmb.Context(**cfg)
    ctx.mapping_create(1, 100000, 800000)
    ctx.pipeline_chain_sweep_create(16 * 1200)
    for m in imu:
        ctx.imu_push(*m)
    oo = oracle.Odometry()
    om = oracle.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4))
    moved = 0.0
    for k, (R, t) in enumerate(synth.trajectory(NF, speed=0.1, yaw_amp=0.02)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=0x5000 + k, cols=1000)
        sweep = synth.organised_to_sweep(fr)
        ok = np.where(np.isfinite(sweep[:, 0]))[0]
        sweep = sweep[ok[0]:ok[-1] + 1]
        scan_time = 100.05 + 0.1 * k
        (gR, gt), (mR, mt), ost, mst = ctx.pipeline_chain_step_sweep(sweep, 0, imu_scan_time=scan_time)
        o, _ = oracle.scanreg_sweep_imu(sweep, 0, scan_time, imu)
        moved = max(moved, float(np.abs(o["cloud"][:, :3] - oracle.scanreg_sweep(sweep, 0)["cloud"][:, :3]).max()))
        oo_ = oo.process(o["sharp"], o["lessSharp"], o["flat"], o["lessFlat"])
        assert np.array_equal(gR, oo_["R"]) and np.array_equal(gt, oo_["t"]), k
        oR, ot, ostm = om.process(oo_["R"], oo_["t"], oo_["corner_last"], oo_["surf_last"])
        assert mst["iterations"] == ostm["iterations"] and np.array_equal(mR, oR) and np.array_equal(mt, ot), k
    assert moved > 1e-3          # the de-skew did move points
    ctx.close()


def test_knn5_full_size_vs_nanoflann(cmb, oracle, synth):
    """Exact 5-NN at the headline size: the ~1M-point surf map of the bench workload, 20,000 queries, against the reference's
    own KD-tree (nanoflann) -- neighbour sets and float distances identical, both map cell sizes (surf and corner default)."""
    bench = importlib.import_module("bench")
    sc = synth.make_scene(seed=bench.SEED & 0xFFFF, extent=125.0, n_boxes=44, n_poles=40)
    _, ms = synth.sample_map(sc, 0.4, seed=2)
    assert len(ms) > 900000
    rng = np.random.default_rng(9)
    q = ms[rng.integers(0, len(ms), 20000), :3] + rng.normal(0, 0.4, (20000, 3)).astype(np.float32)
    q[:500] += 30.0                                                     # far from every surface: gate rejects
    if not oracle.lib().has_nanoflann:
        pytest.skip("oracle/_ref/libcm_ref_nanoflann.so not built")
    ni, nd = oracle.knn(ms, q, 5, nanoflann=True)
    ok = nd[:, 4] < 5.0
    assert 15000 < ok.sum() < 20000
    ctx = cmb.Context()
    for cell in (1.6, 3.2):
        idx, d2 = ctx.knn5(ms, q, cell=cell, gate=5.0)
        assert np.array_equal(d2[ok], nd[ok])                           # bit-exact distances, ascending
        assert np.array_equal(np.sort(idx[ok], 1), np.sort(ni[ok], 1))  # same neighbour sets
        assert np.all(d2[~ok][:, 4] >= 5.0)
    ctx.close()


def test_prefetched_steps_equal_synchronous_steps(cmb, synth):
    """cm_pipeline_prefetch_host / _dev (upload + scan registration issued ahead on side streams) change nothing but timing."""
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    S, NF = 3, 5
    frames, odoms = [], []
    for k in range(NF):
        fr, od = [], []
        for s in range(S):
            R, t = list(synth.trajectory(NF, seed=s, speed=0.4))[k]
            fr.append(synth.simulate_scan(sc, R, t, "VLP-16", seed=900 + 10 * s + k, cols=512))
            od.append((R.astype(np.float32), t.astype(np.float32)))
        frames.append(np.ascontiguousarray(np.stack(fr), np.float32)); odoms.append(od)
    a = cmb.Context(**cfg); a.mapping_create(S, 60000, 300000)
    b = cmb.Context(**cfg); b.mapping_create(S, 60000, 300000)
    mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()
    c = cmb.Context(**cfg); c.mapping_create(S, 60000, 300000)   # deferred prefetches (issued behind the Gauss-Newton submission)
    b.pipeline_prefetch(frames[0]); b.pipeline_prefetch(frames[1])
    c.pipeline_prefetch(frames[0]); c.pipeline_prefetch(frames[1])
    for k in range(NF):
        ref_isos, ref_stats = a.pipeline_step(frames[k], odoms[k])
        if k + 2 < NF:
            b.pipeline_prefetch(frames[k + 2])
            c.pipeline_prefetch(frames[k + 2], deferred=True)
            with pytest.raises(cmb.CoopermapError):              # one registration at a time
                c.pipeline_prefetch(frames[k + 2], deferred=True)
        for ctx_ in (b, c):
            ctx_.pipeline_step_packed(frames[k], ctx_._pack_isos(odoms[k]), mapped, stats)
            for s in range(S):
                assert np.array_equal(mapped[s, :9].reshape(3, 3), ref_isos[s][0]) and np.array_equal(mapped[s, 9:], ref_isos[s][1]), (k, s)
                assert stats[s].iterations == ref_stats[s]["iterations"]
    for s in range(S):
        for cls in (0, 1):
            assert _same(a.map_export_sorted(s, cls)[0], b.map_export_sorted(s, cls)[0])
    # the Gauss-Newton loop of these steps was submitted as a CUDA graph with a conditional WHILE node (CUDA >= 12.4)
    import ctypes as C
    ng = C.c_int(0); wl = C.c_int(0)
    b._check(b.L.cm_debug_graph_info(b.h, C.byref(ng), C.byref(wl)))
    assert ng.value >= 1 and wl.value == 1
    # at most four sweeps in flight
    for k in range(4):
        b.pipeline_prefetch(frames[k])
    with pytest.raises(cmb.CoopermapError):
        b.pipeline_prefetch(frames[4])
    a.close(); b.close(); c.close()
