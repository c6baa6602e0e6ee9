"""GPU parity: device-resident map (insert + voxel merge), the mapping stage and the full pipeline vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAP_CFG = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
ORACLE_MAP = dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _frames(synth, sc, n, model="VLP-16", cols=900, seed=100, speed=0.5):
    out = []
    for k, (R, t) in enumerate(synth.trajectory(n, speed=speed)):
        out.append((R, t, synth.simulate_scan(sc, R, t, model, seed=seed + k, cols=cols)))
    return out


def test_map_insert_matches_oracle_cubes(cmb, oracle, synth):
    sc = synth.make_scene(seed=51, extent=70.0, n_boxes=20, n_poles=10)
    mc, ms = synth.sample_map(sc, 0.23, seed=52)          # denser than the map leaf: many points per voxel
    rng = np.random.default_rng(1)
    ctx = cmb.Context(**MAP_CFG)
    ctx.mapping_create(1, 100000, 1500000)
    om = oracle.Mapping(map_params=ORACLE_MAP)
    om.map_update(np.zeros(3, np.float32))                # valid cubes around the origin (the GPU merges everywhere)
    R, t = synth.pose_matrix(0.3, 0.02, -0.01, (1.0, -2.0, 0.3))
    for part in range(3):                                 # three inserts: later ones merge into resident voxels
        c = mc[rng.permutation(len(mc))[: len(mc) // 2]]; s = ms[rng.permutation(len(ms))[: len(ms) // 3]]
        ctx.map_insert([c], [s], [(R, t)])
        om.map_add(c, s, R, t)
    for cls, which in ((0, 4), (1, 5)):
        g, _ = ctx.map_export_sorted(0, cls)
        assert _same(g, om.cloud(which))


def test_mapping_stage_sequence_bit_exact(cmb, oracle, synth):
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    S = 3
    ctx = cmb.Context(**MAP_CFG)
    ctx.mapping_create(S, 100000, 600000)
    oms = [oracle.Mapping(map_params=ORACLE_MAP) for _ in range(S)]
    frames = [_frames(synth, sc, 6, seed=100 + 50 * s, speed=0.4 + 0.2 * s) for s in range(S)]
    for k in range(6):
        odoms, corners, surfs = [], [], []
        for s in range(S):
            R, t, fr = frames[s][k]
            f = oracle.scanreg_organised(fr)
            odoms.append((R.astype(np.float32), (t + np.array([0.02, -0.02, 0.01]) * k).astype(np.float32)))
            corners.append(f["lessSharp"]); surfs.append(f["lessFlat"])
        isos, stats = ctx.mapping_process(odoms, corners, surfs)
        for s in range(S):
            oR, ot, ost = oms[s].process(odoms[s][0], odoms[s][1], corners[s], surfs[s])
            gR, gt = isos[s]
            assert stats[s]["iterations"] == ost["iterations"], (k, s, stats[s], ost)
            assert (stats[s]["status"] == 1) == bool(ost["tooFewRef"])
            assert stats[s]["rows"] == ost["rows"]
            assert np.all(np.abs(gt - ot) <= 1e-4) and np.all(np.abs(gR - oR) <= 1e-5)      # north-star tolerance
            assert np.array_equal(gR, oR) and np.array_equal(gt, ot)                        # in practice: bit-identical
    for s in range(S):
        for cls, which in ((0, 4), (1, 5)):
            g, _ = ctx.map_export_sorted(s, cls)
            assert _same(g, oms[s].cloud(which))
    assert stats[0]["converged"]


def test_pipeline_step_equals_scanreg_plus_mapping(cmb, oracle, synth):
    sc = synth.make_scene(seed=61, extent=40.0, n_boxes=12, n_poles=10)
    S = 2
    ctx = cmb.Context(**MAP_CFG)
    ctx.mapping_create(S, 100000, 600000)
    oms = [oracle.Mapping(map_params=ORACLE_MAP) for _ in range(S)]
    seqs = [_frames(synth, sc, 4, seed=300 + 40 * s, cols=720) for s in range(S)]
    for k in range(4):
        fr = np.stack([seqs[s][k][2] for s in range(S)])
        odoms = [(seqs[s][k][0].astype(np.float32), seqs[s][k][1].astype(np.float32)) for s in range(S)]
        isos, stats = ctx.pipeline_step(fr, odoms)
        for s in range(S):
            f = oracle.scanreg_organised(fr[s])
            oR, ot, ost = oms[s].process(odoms[s][0], odoms[s][1], f["lessSharp"], f["lessFlat"])
            assert np.array_equal(isos[s][0], oR) and np.array_equal(isos[s][1], ot)
            assert stats[s]["iterations"] == ost["iterations"]


def test_lasermapping_mirror(cmb, oracle, synth):
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    lm = cmb.LaserMapping(max_corner_points=50000, max_surf_points=300000, **MAP_CFG)
    R, t, fr = _frames(synth, sc, 1)[0]
    f = oracle.scanreg_organised(fr)
    gR, gt = lm.process(R.astype(np.float32), t.astype(np.float32), f["lessSharp"], f["lessFlat"])
    assert lm.last_stats["status"] == 1                      # empty map: "reference cloud points too few"
    # 200 m up: the reference shift()s its cube grid (FeatureMap.h:232-245) -- so does the device map
    lm.process(R.astype(np.float32), np.array([0, 0, 200], np.float32), f["lessSharp"], f["lessFlat"])


@pytest.mark.parametrize("step", [4.0, 11.0])
def test_far_cubes_stay_unfiltered_until_valid(cmb, oracle, synth, step):
    """FeatureMap::downsizeValidCloud (FeatureMap.h:289-306) only filters VALID cubes: returns that land in a cube outside the window
    stay raw (several points per voxel, all of them searchable once the cube is valid) until an insert finds the cube valid and
    filters its whole cloud at once.  20 m cubes with a 60 m validity radius and 100 m LiDAR range put a third of every sweep into
    invalid cubes; the sensor then walks into them.  Poses, surround clouds and every stored cube equal the oracle's, frame by frame
    (the larger step also re-centres the grid: FeatureMap::shift on cubes that hold raw points)."""
    sc = synth.make_scene(seed=79, extent=150.0, n_boxes=70, n_poles=40)
    grid = dict(cube_w=13, cube_h=13, cube_d=7, cube_size=20.0, valid_distance=60.0)
    ctx = cmb.Context(**MAP_CFG, **grid)
    ctx.mapping_create(1, 200000, 1200000)
    om = oracle.Mapping(map_params=dict(ORACLE_MAP, cubeW=13, cubeH=13, cubeD=7, cubeSize=20.0, validDistance=60.0))
    raw_seen = 0
    for k in range(9):
        pos = np.array([step * k, 0.6 * step * k, 0.0])
        R, t = synth.pose_matrix(0.04 * np.sin(0.9 * k), 0.0, 0.0, pos)
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=700 + k, cols=900)
        f = oracle.scanreg_organised(fr)
        od = (R.astype(np.float32), (t + np.array([0.02, -0.03, 0.01])).astype(np.float32))
        isos, stats = ctx.mapping_process([od], [f["lessSharp"]], [f["lessFlat"]])
        oR, ot, ost = om.process(od[0], od[1], f["lessSharp"], f["lessFlat"])
        assert stats[0]["iterations"] == ost["iterations"] and stats[0]["rows"] == ost["rows"], (k, stats[0], ost)
        assert np.array_equal(isos[0][0], oR) and np.array_equal(isos[0][1], ot), (k, isos[0], oR, ot)
        gc, gs = ctx.map_surround(0)
        assert _same(gs, om.map_surround(1)) and _same(gc, om.map_surround(0)), k
        # every stored cube, raw ones included: same points (the order inside an unfiltered cube is the push order in the reference
        # and the cell order here, so compare as sorted multisets per cube)
        for cls, which in ((0, 4), (1, 5)):
            g, cubes = ctx.map_export_sorted(0, cls)
            o = om.cloud(which)
            assert len(g) == len(o), (k, cls, len(g), len(o))
            key = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
            assert _same(key(g), key(o)), (k, cls)
            v = np.floor(g[:, :3] * np.float32(1.0 / 0.4)).astype(np.int64)
            cb = np.round(g[:, :3] / np.float32(20.0)).astype(np.int64)
            raw_seen += len(g) - len(np.unique(np.concatenate([v, cb], 1), axis=0))
    assert raw_seen > 100           # there were unfiltered cubes (several points per voxel) along the way
    ctx.close()


@pytest.mark.parametrize("heading", [(1.0, 0.0), (-1.0, 0.0), (-0.8, 0.6), (0.7, -0.7)])
def test_feature_map_shift_matches_literal_reference(cmb, oracle, synth, heading):
    """FeatureMap::shift (FeatureMap.h:354-376) when the sensor walks out of the central cubes of a small grid: towards +x the
    reference's in-place pointer swaps are a proper shift, towards -x (or -x / +y ...) they move every cube the WRONG way (the
    stored clouds end up two cubes from where their coordinates say and partly drop out of the surround window).  The device map
    reproduces both: poses, iterations, row counts, surround clouds and the stored cubes are those of the literal oracle."""
    sc = synth.make_scene(seed=77, extent=220.0, n_boxes=90, n_poles=60)
    grid = dict(cube_w=9, cube_h=9, cube_d=7, cube_size=40.0, valid_distance=120.0)
    ctx = cmb.Context(**MAP_CFG, **grid)
    ctx.mapping_create(1, 200000, 900000)
    om = oracle.Mapping(map_params=dict(ORACLE_MAP, cubeW=9, cubeH=9, cubeD=7, cubeSize=40.0, validDistance=120.0))
    hx, hy = heading
    n_shift_frames = 0
    for k in range(13):
        pos = np.array([hx * 13.0 * k, hy * 13.0 * k, 0.0])
        R, t = synth.pose_matrix(0.05 * np.sin(0.7 * k), 0.0, 0.0, pos)
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=900 + k, cols=900)
        rng_ = np.linalg.norm(fr[..., :3], axis=-1)
        fr[rng_ > 80.0] = np.nan        # every return lands in a VALID cube (80 m + half a cube diagonal < 120 m): see cm_map.cu (1)
        f = oracle.scanreg_organised(fr)
        od = (R.astype(np.float32), (t + np.array([0.03, -0.02, 0.01])).astype(np.float32))
        before = tuple(om.origin())
        isos, stats = ctx.mapping_process([od], [f["lessSharp"]], [f["lessFlat"]])
        oR, ot, ost = om.process(od[0], od[1], f["lessSharp"], f["lessFlat"])
        n_shift_frames += tuple(om.origin()) != before
        assert stats[0]["iterations"] == ost["iterations"] and stats[0]["rows"] == ost["rows"], (k, stats[0], ost)
        assert (stats[0]["status"] == 1) == bool(ost["tooFewRef"]), (k, stats[0], ost)
        assert np.array_equal(isos[0][0], oR) and np.array_equal(isos[0][1], ot), (k, isos[0], oR, ot)
        gc, gs = ctx.map_surround(0)
        assert _same(gc, om.map_surround(0)) and _same(gs, om.map_surround(1)), k
    assert n_shift_frames >= 2
    for cls, which in ((0, 4), (1, 5)):
        g, cubes = ctx.map_export_sorted(0, cls)
        assert _same(g, om.cloud(which))
    ctx.close()


def test_estimate_sized_launches_change_nothing(cmb, synth, monkeypatch):
    """The filtered feature counts never leave the device: the correspondence kernels and the map insertion are launched for
    an ESTIMATE of them.  Forcing a gross under-estimate (search / fit kernels loop over their tiles, the insertion is skipped
    as a whole and repeated with exact sizes) must give the same poses and the same map, bit for bit; and alternating small /
    large / small frames must not make the cached CUDA graphs replay stale buffers (graphs are keyed by the allocation
    generation)."""
    sc = synth.make_scene(seed=61, extent=40.0, n_boxes=12, n_poles=10)
    S = 2
    seq_big = [_frames(synth, sc, 6, seed=300 + 40 * s, cols=900) for s in range(S)]
    seq_small = [_frames(synth, sc, 6, seed=700 + 40 * s, cols=128) for s in range(S)]

    def run(under):
        if under:
            monkeypatch.setenv("COOPERMAP_TEST_UNDERESTIMATE", "1")
        else:
            monkeypatch.delenv("COOPERMAP_TEST_UNDERESTIMATE", raising=False)
        ctx = cmb.Context(**MAP_CFG)
        ctx.mapping_create(S, 100000, 600000)
        out = []
        for k in range(6):
            seqs = seq_small if k in (1, 4) else seq_big                 # small, LARGE, small feature counts across steps
            fr = np.stack([seqs[s][k][2] for s in range(S)])
            odoms = [(seqs[s][k][0].astype(np.float32), seqs[s][k][1].astype(np.float32)) for s in range(S)]
            out.append(ctx.pipeline_step(fr, odoms))
        maps = [ctx.map_export_sorted(s, cls)[0] for s in range(S) for cls in (0, 1)]
        ctx.graph_builds()
        redos = ctx.insert_redos
        ctx.close()
        return out, maps, redos

    ref, maps_ref, redos_ref = run(False)
    und, maps_und, redos_und = run(True)
    assert redos_und >= 4 and redos_ref <= 2
    for (isos_a, st_a), (isos_b, st_b) in zip(ref, und):
        for s in range(S):
            assert np.array_equal(isos_a[s][0], isos_b[s][0]) and np.array_equal(isos_a[s][1], isos_b[s][1])
            assert st_a[s]["iterations"] == st_b[s]["iterations"] and st_a[s]["rows"] == st_b[s]["rows"]
    for a, b in zip(maps_ref, maps_und):
        assert _same(a, b)


def test_surround_and_full_map_clouds(cmb, oracle, synth):
    """'Map cloud out' of the boundary: FeatureMap::getSurroundFeature (valid cubes in (i, j, k) order, VoxelGrid order inside)
    and getFullMap (every cube re-filtered at the full-map leaf), byte for byte."""
    sc = synth.make_scene(seed=41, extent=60.0, n_boxes=16, n_poles=10)
    ctx = cmb.Context(**MAP_CFG)
    ctx.mapping_create(1, 100000, 600000)
    om = oracle.Mapping(map_params=ORACLE_MAP)
    assert all(len(c) == 0 for c in ctx.map_surround(0))            # no update yet: no valid cube
    for k, (R, t, fr) in enumerate(_frames(synth, sc, 4, model="HDL-32", cols=1024, speed=1.5)):
        f = oracle.scanreg_organised(fr)
        od = (R.astype(np.float32), t.astype(np.float32))
        ctx.mapping_process([od], [f["lessSharp"]], [f["lessFlat"]])
        om.process(od[0], od[1], f["lessSharp"], f["lessFlat"])
    gc, gs = ctx.map_surround(0)
    assert len(gs) > 5000
    assert _same(gc, om.map_surround(0)) and _same(gs, om.map_surround(1))
    # getFullMap, restated with the oracle's VoxelGrid on the cube clouds
    leaf = 2.0
    parts = []
    exp = [ctx.map_export_sorted(0, cls) for cls in (0, 1)]
    cubes = sorted(set(exp[0][1].tolist()) | set(exp[1][1].tolist()))
    for c in cubes:
        for cls in (0, 1):
            pts = exp[cls][0][exp[cls][1] == c]
            if len(pts):
                parts.append(oracle.voxel_filter(pts, leaf))
    want = np.concatenate(parts)
    assert _same(ctx.map_full(0, leaf), want)
    ctx.close()


def test_strided_pageable_sweeps_match_packed_entry(cmb, synth):
    """cm_pipeline_prefetch_strided_host / cm_pipeline_step_strided_host: one pageable cloud per stream with the 32-byte point
    stride of pcl::PointXYZI (and 48 / 12 / 22 byte strides -- 22 = a Velodyne PointCloud2 point_step) must give the poses and the map of the packed 16-byte entry, bit for bit,
    prefetched or not."""
    import ctypes as C
    sc = synth.make_scene(seed=91, extent=40.0, n_boxes=12, n_poles=10)
    S, rows, cols, NF = 3, 16, 900, 4
    seqs = [_frames(synth, sc, NF, seed=500 + 40 * s, cols=cols) for s in range(S)]

    def pcl_cloud(fr, stride):
        """(rows, cols, 4) float32 -> flat array of rows*cols points `stride` bytes apart (x, y, z at offset 0, junk elsewhere)"""
        n = fr.shape[0] * fr.shape[1]
        if stride % 4:                                    # a PointCloud2 point_step of 22: the floats are not aligned
            raw = np.full((n, stride), 0x5A, np.uint8)
            raw[:, :12] = np.ascontiguousarray(fr.reshape(n, 4)[:, :3]).view(np.uint8).reshape(n, 12)
            return raw
        buf = np.full((n, stride // 4), 7.25, np.float32)
        buf[:, :3] = fr.reshape(n, 4)[:, :3]
        if stride >= 32:
            buf[:, 4] = fr.reshape(n, 4)[:, 3]        # PointXYZI: intensity behind the padded xyz
        return buf

    def run(mode):
        ctx = cmb.Context(**MAP_CFG)
        ctx.mapping_create(S, 100000, 600000)
        out = []
        mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()
        clouds = [[pcl_cloud(seqs[s][k][2], mode if mode else 32) for s in range(S)] for k in range(NF)]
        for k in range(NF):
            od = ctx._pack_isos([(seqs[s][k][0].astype(np.float32), seqs[s][k][1].astype(np.float32)) for s in range(S)])
            if mode == 0:
                fr = np.ascontiguousarray(np.stack([seqs[s][k][2] for s in range(S)]))
                ctx.pipeline_step_packed(fr, od, mapped, stats)
            else:
                if k == 0:
                    ctx.pipeline_prefetch_strided(clouds[0], rows, cols)
                if k + 1 < NF and k != 2:                          # frame 3 is NOT prefetched: the synchronous strided path
                    ctx.pipeline_prefetch_strided(clouds[k + 1], rows, cols)
                ctx.pipeline_step_strided(clouds[k], rows, cols, od, mapped, stats)
            out.append((mapped.copy(), [stats[s].iterations for s in range(S)]))
        maps = [ctx.map_export_sorted(s, cls)[0] for s in range(S) for cls in (0, 1)]
        ctx.close()
        return out, maps

    ref, maps_ref = run(0)
    for stride in (32, 48, 12, 22):
        got, maps = run(stride)
        for (ma, ia), (mb, ib) in zip(ref, got):
            assert np.array_equal(ma, mb) and ia == ib, stride
        for a, b in zip(maps_ref, maps):
            assert _same(a, b), stride
