"""GPU parity: localisation mode (LaserLocalization::process -> FeatureMap::scanMatchScan, own-cube neighbours) and the
on-disk cube map (FeatureMap::saveCloudToFiles / loadCloudFromFiles) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAP_CFG = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
ORACLE_MAP = dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _build(cmb, oracle, synth, extent=80.0, seed=61):
    """The same prebuilt map in a GPU context and in the oracle (the scene straddles several 50 m cubes)."""
    sc = synth.make_scene(seed=seed, extent=extent, n_boxes=24, n_poles=16)
    mc, ms = synth.sample_map(sc, 0.4, seed=seed + 1)
    ctx = cmb.Context(**MAP_CFG)
    ctx.mapping_create(1, 100000, 1500000)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    ctx.map_insert([mc], [ms], [eye])
    om = oracle.Mapping(map_params=ORACLE_MAP)
    om.map_update(np.zeros(3, np.float32))
    om.map_add(mc, ms, eye[0], eye[1])
    return sc, ctx, om


def test_localization_sequence_bit_exact(cmb, oracle, synth):
    sc, ctx, om = _build(cmb, oracle, synth)
    for cls, which in ((0, 4), (1, 5)):
        assert _same(ctx.map_export_sorted(0, cls)[0], om.cloud(which))
    n_before = [len(ctx.map_export(0, c)[0]) for c in (0, 1)]
    moved = 0.0
    for k, (R, t) in enumerate(synth.trajectory(6, speed=4.0)):          # 4 m per frame: queries cross cube faces (x = +-25 m)
        t = t + np.array([18.0, 3.0, 0.0])
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=300 + k, cols=900)
        f = oracle.scanreg_organised(fr)
        odom = (R.astype(np.float32), (t + np.array([0.05, -0.04, 0.02])).astype(np.float32))
        isos, stats = ctx.localization_process([odom], [f["lessSharp"]], [f["lessFlat"]])
        oR, ot, ost = om.localize(odom[0], odom[1], f["lessSharp"], f["lessFlat"])
        gR, gt = isos[0]
        assert stats[0]["iterations"] == ost["iterations"], (k, stats[0], ost)
        assert stats[0]["rows"] == ost["rows"] and stats[0]["line"] == ost["line"] and stats[0]["plane"] == ost["plane"]
        assert bool(stats[0]["converged"]) == bool(ost["converged"])
        # north-star tolerance: 1e-4 m / 1e-5 rad; in practice bit-identical
        assert np.max(np.abs(gt - ot)) <= 1e-4 and np.max(np.abs(gR - oR)) <= 1e-5
        assert np.array_equal(gR, oR) and np.array_equal(gt, ot), k
        moved = max(moved, float(np.linalg.norm(gt - odom[1])))
    assert moved > 0.01                                                   # the matcher did correct the pose
    assert [len(ctx.map_export(0, c)[0]) for c in (0, 1)] == n_before      # localisation never touches the map
    ctx.close()


def test_localization_sparse_cube_is_skipped(cmb, oracle, synth):
    """Queries whose own cube holds < 5 points are skipped (FeatureMap.h:523); too few rows ends the loop (:585-588)."""
    ctx = cmb.Context(**MAP_CFG)
    ctx.mapping_create(1, 10000, 10000)
    rng = np.random.default_rng(5)
    few_c = np.zeros((3, 4), np.float32); few_c[:, :3] = rng.uniform(-5, 5, (3, 3))
    few_s = np.zeros((4, 4), np.float32); few_s[:, :3] = rng.uniform(-5, 5, (4, 3))
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    ctx.map_insert([few_c], [few_s], [eye])
    om = oracle.Mapping(map_params=ORACLE_MAP)
    om.map_update(np.zeros(3, np.float32)); om.map_add(few_c, few_s, eye[0], eye[1])
    q = np.zeros((200, 4), np.float32); q[:, :3] = rng.uniform(-8, 8, (200, 3))
    isos, stats = ctx.localization_process([eye], [q], [q])
    oR, ot, ost = om.localize(eye[0], eye[1], q, q)
    assert stats[0]["rows"] == ost["rows"] == 0 and stats[0]["iterations"] == ost["iterations"] == 0
    assert stats[0]["status"] == cmb.CM_TOO_FEW_MATCHES and ost["tooFewMatches"] == 1
    assert np.array_equal(isos[0][1], ot)
    ctx.close()


def test_map_files_round_trip(cmb, oracle, synth, tmp_path):
    sc, ctx, om = _build(cmb, oracle, synth, extent=70.0, seed=71)
    d_gpu = tmp_path / "gpu"; d_ora = tmp_path / "oracle"; d_gpu.mkdir(); d_ora.mkdir()
    n_gpu = ctx.map_save(0, d_gpu)
    n_ora = om.save_files(str(d_ora))
    assert n_gpu == n_ora > 4
    assert (d_gpu / "index.txt").read_text() == (d_ora / "index.txt").read_text()         # same cubes, same order, same sizes
    for k in range(n_gpu):
        assert (d_gpu / ("%d.pcd" % k)).read_bytes() == (d_ora / ("%d.pcd" % k)).read_bytes()   # byte-identical PCD files
    # load the oracle-written files into a fresh GPU map and the GPU-written files into a fresh oracle map
    ctx2 = cmb.Context(**MAP_CFG); ctx2.mapping_create(1, 100000, 1500000)
    files, npts, misplaced = ctx2.map_load(0, d_ora)
    assert files == n_ora and misplaced == 0 and npts == sum(len(ctx.map_export(0, c)[0]) for c in (0, 1))
    om2 = oracle.Mapping(map_params=ORACLE_MAP)
    assert om2.load_files(str(d_gpu)) == n_gpu
    for cls, which in ((0, 4), (1, 5)):
        g, _ = ctx2.map_export_sorted(0, cls)
        assert _same(g, om.cloud(which)) and _same(om2.cloud(which), om.cloud(which))
    # an ascii PCD (as pcl::io::savePCDFileASCII would write) is read as well
    pts = oracle.read_pcd(str(d_ora / "0.pcd"))
    with open(d_ora / "0.pcd", "w") as f:
        f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
                "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA ascii\n" % (len(pts), len(pts)))
        for p in pts:
            f.write(" ".join(repr(float(v)) for v in p) + "\n")
    ctx3 = cmb.Context(**MAP_CFG); ctx3.mapping_create(1, 100000, 1500000)
    assert ctx3.map_load(0, d_ora)[0] == n_ora
    for cls, which in ((0, 4), (1, 5)):
        assert _same(ctx3.map_export_sorted(0, cls)[0], om.cloud(which))
    for c in (ctx, ctx2, ctx3):
        c.close()


def test_dynamic_map_paging(cmb, oracle, synth, tmp_path):
    """DynamicFeatureMap paging (DynamicFeatureMap.h:129-161, 504-677): index2.txt with GLOBAL cube indices, a window of cubes
    around the sensor resident, cubes read when they enter the window and dropped when they leave it, the lattice re-centred when
    the window would leave it.  The resident map must always be the union of the catalogued cubes of the current window, and the
    localisation matcher on the paged map must give the pose it gives on the fully loaded map."""
    cube = 20.0
    sc = synth.make_scene(seed=91, extent=260.0, n_boxes=120, n_poles=80)
    mc, ms = synth.sample_map(sc, 0.5, seed=92, region=(-130, 130, -25, 25))
    # the catalogue: one filtered cloud per (cube, class), as FeatureMap::saveCloudToFiles + indexConvert leave them
    files = {}
    lines = []
    count = 0
    for cls, pts in ((0, mc), (1, ms)):
        g = np.round(pts[:, :3] / np.float32(cube)).astype(np.int64)
        for key in sorted(set(map(tuple, g.tolist()))):
            sel = pts[(g == np.array(key)).all(axis=1)]
            fl = oracle.voxel_filter(sel, 0.4)
            oracle.write_pcd_binary(str(tmp_path / ("%d.pcd" % count)), fl)
            files[(cls,) + key] = fl
            lines.append("%d %d %d %d %d %d" % (count, cls, key[0], key[1], key[2], len(fl)))
            count += 1
    (tmp_path / "index2.txt").write_text("\n".join(lines) + "\n")
    grid = dict(cube_w=9, cube_h=9, cube_d=7, cube_size=cube, valid_distance=60.0)
    win = (5, 3, 3)
    ctx = cmb.Context(**MAP_CFG, **grid)
    ctx.mapping_create(1, 100000, 1500000)
    assert ctx.map_page_open(0, tmp_path, win) == count

    def expect(centre, cls):
        parts = []
        for key, fl in files.items():
            if key[0] == cls and all(abs(key[1 + a] - centre[a]) <= win[a] // 2 for a in range(3)):
                parts.append((key[1:], fl))
        parts.sort(key=lambda kv: (kv[0][2], kv[0][1], kv[0][0]))        # lattice index order: k, then j, then i
        return np.concatenate([p for _, p in parts]) if parts else np.zeros((0, 4), np.float32)

    full = cmb.Context(**MAP_CFG, cube_w=21, cube_h=9, cube_d=7, cube_size=cube, valid_distance=60.0)
    full.mapping_create(1, 100000, 1500000)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    full.map_insert([np.concatenate([f for k, f in files.items() if k[0] == 0])], [np.concatenate([f for k, f in files.items() if k[0] == 1])], [eye])
    total_loaded = total_evicted = 0
    for k, x in enumerate(np.linspace(-100.0, 100.0, 21)):                 # 10 m per step: a new cube every other step
        sensor = np.array([x, 2.0, 0.0], np.float32)
        nf, ne, npts = ctx.map_page_update(0, sensor)
        total_loaded += nf; total_evicted += ne
        q = sensor / np.float32(cube)
        centre = (np.sign(q) * np.floor(np.abs(q) + 0.5)).astype(int)      # roundf: halves away from zero (x = -90 -> cube -5)
        for cls in (0, 1):
            got, _ = ctx.map_export_sorted(0, cls)
            assert _same(got, expect(centre, cls)), (k, x, cls, len(got))
        if k % 5 == 2:
            R, t = synth.pose_matrix(0.0, 0.0, 0.0, (float(x), 2.0, 0.0))
            fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=500 + k, cols=900)
            rng_ = np.linalg.norm(fr[..., :3], axis=-1)
            fr[rng_ > 15.0] = np.nan                                         # queries stay inside the resident window (+-1 cube in y and z)
            f = oracle.scanreg_organised(fr)
            od = (R.astype(np.float32), (t + np.array([0.04, -0.03, 0.01])).astype(np.float32))
            a, sa = ctx.localization_process([od], [f["lessSharp"]], [f["lessFlat"]])
            b, sb = full.localization_process([od], [f["lessSharp"]], [f["lessFlat"]])
            assert sa[0]["rows"] == sb[0]["rows"] > 50 and sa[0]["iterations"] == sb[0]["iterations"]
            assert np.array_equal(a[0][0], b[0][0]) and np.array_equal(a[0][1], b[0][1])
    assert total_loaded > 30 and total_evicted > 30
    ctx.close(); full.close()
