"""CPU suite (-m "not gpu"): the oracle against known answers, independent numpy restatements and the committed
golden vectors; the C-ABI library loads and exports every symbol include/coopermap.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------------------------------------------------------
# shared math (restated Eigen routines) against LAPACK
# ---------------------------------------------------------------------------------------------------------------
def test_sincos_correctly_rounded(oracle):
    x = np.random.default_rng(0).uniform(-7, 7, 50000).astype(np.float32)
    s, c = oracle.sincosf(x)
    rs, rc = np.sin(x.astype(np.float64)), np.cos(x.astype(np.float64))
    assert np.all(np.abs(s - rs) <= 0.5000001 * np.spacing(np.abs(rs).astype(np.float32)))
    assert np.all(np.abs(c - rc) <= 0.5000001 * np.spacing(np.abs(rc).astype(np.float32)))


def test_atan_correctly_rounded(oracle):
    """cm_atan2f / cm_atanf (the raw-sweep front end's canonical atan, cm_math.h) against float64 libm rounded to float."""
    rng = np.random.default_rng(5)
    x = (rng.uniform(-1, 1, (200000, 2)) * rng.choice([1e-4, 1e-2, 1.0, 150.0], (200000, 2))).astype(np.float32)
    x[::97, 0] = 0.0; x[::89, 1] = 1e-30
    got = oracle.debug_math(7, x)
    want2 = np.arctan2(x[:, 0].astype(np.float64), x[:, 1].astype(np.float64)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        want1 = np.arctan((x[:, 0] / x[:, 1]).astype(np.float64)).astype(np.float32)
    assert np.array_equal(got[:, 0].view(np.uint32), want2.view(np.uint32))
    assert np.array_equal(got[:, 1].view(np.uint32), want1.view(np.uint32))
    # quadrant conventions at the axes
    ax = np.array([[0, 1], [0, -1], [1, 0], [-1, 0], [0, 0], [-0.0, -1]], np.float32)
    g = oracle.debug_math(7, ax)[:, 0]
    assert np.array_equal(g, np.arctan2(ax[:, 0].astype(np.float64), ax[:, 1].astype(np.float64)).astype(np.float32))


def test_eig3_against_lapack(oracle):
    rng = np.random.default_rng(1)
    for t in range(500):
        B = rng.normal(size=(3, 3)).astype(np.float32) * rng.choice([1e-3, 1, 10])
        A = (B @ B.T).astype(np.float32)
        if t % 5 == 0:
            A = (np.outer(B[0], B[0]) + 1e-6 * np.eye(3)).astype(np.float32)
        w, V = oracle.eig3(A)
        wr = np.linalg.eigvalsh(A.astype(np.float64))
        sc = max(np.abs(wr).max(), 1e-30)
        assert np.all(np.diff(w) >= 0)
        assert np.abs(w - wr).max() <= 2e-6 * sc
        assert np.abs(A.astype(np.float64) @ V - V * w).max() <= 4e-6 * sc
        assert np.abs(V.T @ V - np.eye(3)).max() < 1e-5


def test_eig6_qr_inverse_against_lapack(oracle):
    rng = np.random.default_rng(2)
    for _ in range(200):
        B = rng.normal(size=(20, 6)).astype(np.float32) * np.array([1, 1, 1, 30, 30, 30], np.float32)
        A = (B.T @ B).astype(np.float32)
        w, V = oracle.eig6(A)
        wr = np.linalg.eigvalsh(A.astype(np.float64))
        assert np.abs(w - wr).max() <= 1e-5 * np.abs(wr).max()
        assert np.abs(V.T @ V - np.eye(6)).max() < 1e-4
        b = rng.normal(size=6).astype(np.float32)
        xs = oracle.qr_solve(A, b)
        xr = np.linalg.solve(A.astype(np.float64), b)
        assert np.abs(xs - xr).max() <= 1e-6 * np.linalg.cond(A.astype(np.float64)) * np.abs(xr).max()
        assert np.abs(oracle.inverse6(V) @ V - np.eye(6)).max() < 1e-4
    for _ in range(200):
        A = rng.normal(size=(5, 3)).astype(np.float32) * 5
        b = -np.ones(5, np.float32)
        xs = oracle.qr_solve(A, b)
        xr = np.linalg.lstsq(A.astype(np.float64), b, rcond=None)[0]
        assert np.abs(xs - xr).max() < 1e-3 * max(1.0, np.abs(xr).max())


def test_pose_matrix_is_rz_ry_rx(oracle, synth):
    rng = np.random.default_rng(3)
    for _ in range(100):
        p = (rng.uniform(-1, 1, 6) * np.array([3, 1.5, 3, 10, 10, 10])).astype(np.float32)
        R = oracle.pose_to_matrix(p)
        Rr, _ = synth.pose_matrix(p[2], p[1], p[0])
        assert np.abs(R - Rr).max() < 1e-6
        back = oracle.iso_to_twist(R, p[3:])
        assert np.abs(back - p).max() < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# voxel filter against an independent numpy restatement of PCL VoxelGrid
# ---------------------------------------------------------------------------------------------------------------
def _voxel_numpy(p, leaf):
    p = p[np.isfinite(p[:, :3]).all(1)]
    if len(p) == 0:
        return p
    inv = np.float32(1.0) / np.float32(leaf)
    mn = np.floor(p[:, :3].min(0) * inv).astype(np.int64); mx = np.floor(p[:, :3].max(0) * inv).astype(np.int64)
    div = mx - mn + 1
    ijk = (np.floor(p[:, :3] * inv) - mn.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    out = []
    i = 0
    while i < len(order):
        j = i
        acc = np.zeros(4, np.float32)
        while j < len(order) and idx[order[j]] == idx[order[i]]:
            acc = (acc + p[order[j]]).astype(np.float32); j += 1
        out.append(acc / np.float32(j - i)); i = j
    return np.array(out, np.float32)


@pytest.mark.parametrize("leaf", [0.2, 0.4, 1.0])
def test_voxel_filter_matches_numpy(oracle, leaf):
    rng = np.random.default_rng(4)
    p = np.zeros((2000, 4), np.float32)
    p[:, :3] = rng.uniform(-8, 8, (2000, 3)) * np.array([1, 1, 0.3]); p[:, 3] = rng.uniform(0, 16, 2000)
    p[5, 0] = np.nan
    a = oracle.voxel_filter(p, leaf); b = _voxel_numpy(p, leaf)
    assert a.shape == b.shape and np.array_equal(a, b)


def test_voxel_filter_edge_cases(oracle):
    assert oracle.voxel_filter(np.zeros((0, 4), np.float32), 0.5).shape == (0, 4)
    one = np.array([[1, 2, 3, 4]], np.float32)
    assert np.array_equal(oracle.voxel_filter(one, 0.5), one)
    same = np.repeat(one, 7, 0)
    assert np.array_equal(oracle.voxel_filter(same, 0.5), one)
    huge = np.array([[0, 0, 0, 1], [1e6, 1e6, 1e6, 2]], np.float32)   # index overflow: PCL passes the input through
    assert np.array_equal(oracle.voxel_filter(huge, 0.01), huge)
    out = oracle.voxel_filter(np.random.default_rng(5).uniform(-3, 3, (500, 4)).astype(np.float32), 1.0)
    assert np.array_equal(oracle.voxel_filter(out, 1.0), out)          # idempotent on its own output


# ---------------------------------------------------------------------------------------------------------------
# KNN: the reference's own nanoflann against brute force
# ---------------------------------------------------------------------------------------------------------------
def test_nanoflann_equals_brute_force(oracle):
    if not oracle.lib().has_nanoflann:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(6)
    pts = np.zeros((20000, 4), np.float32); pts[:, :3] = rng.uniform(-30, 30, (20000, 3))
    q = rng.uniform(-30, 30, (2000, 3)).astype(np.float32)
    bi, bd = oracle.knn(pts, q, 5, nanoflann=False)
    ni, nd = oracle.knn(pts, q, 5, nanoflann=True)
    assert np.array_equal(bi, ni) and np.array_equal(bd, nd)
    bi, bd = oracle.knn(pts[:3], q[:10], 5, nanoflann=False)      # fewer than k points: gate must reject
    assert np.all(bi[:, 3:] == -1) and np.all(bd[:, 4] > 1e30)


# ---------------------------------------------------------------------------------------------------------------
# scan registration: known answers
# ---------------------------------------------------------------------------------------------------------------
def _wall_frame(rows=4, cols=400, dist=8.0, corner=False, jump=False):
    """Organised frame looking at a wall x = dist (optionally an L-shaped corner or a near box)."""
    az = np.deg2rad(np.linspace(50, -50, cols))
    el = np.deg2rad(np.linspace(-3, 3, rows))
    fr = np.zeros((rows, cols, 4), np.float32)
    for r in range(rows):
        d = np.stack([np.cos(el[r]) * np.cos(az), np.cos(el[r]) * np.sin(az), np.full(cols, np.sin(el[r]))], -1)
        rng_ = dist / d[:, 0]
        if corner:      # second wall y = -3 closes the corner on the right-hand side
            r2 = -3.0 / np.minimum(d[:, 1], -1e-9)
            rng_ = np.where(d[:, 1] < 0, np.minimum(rng_, r2), rng_)
        if jump:        # a box 3 m in front of the wall over a span of azimuths
            m = (np.abs(az) < np.deg2rad(8))
            rng_ = np.where(m, 4.0 / d[:, 0], rng_)
        fr[r, :, :3] = d * rng_[:, None]
    return fr


def test_wall_gives_flats_and_no_sharps(oracle):
    r = oracle.scanreg_organised(_wall_frame())
    assert len(r["sharpIdx"]) == 0 and len(r["lessSharpIdx"]) == 0
    assert len(r["flatIdx"]) == 4 * 6 * 4          # maxSurfaceFlat per region, 6 regions, 4 rings
    assert np.all(r["picked"][r["flatIdx"]] == 3)
    assert len(r["lessFlat"]) > 0


def test_wall_corner_gives_corner_sharp_at_crease(oracle):
    fr = _wall_frame(corner=True)
    r = oracle.scanreg_organised(fr)
    assert len(r["lessSharpIdx"]) >= 4
    cloud = r["cloud"]
    crease = np.array([8.0, -3.0])
    d = np.linalg.norm(cloud[r["lessSharpIdx"], :2] - crease, axis=1)
    assert np.all(d < 0.35)
    assert set(np.unique(r["classLabel"][r["lessSharpIdx"]])) == {1}   # CORNER_SHARP


def test_range_jump_marks_edge_broken_on_near_side(oracle):
    r = oracle.scanreg_organised(_wall_frame(jump=True))
    edge = np.where(r["picked"] == -2)[0]
    assert len(edge) >= 4
    assert np.all(np.abs(r["cloud"][edge, 0] - 4.0) < 1e-3)      # EDGE_BROKEN sits on the near box, not on the wall
    assert set(edge).issubset(set(r["sharpIdx"]))                # emitted as sharp + lessSharp (ScanRegistration.cpp:297-302)
    assert np.any(r["picked"] == -3)                             # far side is NEAR_BLOCK


def test_short_and_empty_rings_are_skipped(oracle):
    fr = _wall_frame(rows=3, cols=400)
    fr[1, 10:, :] = np.nan                  # 10 valid points: scanEnd <= scanStart + 2*5 -> skipped (ScanRegistration.cpp:205)
    fr[2, :, :] = np.nan                    # empty ring
    r = oracle.scanreg_organised(fr)
    assert r["scanEnd"][1] - r["scanStart"][1] == 9
    ring_of = np.searchsorted(r["scanStart"], r["flatIdx"], side="right") - 1
    assert set(ring_of) == {0}
    assert oracle.scanreg_organised(np.full((2, 50, 4), np.nan, np.float32))["cloud"].shape[0] == 0


def test_organised_and_sweep_entries_agree(oracle, synth):
    sc = synth.make_scene(seed=31, extent=40.0, n_boxes=12, n_poles=8)
    R, t = synth.pose_matrix(0.02, 0.0, 0.0, (0.5, 0.1, 0.0))
    fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=2, cols=900)
    a = oracle.scanreg_organised(fr, params=dict(blindRadius=0.01))
    b = oracle.scanreg_sweep(synth.organised_to_sweep(fr), 0)
    # the sweep entry swaps axes (x,y,z) <- (y,z,x) and bins rings by elevation: same rings, same feature indices
    assert np.array_equal(a["scanStart"], b["scanStart"]) and np.array_equal(a["scanEnd"], b["scanEnd"])
    assert np.array_equal(a["cloud"][:, [1, 2, 0]], b["cloud"][:, :3])
    for k in ("sharpIdx", "lessSharpIdx", "flatIdx", "lessFlatRawIdx", "picked"):
        assert np.array_equal(a[k], b[k]), k


# ---------------------------------------------------------------------------------------------------------------
# golden vectors
# ---------------------------------------------------------------------------------------------------------------
def test_golden_scanreg(oracle):
    g = np.load(os.path.join(GOLD, "scanreg_vlp16_600.npz"))
    r = oracle.scanreg_organised(g["frame"])
    for k in ("scanStart", "scanEnd", "sharpIdx", "lessSharpIdx", "flatIdx", "lessFlatRawIdx"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["picked"], g["picked"].astype(np.int32))
    assert np.array_equal(r["classLabel"], g["classLabel"].astype(np.int32))
    assert np.array_equal(r["curvature"].view(np.uint32), g["curvature"].view(np.uint32))
    assert np.array_equal(r["lessFlat"].view(np.uint32), g["lessFlat"].view(np.uint32))


def test_golden_voxel(oracle):
    g = np.load(os.path.join(GOLD, "voxel_3000.npz"))
    assert np.array_equal(oracle.voxel_filter(g["pts"], 0.4), g["out_0p4"])
    assert np.array_equal(oracle.voxel_filter(g["pts"], 1.0), g["out_1p0"])


def test_golden_match(oracle):
    g = np.load(os.path.join(GOLD, "match_small.npz"))
    for nf in (False, True):
        p, st, log = oracle.scan_match(g["mc"], g["ms"], g["corner"], g["surf"], g["init"], nanoflann=nf, keep_log=True)
        assert st["iterations"] == int(g["iterations"]) and st["converged"] == bool(g["converged"])
        assert np.array_equal(p, g["pose"])
        assert np.array_equal(np.stack([e["AtA"] for e in log]), g["AtA"])
        assert np.array_equal(np.stack([e["x"] for e in log]), g["x"])
        assert np.array_equal(log[0]["nnCorner"], g["nnCorner0"]) and np.array_equal(log[0]["nnSurf"], g["nnSurf0"])


def test_golden_mapping_sequence(oracle):
    """LaserMapping (cube map) and LaserMappingLocal (window; mapped-pose and as-written modes) over the committed 5-frame sequence."""
    g = np.load(os.path.join(GOLD, "mapping_seq5.npz"))
    prm = dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4)
    stages = {"map": oracle.Mapping(map_params=prm), "local": oracle.MappingLocal(map_params=prm, use_mapped_pose=True),
              "literal": oracle.MappingLocal(map_params=prm, use_mapped_pose=False)}
    for k in range(5):
        for name, m in stages.items():
            oR, ot, st = m.process(g["odomR%d" % k], g["odomT%d" % k], g["corner%d" % k], g["surf%d" % k])
            assert np.array_equal(oR, g["%sR%d" % (name, k)]) and np.array_equal(ot, g["%sT%d" % (name, k)]), (name, k)
            assert [st["iterations"], st["rows"], st["tooFewRef"], st["nSurroundCorner"], st["nSurroundSurf"]] == list(g["%sStats%d" % (name, k)])
    assert int(g["mapStats4"][0]) > 0 and int(g["localStats4"][0]) > 0       # the sequence does exercise the solver


# ---------------------------------------------------------------------------------------------------------------
# solver behaviour
# ---------------------------------------------------------------------------------------------------------------
def test_match_recovers_pose_and_gates(oracle, synth, scene_small):
    from conftest import frame_features
    sc, mc, ms = scene_small
    c, s, truth = frame_features(synth, oracle, sc, (0.05, 0.01, -0.008, (3.0, 0.4, 0.1)))
    init = truth + np.array([0.005, -0.004, 0.008, 0.08, -0.06, 0.05], np.float32)
    p, st, _ = oracle.scan_match(mc, ms, c, s, init)
    assert st["converged"] and not st["ok"]          # useScore = false -> returns false on convergence (quirk 5)
    assert np.all(np.abs(p[3:] - truth[3:]) < 0.03) and np.all(np.abs(p[:3] - truth[:3]) < 2e-3)
    p2, st2, _ = oracle.scan_match(mc[:49], ms, c, s, init)
    assert st2["tooFewRef"] and np.array_equal(p2, init)
    far = s.copy(); far[:, :3] += 500
    p3, st3, _ = oracle.scan_match(mc, ms, c[:0], far, init)
    assert st3["tooFewMatches"] and st3["iterations"] == 0 and np.array_equal(p3, init)


def test_scan_against_its_own_map_converges_at_once(oracle, synth):
    """Known answer: a frame matched against a map made of ITS OWN features, starting from the true pose, needs no
    correction -- the first update is below the thresholds and the pose stays put."""
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    R, t = synth.pose_matrix(0.05, 0.0, 0.0, (1.0, -0.5, 0.0))
    f = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, "VLP-16", seed=7, cols=1200))
    c = oracle.voxel_filter(f["lessSharp"], 0.4); s = oracle.voxel_filter(f["lessFlat"], 0.8)

    def to_map(p):   # the same features in the map frame
        q = p.copy(); q[:, :3] = (p[:, :3].astype(np.float64) @ R.T + t).astype(np.float32); return q
    truth = np.array([0.0, 0.0, 0.05, 1.0, -0.5, 0.0], np.float32)
    p, st, log = oracle.scan_match(to_map(f["lessSharp"]), to_map(f["lessFlat"]), c, s, truth, keep_log=True)
    assert not st["tooFewRef"] and st["converged"] and st["iterations"] <= 3, st
    # "~0": the queries are voxel centroids, the map holds the raw (noisy, sigma 1 cm) returns -> residuals of a few mm remain
    assert np.all(np.abs(p[:3] - truth[:3]) < 5e-4) and np.all(np.abs(p[3:] - truth[3:]) < 1e-2), p - truth
    assert np.abs(log[0]["x"]).max() < 1e-2                 # the very first Gauss-Newton step is already ~0


def test_corridor_is_degenerate_along_its_axis(oracle, synth):
    """Known answer: two parallel walls leave the along-corridor translation unobservable; the eigenvalue test at
    iteration 0 (ScanMatch.cpp:211-240) flags it and the projected update does not move the pose along the corridor."""
    sc = synth.make_scene(seed=1, extent=60.0, corridor=True)
    mc, ms = synth.sample_map(sc, 0.4, seed=2)
    pole = np.zeros((80, 4), np.float32); pole[:, 0] = 200.0; pole[:, 2] = np.linspace(-1.8, 6, 80)   # passes the >= 50 corner gate, matches nothing
    mc = np.concatenate([mc, pole])
    R, t = synth.pose_matrix(0.0, 0.0, 0.0, (0.0, 0.0, 0.0))
    r = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, "VLP-16", seed=9))
    c = oracle.voxel_filter(r["lessSharp"], 0.4); s = oracle.voxel_filter(r["lessFlat"], 0.8)
    init = np.array([0.002, -0.003, 0.004, 0.05, 0.04, -0.03], np.float32)
    p, st, log = oracle.scan_match(mc, ms, c, s, init, keep_log=True)
    assert st["degenerate"] and st["iterations"] >= 1
    w = np.linalg.eigvalsh(log[0]["AtA"].astype(np.float64))
    assert w[0] < 100.0 < w[-1]                               # the threshold of ScanMatch.cpp:219


def test_mapping_loop_bootstraps_and_tracks(oracle, synth):
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    m = oracle.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4))
    errs = []
    for k, (R, t) in enumerate(synth.trajectory(6, speed=0.5)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=100 + k, cols=900)
        f = oracle.scanreg_organised(fr)
        odomR = R.astype(np.float32); odomT = (t + np.array([0.02, -0.02, 0.01]) * k).astype(np.float32)   # drifting odometry
        oR, ot, st = m.process(odomR, odomT, f["lessSharp"], f["lessFlat"])
        if k == 0:
            assert st["tooFewRef"] == 1                 # empty map on the first frame: pose = prediction, features inserted
        errs.append(np.linalg.norm(ot - t))
    assert st["nSurroundSurf"] > 1000
    assert errs[-1] < 0.08                              # mapping corrects the 5 x (0.02, 0.02, 0.01) odometry drift


def test_mapping_local_window_literal_and_intended(oracle, synth):
    """LaserMappingLocal over LocalFeatureMap (LaserMappingLocal.cpp:33-78, LocalFeatureMap.h:62-99).  As written the frames
    enter the window with the never-assigned _transformTobeMapped (identity): nothing moves, nothing is dropped.  With the
    mapped pose the window tracks, accumulates FrameUpdater's distance and clean() erases one frame more than it counted."""
    sc = synth.make_scene(seed=41, extent=60.0, n_boxes=16, n_poles=12)
    prm = dict(filterCorner=0.4, filterSurf=0.8)
    lit = oracle.MappingLocal(map_params=prm, use_mapped_pose=False)
    itd = oracle.MappingLocal(map_params=prm, use_mapped_pose=True)
    frames_seen = []
    for k, (R, t) in enumerate(synth.trajectory(6, speed=8.0, yaw_amp=0.02)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=300 + k, cols=600)
        f = oracle.scanreg_organised(fr)
        R = R.astype(np.float32); t = t.astype(np.float32)
        _, _, sl = lit.process(R, t, f["lessSharp"], f["lessFlat"])
        oR, ot, si = itd.process(R, t, f["lessSharp"], f["lessFlat"])
        assert sl["frames"] == k + 1 and sl["accumDistance"] == 0.0          # literal: identity key pose, clean() never fires
        if k == 0:
            assert sl["tooFewRef"] == 1 and si["tooFewRef"] == 1            # empty window on the first frame
            assert np.array_equal(ot, t)
        else:
            assert si["nSurroundSurf"] > 0 and np.linalg.norm(ot - t) < 0.15
        frames_seen.append(si["frames"])
        # window clouds of the literal mode are the voxel-filtered frames, untransformed
        if k == 1:
            c0 = oracle.voxel_filter(f["lessSharp"], 0.4)
            assert np.array_equal(lit.window(0)[-len(c0):], c0)
    # 8 m per frame: accum 0, 8, 16, 24, 32 -> at the fifth push frame 0 is <= 32 - 30, clean() erases TWO frames (deleteNum + 1)
    assert abs(si["accumDistance"] - 40.0) < 1.5
    assert frames_seen[:4] == [1, 2, 3, 4] and frames_seen[4] == 3, frames_seen
    assert frames_seen[5] == 4, frames_seen                                   # accum 40: the oldest frame (16) is > 40 - 30, nothing dropped


# ---------------------------------------------------------------------------------------------------------------
# the C ABI
# ---------------------------------------------------------------------------------------------------------------
def test_capi_exports_every_declared_symbol(cmb):
    hdr = open(os.path.join(ROOT, "include", "coopermap.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 8
    L = ctypes.CDLL(cmb.lib_path())
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_capi_fails_loudly_without_gpu(cmb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cmb.CoopermapError):
        cmb.Context()


def test_oracle_map_files_round_trip(oracle, synth, tmp_path):
    """FeatureMap::saveCloudToFiles -> loadCloudFromFiles (FeatureMap.h:378-462) restated: index.txt order and PCD payload."""
    sc = synth.make_scene(seed=61, extent=70.0, n_boxes=16, n_poles=10)
    mc, ms = synth.sample_map(sc, 0.6, seed=62)
    P = dict(mapFilterCorner=0.4, mapFilterSurf=0.4)
    om = oracle.Mapping(map_params=P)
    om.map_update(np.zeros(3, np.float32))
    om.map_add(mc, ms, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    n = om.save_files(str(tmp_path))
    lines = [l.split() for l in (tmp_path / "index.txt").read_text().splitlines()]
    assert n == len(lines) > 2 and [int(l[0]) for l in lines] == list(range(n))
    order = [(int(l[2]), int(l[3]), int(l[4]), int(l[1])) for l in lines]
    assert order == sorted(order)                                   # i, then j, then k; corner (0) before surf (1)
    raw = (tmp_path / "0.pcd").read_bytes()
    assert raw.startswith(b"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n")
    assert len(raw) - raw.index(b"DATA binary\n") - len(b"DATA binary\n") == 16 * int(lines[0][5])
    om2 = oracle.Mapping(map_params=P)
    assert om2.load_files(str(tmp_path)) == n
    for which in (4, 5):
        a, b = om.cloud(which), om2.cloud(which)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
