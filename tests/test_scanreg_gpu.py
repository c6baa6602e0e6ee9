"""GPU parity: scan registration (mask, curvature, feature selection, per-ring voxel filter) bit-exact vs the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(g, o):
    ncloud = len(o["cloud"])
    assert np.array_equal(g["scanStart"], o["scanStart"]) and np.array_equal(g["scanEnd"], o["scanEnd"])
    assert np.array_equal(g["cloud"][:ncloud].view(np.uint32), o["cloud"].view(np.uint32))
    assert np.array_equal(g["picked"][:ncloud], o["picked"])                        # unreliable-point mask + picks
    assert np.array_equal(g["curvature"][:ncloud].view(np.uint32), o["curvature"].view(np.uint32))
    assert np.array_equal(g["classLabel"][:ncloud], o["classLabel"])                # pointClassify labels
    for k in ("sharpIdx", "lessSharpIdx", "flatIdx", "lessFlatRawIdx"):            # membership AND order of the lists
        assert np.array_equal(g[k], o[k]), k
    for k in ("sharp", "lessSharp", "flat", "lessFlat"):
        assert g[k].shape == o[k].shape and np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32)), k


@pytest.mark.parametrize("model,cols", [("VLP-16", None), ("HDL-64E", None), ("VLP-16", 601), ("HDL-32", 1000)])
def test_scanreg_bit_exact(ctx, oracle, synth, scene_small, model, cols):
    sc, _, _ = scene_small
    R, t = synth.pose_matrix(0.3, 0.01, -0.02, (2.0, -0.5, 0.1))
    fr = synth.simulate_scan(sc, R, t, model, seed=11, cols=cols)
    _check(ctx.scanreg_organised(fr, debug=True), oracle.scanreg_organised(fr))


def test_scanreg_batch_of_streams(ctx, oracle, synth, scene_small):
    sc, _, _ = scene_small
    frames = []
    for k, (R, t) in enumerate(synth.trajectory(5, speed=2.0)):
        frames.append(synth.simulate_scan(sc, R, t, "VLP-16", seed=40 + k, cols=900, dropout=0.02 * k))
    frames[3][5, 12:, :] = np.nan        # ring with 12 points left
    frames[3][7, :, :] = np.nan          # empty ring
    frames[4][:, :, :3] = np.nan         # a stream with no returns at all
    outs = ctx.scanreg_organised(np.stack(frames), debug=True)
    for fr, g in zip(frames, outs):
        _check(g, oracle.scanreg_organised(fr))


def test_scanreg_golden(ctx):
    g = np.load(os.path.join(GOLD, "scanreg_vlp16_600.npz"))
    r = ctx.scanreg_organised(g["frame"], debug=True)
    n = len(g["picked"])
    for k in ("sharpIdx", "lessSharpIdx", "flatIdx", "lessFlatRawIdx"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["picked"][:n], g["picked"].astype(np.int32))
    assert np.array_equal(r["classLabel"][:n], g["classLabel"].astype(np.int32))
    assert np.array_equal(r["lessFlat"].view(np.uint32), g["lessFlat"].view(np.uint32))


def test_scanreg_params(cmb, oracle, synth, scene_small):
    sc, _, _ = scene_small
    R, t = synth.pose_matrix(-0.2, 0, 0, (0, 1, 0))
    fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=13, cols=1200)
    c2 = cmb.Context(surface_curvature_threshold=0.1, max_surface_flat=2, max_corner_sharp=4, n_feature_regions=4,
                     less_flat_filter_size=0.4, blind_radius=1.0, blind_degree_threshold=1.0)
    o = oracle.scanreg_organised(fr, params=dict(surfaceCurvatureThreshold=0.1, maxSurfaceFlat=2, maxCornerSharp=4,
                                                 nFeatureRegions=4, lessFlatFilterSize=0.4, blindRadius=1.0,
                                                 blindDegreeThreshold=1.0))
    _check(c2.scanreg_organised(fr, debug=True), o)
    c2.close()


@pytest.mark.parametrize("model,lidar", [("VLP-16", 0), ("HDL-32", 1), ("HDL-64E", 2), ("Pandar40", 3)])
def test_scanreg_raw_sweep_entry(ctx, oracle, synth, scene_small, model, lidar):
    """MultiScanRegistration::process: DEVICE front end (ring binning, azimuth unwrap, relTime, stable per-ring append) + feature
    extraction == oracle, bit for bit: the ring-major cloud with its curvature field, every ring range, every list."""
    sc, _, _ = scene_small
    R, t = synth.pose_matrix(0.4, 0.0, 0.0, (1.0, 0.5, 0.0))
    fr = synth.simulate_scan(sc, R, t, model, seed=21, cols=1024 if lidar == 2 else None)
    sweep = synth.organised_to_sweep(fr)
    ok = np.where(np.isfinite(sweep[:, 0]))[0]
    sweep = sweep[ok[0]:ok[-1] + 1]      # a driver never starts / ends a sweep on a missing return (startOri / endOri would be NaN)
    _check(ctx.scanreg_sweep(sweep, lidar, debug=True), oracle.scanreg_sweep(sweep, lidar))


def test_scanreg_raw_sweep_with_imu_deskew(cmb, oracle, synth, scene_small):
    """The hasIMUData() branch (ScanRegistration.cpp:89-188): IMU messages integrated by cm_imu_push_host, every point of the sweep
    projected to the sweep start with the state interpolated at its relTime -- on the device, where the reference's forward-only
    _imuIdx becomes a prefix maximum -- and the /imu_trans points; all bit-identical to the oracle."""
    sc, _, _ = scene_small
    R, t = synth.pose_matrix(0.2, 0.0, 0.0, (0.5, 0.3, 0.0))
    fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=33, cols=1200)
    sweep = synth.organised_to_sweep(fr)
    ok = np.where(np.isfinite(sweep[:, 0]))[0]
    sweep = sweep[ok[0]:ok[-1] + 1]
    rng = np.random.default_rng(9)
    scan_time = 100.05
    # 100 Hz IMU from before the sweep to after its end (0.1 s): a gentle turn with accelerations; two messages share a gap of 30 ms
    stamps = np.concatenate([np.arange(99.90, 100.04, 0.01), np.arange(100.07, 100.22, 0.01)])
    imu = np.zeros((len(stamps), 7))
    imu[:, 0] = stamps
    imu[:, 1] = 0.01 * np.sin(3 * stamps); imu[:, 2] = 0.02 * np.cos(2 * stamps); imu[:, 3] = 3.1 + 0.4 * (stamps - 100.0)   # yaw crosses pi
    imu[imu[:, 3] > np.pi, 3] -= 2 * np.pi
    imu[:, 4:] = rng.normal(0, 0.5, (len(stamps), 3)) + np.array([0.0, 0.0, 9.81])
    ctx = cmb.Context()
    for m in imu:
        ctx.imu_push(*m)
    g = ctx.scanreg_sweep(sweep, 0, debug=True, imu_scan_time=scan_time)
    o, tr = oracle.scanreg_sweep_imu(sweep, 0, scan_time, imu)
    _check(g, o)
    assert np.array_equal(g["imu_trans"].view(np.uint32), tr.view(np.uint32))
    plain = oracle.scanreg_sweep(sweep, 0)
    assert np.abs(o["cloud"][:, :3] - plain["cloud"][:, :3]).max() > 1e-3         # the de-skew did move the points
    # an empty history is the plain entry
    ctx.imu_clear()
    _check(ctx.scanreg_sweep(sweep, 0, debug=True, imu_scan_time=scan_time), plain)
    ctx.close()
