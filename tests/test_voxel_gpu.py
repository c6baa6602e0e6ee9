"""GPU parity: batched voxel filter against the oracle's PCL VoxelGrid restatement (bit-exact)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cloud(rng, n, extent=8.0):
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(-extent, extent, (n, 3)) * np.array([1, 1, 0.3]); p[:, 3] = rng.uniform(0, 64, n)
    return p


@pytest.mark.parametrize("leaf", [0.2, 0.4, 1.0])
def test_voxel_single(ctx, oracle, leaf):
    p = _cloud(np.random.default_rng(1), 20000)
    p[17, 1] = np.nan; p[4000, 0] = np.inf
    a = ctx.voxel_filter(p, leaf); b = oracle.voxel_filter(p, leaf)
    assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_voxel_batch_ragged_and_empty(ctx, oracle):
    rng = np.random.default_rng(2)
    clouds = [_cloud(rng, n, e) for n, e in ((5000, 6), (0, 1), (1, 1), (12345, 20), (300, 0.2), (7, 3))]
    clouds.append(np.repeat(_cloud(rng, 1), 50, 0))           # 50 identical points -> one centroid
    outs = ctx.voxel_filter_batch(clouds, 0.5)
    for c, o in zip(clouds, outs):
        ref = oracle.voxel_filter(c, 0.5)
        assert o.shape == ref.shape and np.array_equal(o.view(np.uint32), ref.view(np.uint32))


def test_voxel_overflow_passthrough_and_idempotence(ctx, oracle):
    huge = np.array([[0, 0, 0, 1], [1e6, 1e6, 1e6, 2], [5, 5, 5, 3]], np.float32)
    assert np.array_equal(ctx.voxel_filter(huge, 0.01), huge)          # PCL: "leaf size too small" -> input unchanged
    p = _cloud(np.random.default_rng(3), 8000)
    once = ctx.voxel_filter(p, 1.0)
    assert np.array_equal(ctx.voxel_filter(once, 1.0), once)
    assert np.array_equal(once, oracle.voxel_filter(p, 1.0))


def test_voxel_golden(ctx):
    g = np.load(os.path.join(GOLD, "voxel_3000.npz"))
    assert np.array_equal(ctx.voxel_filter(g["pts"], 0.4), g["out_0p4"])
    assert np.array_equal(ctx.voxel_filter(g["pts"], 1.0), g["out_1p0"])


def test_voxel_full_frame_features(ctx, oracle, synth, scene_small):
    sc, _, _ = scene_small
    R, t = synth.pose_matrix(0.1, 0, 0, (2, 0, 0))
    f = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, "HDL-64E", seed=5))
    for cloud, leaf in ((f["lessFlat"], 0.8), (f["lessSharp"], 0.4), (f["lessFlat"], 1.0)):
        a = ctx.voxel_filter(cloud, leaf); b = oracle.voxel_filter(cloud, leaf)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("n", [0, 1, 31, 1025, 8191, 8192, 70000])
def test_voxel_cluster_kernel_equals_single_cta_kernel(ctx, oracle, monkeypatch, n):
    """The same clouds through vox_cluster_kernel (8 CTAs of a thread-block cluster per cloud, the path for a few clouds) and through
    vox_segment_kernel (one CTA per cloud, the path for many): byte-identical, and equal to the oracle; sizes around the slice and
    tile boundaries of the cluster kernel (1024-point tiles, 8 slices), with non-finite points and long runs of one voxel."""
    rng = np.random.default_rng(100 + n)
    p = _cloud(rng, n, 30.0)
    if n > 40:
        p[5, 0] = np.nan; p[n // 2, 2] = np.inf; p[n - 1, 1] = -np.inf
        p[n // 3:n // 3 + 20] = p[n // 3]                       # a run of identical points across lanes
        p[1023:1027, :3] = p[1023, :3]                          # a run that spans a tile boundary (n > 1027 only)
    outs = {}
    for force in ("1", "0"):
        monkeypatch.setenv("COOPERMAP_VOX_CLUSTER", force)
        outs[force] = [ctx.voxel_filter(p, leaf) for leaf in (0.3, 2.0)] + ctx.voxel_filter_batch([p, p[: n // 2], p[n // 3:]], 0.5)
    monkeypatch.delenv("COOPERMAP_VOX_CLUSTER")
    refs = [oracle.voxel_filter(p, 0.3), oracle.voxel_filter(p, 2.0), oracle.voxel_filter(p, 0.5), oracle.voxel_filter(p[: n // 2], 0.5),
            oracle.voxel_filter(p[n // 3:], 0.5)]
    for a, b, r in zip(outs["1"], outs["0"], refs):
        assert a.shape == b.shape == r.shape
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(a.view(np.uint32), r.view(np.uint32))
