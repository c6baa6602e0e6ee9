"""GPU parity: scan-to-scan odometry (LaserOdometry::process) against the oracle, frame after frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("model,cols,speed", [("VLP-16", None, 0.3), ("HDL-64E", 1024, 0.6)])
def test_odometry_sequence_bit_exact(cmb, oracle, synth, model, cols, speed):
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    ctx = cmb.Context()
    lo = cmb.LaserOdometry(ctx=ctx)
    oo = oracle.Odometry()
    for k, (R, t) in enumerate(synth.trajectory(5, speed=speed)):
        f = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, model, seed=100 + k, cols=cols))
        g = ctx.odometry_process(f["sharp"], f["lessSharp"], f["flat"], f["lessFlat"], trace=True)
        o = oo.process(f["sharp"], f["lessSharp"], f["flat"], f["lessFlat"])
        assert g["initialising"] == (k == 0)
        assert g["iterations"] == o["iterations"], k
        glog = [e for e in g["log"]][:len(o["log"])]
        for it, (a, b) in enumerate(zip(glog, o["log"])):
            assert a["rows"] == b["rows"], (k, it)
            assert np.array_equal(a["pose_in"], b["pose_in"]), (k, it)
            assert np.array_equal(a["x"], b["x"]), (k, it)
        # north-star tolerance (1e-4 m, 1e-5 rad), in practice bit-identical
        assert np.all(np.abs(g["transform"][:3] - o["transform"][:3]) <= 1e-5) and np.all(np.abs(g["transform"][3:] - o["transform"][3:]) <= 1e-4)
        assert np.array_equal(g["transform"], o["transform"])
        assert np.array_equal(g["R"], o["R"]) and np.array_equal(g["t"], o["t"])
        assert _same(g["corner_last"], o["corner_last"]) and _same(g["surf_last"], o["surf_last"])
    assert np.linalg.norm(g["t"]) > 0.5     # it did estimate motion
    ctx.close()


def test_odometry_degenerate_inputs(cmb, oracle):
    ctx = cmb.Context()
    rng = np.random.default_rng(3)
    few = np.zeros((5, 4), np.float32); few[:, :3] = rng.normal(size=(5, 3)) * 5; few[:, 3] = 1.25
    oo = oracle.Odometry()
    for _ in range(3):                      # last clouds too small: scanMatch never runs (LaserOdometry.cpp:338)
        g = ctx.odometry_process(few, few, few, few)
        o = oo.process(few, few, few, few)
        assert not g["matched"] and np.array_equal(g["transform"], o["transform"]) and np.array_equal(g["t"], o["t"])
        assert _same(g["corner_last"], o["corner_last"])
    e = np.zeros((0, 4), np.float32)
    g = ctx.odometry_process(e, e, e, e)
    assert g["iterations"] == 0
    ctx.close()


def test_odometry_batch_equals_per_stream_oracle(cmb, oracle, synth):
    """cm_odometry_batch_process_host: every stream of the batch follows its own LaserOdometry chain bit for bit -- different
    LiDAR motion per stream, a stream whose clouds are too small for scanMatch (LaserOdometry.cpp:338) in the middle of the
    sequence, ragged cloud sizes."""
    sc = synth.make_scene(seed=43, extent=40.0, n_boxes=12, n_poles=10)
    S, NF = 4, 5
    ctx = cmb.Context()
    ctx.odometry_batch_create(S, 4000, 30000, 8000, 30000)
    oos = [oracle.Odometry() for _ in range(S)]
    trajs = [synth.trajectory(NF, seed=s, speed=0.2 + 0.15 * s) for s in range(S)]
    moved = 0.0
    for k in range(NF):
        feats = []
        for s in range(S):
            R, t = trajs[s][k]
            f = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, "VLP-16", seed=200 + 10 * s + k, cols=900 if s % 2 else 1200))
            if s == 2 and k == 2:          # a sweep with almost no features: the NEXT frame of this stream skips scanMatch
                f = {n: f[n][:6] for n in ("sharp", "lessSharp", "flat", "lessFlat")}
            feats.append(f)
        got = ctx.odometry_batch_process([f["sharp"] for f in feats], [f["lessSharp"] for f in feats], [f["flat"] for f in feats],
                                         [f["lessFlat"] for f in feats])
        for s in range(S):
            f = feats[s]
            o = oos[s].process(f["sharp"], f["lessSharp"], f["flat"], f["lessFlat"])
            g = got[s]
            assert g["initialising"] == (k == 0)
            assert g["iterations"] == o["iterations"], (k, s)
            assert np.array_equal(g["transform"], o["transform"]), (k, s)
            assert np.array_equal(g["R"], o["R"]) and np.array_equal(g["t"], o["t"]), (k, s)
            assert _same(g["corner_last"], o["corner_last"]) and _same(g["surf_last"], o["surf_last"]), (k, s)
            if s == 2 and k == 3:
                assert not g["matched"]
            moved = max(moved, float(np.linalg.norm(g["t"])))
    assert moved > 0.5
    ctx.close()


def test_odometry_ring_walks_on_unsorted_last_clouds(cmb, oracle, synth):
    """The ring walks of LaserOdometry::scanMatch (:363-398, 427-476) stop at the FIRST point whose ring index is off by more than 2.5 --
    on a cloud that is not sorted by ring that is an arbitrary place.  The device walks skip whole 32-point chunks through their
    bounding boxes; they must still stop where the reference stops.  Last clouds in shuffled block order (rings interleaved):
    poses, iteration counts and projected clouds equal the oracle's literal walk."""
    sc = synth.make_scene(seed=47, extent=40.0, n_boxes=12, n_poles=10)
    S, NF = 2, 4
    ctx = cmb.Context()
    ctx.odometry_batch_create(S, 4000, 30000, 8000, 30000)
    oos = [oracle.Odometry() for _ in range(S)]
    trajs = [synth.trajectory(NF, seed=10 + s, speed=0.2 + 0.1 * s) for s in range(S)]
    rng = np.random.default_rng(3)

    def shuffle_blocks(c, block):
        """keep runs of `block` consecutive points together, permute the runs: ring indices go up and down along the cloud"""
        nb = (len(c) + block - 1) // block
        order = rng.permutation(nb)
        return np.concatenate([c[b * block:(b + 1) * block] for b in order]) if nb else c

    for k in range(NF):
        feats = []
        for s in range(S):
            R, t = trajs[s][k]
            f = oracle.scanreg_organised(synth.simulate_scan(sc, R, t, "VLP-16", seed=700 + 10 * s + k, cols=1000))
            f = dict(f)
            f["lessSharp"] = shuffle_blocks(f["lessSharp"], 23 if s == 0 else 150)
            f["lessFlat"] = shuffle_blocks(f["lessFlat"], 57 if s == 0 else 400)
            feats.append(f)
        got = ctx.odometry_batch_process([f["sharp"] for f in feats], [f["lessSharp"] for f in feats], [f["flat"] for f in feats],
                                         [f["lessFlat"] for f in feats])
        for s in range(S):
            f = feats[s]
            o = oos[s].process(f["sharp"], f["lessSharp"], f["flat"], f["lessFlat"])
            g = got[s]
            assert g["iterations"] == o["iterations"], (k, s)
            assert np.array_equal(g["transform"], o["transform"]), (k, s)
            assert np.array_equal(g["R"], o["R"]) and np.array_equal(g["t"], o["t"]), (k, s)
            assert _same(g["corner_last"], o["corner_last"]) and _same(g["surf_last"], o["surf_last"]), (k, s)
    ctx.close()
