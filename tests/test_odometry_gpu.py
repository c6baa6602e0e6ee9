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
