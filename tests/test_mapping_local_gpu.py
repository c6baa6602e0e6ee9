"""GPU parity: LaserMappingLocal (sliding-window map, LaserMappingLocal.cpp:33-78 over io_module/LocalFeatureMap.h) vs the oracle.
Poses, counters and the window clouds must be bit-identical, in the reference-as-written mode and with the mapped pose."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = dict(filter_corner=0.4, filter_surf=0.8)
ORACLE = dict(filterCorner=0.4, filterSurf=0.8)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("use_mapped", [False, True])
def test_mapping_local_sequence_bit_exact(cmb, oracle, synth, use_mapped):
    sc = synth.make_scene(seed=41, extent=40.0, n_boxes=12, n_poles=10)
    g = cmb.LaserMappingLocal(use_mapped_pose=use_mapped, **CFG)
    o = oracle.MappingLocal(map_params=ORACLE, use_mapped_pose=use_mapped)
    for k, (R, t) in enumerate(synth.trajectory(7, speed=0.5)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=100 + k, cols=900)
        f = oracle.scanreg_organised(fr)
        R = R.astype(np.float32); t = (t + np.array([0.02, -0.02, 0.01]) * k).astype(np.float32)   # drifting odometry
        gR, gt = g.process(R, t, f["lessSharp"], f["lessFlat"])
        oR, ot, ost = o.process(R, t, f["lessSharp"], f["lessFlat"])
        st = g.last_stats
        assert st["iterations"] == ost["iterations"] and st["rows"] == ost["rows"], (k, st, ost)
        assert (st["status"] == 1) == bool(ost["tooFewRef"])
        assert _same(gR, oR) and _same(gt, ot), (k, gt, ot)
        w = g.ctx.mapping_local_window(clouds=True)
        assert w["frames"] == ost["frames"] and w["nCorner"] == ost["nWindowCorner"] and w["nSurf"] == ost["nWindowSurf"]
        assert w["nSurroundCorner"] == ost["nSurroundCorner"] and w["nSurroundSurf"] == ost["nSurroundSurf"]
        assert w["accumDistance"] == ost["accumDistance"]
        assert _same(w["corner"], o.window(0)) and _same(w["surf"], o.window(1))
    assert ost["nSurroundSurf"] > 1000 and ost["iterations"] > 0
    if use_mapped:
        assert np.linalg.norm(gt - synth.trajectory(7, speed=0.5)[6][1]) < 0.1   # the window corrects the odometry drift
    g.ctx.close()


def test_mapping_local_clean_drops_one_more_than_counted(cmb, oracle, synth):
    """8 m of travel per frame: at the fifth frame clean() counts one frame behind the 30 m threshold and erases two."""
    sc = synth.make_scene(seed=41, extent=60.0, n_boxes=16, n_poles=12)
    g = cmb.LaserMappingLocal(use_mapped_pose=True, **CFG)
    o = oracle.MappingLocal(map_params=ORACLE, use_mapped_pose=True)
    seen = []
    for k, (R, t) in enumerate(synth.trajectory(6, speed=8.0, yaw_amp=0.02)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=300 + k, cols=600)
        f = oracle.scanreg_organised(fr)
        R = R.astype(np.float32); t = t.astype(np.float32)
        gR, gt = g.process(R, t, f["lessSharp"], f["lessFlat"])
        oR, ot, ost = o.process(R, t, f["lessSharp"], f["lessFlat"])
        assert _same(gR, oR) and _same(gt, ot), k
        w = g.ctx.mapping_local_window(clouds=True)
        assert w["frames"] == ost["frames"] and w["accumDistance"] == ost["accumDistance"]
        assert _same(w["surf"], o.window(1))
        seen.append(w["frames"])
    assert seen == [1, 2, 3, 4, 3, 4]
    g.ctx.close()


def test_mapping_local_requires_create_and_handles_empty_clouds(cmb):
    ctx = cmb.Context(**CFG)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    with pytest.raises(cmb.CoopermapError):
        ctx.mapping_local_process(eye, np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32))
    ctx.mapping_local_create(False)
    (R, t), st = ctx.mapping_local_process(eye, np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32))
    assert st["status"] == 1 and np.array_equal(R, eye[0]) and np.array_equal(t, eye[1])   # empty window: too few reference points
    w = ctx.mapping_local_window()
    assert w["frames"] == 1 and w["nCorner"] == 0 and w["nSurf"] == 0
    ctx.close()


def test_golden_mapping_sequence_gpu(cmb):
    """The committed 5-frame sequence (tests/golden/mapping_seq5.npz, made by the oracle): LaserMapping and LaserMappingLocal on the
    GPU reproduce every pose bit for bit."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mapping_seq5.npz"))
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    stages = {"map": cmb.LaserMapping(max_corner_points=100000, max_surf_points=600000, **cfg),
              "local": cmb.LaserMappingLocal(use_mapped_pose=True, **cfg), "literal": cmb.LaserMappingLocal(use_mapped_pose=False, **cfg)}
    for k in range(5):
        for name, m in stages.items():
            R, t = m.process(g["odomR%d" % k], g["odomT%d" % k], g["corner%d" % k], g["surf%d" % k])
            assert _same(R, g["%sR%d" % (name, k)]) and _same(t, g["%sT%d" % (name, k)]), (name, k)
            assert [m.last_stats["iterations"], m.last_stats["rows"]] == list(g["%sStats%d" % (name, k)][:2]), (name, k)
    for m in stages.values():
        m.ctx.close()
