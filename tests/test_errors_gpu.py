"""Error behaviour of the C ABI (SURVEY 8b "Errors"): soft outcomes are status codes >= 0, hard errors are < 0 with a message in
cm_last_error, nothing throws across the ABI and nothing falls back to the CPU."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_map_capacity_is_reported_not_silently_dropped(cmb, synth):
    sc = synth.make_scene(seed=5, extent=40.0, n_boxes=10, n_poles=6)
    mc, ms = synth.sample_map(sc, 0.4, seed=6)
    ctx = cmb.Context(map_filter_corner=0.4, map_filter_surf=0.4)
    ctx.mapping_create(1, 2000, 3000)                       # far too small for this map
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    with pytest.raises(cmb.CoopermapError) as e:
        ctx.map_insert([mc], [ms], [eye])
    assert "capacity" in str(e.value)
    ctx.close()


def test_step_returns_with_the_pose_and_the_insertion_reports_later(cmb, synth):
    """cm_pipeline_step_* returns when the poses are in; the map insertion it enqueued overflows a deliberately small map: the error
    is not lost -- cm_mapping_sync (or the next step) reports it."""
    sc = synth.make_scene(seed=61, extent=40.0, n_boxes=12, n_poles=10)
    R, t = synth.trajectory(1, speed=0.5)[0]
    fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=5, cols=720)[None]
    od = [(R.astype(np.float32), t.astype(np.float32))]
    for how in ("sync", "next_step"):
        ctx = cmb.Context(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
        ctx.mapping_create(1, 50, 200)                      # the first sweep's features do not fit
        isos, stats = ctx.pipeline_step(fr, od)             # empty map: pose = odometry, the insertion runs behind the return
        assert np.array_equal(isos[0][1], od[0][1])
        with pytest.raises(cmb.CoopermapError) as e:
            if how == "sync":
                ctx.mapping_sync()
            else:
                ctx.pipeline_step(fr, od)
        assert "capacity" in str(e.value)
        ctx.mapping_sync()                                  # reported once
        ctx.close()


def test_bad_arguments_return_error_codes(cmb):
    ctx = cmb.Context()
    L = ctx.L
    assert L.cm_mapping_create(ctx.h, C.c_int(0), C.c_size_t(10), C.c_size_t(10)) == -1            # CM_ERR_ARG
    assert L.cm_pipeline_step_host(ctx.h, None, C.c_int(16), C.c_int(100), None, None, None) == -1    # no mapping_create yet
    assert b"cm_mapping_create" in L.cm_last_error(ctx.h)
    pose = np.zeros(6, np.float32)
    assert L.cm_match_stateless_host(ctx.h, None, C.c_size_t(5), None, C.c_size_t(0), None, C.c_size_t(0), None, C.c_size_t(0),
                                     pose.ctypes.data_as(C.c_void_p), None, None, None, None) == -1
    assert L.cm_map_save_host(ctx.h, C.c_int(0), b"/nonexistent-dir", None) == -1
    ctx.mapping_create(1, 1000, 1000)
    assert L.cm_map_load_host(ctx.h, C.c_int(0), b"/nonexistent-dir", None, None, None) == -1
    assert b"index.txt" in L.cm_last_error(ctx.h)
    assert L.cm_map_load_host(ctx.h, C.c_int(3), b"/tmp", None, None, None) == -1                   # stream index out of range
    ctx.close()


def test_sweep_too_wide_for_one_cta_is_unsupported(cmb):
    ctx = cmb.Context()
    ctx.mapping_create(1, 1000, 1000)
    fr = np.full((1, 2, 20000, 4), np.nan, np.float32)
    with pytest.raises(cmb.CoopermapError) as e:
        ctx.pipeline_step(fr, [(np.eye(3, dtype=np.float32), np.zeros(3, np.float32))])
    assert "cols too large" in str(e.value)
    ctx.close()


def test_all_nan_and_empty_inputs_are_soft_outcomes(cmb):
    ctx = cmb.Context()
    ctx.mapping_create(2, 1000, 1000)
    eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
    fr = np.full((2, 16, 256, 4), np.nan, np.float32)                # no returns at all
    isos, stats = ctx.pipeline_step(fr, [eye, eye])
    assert [s["status"] for s in stats] == [cmb.CM_TOO_FEW_REF] * 2   # "reference cloud points too few.": pose = prediction
    assert np.array_equal(isos[0][0], eye[0]) and np.array_equal(isos[1][1], eye[1])
    e = np.zeros((0, 4), np.float32)
    pose, st, _ = ctx.match_stateless(e, e, e, e, np.zeros(6, np.float32))
    assert st["status"] == cmb.CM_TOO_FEW_REF and np.array_equal(pose, np.zeros(6, np.float32))
    out = ctx.voxel_filter(e, 0.4)
    assert out.shape == (0, 4)
    ctx.close()
