import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("the-cooper-mapper_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def cmb():
    return importlib.import_module("the-cooper-mapper_b200")


@pytest.fixture(scope="session")
def ctx(cmb):
    c = cmb.Context()
    yield c
    c.close()


@pytest.fixture(scope="session")
def scene_small(synth):
    """A 60 m scene with its sampled map (corner, surf) at 0.4 m."""
    sc = synth.make_scene(seed=11, extent=60.0, n_boxes=24, n_poles=20)
    mc, ms = synth.sample_map(sc, 0.4, seed=12)
    return sc, mc, ms


def frame_features(synth, oracle, scene, pose_args, model="VLP-16", seed=3, leaf_c=0.4, leaf_s=0.8):
    """Simulate one frame and run the ORACLE front end: returns (corner, surf) query clouds in the sensor frame."""
    yaw, pitch, roll, t = pose_args
    R, tt = synth.pose_matrix(yaw, pitch, roll, t)
    fr = synth.simulate_scan(scene, R, tt, model, seed=seed)
    r = oracle.scanreg_organised(fr)
    c = oracle.voxel_filter(r["lessSharp"], leaf_c)
    s = oracle.voxel_filter(r["lessFlat"], leaf_s)
    truth = np.array([roll, pitch, yaw, t[0], t[1], t[2]], np.float32)
    return c, s, truth
