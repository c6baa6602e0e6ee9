"""The C++ facade (include/coopermap.hpp: the reference's stage classes over the C ABI).  CPU: it compiles and links against
libcoopermap.so with the host compiler alone.  GPU: the chain OrganisedScanRegistration -> LaserOdometry -> LaserMapping driven
from C++ gives the poses of the (oracle-checked) ctypes path, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "the-cooper-mapper_b200")


def _build(tmp_path):
    exe = str(tmp_path / "facade_chain")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_chain.cpp"),
           "-o", exe, "-L" + PKG, "-lcoopermap", "-Wl,-rpath," + PKG]
    subprocess.check_call(cmd)
    return exe


def test_facade_compiles_and_links(cmb, tmp_path):
    cmb.lib_path()
    if not os.path.exists(os.path.join(PKG, "libcoopermap.so")):
        pytest.skip("libcoopermap.so not built")
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    # without arguments the program exits with its usage code before touching CUDA
    assert subprocess.run([exe]).returncode == 2


@pytest.mark.gpu
def test_facade_chain_equals_ctypes_chain(cmb, synth, tmp_path):
    exe = _build(tmp_path)
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    NF, rows, cols = 8, 16, 900
    frames = np.stack([synth.simulate_scan(sc, R, t, "VLP-16", seed=0x2000 + k, cols=cols)
                       for k, (R, t) in enumerate(synth.trajectory(NF, speed=0.1, yaw_amp=0.02))]).astype(np.float32)
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([NF, rows, cols], np.int32).tofile(f); frames.tofile(f)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(tmp_path / "out.bin", np.float32).reshape(NF, 14)
    cfg = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
    c_sr = cmb.Context(**cfg); c_od = cmb.Context(**cfg); c_od.odometry_reset(); c_mp = cmb.Context(**cfg); c_mp.mapping_create(1, 100000, 800000)
    for k in range(NF):
        g = c_sr.scanreg_organised(frames[k], debug=True)
        go = c_od.odometry_process(g["sharp"], g["lessSharp"], g["flat"], g["lessFlat"])
        isos, stats = c_mp.mapping_process([(go["R"], go["t"])], [go["corner_last"]], [go["surf_last"]])
        assert np.array_equal(out[k, :9].reshape(3, 3), isos[0][0]) and np.array_equal(out[k, 9:12], isos[0][1]), k
        assert int(out[k, 12]) == int(g["scanEnd"][-1]) + 1 and int(out[k, 13]) == stats[0]["status"]   # valid returns of the sweep
    for c in (c_sr, c_od, c_mp):
        c.close()
