"""Generates tests/golden/*.npz from the CPU oracle on seeded synthetic inputs.

The reference ships no tests, golden vectors or fixtures, so these pin the ORACLE's own behaviour
(regression guard) and give the GPU tests inputs/outputs that do not depend on the generator's code.
Run from the repo root:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

synth = importlib.import_module("the-cooper-mapper_b200.synth")
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sc = synth.make_scene(seed=0x5EED, extent=40.0, n_boxes=14, n_poles=12)
    R, t = synth.pose_matrix(0.04, 0.005, -0.004, (1.5, 0.3, 0.05))
    # --- scan registration: a reduced VLP-16-shaped frame (16 x 600) -------------------------------------------
    fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=1, cols=600)
    r = O.scanreg_organised(fr)
    np.savez_compressed(os.path.join(OUT, "scanreg_vlp16_600.npz"), frame=fr, scanStart=r["scanStart"], scanEnd=r["scanEnd"],
                        sharpIdx=r["sharpIdx"], lessSharpIdx=r["lessSharpIdx"], flatIdx=r["flatIdx"],
                        lessFlatRawIdx=r["lessFlatRawIdx"], picked=r["picked"].astype(np.int8),
                        curvature=r["curvature"], classLabel=r["classLabel"].astype(np.int8), lessFlat=r["lessFlat"])
    # --- voxel filter -------------------------------------------------------------------------------------------
    rng = np.random.default_rng(7)
    pts = np.zeros((3000, 4), np.float32)
    pts[:, :3] = rng.uniform(-6, 6, (3000, 3)) * np.array([1, 1, 0.2]); pts[:, 3] = rng.uniform(0, 16, 3000)
    np.savez_compressed(os.path.join(OUT, "voxel_3000.npz"), pts=pts, out_0p4=O.voxel_filter(pts, 0.4), out_1p0=O.voxel_filter(pts, 1.0))
    # --- scan-to-map: small map, per-iteration log ----------------------------------------------------------------
    mc, ms = synth.sample_map(sc, 0.4, seed=2, region=(-25, 25, -25, 25))
    full = synth.simulate_scan(sc, R, t, "VLP-16", seed=3)
    f = O.scanreg_organised(full)
    corner = O.voxel_filter(f["lessSharp"], 0.4); surf = O.voxel_filter(f["lessFlat"], 1.0)
    truth = np.array([-0.004, 0.005, 0.04, 1.5, 0.3, 0.05], np.float32)
    init = truth + np.array([0.004, -0.003, 0.006, 0.06, -0.05, 0.04], np.float32)
    p, st, log = O.scan_match(mc, ms, corner, surf, init, keep_log=True)
    np.savez_compressed(os.path.join(OUT, "match_small.npz"), mc=mc, ms=ms, corner=corner, surf=surf, init=init, pose=p,
                        iterations=st["iterations"], converged=st["converged"],
                        AtA=np.stack([e["AtA"] for e in log]), AtB=np.stack([e["AtB"] for e in log]),
                        x=np.stack([e["x"] for e in log]), counts=np.stack([e["counts"] for e in log]),
                        nnCorner0=log[0]["nnCorner"], nnSurf0=log[0]["nnSurf"])
    # --- mapping stages over a 5-frame sequence: LaserMapping (cube map) and LaserMappingLocal (sliding window) -----------------
    seq = {}
    lm = O.Mapping(map_params=dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4))
    ll = O.MappingLocal(map_params=dict(filterCorner=0.4, filterSurf=0.8), use_mapped_pose=True)
    lw = O.MappingLocal(map_params=dict(filterCorner=0.4, filterSurf=0.8), use_mapped_pose=False)
    for k, (Rk, tk) in enumerate(synth.trajectory(5, speed=0.5)):
        frk = synth.simulate_scan(sc, Rk, tk, "VLP-16", seed=50 + k, cols=1200)
        fk = O.scanreg_organised(frk)
        odomR = Rk.astype(np.float32); odomT = (tk + np.array([0.02, -0.02, 0.01]) * k).astype(np.float32)
        seq["corner%d" % k] = fk["lessSharp"]; seq["surf%d" % k] = fk["lessFlat"]; seq["odomR%d" % k] = odomR; seq["odomT%d" % k] = odomT
        for name, m in (("map", lm), ("local", ll), ("literal", lw)):
            oR, ot, st = m.process(odomR, odomT, fk["lessSharp"], fk["lessFlat"])
            seq["%sR%d" % (name, k)] = oR; seq["%sT%d" % (name, k)] = ot
            seq["%sStats%d" % (name, k)] = np.array([st["iterations"], st["rows"], st["tooFewRef"], st["nSurroundCorner"], st["nSurroundSurf"]], np.int32)
    np.savez_compressed(os.path.join(OUT, "mapping_seq5.npz"), **seq)
    for n in sorted(os.listdir(OUT)):
        if n.endswith(".npz"):
            print(n, os.path.getsize(os.path.join(OUT, n)))


if __name__ == "__main__":
    main()
