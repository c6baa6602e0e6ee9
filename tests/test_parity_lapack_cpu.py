"""How far is "bit-exact against our restatement of Eigen" from "the reference linked against a real LAPACK-class library"?

The reference's eigen-solvers, QR and GEMM live in Eigen (not under /root/reference, not installable here), so the oracle
restates them (csrc/cm_math.h) and the GPU is bit-compared with that restatement.  This file measures the only thing that can
be measured here about the gap: every *decision* of the path that goes through such a routine is re-taken with an INDEPENDENT
float32 implementation -- numpy.linalg (LAPACK ssyevd / sgelsd / sgesv) and a float32 BLAS GEMM -- and the test counts how
many decisions flip and how far the final pose moves:

  * pointClassify (ScanRegistration.cpp:547-666): labels of every classified point of a sweep,
  * findLine / findPlane (feature_utils.h:108-204) at the first Gauss-Newton evaluation: accepted / rejected,
  * the degeneracy decision of iteration 0 (ScanMatch.cpp:211-235),
  * the whole Gauss-Newton loop (ScanMatch.cpp:91-260): iterations and final pose.

The neighbour sets come from the reference's own nanoflann (pinned), so they are shared.  The numbers printed here are quoted in
DESIGN.md section 2; the asserts are the north-star tolerance (1e-4 m, 1e-5 rad) on the pose and small flip-rate bounds.
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = np.float32

CORNER_SHARP, SURFACE_FLAT, ONESIDE_FLAT, MESSY, NONE = 1, -1, 5, 9, 0x7f


# ---- pointClassify with LAPACK -------------------------------------------------------------------------------------------
def _window_line(P):
    """P: (W, 6, 3) float32 windows (points in the reference's summation order) -> (is_line, direction)."""
    c = P.sum(axis=1, dtype=F) / F(P.shape[1])
    D = P - c[:, None, :]
    A = np.einsum("wki,wkj->wij", D, D).astype(F) / F(P.shape[1])
    w, V = np.linalg.eigh(A)                       # float32 in -> LAPACK ssyevd, ascending eigenvalues
    line = (w[:, 2] > F(100) * w[:, 1]) & (w[:, 2] > F(10000) * w[:, 0])
    v = V[:, :, 2]
    k = np.cross(D, v[:, None, :]).astype(F)
    dist = np.sqrt((k * k).sum(axis=2, dtype=F)) / np.sqrt((v * v).sum(axis=1, dtype=F))[:, None]
    line &= ~(np.abs(dist.astype(np.float64)) > 0.08).any(axis=1)
    return line, v


def classify_lapack(xyz, idx, R=5):
    """Labels of the cloud points `idx` (pointClassify with numpy.linalg.eigh in place of Eigen's SelfAdjointEigenSolver)."""
    back = np.stack([xyz[idx - j] for j in range(R + 1)], axis=1)            # c, c-1, .., c-R
    fwd = np.stack([xyz[idx - j] for j in range(-R, 1)], axis=1)             # c+R, .., c
    l1, v1 = _window_line(back)
    l2, v2 = _window_line(fwd)
    cosd = (v1 * v2).sum(axis=1, dtype=F) / (np.sqrt((v1 * v1).sum(axis=1, dtype=F)) * np.sqrt((v2 * v2).sum(axis=1, dtype=F)))
    cosd = cosd.astype(np.float64)
    lab = np.full(len(idx), MESSY, np.int32)
    lab[l1 | l2] = ONESIDE_FLAT
    both = l1 & l2
    flat = both & ((cosd < np.cos(np.deg2rad(175.0))) | (cosd > np.cos(np.deg2rad(5.0))))
    sharp = both & ~flat & (cosd > np.cos(np.deg2rad(135.0))) & (cosd < np.cos(np.deg2rad(45.0)))
    lab[flat] = SURFACE_FLAT
    lab[sharp] = CORNER_SHARP
    return lab


def _classify_flips(oracle, frame):
    r = oracle.scanreg_organised(frame)
    lab = r["classLabel"]
    idx = np.nonzero(lab != NONE)[0]
    got = classify_lapack(r["cloud"][:, :3].astype(F), idx)
    flips = int((got != lab[idx]).sum())
    # a flip that touches a feature list: anything but MESSY <-> MESSY changes lessFlat / lessSharp membership
    return dict(classified=int(len(idx)), flips=flips, by_label={int(k): int(((lab[idx] == k) & (got != k)).sum()) for k in (1, -1, 5, 9)})


# ---- scan-to-map with LAPACK ----------------------------------------------------------------------------------------------
def _pose_R(p):
    rx, ry, rz = (np.float64(p[0]), np.float64(p[1]), np.float64(p[2]))
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return (Rz @ Ry @ Rx).astype(F)          # transform_utils.h:288-299: Rz * Ry * Rx


def _find_line(nb):
    c = nb.sum(axis=1, dtype=F) / F(5)
    D = nb - c[:, None, :]
    A = np.einsum("qki,qkj->qij", D, D).astype(F) / F(5)
    w, V = np.linalg.eigh(A)
    ok = w[:, 2] > F(5) * w[:, 1]
    v = V[:, :, 2]
    return ok, c - v * F(0.1), c + v * F(0.1)


def _find_plane(nb, max_dist):
    ok = np.zeros(len(nb), bool)
    pl = np.zeros((len(nb), 4), F)
    b = np.full(5, -1, F)
    for i in range(len(nb)):
        x = np.linalg.lstsq(nb[i], b, rcond=None)[0].astype(F)      # LAPACK sgelsd in place of colPivHouseholderQr().solve
        n = np.sqrt((x * x).sum(dtype=F))
        x = x / n
        c = nb[i].sum(axis=0, dtype=F) / F(5)
        d = -(x * c).sum(dtype=F)
        dist = (nb[i] * x).sum(axis=1, dtype=F) + d
        ok[i] = not (np.abs(dist.astype(np.float64)) > max_dist).any()
        pl[i, :3] = x; pl[i, 3] = d
    return ok, pl


def scan_match_lapack(oracle, mc, ms, corner, surf, pose, max_iterations=10, align_signs=False):
    """ScanMatch::scanMatchScan (ScanMatch.cpp:51-260) with numpy.linalg / BLAS float32 in place of Eigen; neighbours from the
    reference's nanoflann.  Returns (pose, iterations, per-iteration decision log).
    align_signs: give LAPACK's eigenvectors of AtA the signs the restatement of Eigen's solver produces (the degeneracy projector
    matV.inverse() * matV2 of ScanMatch.cpp:223-234 zeroes ROWS of the column-eigenvector matrix, so it is NOT invariant under a
    sign flip of an eigenvector: V -> V S gives S P S)."""
    p = np.array(pose, F)
    log = []
    P = None
    degenerate = False
    it_done = 0
    for it in range(max_iterations):
        R = _pose_R(p); t = p[3:6]
        rows, rhs = [], []
        dec = {}
        for name, q, ref in (("corner", corner, mc), ("surf", surf, ms)):
            sel = (q[:, :3] @ R.T + t).astype(F)
            idx, d2 = oracle.knn(ref, sel, 5, nanoflann=True)
            gate = d2[:, 4] < F(5.0)
            nb = ref[idx[gate]][:, :, :3].astype(F)
            X = sel[gate]; ori = q[gate][:, :3]
            if name == "corner":
                ok, A, B = _find_line(nb)
                b_ = X - B; a_ = X - A
                k = np.cross(b_, a_).astype(F)
                kn = np.sqrt((k * k).sum(axis=1, dtype=F)); lab = np.sqrt(((A - B) ** 2).sum(axis=1, dtype=F))
                u = np.cross(k, B - A).astype(F)
                dirv = -u / (kn * lab)[:, None]
                dist = kn / lab
                w = (1 - F(0.9) * np.abs(dist.astype(np.float64))).astype(F)
            else:
                ok, pl = _find_plane(nb, 0.2)
                dist = (pl[:, :3] * X).sum(axis=1, dtype=F) + pl[:, 3]
                xn = np.sqrt((X * X).sum(axis=1, dtype=F))
                w = (1 - 0.9 * np.abs(dist.astype(np.float64)) / np.sqrt(xn.astype(np.float64))).astype(F)
                dirv = pl[:, :3]
            keep = ok & (w.astype(np.float64) > 0.1)
            full = np.zeros(len(q), np.int8); full[np.nonzero(gate)[0][ok]] = 1; full[np.nonzero(gate)[0][keep]] = 3
            dec[name] = full
            co = np.concatenate([dirv * w[:, None], (dist * w)[:, None]], axis=1)[keep].astype(F)
            o = ori[keep].astype(F)
            srx, crx, sry, cry, srz, crz = (F(np.sin(np.float64(p[0]))), F(np.cos(np.float64(p[0]))), F(np.sin(np.float64(p[1]))),
                                            F(np.cos(np.float64(p[1]))), F(np.sin(np.float64(p[2]))), F(np.cos(np.float64(p[2]))))
            x, y, z = o[:, 0], o[:, 1], o[:, 2]
            cx_, cy_, cz_ = co[:, 0], co[:, 1], co[:, 2]
            arx = ((crz * sry * crx + srz * srx) * y + (srz * crx - crz * sry * srx) * z) * cx_ + \
                  ((srz * sry * crx - crz * srx) * y - (srz * sry * srx + crz * crx) * z) * cy_ + (cry * crx * y - cry * srx * z) * cz_
            ary = (-crz * sry * x + crz * cry * srx * y + crz * cry * crx * z) * cx_ + \
                  (-srz * sry * x + srz * cry * srx * y + srz * cry * crx * z) * cy_ + (-cry * x - sry * srx * y - sry * crx * z) * cz_
            arz = (-srz * cry * x - (srz * sry * srx + crz * crx) * y + (crz * srx - srz * sry * crx) * z) * cx_ + \
                  (crz * cry * x + (crz * sry * srx - srz * crx) * y + crz * sry * crx + srz * srx * z) * cy_      # ScanMatch.cpp:193-195 as written
            rows.append(np.stack([arx, ary, arz, cx_, cy_, cz_], axis=1).astype(F)); rhs.append(-co[:, 3])
        A = np.concatenate(rows); b = np.concatenate(rhs)
        if len(A) < 50:
            break
        AtA = (A.T @ A).astype(F); AtB = (A.T @ b).astype(F)      # float32 BLAS, its own summation order
        x = np.linalg.solve(AtA, AtB).astype(F)
        if it == 0:
            w, V = np.linalg.eigh(AtA)                                # columns = eigenvectors, like esolver.eigenvectors()
            if align_signs:
                _, Vo = oracle.eig6(AtA)
                V = (V * np.sign((V * Vo).sum(axis=0))[None, :]).astype(F)
            V2 = V.copy(); degenerate = False
            for i in range(6):
                if w[i] < F(100):
                    V2[i] = 0; degenerate = True
                else:
                    break
            P = (np.linalg.inv(V) @ V2).astype(F)
        if degenerate:
            x = (P @ x).astype(F)
        p = (p + x).astype(F)
        it_done = it + 1
        log.append(dict(dec=dec, rows=len(A), degenerate=degenerate, x=x))
        dR = np.sqrt(((np.rad2deg(x[:3].astype(np.float64))) ** 2).sum()); dT = np.sqrt(((x[3:].astype(np.float64) * 100) ** 2).sum())
        if dR < 0.1 and dT < 0.1:
            break
    return p, it_done, log


def _oracle_flags(oracle, mc, ms, corner, surf, pose):
    """accepted / kept flags of the oracle's first evaluation, from its own neighbour log (find_line / find_plane through cm_math.h)."""
    p, st, log = oracle.scan_match(mc, ms, corner, surf, pose, keep_log=True)
    return p, st, log


def _match_gap(oracle, mc, ms, corner, surf, init):
    po, so, lo = _oracle_flags(oracle, mc, ms, corner, surf, init)
    pl, itl, ll = scan_match_lapack(oracle, mc, ms, corner, surf, init)
    pa, ita, la = scan_match_lapack(oracle, mc, ms, corner, surf, init, align_signs=True)
    out = dict(aligned_dpos_m=float(np.abs(pa[3:].astype(np.float64) - po[3:]).max()),
               aligned_drot_rad=float(np.abs(pa[:3].astype(np.float64) - po[:3]).max()), iterations_aligned=ita,
               iterations_oracle=so["iterations"], iterations_lapack=itl,
               rows_iter0_oracle=int(lo[0]["counts"][0]) if len(lo) else None, rows_iter0_lapack=ll[0]["rows"] if ll else None,
               degenerate_oracle=bool(so["degenerate"]), degenerate_lapack=bool(ll[0]["degenerate"]) if ll else None,
               dpos_m=float(np.abs(pl[3:].astype(np.float64) - po[3:]).max()), drot_rad=float(np.abs(pl[:3].astype(np.float64) - po[:3]).max()),
               dx_iter0=float(np.abs(ll[0]["x"].astype(np.float64) - lo[0]["x"]).max()) if ll and lo else None)
    return out


def test_pointclassify_labels_with_lapack_eigh(oracle, synth, capsys):
    g = np.load(os.path.join(GOLD, "scanreg_vlp16_600.npz"))
    res = {"golden_vlp16_600": _classify_flips(oracle, g["frame"])}
    sc = synth.make_scene(seed=0x5EED0002 & 0xFFFF, extent=60.0, n_boxes=30, n_poles=24)
    R, t = synth.pose_matrix(0.02, 0.0, 0.0, (1.0, 0.5, 0.0))
    res["hdl64_2048"] = _classify_flips(oracle, synth.simulate_scan(sc, R, t, "HDL-64E", seed=5))
    with capsys.disabled():
        print("\n[parity-gap] pointClassify, oracle (cm_math.h eig3_sym) vs numpy.linalg.eigh float32:", json.dumps(res))
    for k, r in res.items():
        assert r["classified"] > 100, (k, r)
        assert r["flips"] <= max(2, r["classified"] // 200), (k, r)     # <= 0.5 % of the classified points


def test_scan_match_with_lapack(oracle, synth, capsys):
    g = np.load(os.path.join(GOLD, "match_small.npz"))
    res = {"golden_match_small": _match_gap(oracle, g["mc"], g["ms"], g["corner"], g["surf"], g["init"])}
    sc = synth.make_scene(seed=77, extent=40.0, n_boxes=14, n_poles=12)
    mc, ms = synth.sample_map(sc, 0.4, seed=78)
    for k, (Rk, tk) in enumerate(synth.trajectory(3, speed=0.8)):
        fr = synth.simulate_scan(sc, Rk, tk, "VLP-16", seed=90 + k, cols=900)
        f = oracle.scanreg_organised(fr)
        corner = oracle.voxel_filter(f["lessSharp"], 0.4); surf = oracle.voxel_filter(f["lessFlat"], 0.8)
        Rf, tf = Rk.astype(F), tk.astype(F)
        tw = oracle.iso_to_twist(Rf, tf)
        init = (np.asarray(tw, F) + np.array([0.004, -0.003, 0.005, 0.05, -0.04, 0.03], F)).astype(F)
        res["frame%d" % k] = _match_gap(oracle, mc, ms, corner, surf, init)
    with capsys.disabled():
        print("\n[parity-gap] scanMatchScan, oracle (cm_math.h) vs numpy.linalg / BLAS float32:", json.dumps(res))
    for k, r in res.items():
        assert r["degenerate_oracle"] == r["degenerate_lapack"], (k, r)
        assert abs(r["rows_iter0_oracle"] - r["rows_iter0_lapack"]) <= max(2, r["rows_iter0_oracle"] // 200), (k, r)
        # north-star tolerance between two float32 implementations of the same loop.  Degenerate scenes are the exception: the
        # reference's projector depends on the SIGN of each eigenvector (see scan_match_lapack), which is a property of the solver's
        # algorithm, not of the mathematics -- with LAPACK's signs the pose differs by centimetres, with the signs of the restated
        # Eigen algorithm it agrees again.
        assert r["aligned_dpos_m"] <= 1e-4 and r["aligned_drot_rad"] <= 1e-5, (k, r)
        if not r["degenerate_oracle"]:
            assert r["dpos_m"] <= 1e-4 and r["drot_rad"] <= 1e-5, (k, r)


def test_degenerate_projector_depends_on_eigenvector_signs(oracle, capsys, monkeypatch):
    """The one place where "which eigen-solver" matters beyond rounding: matP = matV.inverse() * matV2 (ScanMatch.cpp:223-234) zeroes
    ROWS of the column-eigenvector matrix, so flipping the sign of an eigenvector (any solver may) changes the projected step.
    Measured on the degenerate golden case: two flipped signs move the final pose by centimetres.  The oracle / GPU follow Eigen
    3.3's tridiagonal-QR algorithm step by step (cm_math.h eig_sym), and LAPACK's ssyevd happens to return the same signs here."""
    g = np.load(os.path.join(GOLD, "match_small.npz"))
    po, so, _ = oracle.scan_match(g["mc"], g["ms"], g["corner"], g["surf"], g["init"])
    assert so["degenerate"]
    orig = np.linalg.eigh

    def flipped(A):
        w, V = orig(A)
        if A.shape == (6, 6):
            V = V.copy(); V[:, 5] *= -1; V[:, 2] *= -1
        return w, V
    monkeypatch.setattr(np.linalg, "eigh", flipped)
    pl, _, _ = scan_match_lapack(oracle, g["mc"], g["ms"], g["corner"], g["surf"], g["init"])
    d = float(np.abs(pl[3:].astype(np.float64) - po[3:]).max())
    with capsys.disabled():
        print("\n[parity-gap] degenerate case, two eigenvector signs flipped: final position moves by %.4f m" % d)
    assert d > 1e-3


# ---- the canonical trigonometry (cm_sincosf / cm_atanf / cm_atan2f) against the reference's own libm calls -----------------------------
def test_canonical_trig_vs_libm_at_the_pose_level():
    """DESIGN.md section 2, canonical choices 5 and 7: Angle's sin / cos (util/Angle.h:19-20) and the front end's atan / atan2
    (MultiScanRegistration.cpp:103-156) are libm FLOAT calls in the reference; GPU and oracle use correctly rounded cm_* definitions
    instead (glibc differs from them by 1 ulp on 1-15 % of the inputs).  liboracle_libm.so is the same oracle built with libm: this
    runs the whole chain raw sweep -> features -> odometry -> mapping with both and reports what the choice changes -- ring / list
    membership of the front end and the poses of six sweeps."""
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_py as O
    synth = importlib.import_module("the-cooper-mapper_b200.synth")
    sc = synth.make_scene(seed=0x5EED0001 & 0xFFFF, extent=60.0, n_boxes=24, n_poles=20)
    mp = dict(filterCorner=0.4, filterSurf=0.8, mapFilterCorner=0.4, mapFilterSurf=0.4)
    chains = {v: (O.Odometry(fast=v), O.Mapping(map_params=mp, fast=v)) for v in (False, "libm")}
    ring_flips = reltime_diffs = npts = 0
    list_diff = 0
    dt = dr = 0.0
    its = []
    for k, (R, t) in enumerate(synth.trajectory(6, speed=0.1, yaw_amp=0.02)):
        fr = synth.simulate_scan(sc, R, t, "VLP-16", seed=0x6000 + k, cols=1200)
        sweep = synth.organised_to_sweep(fr)
        ok = np.where(np.isfinite(sweep[:, 0]))[0]
        sweep = sweep[ok[0]:ok[-1] + 1]
        out = {}
        for v, (od, mapping) in chains.items():
            f = O.scanreg_sweep(sweep, 0, fast=v)
            o = od.process(f["sharp"], f["lessSharp"], f["flat"], f["lessFlat"])
            mR, mt, st = mapping.process(o["R"], o["t"], o["corner_last"], o["surf_last"])
            out[v] = (f, mR, mt, st)
        fa, fb = out[False][0], out["libm"][0]
        assert len(fa["cloud"]) == len(fb["cloud"])                      # same points accepted
        ca, cb = fa["cloud"][:, 4], fb["cloud"][:, 4]                    # the curvature field = ring + relTime
        npts += len(ca)
        ring_flips += int((np.floor(ca) != np.floor(cb)).sum())
        reltime_diffs += int((ca != cb).sum())
        list_diff += sum(abs(len(fa[n]) - len(fb[n])) for n in ("sharp", "lessSharp", "flat", "lessFlat"))
        dt = max(dt, float(np.abs(out[False][2] - out["libm"][2]).max()))
        dr = max(dr, float(np.abs(out[False][1] - out["libm"][1]).max()))
        its.append((out[False][3]["iterations"], out["libm"][3]["iterations"]))
    print("[parity-gap] canonical vs libm trig over 6 raw VLP-16 sweeps: %d ring flips and %d relTime last-bit differences in %d points, "
          "feature-list size differences %d, max |dt| %.2e m, max |dR| %.2e, iterations %s" % (ring_flips, reltime_diffs, npts, list_diff, dt, dr, its))
    assert ring_flips == 0 and 0 < reltime_diffs < 0.05 * npts          # the variant does differ, in the last bit of a per cent of the tags
    assert dt <= 1e-4 and dr <= 1e-5                                    # the north-star tolerance
