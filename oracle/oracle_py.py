"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/cm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}
REF_SO = os.path.join(_HERE, "_ref", "libcm_ref_nanoflann.so")


def build(force=False):
    """Compile liboracle.so / liboracle_fast.so (and oracle/_ref when /root/reference is present)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so", "liboracle_libm.so"))
    if need or (os.path.isdir("/root/reference") and not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def lib(fast=False):
    # fast="libm": the variant built with the reference's libm trigonometry (a measuring instrument, not a parity target)
    name = "liboracle_libm.so" if fast == "libm" else ("liboracle_fast.so" if fast else "liboracle.so")
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.cmo_scanreg_organised.restype = C.c_void_p
    L.cmo_scanreg_sweep.restype = C.c_void_p
    L.cmo_scanreg_size.restype = C.c_size_t
    L.cmo_voxel_filter.restype = C.c_size_t
    L.cmo_scan_match.restype = C.c_void_p
    L.cmo_mapping_create.restype = C.c_void_p
    L.cmo_odom_create.restype = C.c_void_p
    L.cmo_odom_indices.restype = C.c_size_t
    L.cmo_mapping_cloud.restype = C.c_size_t
    L.cmo_mapping_file_order.restype = C.c_size_t
    L.cmo_mapping_cube.restype = C.c_size_t
    L.cmo_mapping_map_surround.restype = C.c_size_t
    L.cmo_mapping_local_create.restype = C.c_void_p
    L.cmo_mapping_local_window.restype = C.c_size_t
    L.has_nanoflann = bool(os.path.exists(REF_SO) and L.cmo_load_nanoflann(REF_SO.encode()))
    _libs[name] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- math hooks -------------------------------------------------------------------------------------------
def sincosf(x):
    x = _f32(x).ravel()
    s = np.empty_like(x); c = np.empty_like(x)
    lib().cmo_sincosf(_p(x), C.c_int(x.size), _p(s), _p(c))
    return s, c


def eig3(A):
    """A: 3x3 symmetric -> (w ascending, V columns)."""
    A = _f32(A)
    a6 = _f32([A[0, 0], A[1, 0], A[2, 0], A[1, 1], A[2, 1], A[2, 2]])
    w = np.empty(3, np.float32); V = np.empty((3, 3), np.float32)
    lib().cmo_eig3(_p(a6), _p(w), _p(V))
    return w, V


def eig6(A):
    A = _f32(A)
    w = np.empty(6, np.float32); V = np.empty((6, 6), np.float32)
    lib().cmo_eig6(_p(A), _p(w), _p(V))
    return w, V


def qr_solve(A, b):
    A = _f32(A); b = _f32(b)
    x = np.empty(A.shape[1], np.float32)
    if A.shape == (5, 3):
        lib().cmo_qr_solve_5x3(_p(A), _p(b), _p(x))
    elif A.shape == (6, 6):
        lib().cmo_qr_solve_6x6(_p(A), _p(b), _p(x))
    else:
        raise ValueError(A.shape)
    return x


def inverse6(A):
    A = _f32(A); inv = np.empty((6, 6), np.float32)
    ok = lib().cmo_inverse6(_p(A), _p(inv))
    return inv if ok else None


def pose_to_matrix(pose):
    pose = _f32(pose); R = np.empty((3, 3), np.float32)
    lib().cmo_pose_to_matrix(_p(pose), _p(R))
    return R


def iso_to_twist(R, t):
    R = _f32(R); t = _f32(t); pose = np.empty(6, np.float32)
    lib().cmo_iso_to_twist(_p(R), _p(t), _p(pose))
    return pose


MATH_DIMS = {0: (42, 6), 1: (20, 3), 2: (6, 12), 3: (36, 42), 4: (36, 6), 5: (36, 36), 6: (6, 15), 7: (2, 2)}


def debug_math(op, inputs):
    """Batch form of the hooks above, same op codes as cm_debug_math_host."""
    nin, nout = MATH_DIMS[op]
    a = _f32(inputs).reshape(-1, nin)
    out = np.empty((len(a), nout), np.float32)
    lib().cmo_debug_math(C.c_int(op), _p(a), C.c_size_t(len(a)), _p(out))
    return out


# ---- scan registration ------------------------------------------------------------------------------------
_SR_FIELDS = [("cloud", np.float32, 5), ("scanStart", np.int32, 1), ("scanEnd", np.int32, 1), ("sharp", np.float32, 4),
              ("lessSharp", np.float32, 4), ("flat", np.float32, 4), ("lessFlat", np.float32, 4),
              ("sharpIdx", np.int32, 1), ("lessSharpIdx", np.int32, 1), ("flatIdx", np.int32, 1),
              ("lessFlatRawIdx", np.int32, 1), ("lessFlatRawRing", np.int32, 1), ("picked", np.int32, 1),
              ("curvature", np.float32, 1), ("classLabel", np.int32, 1), ("dbgBlind", np.float32, 4),
              ("dbgBlock", np.float32, 4), ("dbgSlop", np.float32, 4), ("dbgCurv", np.float32, 4)]


def _sr_params(params):
    d = dict(scanPeriod=0.1, lessFlatFilterSize=0.2, surfaceCurvatureThreshold=0.02, blindDegreeThreshold=0.5,
             blindRadius=2.5, nFeatureRegions=6, curvatureRegion=5, maxCornerSharp=2, maxSurfaceFlat=4)
    d.update(params or {})
    f = _f32([d["scanPeriod"], d["lessFlatFilterSize"], d["surfaceCurvatureThreshold"], d["blindDegreeThreshold"],
              d["blindRadius"]])
    i = np.array([d["nFeatureRegions"], d["curvatureRegion"], d["maxCornerSharp"], d["maxSurfaceFlat"]], np.int32)
    return f, i


def _sr_collect(L, h):
    out = {}
    for fid, (name, dt, w) in enumerate(_SR_FIELDS):
        n = L.cmo_scanreg_size(C.c_void_p(h), C.c_int(fid))
        a = np.empty((n, w) if w > 1 else (n,), dt)
        if n:
            L.cmo_scanreg_copy(C.c_void_p(h), C.c_int(fid), _p(a))
        out[name] = a
    L.cmo_scanreg_free(C.c_void_p(h))
    return out


def scanreg_organised(xyzi, params=None, fast=False):
    """xyzi: (rows, cols, 4) float32 organised cloud -> dict of oracle outputs."""
    L = lib(fast)
    xyzi = _f32(xyzi)
    rows, cols = xyzi.shape[0], xyzi.shape[1]
    f, i = _sr_params(params)
    h = L.cmo_scanreg_organised(_p(f), _p(i), _p(xyzi), C.c_int(rows), C.c_int(cols))
    return _sr_collect(L, h)


def scanreg_sweep(xyzi, lidar, params=None, fast=False):
    L = lib(fast)
    xyzi = _f32(xyzi)
    f, i = _sr_params(params)
    h = L.cmo_scanreg_sweep(_p(f), _p(i), _p(xyzi), C.c_int(xyzi.shape[0]), C.c_int(lidar))
    return _sr_collect(L, h)


def scanreg_sweep_imu(xyzi, lidar, scan_time, imu, params=None):
    """MultiScanRegistration::process with IMU data: imu = (k, 7) float64 rows (stamp, roll, pitch, yaw, ax, ay, az) in arrival order.
    Returns (outputs, imu_trans (4, 3))."""
    L = lib()
    L.cmo_scanreg_sweep_imu.restype = C.c_void_p
    xyzi = _f32(xyzi)
    imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
    f, i = _sr_params(params)
    tr = np.zeros(12, np.float32)
    h = L.cmo_scanreg_sweep_imu(_p(f), _p(i), _p(xyzi), C.c_int(xyzi.shape[0]), C.c_int(lidar), C.c_double(scan_time), _p(imu),
                                C.c_int(len(imu)), _p(tr))
    return _sr_collect(L, h), tr.reshape(4, 3)


# ---- voxel filter -----------------------------------------------------------------------------------------
def voxel_filter(xyzi, leaf, fast=False):
    L = lib(fast)
    xyzi = _f32(xyzi).reshape(-1, 4)
    out = np.empty((max(len(xyzi), 1), 4), np.float32)
    n = L.cmo_voxel_filter(_p(xyzi), C.c_size_t(len(xyzi)), C.c_float(leaf), _p(out), C.c_size_t(len(out)))
    return out[:n].copy()


# ---- KNN ---------------------------------------------------------------------------------------------------
def knn(pts, q, k=5, nanoflann=False, fast=False):
    L = lib(fast)
    pts = _f32(pts).reshape(-1, 4); q = _f32(q).reshape(-1, 3)
    idx = np.empty((len(q), k), np.int32); d2 = np.empty((len(q), k), np.float32)
    L.cmo_knn(C.c_int(int(nanoflann)), _p(pts), C.c_size_t(len(pts)), _p(q), C.c_size_t(len(q)), C.c_int(k), _p(idx), _p(d2))
    return idx, d2


# ---- scan-to-map solver -----------------------------------------------------------------------------------
def _match_params(params):
    d = dict(deltaTAbort=0.1, deltaRAbort=0.1, knnGate=5.0, planeMaxDistance=0.2, maxIterations=10, useScore=0)
    d.update(params or {})
    f = _f32([d["deltaTAbort"], d["deltaRAbort"], d["knnGate"], d["planeMaxDistance"]])
    i = np.array([d["maxIterations"], int(d["useScore"])], np.int32)
    return f, i


def scan_match(ref_corner, ref_surf, corner, surf, pose, params=None, nanoflann=True, keep_log=False, fast=False):
    """ScanMatch::scanMatchScan(Twist).  pose = (rx, ry, rz, tx, ty, tz).  Returns (pose_out, stats, log)."""
    L = lib(fast)
    rc = _f32(ref_corner).reshape(-1, 4); rs = _f32(ref_surf).reshape(-1, 4)
    c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
    p = _f32(pose).copy()
    f, i = _match_params(params)
    use_nf = int(nanoflann and L.has_nanoflann)
    h = L.cmo_scan_match(_p(f), _p(i), C.c_int(use_nf), _p(rc), C.c_size_t(len(rc)), _p(rs), C.c_size_t(len(rs)),
                         _p(c), C.c_size_t(len(c)), _p(s), C.c_size_t(len(s)), _p(p), C.c_int(int(keep_log)))
    st = np.zeros(10, np.int32); score = C.c_double(0)
    L.cmo_match_stats(C.c_void_p(h), _p(st), C.byref(score))
    stats = dict(ok=bool(st[0]), converged=bool(st[1]), tooFewRef=bool(st[2]), tooFewMatches=bool(st[3]),
                 degenerate=bool(st[4]), iterations=int(st[5]), rows=int(st[6]), line=int(st[7]), plane=int(st[8]),
                 score=score.value)
    log = []
    for it in range(int(st[9])):
        e = dict(pose_in=np.empty(6, np.float32), AtA=np.empty((6, 6), np.float32), AtB=np.empty(6, np.float32),
                 x=np.empty(6, np.float32), counts=np.empty(4, np.int32), nnCorner=np.empty((len(c), 5), np.int32),
                 nnSurf=np.empty((len(s), 5), np.int32))
        L.cmo_match_log(C.c_void_p(h), C.c_int(it), _p(e["pose_in"]), _p(e["AtA"]), _p(e["AtB"]), _p(e["x"]),
                        _p(e["counts"]), _p(e["nnCorner"]), _p(e["nnSurf"]))
        log.append(e)
    L.cmo_match_free(C.c_void_p(h))
    return p, stats, log


def scan_match_local(ref_corner, ref_surf, corner, surf, pose, params=None, nanoflann=True):
    """ScanMatch::scanMatchLocal (ScanMatch.cpp:375-398): voxel-filter the four clouds (corner 0.2, surf 0.4, :29-30), then
    scanMatchScan with the class defaults (use score, abort thresholds 0.05, ScanMatch.cpp:22-24)."""
    d = dict(deltaTAbort=0.05, deltaRAbort=0.05, useScore=1)
    d.update(params or {})
    return scan_match(voxel_filter(ref_corner, 0.2), voxel_filter(ref_surf, 0.4), voxel_filter(corner, 0.2), voxel_filter(surf, 0.4),
                      pose, params=d, nanoflann=nanoflann)


# ---- mapping loop -----------------------------------------------------------------------------------------
def write_pcd_binary(path, pts):
    """pcl::io::savePCDFileBinary of a pcl::PointCloud<pcl::PointXYZI>: v0.7 header, 16 packed bytes per point."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (len(pts), len(pts)))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii")); f.write(pts.tobytes())


def read_pcd(path):
    """x, y, z, intensity of a binary / ascii PCD file with float32 fields."""
    raw = open(path, "rb").read()
    pos = 0; hdr = {}
    while True:
        end = raw.index(b"\n", pos); line = raw[pos:end].decode("ascii").strip(); pos = end + 1
        if line.startswith("#") or not line:
            continue
        k, *v = line.split(); hdr[k] = v
        if k == "DATA":
            break
    fields = hdr["FIELDS"]; size = [int(x) for x in hdr["SIZE"]]; cnt = [int(x) for x in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0])
    if hdr["DATA"][0] == "binary":
        dt = np.dtype([(f if f != "_" else "_%d" % i, "<f4" if s == 4 else "V%d" % (s * c), ()) if (s == 4 and c == 1) else ("pad%d" % i, "V%d" % (s * c))
                       for i, (f, s, c) in enumerate(zip(fields, size, cnt))])
        a = np.frombuffer(raw, dt, count=n, offset=pos)
        out = np.zeros((n, 4), np.float32)
        for col, name in enumerate(["x", "y", "z", "intensity"]):
            if name in a.dtype.names:
                out[:, col] = a[name]
        return out
    rows = np.array(raw[pos:].split(), np.float64).reshape(n, -1)
    out = np.zeros((n, 4), np.float32)
    for col, name in enumerate(["x", "y", "z", "intensity"]):
        if name in fields:
            out[:, col] = rows[:, fields.index(name)]
    return out


class Mapping:
    """LaserMapping::process restated (oracle_map.cpp)."""

    def __init__(self, map_params=None, match_params=None, nanoflann=True, fast=False):
        self.L = lib(fast)
        d = dict(cubeSize=50.0, validDistance=150.0, mapFilterCorner=1.0, mapFilterSurf=1.0, filterCorner=1.0,
                 filterSurf=1.0, cubeW=121, cubeH=121, cubeD=11)
        d.update(map_params or {})
        mf = _f32([d["cubeSize"], d["validDistance"], d["mapFilterCorner"], d["mapFilterSurf"], d["filterCorner"],
                   d["filterSurf"]])
        mi = np.array([d["cubeW"], d["cubeH"], d["cubeD"]], np.int32)
        sf, si = _match_params(match_params)
        self.dims = (d["cubeW"], d["cubeH"], d["cubeD"])
        self.h = self.L.cmo_mapping_create(_p(mf), _p(mi), _p(sf), _p(si), C.c_int(int(nanoflann and self.L.has_nanoflann)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cmo_mapping_free(C.c_void_p(self.h)); self.h = None

    def process(self, odom_R, odom_t, corner, surf):
        R = _f32(odom_R); t = _f32(odom_t)
        c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        oR = np.empty((3, 3), np.float32); ot = np.empty(3, np.float32); st = np.zeros(13, np.int32)
        self.L.cmo_mapping_process(C.c_void_p(self.h), _p(R), _p(t), _p(c), C.c_size_t(len(c)), _p(s), C.c_size_t(len(s)),
                                   _p(oR), _p(ot), _p(st))
        keys = ["ok", "converged", "tooFewRef", "tooFewMatches", "degenerate", "iterations", "rows", "line", "plane",
                "nCornerDS", "nSurfDS", "nSurroundCorner", "nSurroundSurf"]
        return oR, ot, dict(zip(keys, [int(v) for v in st]))

    def localize(self, odom_R, odom_t, corner, surf):
        """LaserLocalization::process: FeatureMap::scanMatchScan against the map, no map update."""
        R = _f32(odom_R); t = _f32(odom_t)
        c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        oR = np.empty((3, 3), np.float32); ot = np.empty(3, np.float32); st = np.zeros(13, np.int32)
        self.L.cmo_mapping_localize(C.c_void_p(self.h), _p(R), _p(t), _p(c), C.c_size_t(len(c)), _p(s), C.c_size_t(len(s)),
                                    _p(oR), _p(ot), _p(st))
        keys = ["ok", "converged", "tooFewRef", "tooFewMatches", "degenerate", "iterations", "rows", "line", "plane",
                "nCornerDS", "nSurfDS", "nSurroundCorner", "nSurroundSurf"]
        return oR, ot, dict(zip(keys, [int(v) for v in st]))

    # ---- FeatureMap::saveCloudToFiles / loadCloudFromFiles (FeatureMap.h:378-462), file format restated in numpy ----
    def save_files(self, directory):
        """index.txt + <count>.pcd (pcl::io::savePCDFileBinary layout for PointXYZI); returns the number of files."""
        n = self.L.cmo_mapping_file_order(C.c_void_p(self.h), None, None, None, None, C.c_size_t(0))
        ty = np.zeros(max(n, 1), np.int32); ci = np.zeros_like(ty); cj = np.zeros_like(ty); ck = np.zeros_like(ty)
        self.L.cmo_mapping_file_order(C.c_void_p(self.h), _p(ty), _p(ci), _p(cj), _p(ck), C.c_size_t(len(ty)))
        dims = np.array(self.dims, np.int32)
        with open(os.path.join(directory, "index.txt"), "w") as idx:
            for count in range(n):
                m = self.L.cmo_mapping_cube(C.c_void_p(self.h), C.c_int(int(ty[count])), C.c_int(int(ci[count])), C.c_int(int(cj[count])),
                                            C.c_int(int(ck[count])), _p(dims), None, C.c_size_t(0))
                pts = np.empty((m, 4), np.float32)
                self.L.cmo_mapping_cube(C.c_void_p(self.h), C.c_int(int(ty[count])), C.c_int(int(ci[count])), C.c_int(int(cj[count])),
                                        C.c_int(int(ck[count])), _p(dims), _p(pts), C.c_size_t(m))
                write_pcd_binary(os.path.join(directory, "%d.pcd" % count), pts)
                idx.write("%d %d %d %d %d %d\n" % (count, ty[count], ci[count], cj[count], ck[count], m))
        return n

    def load_files(self, directory):
        n = 0
        for line in open(os.path.join(directory, "index.txt")):
            f = line.split()
            if len(f) != 6:
                continue
            count, ty, i, j, k, _size = [int(v) for v in f]
            pts = read_pcd(os.path.join(directory, "%d.pcd" % count))
            self.L.cmo_mapping_load_cube(C.c_void_p(self.h), C.c_int(ty), C.c_int(i), C.c_int(j), C.c_int(k), _p(pts), C.c_size_t(len(pts)))
            n += 1
        return n

    def cloud(self, which):
        n = self.L.cmo_mapping_cloud(C.c_void_p(self.h), C.c_int(which), None, C.c_size_t(0))
        out = np.empty((max(n, 1), 4), np.float32)
        self.L.cmo_mapping_cloud(C.c_void_p(self.h), C.c_int(which), _p(out), C.c_size_t(len(out)))
        return out[:n].copy()

    def origin(self):
        o = np.zeros(3, np.int32)
        self.L.cmo_mapping_origin(C.c_void_p(self.h), _p(o))
        return o.tolist()

    def map_update(self, sensor):
        s = _f32(sensor)
        self.L.cmo_mapping_map_update(C.c_void_p(self.h), _p(s))

    def map_add(self, corner, surf, R, t):
        c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        self.L.cmo_mapping_map_add(C.c_void_p(self.h), _p(c), C.c_size_t(len(c)), _p(s), C.c_size_t(len(s)), _p(_f32(R)), _p(_f32(t)))

    def map_surround(self, which):
        n = self.L.cmo_mapping_map_surround(C.c_void_p(self.h), C.c_int(which), None, C.c_size_t(0))
        out = np.empty((max(n, 1), 4), np.float32)
        self.L.cmo_mapping_map_surround(C.c_void_p(self.h), C.c_int(which), _p(out), C.c_size_t(len(out)))
        return out[:n].copy()


class MappingLocal:
    """LaserMappingLocal::process over LocalFeatureMap restated (oracle_map.cpp): sliding window of voxel-filtered frames.
    use_mapped_pose=False is the literal reference (frames placed with the never-assigned _transformTobeMapped = identity)."""

    def __init__(self, map_params=None, match_params=None, nanoflann=True, use_mapped_pose=False, fast=False):
        self.L = lib(fast)
        d = dict(filterCorner=1.0, filterSurf=1.0)
        d.update(map_params or {})
        mf = _f32([50.0, 150.0, 1.0, 1.0, d["filterCorner"], d["filterSurf"]])
        sf, si = _match_params(match_params)
        self.h = self.L.cmo_mapping_local_create(_p(mf), _p(sf), _p(si), C.c_int(int(nanoflann and self.L.has_nanoflann)),
                                                 C.c_int(int(use_mapped_pose)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cmo_mapping_local_free(C.c_void_p(self.h)); self.h = None

    def process(self, odom_R, odom_t, corner, surf):
        R = _f32(odom_R); t = _f32(odom_t)
        c = _f32(corner).reshape(-1, 4); s = _f32(surf).reshape(-1, 4)
        oR = np.empty((3, 3), np.float32); ot = np.empty(3, np.float32); st = np.zeros(16, np.int32); acc = C.c_double(0.0)
        self.L.cmo_mapping_local_process(C.c_void_p(self.h), _p(R), _p(t), _p(c), C.c_size_t(len(c)), _p(s), C.c_size_t(len(s)),
                                         _p(oR), _p(ot), _p(st), C.byref(acc))
        keys = ["ok", "converged", "tooFewRef", "tooFewMatches", "degenerate", "iterations", "rows", "line", "plane",
                "nCornerDS", "nSurfDS", "nSurroundCorner", "nSurroundSurf", "frames", "nWindowCorner", "nWindowSurf"]
        d = dict(zip(keys, [int(v) for v in st])); d["accumDistance"] = float(acc.value)
        return oR, ot, d

    def window(self, which):
        n = self.L.cmo_mapping_local_window(C.c_void_p(self.h), C.c_int(which), None, C.c_size_t(0))
        out = np.empty((max(n, 1), 4), np.float32)
        self.L.cmo_mapping_local_window(C.c_void_p(self.h), C.c_int(which), _p(out), C.c_size_t(len(out)))
        return out[:n].copy()


# ---- scan-to-scan odometry -----------------------------------------------------------------------------------
class Odometry:
    """LaserOdometry::process restated (oracle_odom.cpp)."""

    def __init__(self, nanoflann=True, fast=False):
        self.L = lib(fast)
        self.h = self.L.cmo_odom_create(C.c_int(int(nanoflann and self.L.has_nanoflann)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cmo_odom_free(C.c_void_p(self.h)); self.h = None

    def process(self, sharp, less_sharp, flat, less_flat):
        a = [_f32(x).reshape(-1, 4) for x in (sharp, less_sharp, flat, less_flat)]
        tf = np.empty(6, np.float32); R = np.empty((3, 3), np.float32); t = np.empty(3, np.float32); cnt = np.zeros(5, np.int32)
        self.L.cmo_odom_process(C.c_void_p(self.h), _p(a[0]), C.c_size_t(len(a[0])), _p(a[1]), C.c_size_t(len(a[1])), _p(a[2]),
                                C.c_size_t(len(a[2])), _p(a[3]), C.c_size_t(len(a[3])), _p(tf), _p(R), _p(t), _p(cnt))
        corner = np.empty((max(int(cnt[2]), 1), 4), np.float32); surf = np.empty((max(int(cnt[3]), 1), 4), np.float32)
        self.L.cmo_odom_last_clouds(C.c_void_p(self.h), _p(corner), _p(surf))
        log = []
        for it in range(int(cnt[4])):
            pin = np.empty(6, np.float32); x = np.empty(6, np.float32); rows = C.c_int(0)
            self.L.cmo_odom_log(C.c_void_p(self.h), C.c_int(it), _p(pin), _p(x), C.byref(rows))
            log.append(dict(pose_in=pin, x=x, rows=rows.value))
        n = self.L.cmo_odom_indices(C.c_void_p(self.h), None, C.c_size_t(0))
        ind = np.empty(max(n, 1), np.int32)
        self.L.cmo_odom_indices(C.c_void_p(self.h), _p(ind), C.c_size_t(len(ind)))
        return dict(transform=tf, R=R, t=t, iterations=int(cnt[0]), rows=int(cnt[1]), corner_last=corner[:cnt[2]].copy(),
                    surf_last=surf[:cnt[3]].copy(), log=log, ind=ind[:n].copy())
