"""ctypes binding of include/coopermap.h and host-side mirrors of the reference's operator classes.

Mirrors (same names, argument meaning and error behaviour as the reference):
  ScanMatch.scanMatchScan   <- lidar_slam::ScanMatch::scanMatchScan   (ScanMatch.h:47-55)
"""
import ctypes as C
import os

import numpy as np

CM_OK, CM_TOO_FEW_REF, CM_TOO_FEW_MATCHES, CM_NOT_CONVERGED, CM_LOW_SCORE = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))


class CoopermapError(RuntimeError):
    pass


def lib_path():
    # COOPERMAP_LIB: development override (kernel variants built side by side); the product library is in-tree
    return os.environ.get("COOPERMAP_LIB") or os.path.join(_HERE, "libcoopermap.so")


class Config(C.Structure):
    """cm_config (include/coopermap.h)."""
    _fields_ = [("device", C.c_int), ("scan_period", C.c_float), ("n_feature_regions", C.c_int),
                ("curvature_region", C.c_int), ("max_corner_sharp", C.c_int), ("max_surface_flat", C.c_int),
                ("less_flat_filter_size", C.c_float), ("surface_curvature_threshold", C.c_float),
                ("blind_degree_threshold", C.c_float), ("blind_radius", C.c_float), ("max_iterations", C.c_int),
                ("delta_t_abort", C.c_float), ("delta_r_abort", C.c_float), ("use_score", C.c_int),
                ("score_threshold", C.c_double), ("match_percentage_threshold", C.c_float),
                ("filter_corner", C.c_float), ("filter_surf", C.c_float), ("map_filter_corner", C.c_float),
                ("map_filter_surf", C.c_float), ("cube_w", C.c_int), ("cube_h", C.c_int), ("cube_d", C.c_int),
                ("cube_size", C.c_float), ("valid_distance", C.c_float), ("cell_corner", C.c_float),
                ("cell_surf", C.c_float), ("gn_groups", C.c_int)]


class MatchStats(C.Structure):
    _fields_ = [("status", C.c_int), ("ret", C.c_int), ("converged", C.c_int), ("degenerate", C.c_int),
                ("iterations", C.c_int), ("rows", C.c_int), ("line_matches", C.c_int), ("plane_matches", C.c_int),
                ("score", C.c_double)]


class ScanRegOut(C.Structure):
    _fields_ = [("pts", C.c_void_p * 4), ("cap", C.c_int * 4), ("n", C.c_void_p), ("cloud", C.c_void_p),
                ("cloud_curvature", C.c_void_p), ("scan_ranges", C.c_void_p), ("idx", C.c_void_p * 4),
                ("picked", C.c_void_p), ("curvature", C.c_void_p), ("label", C.c_void_p)]


class OdomStats(C.Structure):
    _fields_ = [("initialising", C.c_int), ("matched", C.c_int), ("iterations", C.c_int), ("rows", C.c_int),
                ("converged", C.c_int), ("degenerate", C.c_int)]


class IterTrace(C.Structure):
    _fields_ = [("pose_in", C.c_float * 6), ("AtA", C.c_float * 36), ("AtB", C.c_float * 6), ("x", C.c_float * 6),
                ("rows", C.c_int), ("line_matches", C.c_int), ("plane_matches", C.c_int), ("degenerate", C.c_int)]


_lib = None


def load_library():
    """Load libcoopermap.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise CoopermapError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    L.cm_last_error.restype = C.c_char_p
    L.cm_launch_count.restype = C.c_ulonglong
    _lib = L
    return L


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """cm_ctx: one context <-> one CUDA stream <-> one host thread."""

    def __init__(self, **kw):
        L = load_library()
        self.L = L
        self.cfg = Config()
        L.cm_config_default(C.byref(self.cfg))
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise CoopermapError("unknown config field %s" % k)
            setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        rc = L.cm_ctx_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise CoopermapError("cm_ctx_create failed (%d): no usable CUDA device; there is no CPU fallback" % rc)

    def close(self):
        if getattr(self, "h", None):
            self.L.cm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise CoopermapError("coopermap error %d: %s" % (rc, self.L.cm_last_error(self.h).decode()))
        return rc

    def launch_count(self):
        return int(self.L.cm_launch_count(self.h))

    # ---- device self-test of the shared math ------------------------------------------------------------------
    MATH_DIMS = {0: (42, 6), 1: (20, 3), 2: (6, 12), 3: (36, 42), 4: (36, 6), 5: (36, 36), 6: (6, 15), 7: (2, 2), 8: (42, 6), 9: (37, 1)}

    def debug_math(self, op, inputs):
        nin, nout = self.MATH_DIMS[op]
        a = _f32(inputs, nin)
        out = np.empty((len(a), nout), np.float32)
        self._check(self.L.cm_debug_math_host(self.h, C.c_int(op), _ptr(a), C.c_size_t(len(a)), _ptr(out)))
        return out

    # ---- one map over several GPUs (cm_dist_*) -------------------------------------------------------------------------------
    @staticmethod
    def dist_unique_id():
        """128-byte NCCL id (call on rank 0, hand the bytes to every rank)."""
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.cm_dist_unique_id(buf)
        if rc != 0:
            raise CoopermapError("cm_dist_unique_id failed (%d): libnccl.so.2 not loadable?" % rc)
        return buf.raw

    def dist_init(self, id128, rank, nranks):
        """Join `nranks` contexts (one per process / GPU) into one sharded map; call before mapping_create."""
        self._check(self.L.cm_dist_init(self.h, C.c_char_p(id128), C.c_int(rank), C.c_int(nranks)))

    def dist_info(self):
        r = C.c_int(0); n = C.c_int(1); p = C.c_int(0)
        self._check(self.L.cm_dist_info(self.h, C.byref(r), C.byref(n), C.byref(p)))
        return dict(rank=r.value, nranks=n.value, p2p=bool(p.value))

    def dist_allreduce(self, vec, repeat=1):
        """Sum a float64 vector over the ranks with the library's exchange kernel -> (total, ms per call)."""
        v = np.ascontiguousarray(vec, np.float64).copy()
        ms = C.c_float(0)
        self._check(self.L.cm_dist_allreduce_host(self.h, _ptr(v), C.c_int(len(v)), C.c_int(repeat), C.byref(ms)))
        return v, ms.value

    # ---- mapping stage (device-resident map) -------------------------------------------------------------------------
    def mapping_create(self, nstreams, max_corner_points=200000, max_surf_points=2000000):
        self.nstreams = int(nstreams)
        self._check(self.L.cm_mapping_create(self.h, C.c_int(nstreams), C.c_size_t(max_corner_points), C.c_size_t(max_surf_points)))

    @staticmethod
    def _pack_isos(isos):
        a = np.empty((len(isos), 12), np.float32)
        for i, (R, t) in enumerate(isos):
            a[i, :9] = np.asarray(R, np.float32).ravel(); a[i, 9:] = np.asarray(t, np.float32).ravel()
        return a

    @staticmethod
    def _pack_clouds(clouds):
        n = np.array([len(c) for c in clouds], np.int32)
        cap = max(int(n.max()), 1)
        buf = np.zeros((len(clouds), cap, 4), np.float32)
        for i, c in enumerate(clouds):
            buf[i, :len(c)] = _f32(c, 4)
        return buf, n, cap

    def _unpack_results(self, mapped, stats):
        out_iso = [(mapped[i, :9].reshape(3, 3).copy(), mapped[i, 9:].copy()) for i in range(len(mapped))]
        out_st = [dict(status=s.status, ret=bool(s.ret), converged=bool(s.converged), degenerate=bool(s.degenerate),
                       iterations=s.iterations, rows=s.rows, line=s.line_matches, plane=s.plane_matches, score=s.score)
                  for s in stats]
        return out_iso, out_st

    def mapping_process(self, odoms, corners, surfs):
        """LaserMapping::process for one frame per stream.  odoms: [(R, t)], corners / surfs: lists of (n, 4) clouds."""
        S = self.nstreams
        od = self._pack_isos(odoms)
        cb, cn, ccap = self._pack_clouds(corners); sb, sn, scap = self._pack_clouds(surfs)
        mapped = np.empty((S, 12), np.float32); stats = (MatchStats * S)()
        self._check(self.L.cm_mapping_process_host(self.h, _ptr(od), _ptr(cb), _ptr(cn), C.c_int(ccap), _ptr(sb), _ptr(sn),
                                                   C.c_int(scap), _ptr(mapped), stats))
        return self._unpack_results(mapped, stats)

    def localization_process(self, odoms, corners, surfs):
        """LaserLocalization::process for one frame per stream (FeatureMap::scanMatchScan against the resident map)."""
        S = self.nstreams
        od = self._pack_isos(odoms)
        cb, cn, ccap = self._pack_clouds(corners); sb, sn, scap = self._pack_clouds(surfs)
        mapped = np.empty((S, 12), np.float32); stats = (MatchStats * S)()
        self._check(self.L.cm_localization_process_host(self.h, _ptr(od), _ptr(cb), _ptr(cn), C.c_int(ccap), _ptr(sb), _ptr(sn),
                                                        C.c_int(scap), _ptr(mapped), stats))
        return self._unpack_results(mapped, stats)

    def mapping_local_create(self, use_mapped_pose=False):
        self._check(self.L.cm_mapping_local_create(self.h, C.c_int(int(use_mapped_pose))))

    def mapping_local_process(self, odom, corner, surf):
        """LaserMappingLocal::process for one frame: odom (R, t), corner / surf (n, 4) clouds -> ((R, t), stats)."""
        od = self._pack_isos([odom]); c = _f32(corner, 4); s = _f32(surf, 4)
        mapped = np.empty((1, 12), np.float32); stats = (MatchStats * 1)()
        self._check(self.L.cm_mapping_local_process_host(self.h, _ptr(od), _ptr(c), C.c_size_t(len(c)), _ptr(s), C.c_size_t(len(s)),
                                                         _ptr(mapped), stats))
        isos, st = self._unpack_results(mapped, stats)
        return isos[0], st[0]

    def mapping_local_window(self, clouds=False):
        """State of LocalFeatureMap's queue: dict(frames, nCorner, nSurf, nSurroundCorner, nSurroundSurf, accumDistance[, corner, surf])."""
        nf = C.c_int(0); nc = C.c_size_t(0); ns = C.c_size_t(0); sur = (C.c_int * 2)(); acc = C.c_double(0.0)
        self._check(self.L.cm_mapping_local_window_host(self.h, C.byref(nf), C.byref(nc), C.byref(ns), sur, C.byref(acc), None,
                                                        C.c_size_t(0), None, C.c_size_t(0)))
        d = dict(frames=nf.value, nCorner=nc.value, nSurf=ns.value, nSurroundCorner=sur[0], nSurroundSurf=sur[1], accumDistance=acc.value)
        if clouds:
            c = np.empty((max(nc.value, 1), 4), np.float32); s = np.empty((max(ns.value, 1), 4), np.float32)
            self._check(self.L.cm_mapping_local_window_host(self.h, None, None, None, None, None, _ptr(c), C.c_size_t(len(c)), _ptr(s),
                                                            C.c_size_t(len(s))))
            d["corner"] = c[:nc.value].copy(); d["surf"] = s[:ns.value].copy()
        return d

    def map_save(self, stream, directory):
        """FeatureMap::saveCloudToFiles: index.txt + <count>.pcd; returns the number of files."""
        n = C.c_int(0)
        self._check(self.L.cm_map_save_host(self.h, C.c_int(stream), str(directory).encode(), C.byref(n)))
        return n.value

    def map_load(self, stream, directory):
        """FeatureMap::loadCloudFromFiles -> (files, points read, points whose coordinates disagree with the index line)."""
        n = C.c_int(0); npts = C.c_size_t(0); bad = C.c_size_t(0)
        self._check(self.L.cm_map_load_host(self.h, C.c_int(stream), str(directory).encode(), C.byref(n), C.byref(npts), C.byref(bad)))
        return n.value, npts.value, bad.value

    def map_page_open(self, stream, directory, window=(21, 11, 21)):
        """DynamicFeatureMap::setupFilesDirectory: read <dir>/index2.txt (global cube indices) -> number of catalogue lines."""
        n = C.c_int(0)
        self._check(self.L.cm_map_page_open_host(self.h, C.c_int(stream), str(directory).encode(), C.c_int(window[0]), C.c_int(window[1]),
                                                 C.c_int(window[2]), C.byref(n)))
        return n.value

    def map_page_update(self, stream, sensor):
        """DynamicFeatureMap::update -> (files read, cubes dropped, points read)."""
        sx = _f32(sensor)
        nf = C.c_int(0); ne = C.c_int(0); npts = C.c_size_t(0)
        self._check(self.L.cm_map_page_update_host(self.h, C.c_int(stream), _ptr(sx), C.byref(nf), C.byref(ne), C.byref(npts)))
        return nf.value, ne.value, npts.value

    def pipeline_step(self, frames, odoms):
        """Scan registration + mapping for one organised sweep per stream: frames (S, rows, cols, 4)."""
        fr = _f32(frames)
        S, rows, cols = fr.shape[:3]
        od = self._pack_isos(odoms)
        mapped = np.empty((S, 12), np.float32); stats = (MatchStats * S)()
        self._check(self.L.cm_pipeline_step_host(self.h, _ptr(fr), C.c_int(rows), C.c_int(cols), _ptr(od), _ptr(mapped), stats))
        return self._unpack_results(mapped, stats)

    def pipeline_step_dev(self, frames_dev_ptr, rows, cols, odoms_packed, mapped_out, stats_out):
        """Same with the frames already resident in device memory (raw pointer); pre-packed host arrays, no allocation."""
        return self._check(self.L.cm_pipeline_step_dev(self.h, C.c_void_p(frames_dev_ptr), C.c_int(rows), C.c_int(cols),
                                                       _ptr(odoms_packed), _ptr(mapped_out), stats_out))

    def pipeline_prefetch(self, frames, deferred=False):
        """Start the host-to-device upload of the NEXT step's sweeps (pinned (S, rows, cols, 4) float32 array).
        deferred: only registered now, issued by the next pipeline_step right after it has submitted its Gauss-Newton loop."""
        fn = self.L.cm_pipeline_prefetch_deferred_host if deferred else self.L.cm_pipeline_prefetch_host
        return self._check(fn(self.h, _ptr(frames), C.c_int(frames.shape[1]), C.c_int(frames.shape[2])))

    def pipeline_prefetch_dev(self, frames_dev_ptr, rows, cols, deferred=False):
        """Issue scan registration of the NEXT step's device-resident sweeps on the side stream."""
        fn = self.L.cm_pipeline_prefetch_deferred_dev if deferred else self.L.cm_pipeline_prefetch_dev
        return self._check(fn(self.h, C.c_void_p(frames_dev_ptr), C.c_int(rows), C.c_int(cols)))

    def pipeline_step_packed(self, frames, odoms_packed, mapped_out, stats_out):
        fr = frames
        return self._check(self.L.cm_pipeline_step_host(self.h, _ptr(fr), C.c_int(fr.shape[1]), C.c_int(fr.shape[2]),
                                                        _ptr(odoms_packed), _ptr(mapped_out), stats_out))

    @staticmethod
    def _cloud_ptrs(clouds):
        """clouds: one array per stream whose rows are points (x, y, z first); returns (void*[S], stride in bytes)."""
        stride = int(clouds[0].strides[-2]) if clouds[0].ndim >= 2 else int(clouds[0].strides[0])
        arr = (C.c_void_p * len(clouds))(*[c.ctypes.data for c in clouds])
        return arr, stride

    def cloud_ptrs(self, clouds):
        """(void*[S], stride) of a list of per-stream clouds, for the *_ptrs variants below (built once, used every step)."""
        return self._cloud_ptrs(clouds)

    def pipeline_prefetch_strided_ptrs(self, arr, stride, rows, cols):
        return self._check(self.L.cm_pipeline_prefetch_strided_host(self.h, arr, C.c_size_t(stride), C.c_int(rows), C.c_int(cols)))

    def pipeline_step_strided_ptrs(self, arr, stride, rows, cols, odoms_packed, mapped_out, stats_out):
        return self._check(self.L.cm_pipeline_step_strided_host(self.h, arr, C.c_size_t(stride), C.c_int(rows), C.c_int(cols),
                                                                _ptr(odoms_packed), _ptr(mapped_out), stats_out))

    def pipeline_prefetch_strided(self, clouds, rows, cols):
        """cm_pipeline_prefetch_strided_host: one host cloud per stream, any point stride (pcl::PointXYZI = 32 bytes), pageable memory."""
        arr, stride = self._cloud_ptrs(clouds)
        return self._check(self.L.cm_pipeline_prefetch_strided_host(self.h, arr, C.c_size_t(stride), C.c_int(rows), C.c_int(cols)))

    def pipeline_step_strided(self, clouds, rows, cols, odoms_packed, mapped_out, stats_out):
        arr, stride = self._cloud_ptrs(clouds)
        return self._check(self.L.cm_pipeline_step_strided_host(self.h, arr, C.c_size_t(stride), C.c_int(rows), C.c_int(cols),
                                                                _ptr(odoms_packed), _ptr(mapped_out), stats_out))

    def map_update(self, stream, sensor):
        """FeatureMap::update(sensorPose) for one stream: shift if needed + the valid-cube window."""
        sx = _f32(sensor)
        self._check(self.L.cm_map_update_host(self.h, C.c_int(stream), _ptr(sx)))

    def map_insert(self, corners, surfs, tfs):
        """FeatureMap::addFeatureCloud per stream."""
        cb, cn, ccap = self._pack_clouds(corners); sb, sn, scap = self._pack_clouds(surfs)
        tf = self._pack_isos(tfs)
        self._check(self.L.cm_map_insert_host(self.h, _ptr(cb), _ptr(cn), C.c_int(ccap), _ptr(sb), _ptr(sn), C.c_int(scap), _ptr(tf)))

    def map_export(self, stream, cls):
        """All resident points of one stream's map (cls 0 corner / 1 surf) + cube index, in storage order."""
        n = C.c_size_t(0)
        self._check(self.L.cm_map_export_host(self.h, C.c_int(stream), C.c_int(cls), None, None, C.c_size_t(0), C.byref(n)))
        pts = np.empty((max(n.value, 1), 4), np.float32); cube = np.empty(max(n.value, 1), np.int32)
        self._check(self.L.cm_map_export_host(self.h, C.c_int(stream), C.c_int(cls), _ptr(pts), _ptr(cube), C.c_size_t(len(pts)), C.byref(n)))
        return pts[:n.value].copy(), cube[:n.value].copy()

    def map_surround(self, stream):
        """FeatureMap::getSurroundFeature -> (corner, surf) clouds of the valid cubes, in the reference's order."""
        n = (C.c_size_t * 2)()
        self._check(self.L.cm_map_surround_host(self.h, C.c_int(stream), None, C.c_size_t(0), None, C.c_size_t(0), n))
        c = np.empty((max(n[0], 1), 4), np.float32); s = np.empty((max(n[1], 1), 4), np.float32)
        self._check(self.L.cm_map_surround_host(self.h, C.c_int(stream), _ptr(c), C.c_size_t(len(c)), _ptr(s), C.c_size_t(len(s)), n))
        return c[:n[0]].copy(), s[:n[1]].copy()

    def map_full(self, stream, leaf):
        """FeatureMap::getFullMap: every cube's corner and surf cloud re-filtered with `leaf`, cubes in index order."""
        n = C.c_size_t(0)
        self._check(self.L.cm_map_full_host(self.h, C.c_int(stream), C.c_float(leaf), None, C.c_size_t(0), C.byref(n)))
        out = np.empty((max(n.value, 1), 4), np.float32)
        self._check(self.L.cm_map_full_host(self.h, C.c_int(stream), C.c_float(leaf), _ptr(out), C.c_size_t(len(out)), C.byref(n)))
        return out[:n.value].copy()

    def map_export_sorted(self, stream, cls):
        """Export ordered like the reference's cube clouds: by cube index, then by voxel (z, y, x)."""
        pts, cube = self.map_export(stream, cls)
        leaf = self.cfg.map_filter_corner if cls == 0 else self.cfg.map_filter_surf
        inv = np.float32(1.0) / np.float32(leaf)
        v = np.floor(pts[:, :3] * inv).astype(np.int64)
        order = np.lexsort((v[:, 0], v[:, 1], v[:, 2], cube))
        return pts[order], cube[order]

    # ---- scan-to-scan odometry -------------------------------------------------------------------------------------
    def odometry_reset(self):
        self._check(self.L.cm_odometry_reset(self.h))

    def odometry_process(self, sharp, less_sharp, flat, less_flat, trace=False):
        """LaserOdometry::process for one frame -> dict(R, t, transform, corner_last, surf_last, stats, log)."""
        a = [_f32(x, 4) for x in (sharp, less_sharp, flat, less_flat)]
        iso = np.empty(12, np.float32); tf = np.empty(6, np.float32); st = OdomStats()
        cl = np.empty((max(len(a[1]), 1), 4), np.float32); sl = np.empty((max(len(a[3]), 1), 4), np.float32)
        tr = (IterTrace * 25)() if trace else None
        self._check(self.L.cm_odometry_process_host(self.h, _ptr(a[0]), C.c_int(len(a[0])), _ptr(a[1]), C.c_int(len(a[1])), _ptr(a[2]),
                                                    C.c_int(len(a[2])), _ptr(a[3]), C.c_int(len(a[3])), _ptr(iso), _ptr(tf), _ptr(cl), _ptr(sl),
                                                    C.byref(st), tr))
        log = []
        if trace:
            for it in range(25):
                t = tr[it]
                log.append(dict(pose_in=np.array(t.pose_in[:], np.float32), x=np.array(t.x[:], np.float32), rows=t.rows))
        return dict(R=iso[:9].reshape(3, 3).copy(), t=iso[9:].copy(), transform=tf, corner_last=cl[:len(a[1])].copy(),
                    surf_last=sl[:len(a[3])].copy(), iterations=st.iterations, rows=st.rows, matched=bool(st.matched),
                    initialising=bool(st.initialising), converged=bool(st.converged), log=log)

    def odometry_batch_create(self, nstreams, cap_sharp, cap_less_sharp, cap_flat, cap_less_flat):
        self._ob = (nstreams, cap_sharp, cap_less_sharp, cap_flat, cap_less_flat)
        self._check(self.L.cm_odometry_batch_create(self.h, *[C.c_int(v) for v in self._ob]))

    def odometry_batch_process(self, sharps, less_sharps, flats, less_flats, want_clouds=True):
        """LaserOdometry::process for one frame of EVERY stream (lists of (n, 4) clouds) -> list of dicts like odometry_process."""
        S, caps = self._ob[0], self._ob[1:]
        bufs, cnts = [], []
        for clouds, cap in zip((sharps, less_sharps, flats, less_flats), caps):
            b = np.zeros((S, cap, 4), np.float32); n = np.zeros(S, np.int32)
            for s, c in enumerate(clouds):
                c = _f32(c, 4); n[s] = len(c); b[s, :len(c)] = c
            bufs.append(b); cnts.append(n)
        iso = np.empty((S, 12), np.float32); tf = np.empty((S, 6), np.float32); st = (OdomStats * S)()
        cl = np.empty((S, caps[1], 4), np.float32) if want_clouds else None
        sl = np.empty((S, caps[3], 4), np.float32) if want_clouds else None
        self._check(self.L.cm_odometry_batch_process_host(self.h, _ptr(bufs[0]), _ptr(cnts[0]), _ptr(bufs[1]), _ptr(cnts[1]), _ptr(bufs[2]),
                                                          _ptr(cnts[2]), _ptr(bufs[3]), _ptr(cnts[3]), _ptr(iso), _ptr(tf),
                                                          _ptr(cl) if want_clouds else None, _ptr(sl) if want_clouds else None, st))
        out = []
        for s in range(S):
            out.append(dict(R=iso[s, :9].reshape(3, 3).copy(), t=iso[s, 9:].copy(), transform=tf[s].copy(),
                            corner_last=cl[s, :cnts[1][s]].copy() if want_clouds else None,
                            surf_last=sl[s, :cnts[3][s]].copy() if want_clouds else None, iterations=st[s].iterations, rows=st[s].rows,
                            matched=bool(st[s].matched), initialising=bool(st[s].initialising), converged=bool(st[s].converged)))
        return out

    def pipeline_chain_create(self, rows, cols):
        """size the odometry stage of the three-stage chain for rows x cols sweeps (after mapping_create)"""
        self._check(self.L.cm_pipeline_chain_create(self.h, C.c_int(rows), C.c_int(cols)))

    def pipeline_chain_step(self, frames):
        """scan registration -> laserOdometry -> laserMapping for one organised sweep per stream, frames (S, rows, cols, 4);
        returns (odometry poses, mapped poses, odometry stats, mapping stats)"""
        fr = _f32(frames)
        S, rows, cols = fr.shape[:3]
        od = np.empty((S, 12), np.float32); mapped = np.empty((S, 12), np.float32)
        ost = (OdomStats * S)(); mst = (MatchStats * S)()
        self._check(self.L.cm_pipeline_chain_step_host(self.h, _ptr(fr), C.c_int(rows), C.c_int(cols), _ptr(od), _ptr(mapped), ost, mst))
        isos, stats = self._unpack_results(mapped, mst)
        odoms = [(od[s, :9].reshape(3, 3).copy(), od[s, 9:].copy()) for s in range(S)]
        ostats = [dict(iterations=ost[s].iterations, rows=ost[s].rows, matched=bool(ost[s].matched), initialising=bool(ost[s].initialising),
                       converged=bool(ost[s].converged)) for s in range(S)]
        return odoms, isos, ostats, stats

    def pipeline_chain_sweep_create(self, max_points):
        """the chain for ONE stream fed with raw (unorganised) sweeps of at most max_points points"""
        self._check(self.L.cm_pipeline_chain_sweep_create(self.h, C.c_size_t(max_points)))

    def pipeline_chain_step_sweep(self, sweep, lidar, imu_scan_time=None):
        """MultiScanRegistration front end -> feature extraction -> odometry -> mapping for one raw sweep (n, 4);
        returns ((R, t) odometry, (R, t) mapped, odometry stats dict, mapping stats dict)"""
        sw = _f32(sweep, 4)
        od = np.empty((1, 12), np.float32); mapped = np.empty((1, 12), np.float32)
        ost = (OdomStats * 1)(); mst = (MatchStats * 1)()
        self._check(self.L.cm_pipeline_chain_step_sweep_host(self.h, _ptr(sw), C.c_size_t(len(sw)), C.c_int(lidar),
                                                             C.c_double(-1.0 if imu_scan_time is None else imu_scan_time), _ptr(od), _ptr(mapped), ost, mst))
        isos, stats = self._unpack_results(mapped, mst)
        return ((od[0, :9].reshape(3, 3).copy(), od[0, 9:].copy()), isos[0],
                dict(iterations=ost[0].iterations, rows=ost[0].rows, matched=bool(ost[0].matched), initialising=bool(ost[0].initialising)), stats[0])

    def pipeline_chain_step_dev(self, frames_dev_ptr, rows, cols, odom_out, mapped_out, ostats_out, mstats_out):
        """the same with the sweeps already in device memory ([S][rows][cols] float4, raw pointer)"""
        return self._check(self.L.cm_pipeline_chain_step_dev(self.h, C.c_void_p(frames_dev_ptr), C.c_int(rows), C.c_int(cols), _ptr(odom_out),
                                                             _ptr(mapped_out), ostats_out, mstats_out))

    def pipeline_chain_step_packed(self, frames, odom_out, mapped_out, ostats_out, mstats_out):
        fr = frames
        return self._check(self.L.cm_pipeline_chain_step_host(self.h, _ptr(fr), C.c_int(fr.shape[1]), C.c_int(fr.shape[2]), _ptr(odom_out),
                                                              _ptr(mapped_out), ostats_out, mstats_out))

    # ---- measurement helpers -------------------------------------------------------------------------------------
    def timer_record(self, which):
        self._check(self.L.cm_timer_record(self.h, C.c_int(which)))

    def timer_record_side(self, which):
        self._check(self.L.cm_timer_record_side(self.h, C.c_int(which)))

    def mapping_sync(self):
        """wait for the map insertion of the last mapping / pipeline step and report what it hit"""
        self._check(self.L.cm_mapping_sync(self.h))

    def pipeline_wait(self):
        self._check(self.L.cm_pipeline_wait(self.h))

    def pipeline_discard(self, frames_ptr):
        self._check(self.L.cm_pipeline_discard(self.h, C.c_void_p(frames_ptr)))

    def graph_builds(self):
        """(Gauss-Newton loop graphs built, filter / insert chain graphs captured) so far."""
        a = (C.c_ulonglong * 4)()
        self._check(self.L.cm_debug_graph_builds(self.h, a))
        self.insert_redos = int(a[2])
        return int(a[0]), int(a[1])

    def timer_elapsed_ms(self):
        ms = C.c_float(0)
        self._check(self.L.cm_timer_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def prof_enable(self, on):
        self._check(self.L.cm_prof_enable(self.h, C.c_int(int(on))))

    def prof_drain(self):
        ms = C.c_double(0); n = C.c_int(0)
        self._check(self.L.cm_prof_drain(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def prof_drain_scanreg(self):
        ms = C.c_double(0); n = C.c_int(0)
        self._check(self.L.cm_prof_drain_scanreg(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timeline_enable(self, on):
        self._check(self.L.cm_timeline_enable(self.h, C.c_int(int(on))))

    def timeline_report(self):
        buf = C.create_string_buffer(1 << 16)
        self._check(self.L.cm_timeline_report(self.h, buf, C.c_size_t(len(buf))))
        return buf.value.decode()

    def last_step_counters(self):
        a = (C.c_ulonglong * 4)()
        self._check(self.L.cm_last_step_counters(self.h, a))
        return dict(query_iters=a[0], queries=a[1], inserted=a[2], features=a[3])

    # ---- scan registration -----------------------------------------------------------------------------------------
    def scanreg_organised(self, frames, debug=False):
        """frames: (S, rows, cols, 4) or (rows, cols, 4) organised sweeps -> list of per-stream dicts with the four
        feature clouds (sharp, lessSharp, flat, lessFlat); debug=True adds the full cloud, ring ranges, index lists,
        mask, curvature and pointClassify labels (what the oracle returns)."""
        fr = _f32(frames)
        single = fr.ndim == 3
        if single:
            fr = fr[None]
        S, rows, cols = fr.shape[:3]
        npts = rows * cols
        out = ScanRegOut()
        bufs = [np.empty((S, npts, 4), np.float32) for _ in range(4)]
        n = np.zeros((S, 5), np.int32)
        for k in range(4):
            out.pts[k] = bufs[k].ctypes.data; out.cap[k] = npts
        out.n = n.ctypes.data
        dbg = {}
        if debug:
            dbg = dict(cloud=np.empty((S, npts, 4), np.float32), ccurv=np.empty((S, npts), np.float32),
                       ranges=np.empty((S, rows, 2), np.int32), idx=[np.empty((S, npts), np.int32) for _ in range(4)],
                       picked=np.empty((S, npts), np.int8), curvature=np.empty((S, npts), np.float32),
                       label=np.empty((S, npts), np.int8))
            out.cloud = dbg["cloud"].ctypes.data; out.cloud_curvature = dbg["ccurv"].ctypes.data
            out.scan_ranges = dbg["ranges"].ctypes.data
            for k in range(4):
                out.idx[k] = dbg["idx"][k].ctypes.data
            out.picked = dbg["picked"].ctypes.data; out.curvature = dbg["curvature"].ctypes.data; out.label = dbg["label"].ctypes.data
        self._check(self.L.cm_scanreg_organised_host(self.h, _ptr(fr), C.c_int(S), C.c_int(rows), C.c_int(cols), C.byref(out)))
        res = []
        names = ["sharp", "lessSharp", "flat", "lessFlat"]
        for s_ in range(S):
            d = {names[k]: bufs[k][s_, :n[s_, k]].copy() for k in range(4)}
            if debug:
                ncloud = int(dbg["ranges"][s_, -1, 1]) + 1 if dbg["ranges"][s_, :, 1].max() > 0 or n[s_].sum() > 0 else 0
                total = int(sum(max(0, int(e) - int(b) + 1) for b, e in dbg["ranges"][s_] if not (b == 0 and e == 0)))
                # ring sizes: a ring is empty when its range is (first, first-1) collapsed to (first, max(first-1, 0))
                d["scanStart"] = dbg["ranges"][s_, :, 0].copy(); d["scanEnd"] = dbg["ranges"][s_, :, 1].copy()
                d["cloud"] = np.concatenate([dbg["cloud"][s_], dbg["ccurv"][s_][:, None]], 1)
                d["sharpIdx"] = dbg["idx"][0][s_, :n[s_, 0]].copy(); d["lessSharpIdx"] = dbg["idx"][1][s_, :n[s_, 1]].copy()
                d["flatIdx"] = dbg["idx"][2][s_, :n[s_, 2]].copy(); d["lessFlatRawIdx"] = dbg["idx"][3][s_, :n[s_, 4]].copy()
                d["picked"] = dbg["picked"][s_].astype(np.int32); d["curvature"] = dbg["curvature"][s_].copy()
                d["classLabel"] = dbg["label"][s_].astype(np.int32)
            res.append(d)
        return res[0] if single else res

    def imu_push(self, stamp, roll, pitch, yaw, ax, ay, az):
        """ScanRegistration::handleIMUMessage."""
        m = (C.c_double * 7)(stamp, roll, pitch, yaw, ax, ay, az)
        self._check(self.L.cm_imu_push_host(self.h, m))

    def imu_clear(self):
        self._check(self.L.cm_imu_clear(self.h))

    def scanreg_sweep(self, sweep, lidar, debug=False, imu_scan_time=None):
        """MultiScanRegistration::process for one raw azimuth-major sweep (n, 4); lidar 0 VLP-16, 1 HDL-32, 2 HDL-64E, 3 Pandar40."""
        sw = _f32(sweep, 4)
        n = max(len(sw), 1)
        out = ScanRegOut()
        bufs = [np.empty((n, 4), np.float32) for _ in range(4)]
        cnt = np.zeros(5, np.int32)
        for k in range(4):
            out.pts[k] = bufs[k].ctypes.data; out.cap[k] = n
        out.n = cnt.ctypes.data
        dbg = {}
        if debug:
            rings = {0: 16, 1: 32, 2: 64, 3: 40}[lidar]
            dbg = dict(cloud=np.empty((n, 4), np.float32), ccurv=np.empty(n, np.float32), ranges=np.empty((rings, 2), np.int32),
                       idx=[np.empty(n, np.int32) for _ in range(4)], picked=np.empty(n, np.int8),
                       curvature=np.empty(n, np.float32), label=np.empty(n, np.int8))
            out.cloud = dbg["cloud"].ctypes.data; out.cloud_curvature = dbg["ccurv"].ctypes.data
            out.scan_ranges = dbg["ranges"].ctypes.data
            for k in range(4):
                out.idx[k] = dbg["idx"][k].ctypes.data
            out.picked = dbg["picked"].ctypes.data; out.curvature = dbg["curvature"].ctypes.data; out.label = dbg["label"].ctypes.data
        rows = C.c_int(0); cols = C.c_int(0)
        imu_trans = np.zeros(12, np.float32)
        if imu_scan_time is None:
            self._check(self.L.cm_scanreg_sweep_host(self.h, _ptr(sw), C.c_size_t(len(sw)), C.c_int(lidar), C.byref(out), C.byref(rows), C.byref(cols)))
        else:   # de-skew with the IMU states pushed so far (cm_imu_push_host)
            self._check(self.L.cm_scanreg_sweep_imu_host(self.h, _ptr(sw), C.c_size_t(len(sw)), C.c_int(lidar), C.c_double(imu_scan_time),
                                                         C.byref(out), C.byref(rows), C.byref(cols), _ptr(imu_trans)))
        names = ["sharp", "lessSharp", "flat", "lessFlat"]
        d = {names[k]: bufs[k][:cnt[k]].copy() for k in range(4)}
        if debug:
            d["scanStart"] = dbg["ranges"][:, 0].copy(); d["scanEnd"] = dbg["ranges"][:, 1].copy()
            d["cloud"] = np.concatenate([dbg["cloud"], dbg["ccurv"][:, None]], 1)
            d["sharpIdx"] = dbg["idx"][0][:cnt[0]].copy(); d["lessSharpIdx"] = dbg["idx"][1][:cnt[1]].copy()
            d["flatIdx"] = dbg["idx"][2][:cnt[2]].copy(); d["lessFlatRawIdx"] = dbg["idx"][3][:cnt[4]].copy()
            d["picked"] = dbg["picked"].astype(np.int32); d["curvature"] = dbg["curvature"].copy()
            d["classLabel"] = dbg["label"].astype(np.int32)
        if imu_scan_time is not None:
            d["imu_trans"] = imu_trans.reshape(4, 3).copy()
        return d

    # ---- voxel filter --------------------------------------------------------------------------------------------
    def voxel_filter_batch(self, clouds, leaf):
        """clouds: list of (n_i, 4) arrays -> list of filtered (m_i, 4) arrays (pcl::VoxelGrid semantics)."""
        nseg = len(clouds)
        n_in = np.array([len(c) for c in clouds], np.int32)
        cap = max(int(n_in.max()), 1)
        buf = np.zeros((nseg, cap, 4), np.float32)
        for i, c in enumerate(clouds):
            buf[i, :len(c)] = _f32(c, 4)
        out = np.empty((nseg, cap, 4), np.float32); n_out = np.zeros(nseg, np.int32)
        self._check(self.L.cm_voxel_filter_host(self.h, _ptr(buf), C.c_int(nseg), _ptr(n_in), C.c_int(cap), C.c_float(leaf),
                                                _ptr(out), _ptr(n_out), C.c_int(cap)))
        return [out[i, :n_out[i]].copy() for i in range(nseg)]

    def voxel_filter(self, cloud, leaf):
        return self.voxel_filter_batch([cloud], leaf)[0]

    # ---- exact 5-NN ---------------------------------------------------------------------------------------
    def knn5(self, map_pts, queries, cell=1.2, gate=5.0):
        m = _f32(map_pts, 4); q = _f32(queries, 3)
        idx = np.empty((len(q), 5), np.int32); d2 = np.empty((len(q), 5), np.float32)
        self._check(self.L.cm_knn5_host(self.h, _ptr(m), C.c_size_t(len(m)), C.c_float(cell), C.c_float(gate), _ptr(q),
                                        C.c_size_t(len(q)), _ptr(idx), _ptr(d2)))
        return idx, d2

    # ---- stateless scan-to-map --------------------------------------------------------------------------------
    def match_stateless(self, ref_corner, ref_surf, corner, surf, pose, trace=False):
        rc_ = _f32(ref_corner, 4); rs = _f32(ref_surf, 4); c = _f32(corner, 4); s = _f32(surf, 4)
        p = _f32(pose).copy()
        st = MatchStats()
        iters = self.cfg.max_iterations
        tr = (IterTrace * iters)() if trace else None
        nnc = np.empty((iters, len(c), 5), np.int32) if trace else None
        nns = np.empty((iters, len(s), 5), np.int32) if trace else None
        rc = self._check(self.L.cm_match_stateless_host(
            self.h, _ptr(rc_), C.c_size_t(len(rc_)), _ptr(rs), C.c_size_t(len(rs)), _ptr(c), C.c_size_t(len(c)), _ptr(s),
            C.c_size_t(len(s)), _ptr(p), C.byref(st), tr, _ptr(nnc), _ptr(nns)))
        stats = dict(status=rc, ret=bool(st.ret), converged=bool(st.converged), degenerate=bool(st.degenerate),
                     iterations=st.iterations, rows=st.rows, line=st.line_matches, plane=st.plane_matches, score=st.score)
        log = []
        if trace:
            n_eval = st.iterations + (1 if rc == CM_TOO_FEW_MATCHES else 0)
            for it in range(n_eval):
                t = tr[it]
                log.append(dict(pose_in=np.array(t.pose_in[:], np.float32), AtA=np.array(t.AtA[:], np.float32).reshape(6, 6),
                                AtB=np.array(t.AtB[:], np.float32), x=np.array(t.x[:], np.float32),
                                counts=np.array([t.rows, t.line_matches, t.plane_matches, t.degenerate], np.int32),
                                nnCorner=nnc[it], nnSurf=nns[it]))
        return p, stats, log

    def match_local(self, ref_corner, ref_surf, corner, surf, pose):
        """ScanMatch::scanMatchLocal: voxel-filter the four clouds (0.2 / 0.4) then scanMatchScan."""
        rc_ = _f32(ref_corner, 4); rs = _f32(ref_surf, 4); c = _f32(corner, 4); s = _f32(surf, 4)
        p = _f32(pose).copy(); st = MatchStats()
        rc = self._check(self.L.cm_match_local_host(self.h, _ptr(rc_), C.c_size_t(len(rc_)), _ptr(rs), C.c_size_t(len(rs)), _ptr(c),
                                                    C.c_size_t(len(c)), _ptr(s), C.c_size_t(len(s)), _ptr(p), C.byref(st)))
        return p, dict(status=rc, ret=bool(st.ret), converged=bool(st.converged), degenerate=bool(st.degenerate),
                       iterations=st.iterations, rows=st.rows, line=st.line_matches, plane=st.plane_matches, score=st.score)

    def match_stateless_iso(self, ref_corner, ref_surf, corner, surf, R, t):
        rc_ = _f32(ref_corner, 4); rs = _f32(ref_surf, 4); c = _f32(corner, 4); s = _f32(surf, 4)
        iso = np.concatenate([_f32(R).ravel(), _f32(t).ravel()]).astype(np.float32)
        st = MatchStats()
        rc = self._check(self.L.cm_match_stateless_iso_host(
            self.h, _ptr(rc_), C.c_size_t(len(rc_)), _ptr(rs), C.c_size_t(len(rs)), _ptr(c), C.c_size_t(len(c)), _ptr(s),
            C.c_size_t(len(s)), _ptr(iso), C.byref(st)))
        return iso[:9].reshape(3, 3).copy(), iso[9:].copy(), rc, st


class LaserMapping:
    """Mirror of lidar_slam::LaserMapping (LaserMapping.h / LaserMatcher.h): the scan-to-map stage with its own map.

    process(odom, corner, surf) -> (R, t) of /aft_mapped_to_init.  One instance = one LiDAR stream; use
    Context.mapping_process directly to drive many streams in one batch."""

    def __init__(self, ctx=None, max_corner_points=200000, max_surf_points=2000000, **cfg):
        self.ctx = ctx or Context(**cfg)
        self.ctx.mapping_create(1, max_corner_points, max_surf_points)
        self.last_stats = None

    def process(self, odom_R, odom_t, laserCloudCornerLast, laserCloudSurfLast):
        isos, stats = self.ctx.mapping_process([(odom_R, odom_t)], [laserCloudCornerLast], [laserCloudSurfLast])
        self.last_stats = stats[0]
        return isos[0]


class LaserMappingLocal:
    """Mirror of lidar_slam::LaserMappingLocal (LaserMappingLocal.h): the mapping stage over LocalFeatureMap, a sliding
    window of voxel-filtered frames (30 m of travel).  use_mapped_pose=False is the reference as written (frames placed with
    the never-assigned _transformTobeMapped, i.e. the identity); True places them with the mapped pose."""

    def __init__(self, ctx=None, use_mapped_pose=False, **cfg):
        self.ctx = ctx or Context(**cfg)
        self.ctx.mapping_local_create(use_mapped_pose)
        self.last_stats = None

    def process(self, odom_R, odom_t, laserCloudCornerLast, laserCloudSurfLast):
        iso, self.last_stats = self.ctx.mapping_local_process((odom_R, odom_t), laserCloudCornerLast, laserCloudSurfLast)
        return iso


class LaserLocalization:
    """Mirror of lidar_slam::LaserLocalization (LaserLocalization.h): localisation against a prebuilt cube map.

    load(directory) reads the index.txt / PCD layout FeatureMap::saveCloudToFiles writes; process(odom, corner, surf)
    refines the pose with FeatureMap::scanMatchScan (own-cube neighbours) and leaves the map untouched."""

    def __init__(self, ctx=None, max_corner_points=200000, max_surf_points=2000000, **cfg):
        self.ctx = ctx or Context(**cfg)
        self.ctx.mapping_create(1, max_corner_points, max_surf_points)
        self.last_stats = None

    def load(self, directory):
        return self.ctx.map_load(0, directory)

    def process(self, odom_R, odom_t, laserCloudCornerLast, laserCloudSurfLast):
        isos, stats = self.ctx.localization_process([(odom_R, odom_t)], [laserCloudCornerLast], [laserCloudSurfLast])
        self.last_stats = stats[0]
        return isos[0]


class LaserOdometry:
    """Mirror of lidar_slam::LaserOdometry (LaserOdometry.h): frame-to-frame odometry on the four feature clouds."""

    def __init__(self, ctx=None, **cfg):
        self.ctx = ctx or Context(**cfg)
        self.ctx.odometry_reset()

    def process(self, cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat):
        return self.ctx.odometry_process(cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat)


class ScanMatch:
    """Mirror of lidar_slam::ScanMatch (ScanMatch.h:22-86): the scan-to-map Gauss-Newton operator.

    scanMatchScan(referenceCornerCloud, referenceSurfCloud, CornerCloud, SurfCloud, transform) -> (bool, transform)
    `transform` is a Twist (rx, ry, rz, tx, ty, tz); it is written back even when the call returns False, exactly
    like the reference (ScanMatch.cpp:342-346).
    """

    def __init__(self, maxIterations=10, ctx=None, **cfg):
        cfg.setdefault("delta_t_abort", 0.05)   # class defaults, ScanMatch.cpp:22
        cfg.setdefault("delta_r_abort", 0.05)
        cfg.setdefault("use_score", 1)          # ScanMatch.cpp:23
        cfg["max_iterations"] = int(maxIterations)
        self.ctx = ctx or Context(**cfg)
        self._match_count = 0
        self._fail_match_count = 0
        self._total_score = 0.0
        self.last_stats = None

    def setConvergeThreshold(self, deltaTAbort, deltaRAbort):   # ScanMatch.h:29-32
        self._rebuild(delta_t_abort=deltaTAbort, delta_r_abort=deltaRAbort)

    def setUseCore(self, useScore):   # ScanMatch.h:34 (sic)
        self._rebuild(use_score=int(bool(useScore)))

    def setScoreThreshold(self, score):   # ScanMatch.h:25
        self._rebuild(score_threshold=float(score))

    def _rebuild(self, **kw):
        vals = {f[0]: getattr(self.ctx.cfg, f[0]) for f in Config._fields_}
        vals.update(kw)
        self.ctx.close()
        self.ctx = Context(**vals)

    def scanMatchScan(self, referenceCornerCloud, referenceSurfCloud, CornerCloud, SurfCloud, transform):
        pose, stats, _ = self.ctx.match_stateless(referenceCornerCloud, referenceSurfCloud, CornerCloud, SurfCloud, transform)
        self.last_stats = stats
        if stats["ret"]:
            self._match_count += 1
            self._total_score += stats["score"]
        elif stats["status"] != CM_TOO_FEW_REF:
            self._fail_match_count += 1
        return bool(stats["ret"]), pose

    def scanMatchLocal(self, referenceCornerCloud, referenceSurfCloud, CornerCloud, SurfCloud, transform):   # ScanMatch.h:38-46
        pose, stats = self.ctx.match_local(referenceCornerCloud, referenceSurfCloud, CornerCloud, SurfCloud, transform)
        self.last_stats = stats
        if stats["ret"]:
            self._match_count += 1
            self._total_score += stats["score"]
        elif stats["status"] != CM_TOO_FEW_REF:
            self._fail_match_count += 1
        return bool(stats["ret"]), pose

    def getAverageScore(self):   # ScanMatch.h:59-61
        return self._total_score / self._match_count if self._match_count > 0 else 0.0
