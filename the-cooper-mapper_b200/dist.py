"""Multi-GPU plumbing for the two ways the path shards.

1. Independent LiDAR streams (BASELINE config 3): stream i -> rank i mod G, no data-path collective.
2. One large map split over ranks (BASELINE config 4): x-slabs with a sqrt(5) m halo; every rank evaluates the queries
   that fall in its slab, the 32 double partial sums of the normal equations are all-reduced per Gauss-Newton iteration
   (NCCL over NVLink through torch.distributed; gloo in the CPU tests), every rank solves redundantly.

PyTorch is plumbing here (process groups, one collective of 256 bytes per iteration); the compute is the C ABI.
"""
import ctypes as C

import numpy as np

from .api import Context, MatchStats, _f32, _ptr

GATE = 5.0                       # ScanMatch.cpp:102,120
HALO = float(np.sqrt(GATE)) * 1.001 + 1e-3
BIG = np.float32(3.0e38)


def stream_shard(n_streams, rank, world):
    """Indices of the streams rank `rank` drives (stream i -> rank i mod world)."""
    return list(range(rank, n_streams, world))


def slab_bounds(points, world, axis=0):
    """world + 1 slab boundaries along `axis` balancing the point count; the outer ones are -inf / +inf."""
    x = np.sort(np.asarray(points, np.float32)[:, axis])
    cuts = [x[min(len(x) - 1, (len(x) * r) // world)] for r in range(1, world)] if len(x) else [0.0] * (world - 1)
    return np.array([-BIG] + [np.float32(c) for c in cuts] + [BIG], np.float32)


def shard_cloud(points, bounds, rank, axis=0, halo=HALO):
    """The part of a map cloud rank `rank` must hold: its slab plus the halo that makes every accepted 5-NN local."""
    p = _f32(points, 4)
    lo, hi = bounds[rank], bounds[rank + 1]
    m = (p[:, axis] >= lo - np.float32(halo)) & (p[:, axis] < hi + np.float32(halo))
    return p[m]


def own_box(bounds, rank, axis=0):
    lo = np.full(3, -BIG, np.float32); hi = np.full(3, BIG, np.float32)
    lo[axis] = bounds[rank]; hi[axis] = bounds[rank + 1]
    return lo, hi


def allreduce_sums(sums, group=None, device=None):
    """Sum a (32,) float64 array over the process group.  NCCL needs a CUDA tensor, gloo takes the host tensor."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sums
    t = torch.from_numpy(np.ascontiguousarray(sums, np.float64))
    if dist.get_backend(group) == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class ShardedScanMatch:
    """ScanMatch::scanMatchScan (ScanMatch.cpp:51-347) against a reference map that is split over `world` ranks.

    Every rank constructs one with ITS shard (shard_cloud) and calls scanMatchScan with the same queries and pose;
    `reduce_fn` sums the partial normal equations over the ranks (default: torch.distributed all-reduce).
    """

    def __init__(self, ctx, corner_shard, surf_shard, total_ref_corner, total_ref_surf, lo, hi, reduce_fn=None):
        self.ctx = ctx
        self.L = ctx.L
        c = _f32(corner_shard, 4); s = _f32(surf_shard, 4)
        ctx._check(self.L.cm_shard_set_map_host(ctx.h, _ptr(c), C.c_size_t(len(c)), _ptr(s), C.c_size_t(len(s))))
        self.tot = (int(total_ref_corner), int(total_ref_surf))
        self.lo = _f32(lo); self.hi = _f32(hi)
        self.reduce_fn = reduce_fn or allreduce_sums

    def scanMatchScan(self, corner, surf, transform):
        ctx = self.ctx
        c = _f32(corner, 4); s = _f32(surf, 4); p = _f32(transform).copy()
        ctx._check(self.L.cm_shard_begin_host(ctx.h, _ptr(c), C.c_size_t(len(c)), _ptr(s), C.c_size_t(len(s)), _ptr(p),
                                              C.c_size_t(self.tot[0]), C.c_size_t(self.tot[1])))
        st = MatchStats(); done = C.c_int(0); sums = np.zeros(32, np.float64)
        if self.tot[0] < 50 or self.tot[1] < 100:          # "reference cloud points too few.", ScanMatch.cpp:57-61
            return False, p, dict(status=1, iterations=0, converged=False)
        for it in range(ctx.cfg.max_iterations):
            ctx._check(self.L.cm_shard_partial_host(ctx.h, C.c_int(it), _ptr(self.lo), _ptr(self.hi), _ptr(sums)))
            total = np.ascontiguousarray(self.reduce_fn(sums.copy()), np.float64)
            ctx._check(self.L.cm_shard_solve_host(ctx.h, C.c_int(it), _ptr(total), _ptr(p), C.byref(done), C.byref(st)))
            if done.value:
                break
        stats = dict(status=st.status, ret=bool(st.ret), converged=bool(st.converged), degenerate=bool(st.degenerate),
                     iterations=st.iterations, rows=st.rows, line=st.line_matches, plane=st.plane_matches, score=st.score)
        return bool(st.ret), p, stats
