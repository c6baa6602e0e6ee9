"""World-size-2 gloo tests of the host-side multi-GPU logic (stream sharding, map slabs + halo, the per-iteration
all-reduce of the normal-equation partial sums, max-over-ranks timing).  No GPU needed."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

dmod = importlib.import_module("the-cooper-mapper_b200.dist")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)                         # same data on every rank
    pts = np.zeros((5000, 4), np.float32); pts[:, :3] = rng.uniform(-60, 60, (5000, 3))
    q = rng.uniform(-60, 60, (800, 3)).astype(np.float32)
    bounds = dmod.slab_bounds(pts, world)
    shard = dmod.shard_cloud(pts, bounds, rank)
    lo, hi = dmod.own_box(bounds, rank)
    mine = (q[:, 0] >= lo[0]) & (q[:, 0] < hi[0])
    # every point within sqrt(5) of an owned query is in this rank's shard (exactness of the sharded 5-NN)
    shard_set = set(map(bytes, shard))
    for qq in q[mine][:200]:
        d2 = ((pts[:, :3] - qq) ** 2).sum(1)
        for p in pts[d2 < 5.0]:
            assert bytes(p) in shard_set
    owned = torch.tensor([int(mine.sum())]); dist.all_reduce(owned)
    # partial "normal equation" sums: each rank sums its own rows; the all-reduce equals the global sum
    rows = rng.normal(size=(800, 32))
    part = rows[mine].sum(0)
    tot = dmod.allreduce_sums(part)
    # timing aggregation: the bench takes the MAX over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    streams = dmod.stream_shard(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, streams)
    if rank == 0:
        out.put(dict(owned=int(owned.item()), tot=tot, ref=rows.sum(0), tmax=float(t.item()), streams=gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["owned"] == 800                                     # ownership is a partition of the queries
    assert np.allclose(res["tot"], res["ref"], rtol=1e-12, atol=1e-9)
    assert res["tmax"] == 11.0
    flat = sorted(i for s in res["streams"] for i in s)
    assert flat == list(range(11)) and res["streams"][0] == [0, 2, 4, 6, 8, 10]


def test_slab_bounds_edge_cases():
    pts = np.zeros((10, 4), np.float32); pts[:, 0] = np.arange(10)
    b = dmod.slab_bounds(pts, 1)
    assert len(b) == 2 and b[0] < -1e30 and b[1] > 1e30
    b = dmod.slab_bounds(pts, 4)
    assert np.all(np.diff(b) >= 0) and len(b) == 5
    assert sum(len(dmod.shard_cloud(pts, b, r, halo=0.0)) for r in range(4)) == 10
    assert len(dmod.slab_bounds(np.zeros((0, 4), np.float32), 3)) == 4
