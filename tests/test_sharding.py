"""CPU tests of the tensor-parallel sharding logic with a real world_size-2 gloo group:
column-parallel shards concatenate, row-parallel shards all-reduce, both to the unsharded oracle result."""
import os
import socket

import numpy as np
import pytest

from effort_b200 import sharding
from oracle import oracle as O
from tests.util import make_v, make_w, rel_err

IN, OUT, EFFORT = 4096, 4096, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = make_w(OUT, IN, 1234)
        t = O.bucketize(w)
        v = make_v(IN, 42)
        full = O.bucket_mul(v, t["buckets"], t["bucket.stats"], t["probes"], IN, OUT, EFFORT)
        cutoff = full["cutoff"]
        # ---- column-parallel: same cutoff/selection on every rank, disjoint output slices, all_gather ----
        sc = sharding.shard_columns(t, IN, OUT, rank, world)
        rc = O.bucket_mul(v, sc["buckets"], sc["bucket.stats"], sc["probes"], IN, sc["out"], EFFORT)
        assert rc["n_selected"] == full["n_selected"] and rc["cutoff"] == cutoff
        parts = [torch.zeros(sc["out"], dtype=torch.float32) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(rc["out32"]))
        col = torch.cat(parts).numpy()
        # ---- row-parallel: local rows + local slice of v, cutoff from the gathered first 4096 dims of v ----
        sr = sharding.shard_rows(t, IN, OUT, rank, world)
        v_loc = torch.from_numpy(v[sr["in_offset"]: sr["in_offset"] + sr["in"]].copy())
        gathered = [torch.zeros(sr["in"], dtype=torch.float32) for _ in range(world)]
        dist.all_gather(gathered, v_loc)                       # the 16 KB all-gather of v[:4096]
        v_cut = torch.cat(gathered).numpy()[:4096]
        c2 = O.find_cutoff(v_cut, sr["probes"], EFFORT)
        assert c2 == cutoff
        disp = O.prepare_dispatch(v_loc.numpy(), sr["bucket.stats"], c2, sr["in"], OUT // 16, 16 * sr["in"])
        o32, _ = O.bucket_mul_dispatch(sr["buckets"], disp, OUT // 16)
        n_sel = torch.tensor([disp.shape[0]])
        part = torch.from_numpy(o32)
        dist.all_reduce(part)                                  # the all-reduce on the row-parallel output
        dist.all_reduce(n_sel)
        if rank == 0:
            q.put({"col_equal": bool(np.array_equal(col, full["out32"])),
                   "row_err": rel_err(part.numpy(), full["out64"]), "row_nsel": int(n_sel), "nsel": full["n_selected"]})
    finally:
        dist.destroy_process_group()


def test_tp2_gloo_matches_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = q.get(timeout=10)
    assert res["col_equal"]                       # column shards: bit-identical to the unsharded fp32 result
    assert res["row_nsel"] == res["nsel"]         # row shards select exactly the same rows...
    assert res["row_err"] <= 2e-6                 # ...and sum to the same vector (fp32 reorder only)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_shapes_and_coverage(world):
    w = make_w(1024, 4096, 5)
    t = O.bucketize(w)
    b = np.ascontiguousarray(t["buckets"]).view(np.uint16)
    cols = [np.ascontiguousarray(sharding.shard_columns(t, 4096, 1024, r, world)["buckets"]).view(np.uint16) for r in range(world)]
    assert np.array_equal(np.concatenate(cols, axis=1), b)
    rows = [sharding.shard_rows(t, 4096, 1024, r, world) for r in range(world)]
    rb = np.concatenate([np.ascontiguousarray(x["buckets"]).view(np.uint16).reshape(16, x["in"], 64) for x in rows], axis=1)
    assert np.array_equal(rb.reshape(16 * 4096, 64), b)
    assert sum(x["in"] for x in rows) == 4096 and [x["in_offset"] for x in rows] == [4096 * r // world for r in range(world)]
    with pytest.raises(ValueError):
        sharding.shard_columns(t, 4096, 1024, 0, 3)
