"""world_size-2 gloo test on CPU of the host-side logic of the multi-GPU path: the host-sharding rule and the collective recipe
(SUM over count-min / histogram arrays, MAX over HLL registers, all-gather + rank-ascending merge of t-digests). Each rank folds its
shard with the CPU oracle, the ranks exchange with torch.distributed (gloo), and rank 0 compares with a single oracle engine."""
import os
import socket

import numpy as np
import pytest

from gyeeta_b200 import dist as gd
from gyeeta_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(44)
        ev = synth.gen_mixed(rng, 40_000, 120, ntask=8, nhosts=16, nclients=2000)
        mine = ev[gd.shard_of_host(ev["host_idx"], world) == rank]
        eng = po.OracleEngine(max_svcs=256, max_tasks=32, cms_log2_width=10)
        eng.ingest(mine)
        ids = synth.service_ids(120)
        # SUM: count-min + histograms (viewed as int64: two's complement sums are the u64 sums)
        cms = torch.from_numpy(eng.cms().view(np.int64).copy())
        hist = np.zeros((len(ids), 15, 2), dtype=np.int64)
        hll = np.zeros((len(ids), 4096), dtype=np.uint8)
        for i, id_ in enumerate(ids):
            h = eng.export_hist(int(id_), 0)
            if h is not None:
                hist[i, :, 0] = h[0]["count"].view(np.int64); hist[i, :, 1] = h[0]["sum"]
                hll[i] = eng.export_hll(int(id_))
        th, tl = torch.from_numpy(hist), torch.from_numpy(hll)
        dist.all_reduce(cms, op=dist.ReduceOp.SUM)
        dist.all_reduce(th, op=dist.ReduceOp.SUM)
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        if rank == 0:
            one = po.OracleEngine(max_svcs=256, max_tasks=32, cms_log2_width=10)
            one.ingest(ev)
            ok = np.array_equal(cms.numpy().view(np.uint64), one.cms())
            for i, id_ in enumerate(ids):
                h = one.export_hist(int(id_), 0)
                if h is None:
                    ok &= not th[i].any().item()
                    continue
                ok &= np.array_equal(th[i, :, 0].numpy().view(np.uint64), h[0]["count"]) and np.array_equal(th[i, :, 1].numpy(), h[0]["sum"])
                ok &= np.array_equal(tl[i].numpy(), one.export_hll(int(id_)))
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shard_and_merge():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_rule():
    h = np.arange(100)
    for w in (1, 2, 4, 8):
        s = gd.shard_of_host(h, w)
        assert s.min() == 0 and s.max() == w - 1 and np.all(s == h % w)
