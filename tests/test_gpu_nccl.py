"""2-GPU NCCL run of the merge step (skipped on a 1-GPU box): one process per GPU, host-sharded ingest, gyeeta_b200.dist.merge_global
over NCCL, and the merged logical-service answers compared with the CPU oracle over the whole stream."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from gyeeta_b200 import dist as gd
    from gyeeta_b200 import engine as ge
    from gyeeta_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(5)
        nsvc = 256
        ev = synth.gen_mixed(rng, 120_000, nsvc, ntask=16, nhosts=32, nclients=3000)
        eng = ge.Engine(device=rank, max_svcs=1024, max_tasks=64, max_batch=1 << 17, cms_log2_width=14, rank=rank, world=world)
        eng.ingest_events(ev); eng.sync(); eng.flush(5)
        ids = synth.service_ids(nsvc)
        logical = np.arange(nsvc, dtype=np.uint64) // np.uint64(16) + np.uint64(7000)
        eng.set_logical_map(ids, logical)
        ms = gd.merge_global(eng, torch, dist, torch.device("cuda", rank))
        out = eng.query_logical(np.unique(logical))
        keys = np.unique(ev["flow_key"][(ev["type"] >= 1) & (ev["type"] <= 4)])[:500]
        flows = eng.query_flows_global(keys, last_window=True)
        if rank == 0:
            from oracle import pyoracle as po
            orc = po.OracleEngine(max_svcs=1024, max_tasks=64, cms_log2_width=14)
            orc.ingest(ev); orc.flush(5)
            ok = True
            for o, lid in zip(out, np.unique(logical)):
                members = ids[logical == lid]
                hs = [orc.export_hist(int(m), 1) for m in members]
                tot = sum(h[1] for h in hs if h is not None)
                ok &= (o["nqrys_5s"] == tot)
                regs = np.zeros(4096, dtype=np.uint8)
                for m in members:
                    r = orc.export_hll(int(m))
                    if r is not None:
                        regs = np.maximum(regs, r)
                ok &= (o["distinct_clients"] == po.lib().gyo_hll_estimate(po._p(regs), 12))
            tbl = orc.cms(last_window=True).reshape(4, -1)
            for k, f in zip(keys[:100], flows[:100]):
                cells = [tbl[r, po.lib().gyo_cms_index(int(k), r, 14)] for r in range(4)]
                ok &= (f["count"] == min(int(c) & 0xFFFFFFFF for c in cells))
            q.put((bool(ok), float(ms)))
    finally:
        dist.destroy_process_group()


def test_nccl_two_gpu_merge():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    ok, ms = q.get(timeout=5)
    assert ok is True
