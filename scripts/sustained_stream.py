"""BASELINE.json configs[4] (SURVEY.md §8d "Config 5"): a sustained stream at the rate of 1 B events/s over 5-s windows — 5 B events per
window over all GPUs — 1 M services Zipf(1.0) in total, idle-service eviction after 300 s, one sketch merge and a replay of a fixed
list of 10 K global (logical-service) queries per window. Events come from the on-device Philox source (libgysynth.so), so the host
link is not in the way; the run is time-compressed: window w is stamped tsec = T0 + 5 w and processed as fast as the GPUs go, and the
line reports how many times faster than the 1 B events/s real-time rate that was.

    python scripts/sustained_stream.py [--windows 180] [--events-per-window 5e9] [--services 1000000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sustained_stream.py

Retention: the engine mirrors the reference's levels (common/gy_statistics.h:1105,1548): a 300-s level of 10 ring slots (30 s each) and a
5-day level next to the open / last 5-s windows, not 60 separate 5-s slots. Service churn: the upper half of the Zipf ranks is split
into 5 groups, each alive for 30 windows (150 s) out of 150, so a silent group passes the 300-s idle limit, is evicted once its listeners
are 600 s old (gy_socket_stat.cc:3968-3982) — group 0 at window 120 — and registers again into recycled slots when it comes back (window
150); the default 180 windows = 15 min of stream cover one full cycle."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gyeeta_b200 import dist as gd  # noqa: E402
from gyeeta_b200 import engine as ge  # noqa: E402
from gyeeta_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=180)
ap.add_argument("--events-per-window", type=float, default=5e9, help="over all GPUs: 5 s at 1 B events/s")
ap.add_argument("--services", type=int, default=1_000_000, help="over all GPUs")
ap.add_argument("--batch", type=int, default=100_000_000)
ap.add_argument("--queries", type=int, default=10_000)
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)

nsvc = args.services // world
bench.NSVC = nsvc
ids = bench.rank_service_ids(rank)
task_ids = synth.splitmix64(np.arange(1, bench.NTASK + 1, dtype=np.uint64) + np.uint64((1 << 40) + rank * bench.NTASK))
per_rank = int(args.events_per_window) // world
free0 = torch.cuda.mem_get_info()[0]
cap = 1 << int(np.ceil(np.log2(nsvc * 1.3)))
eng = ge.Engine(device=local, max_svcs=cap, max_tasks=1 << 15, max_batch=(1 << 27) - 1, idle_evict_secs=300, rank=rank, world=world)
# logical service = 16 instances spread over the ranks (global index g = local index * world + rank; logical id = g // 16)
ids_all = np.concatenate([bench.rank_service_ids(r) for r in range(world)])
g_all = np.concatenate([np.arange(nsvc, dtype=np.uint64) * np.uint64(world) + np.uint64(r) for r in range(world)])
eng.set_logical_map(ids_all, g_all // np.uint64(16) + np.uint64(1))
if world > 1:
    gd.nccl_comm_init(eng, dist)
held = free0 - torch.cuda.mem_get_info()[0]
src = synth.DeviceSynth(torch, dev, ids, task_ids, 1.0, rank=rank, world=world, nhosts=bench.NHOSTS, nclients=bench.NCLIENTS, seed=20240 + rank,
                        tail_start=nsvc // 2, churn_groups=5, churn_epoch=30)
buf = torch.empty((args.batch, 4), dtype=torch.int64, device=dev)
stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
rng = np.random.default_rng(5)
nlogical = int(g_all.max() // 16 + 1)
lids = (rng.choice(nlogical, size=min(args.queries, nlogical), replace=False).astype(np.uint64) + np.uint64(1))
qout = (ge.SvcSummary * len(lids))()
qids_p = lids.ctypes.data_as(C.c_void_p)


def ev_pair():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    return a, b


# generator cost alone (one batch), to be read next to the window times
g0, g1 = ev_pair()
with torch.cuda.stream(stream):
    g0.record()
src.fill(buf.data_ptr(), args.batch, 1 << 60, eng.stream())
with torch.cuda.stream(stream):
    g1.record()
eng.sync()
gen_ms_per_batch = g0.elapsed_time(g1)

if world > 1:
    dist.barrier()
torch.cuda.synchronize()
T0 = 1_700_000_000
rows = []
counter = 0
wall0 = time.perf_counter()
for w in range(args.windows):
    tsec = T0 + 5 * (w + 1)
    a, b = ev_pair(); c, d = ev_pair()
    with torch.cuda.stream(stream):
        a.record()
    left = per_rank
    while left > 0:
        m = min(left, args.batch)
        src.fill(buf.data_ptr(), m, counter, eng.stream(), window=w, tsec=tsec)
        eng.ingest_device_ptr(buf.data_ptr(), m)
        counter += m; left -= m
    with torch.cuda.stream(stream):
        b.record()
    eng.flush(tsec)
    with torch.cuda.stream(stream):
        c.record()
    if world > 1:
        eng.merge_global()
    else:
        eng.merge_prepare(); eng.merge_finish(None, 1)
    with torch.cuda.stream(stream):
        d.record()
    tq = time.perf_counter()
    rc = eng.L.gysk_query_logical(eng.h, qids_p, len(lids), qout)
    assert rc == 0, rc
    tq = (time.perf_counter() - tq) * 1e3                       # includes waiting for the window's device work (the call syncs)
    nev = len(eng.evicted_ids(1 << 20)) if w % 5 == 4 or w == args.windows - 1 else -1
    rows.append((a, b, c, d, tq, nev))
eng.sync()
torch.cuda.synchronize()
wall = time.perf_counter() - wall0
tw = torch.tensor([wall], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
wall = float(tw.item())
st = eng.stats()
mine = {"ingest_ms": [r[0].elapsed_time(r[1]) for r in rows], "flush_ms": [r[1].elapsed_time(r[2]) for r in rows],
        "merge_ms": [r[2].elapsed_time(r[3]) for r in rows], "query_ms_wall": [r[4] for r in rows]}
found = int(sum(1 for o in qout if o.found))
stats_t = torch.tensor([st["nsvcs"], st["svcs_evicted"], st["events_dropped"], st["events_in"]], device=dev, dtype=torch.int64)
if world > 1:
    dist.all_reduce(stats_t)
if rank == 0:
    def pct(v, q):
        return float(np.percentile(np.array(v), q))
    total_events = per_rank * world * args.windows
    stream_secs = 5.0 * args.windows
    out = {"config": "BASELINE configs[4]: sustained stream, time-compressed", "n_gpus": world, "windows": args.windows,
           "stream_seconds_covered": stream_secs, "events_per_window": per_rank * world, "services_total": nsvc * world,
           "events_total": total_events, "wall_s": wall, "events_per_s_sustained": total_events / wall,
           "times_faster_than_1B_per_s_realtime": (total_events / wall) / 1e9,
           "includes": "on-device Philox generation, ingest + sort + merge chain, 5-s flush (window roll, rolling levels, idle eviction), "
                       "one gysk_merge_global per window, 10 K logical-service queries per window (host wall, rank 0 shown)",
           "generator_ms_per_100M_events": gen_ms_per_batch * 1e8 / args.batch,
           "rank0_window_ms": {k: {"p50": pct(v, 50), "p95": pct(v, 95), "max": float(max(v))} for k, v in mine.items()},
           "queries_per_window": len(lids), "queries_found_last_window": found,
           "live_services_end": int(stats_t[0]), "services_evicted_total": int(stats_t[1]), "events_dropped": int(stats_t[2]),
           "events_in": int(stats_t[3]), "engine_device_bytes_rank0": int(held),
           "evicted_per_sampled_window_rank0": [r[5] for r in rows if r[5] >= 0]}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
