"""BASELINE.json configs[4] on ONE GPU's share (python scripts/sustained_probe.py [windows]): 1 M services (Zipf 1.0), one window =
one device batch of 100 M mixed events, then the 5-s tick (flush: window roll, rolling levels, idle eviction, task windows) and a
fixed list of 10 K service queries replayed every window. Events are generated on the device; prints one JSON line with the
per-window device times (CUDA events on the engine's stream) and the memory the engine holds."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gyeeta_b200 import engine as ge  # noqa: E402

NW = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
bench.NSVC, bench.ZIPF_S = 1_000_000, 1.0
n = 100_000_000
free0 = torch.cuda.mem_get_info()[0]
eng = ge.Engine(device=0, max_svcs=1 << 20, max_tasks=1 << 15, max_batch=(1 << 27) - 1, stage_batch=1 << 22, idle_evict_secs=300)
held = free0 - torch.cuda.mem_get_info()[0]
ev = bench.gen_events_gpu(torch, n, 77, 0, 1, dev)
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
qids = bench.rank_service_ids(0)[:10_000].copy()
rows = []
tsec = 1000
for w in range(NW):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with torch.cuda.stream(stream):
        e[0].record()
    eng.ingest_device_ptr(ev.data_ptr(), n)
    with torch.cuda.stream(stream):
        e[1].record()
    tsec += 5 if w != NW - 2 else 700            # one jump past the idle limit: everything silent since then would go
    eng.flush(tsec)
    with torch.cuda.stream(stream):
        e[2].record()
    t0 = time.perf_counter()
    sm = eng.query_svcs(qids)
    tq = (time.perf_counter() - t0) * 1e3
    rows.append({"ingest_chain_ms": e[0].elapsed_time(e[1]), "flush_ms": e[1].elapsed_time(e[2]), "query_10k_ms_wall": tq,
                 "found": int(sum(s["found"] for s in sm)), "evicted": int(len(eng.evicted_ids(1 << 20)))})
st = eng.stats()
print(json.dumps({"services": bench.NSVC, "events_per_window": n, "engine_device_bytes": int(held), "windows": rows,
                  "nsvcs": st["nsvcs"], "svcs_evicted": st["svcs_evicted"], "events_dropped": st["events_dropped"]}))
