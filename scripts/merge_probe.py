"""one merge step (fold -> finish) on one GPU at the bench's shape, for an ncu launch list: python scripts/merge_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gyeeta_b200 import engine as ge  # noqa: E402

dev = torch.device("cuda", 0)
n = 50_000_000
eng = ge.Engine(device=0, max_svcs=1 << 17, max_tasks=1 << 15, max_batch=(1 << 27) - 1, stage_batch=1 << 23)
ev = bench.gen_events_gpu(torch, n, 1234, 0, 1, dev)
torch.cuda.synchronize()
eng.ingest_device_ptr(ev.data_ptr(), n)
eng.sync()
eng.flush(5)
ids = bench.rank_service_ids(0)
eng.set_logical_map(ids, np.arange(bench.NSVC, dtype=np.uint64) // np.uint64(16) + np.uint64(1))
for _ in range(3):
    eng.merge_prepare()
    eng.merge_finish(None, 1)
eng.sync()
print("ok")
