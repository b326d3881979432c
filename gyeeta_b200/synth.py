"""Seeded synthetic event streams for the BASELINE.json configs. numpy only — used by tests and by
bench.py for the host-side sample; bench.py generates its large device-resident batches with the same formulas on the GPU."""
import numpy as np

from .engine import EVENT_DTYPE, EV_ACCEPT, EV_CLOSE_SER, EV_CONNECT, EV_RESP, EV_TASK


def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def zipf_cdf(n, s):
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def service_ids(n):
    """ids = splitmix64(i), never 0"""
    ids = splitmix64(np.arange(1, n + 1, dtype=np.uint64))
    ids[ids == 0] = 1
    return ids


def task_ids(n):
    ids = splitmix64(np.arange(1, n + 1, dtype=np.uint64) + np.uint64(1 << 40))
    ids[ids == 0] = 1
    return ids


def gen_tcp(rng, n, nsvc, zipf_s=1.1, nclients=1_000_000, nhosts=512):
    """config 2: TCP conn events; svc ~ Zipf, each client bound to <= 8 services, bytes ~ LogNormal(ln 4096, 2)"""
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    rank = np.searchsorted(zipf_cdf(nsvc, zipf_s), rng.random(n), side="left").astype(np.uint64)
    ids = service_ids(nsvc)
    ev["svc_id"] = ids[rank]
    # a client key is a function of (service rank, one of 8 per-service client groups, draw) so clients stay bound to few services
    cli = rng.integers(0, max(nclients // 8, 1), n, dtype=np.uint64) * np.uint64(8) + (rank % np.uint64(8))
    ev["flow_key"] = splitmix64(cli + np.uint64(1 << 48))
    ev["value"] = np.minimum(np.exp(rng.normal(np.log(4096.0), 2.0, n)), 4.0e9).astype(np.uint32)
    u = rng.random(n)
    ev["type"] = np.where(u < 0.45, EV_ACCEPT, np.where(u < 0.90, EV_CLOSE_SER, EV_CONNECT)).astype(np.uint16)
    ev["host_idx"] = (rank % np.uint64(nhosts)).astype(np.uint32)
    ev["tsec"] = 1
    return ev


def gen_mixed(rng, n, nsvc, ntask=None, zipf_s=1.05, nhosts=4096, nclients=1_000_000):
    """config 3/4/5: 70 % RESP (usec ~ LogNormal(ln 2000, 1.5)), 20 % TCP as config 2, 10 % TASK"""
    ntask = ntask or max(nsvc // 4, 1)
    ev = gen_tcp(rng, n, nsvc, zipf_s=zipf_s, nclients=nclients, nhosts=nhosts)
    u = rng.random(n)
    resp = u < 0.70
    task = u >= 0.90
    nr, nt = int(resp.sum()), int(task.sum())
    ev["type"][resp] = EV_RESP
    ev["value"][resp] = np.minimum(np.exp(rng.normal(np.log(2000.0), 1.5, nr)), 9.0e8).astype(np.uint32)
    tids = task_ids(ntask)
    trank = np.searchsorted(zipf_cdf(ntask, zipf_s), rng.random(nt), side="left")
    ev["type"][task] = EV_TASK
    ev["svc_id"][task] = tids[trank]
    ev["value"][task] = np.minimum(rng.gamma(1.2, 20.0, nt), 3200).astype(np.uint32)          # cpu pct
    cpu_delay = np.minimum(np.exp(rng.normal(np.log(30.0), 2.0, nt)), 1.0e5).astype(np.uint64)
    blkio = np.minimum(np.exp(rng.normal(np.log(5.0), 2.5, nt)), 1.0e5).astype(np.uint64)
    ev["flow_key"][task] = cpu_delay | (blkio << np.uint64(32))
    return ev


def gen_resp_config1(rng, n=1_000_000, svc_id=None):
    """config 1: one service, msec ~ round(LogNormal(ln 20, 1.2)) clipped to [0, 100000]; value field carries usec"""
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    ms = np.clip(np.round(np.exp(rng.normal(np.log(20.0), 1.2, n))), 0, 100000).astype(np.uint64)
    ev["svc_id"] = svc_id if svc_id is not None else service_ids(1)[0]
    ev["flow_key"] = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    ev["value"] = (ms * np.uint64(1000) + rng.integers(0, 1000, n, dtype=np.uint64)).astype(np.uint32)
    ev["type"] = EV_RESP
    ev["tsec"] = 1
    return ev


def shard_by_host(ev, world):
    """the partition of SURVEY.md §8e: events of host_idx % world == rank belong to rank"""
    return [ev[(ev["host_idx"] % world) == r] for r in range(world)]
