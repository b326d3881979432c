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


class DeviceSynth:
    """ctypes binding of libgysynth.so (csrc/gysk_synth.cu): the on-device Philox event source of the sustained-stream run. The id and
    CDF arrays are torch device tensors kept alive by this object; fill() enqueues one generator launch on `stream`."""

    def __init__(self, torch, dev, svc_ids, task_ids, zipf_s, rank=0, world=1, nhosts=4096, nclients=1_000_000, seed=1,
                 tail_start=0, churn_groups=0, churn_epoch=1, resp_mu=float(np.log(2000.0)), resp_sigma=1.5):
        import ctypes as C
        import os

        class Params(C.Structure):
            _fields_ = [("seed", C.c_uint64), ("rank", C.c_uint32), ("world", C.c_uint32), ("nsvc", C.c_uint32), ("ntask", C.c_uint32),
                        ("d_svc_ids", C.c_void_p), ("d_task_ids", C.c_void_p), ("d_cdf_svc", C.c_void_p), ("d_cdf_task", C.c_void_p),
                        ("nhosts", C.c_uint32), ("nclients", C.c_uint32), ("tsec", C.c_uint32), ("tail_start", C.c_uint32),
                        ("churn_groups", C.c_uint32), ("churn_epoch", C.c_uint32), ("window", C.c_uint32), ("resp_mu", C.c_float),
                        ("resp_sigma", C.c_float)]

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgysynth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -m gyeeta_b200.build`")
        self.L = C.CDLL(path)
        self.L.gysyn_fill.restype = C.c_int
        self.L.gysyn_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(Params), C.c_void_p]
        self._keep = [torch.from_numpy(np.ascontiguousarray(svc_ids).view(np.int64)).to(dev),
                      torch.from_numpy(np.ascontiguousarray(task_ids).view(np.int64)).to(dev),
                      torch.from_numpy(zipf_cdf(len(svc_ids), zipf_s)).to(dev), torch.from_numpy(zipf_cdf(len(task_ids), zipf_s)).to(dev)]
        self.p = Params(seed=seed, rank=rank, world=world, nsvc=len(svc_ids), ntask=len(task_ids), d_svc_ids=self._keep[0].data_ptr(),
                        d_task_ids=self._keep[1].data_ptr(), d_cdf_svc=self._keep[2].data_ptr(), d_cdf_task=self._keep[3].data_ptr(),
                        nhosts=nhosts, nclients=nclients, tsec=0, tail_start=tail_start, churn_groups=churn_groups, churn_epoch=churn_epoch,
                        window=0, resp_mu=resp_mu, resp_sigma=resp_sigma)
        self._C = C

    def fill(self, dptr, n, counter_base, stream, window=0, tsec=0):
        self.p.window, self.p.tsec = window, tsec
        rc = self.L.gysyn_fill(self._C.c_void_p(dptr), n, counter_base, self._C.byref(self.p), self._C.c_void_p(stream))
        if rc:
            raise RuntimeError(f"gysyn_fill failed: {rc}")
