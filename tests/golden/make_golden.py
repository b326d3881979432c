"""Generates tests/golden/hist_golden.npz and jhash_golden.npz by RUNNING THE REFERENCE's own code
(oracle/_ref/libgyref.so, compiled from /root/reference by oracle/Makefile). Run only in the build
container, where /root/reference exists:  python tests/golden/make_golden.py
The fixtures pin oracle/gysk_oracle.c on machines that have no reference tree (the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PCTS = np.array([25, 50, 75, 90, 95, 99, 99.999, 0.001, 100], dtype=np.float32)


def inputs_for(cls_name, rng):
    edge = np.array([-(2 ** 40), -(2 ** 31) - 1, -(2 ** 31), -1000, -16, -15, -3, -2, -1, 0, 1, 2, 5, 8, 9, 10, 13, 14, 26, 27,
                     99, 100, 101, 250, 251, 3000, 3001, 5000, 5001, 15000, 15001, 65000, 65001, 150000, 150001,
                     5000000, 5000001, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 2 ** 32, 2 ** 32 + 7, 2 ** 40], dtype=np.int64)
    ln = np.round(np.exp(rng.normal(np.log(20), 1.2, 4000))).astype(np.int64)
    uni = rng.integers(-50, 200000, 2000, dtype=np.int64)
    big = rng.integers(0, 2 ** 33, 500, dtype=np.int64)
    return np.concatenate([edge, ln, uni, big])


def main():
    R = po.ref()
    assert R is not None, "reference library not built"
    rng = np.random.default_rng(20260922)
    out = {}
    for name, cls in po.CLS.items():
        tkinds = {"FD_I8_9_26_5": [po.T_INT8], "FD_INT_M15_M3_4": [po.T_INT]}.get(name, [po.T_INT64, po.T_INT])
        for tk in tkinds:
            vals = inputs_for(name, rng)
            if tk == po.T_INT8:
                vals = rng.integers(-128, 128, 3000, dtype=np.int64)
            r = po.hist_run(R, "gyref_hist_run", cls, tk, vals, PCTS)
            key = f"{name}__{tk}"
            out[key + "__vals"] = vals
            out[key + "__buckets"] = r["buckets"].astype(np.int16)
            out[key + "__count"] = r["stats"]["count"]
            out[key + "__sum"] = r["stats"]["sum"]
            out[key + "__total_max"] = np.array([r["total"], r["max"]], dtype=np.int64)
            out[key + "__pct"] = r["pct"]
            out[key + "__avg"] = np.array([r["avg"]], dtype=np.float32)
    # percentile cut-off uses a float multiplier on size_t (gy_statistics.h:753): exercise large counts
    for total_pow in (24, 25, 31, 40):
        stats = np.zeros(16, dtype=po.SERIAL_DTYPE)
        base = (1 << total_pow) // 15
        stats["count"][:15] = base + np.arange(15) * 3 + 1
        stats["sum"][:15] = stats["count"][:15] * 7
        total = int(stats["count"].sum())
        pct = np.zeros(len(PCTS), dtype=np.int64)
        avg = po.C.c_float()
        R.gyref_hist_pct_from_serial(0, 0, po._p(stats), total, 12345, po._p(PCTS), len(PCTS), po._p(pct), po.C.byref(avg))
        out[f"bigcount_{total_pow}__count"] = stats["count"].copy()
        out[f"bigcount_{total_pow}__sum"] = stats["sum"].copy()
        out[f"bigcount_{total_pow}__pct"] = pct
        out[f"bigcount_{total_pow}__avg"] = np.array([avg.value], dtype=np.float32)
    out["pcts"] = PCTS
    np.savez_compressed(os.path.join(HERE, "hist_golden.npz"), **out)

    keys = np.concatenate([np.array([0, 1, 42, 2 ** 32 - 1, 2 ** 32, 2 ** 63, 2 ** 64 - 1], dtype=np.uint64),
                           rng.integers(0, 2 ** 64, 4096, dtype=np.uint64)])
    h64 = np.array([R.gyref_uint64_hash(int(k)) for k in keys], dtype=np.uint32)
    seeds = rng.integers(0, 2 ** 32, len(keys), dtype=np.uint32)
    h2w = np.array([R.gyref_jhash_2words(int(k & np.uint64(0xFFFFFFFF)), int(k >> np.uint64(32)), int(s))
                    for k, s in zip(keys, seeds)], dtype=np.uint32)
    blob = rng.integers(0, 256, 64, dtype=np.uint8)
    hbytes = np.array([R.gyref_jhash(po._p(blob), n, 0xceedfead) for n in range(0, 41)], dtype=np.uint32)
    words = rng.integers(0, 2 ** 32, 16, dtype=np.uint32)
    hwords = np.array([R.gyref_jhash2(po._p(words), n, 0xceedfead) for n in range(0, 13)], dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "jhash_golden.npz"), keys=keys, h64=h64, seeds=seeds, h2w=h2w, blob=blob,
                        hbytes=hbytes, words=words, hwords=hwords)
    print("golden fixtures written:", len(out), "hist arrays")


if __name__ == "__main__":
    main()
