#!/usr/bin/env python
"""Equivalence classes of a hit matrix: how many reads share exactly the same transcript list?

Evidence for DESIGN.md section 8 item 1(a) (dense per-class blocks for K2).  Works on a `.dat` file written by
rsem-parse-alignments / tools/gen_dataset, or on the synthetic C3-shaped matrix of bench.py at a reduced size with the
same reads-per-transcript ratio.

    python tools/class_stats.py --dat sample.temp/sample.dat
    python tools/class_stats.py --synthetic 500000 2000 20
"""
import argparse
import hashlib
import sys

import numpy as np


def classes_from_rows(row_ptr, sid):
    """returns the class size of every row (rows with identical |sid| lists form a class)"""
    keys = {}
    size = np.zeros(len(row_ptr) - 1, np.int64)
    ids = np.empty(len(row_ptr) - 1, np.int64)
    a = np.abs(sid).astype(np.int32)
    for i in range(len(row_ptr) - 1):
        k = hashlib.blake2b(a[row_ptr[i]:row_ptr[i + 1]].tobytes(), digest_size=12).digest()
        ids[i] = keys.setdefault(k, len(keys))
    counts = np.bincount(ids)
    size[:] = counts[ids]
    return size, counts


def load_dat(path):
    with open(path) as f:
        n1, nh, rt = (int(x) for x in f.readline().split())
        paired = rt >= 2
        row_ptr, sid = [0], []
        for line in f:
            t = line.split()
            k = int(t[0])
            step = 3 if paired else 2
            sid += [int(x) for x in t[1:1 + k * step:step]]
            row_ptr.append(len(sid))
    return np.array(row_ptr, np.int64), np.array(sid, np.int32)


def synthetic(N, M, deg, seed=1234):
    rng = np.random.default_rng(seed)
    degs = np.minimum(1 + rng.poisson(deg - 1, N), M)
    row_ptr = np.concatenate([[0], np.cumsum(degs)])
    start = np.minimum(rng.integers(1, M + 1, N), M - degs + 1).clip(1)
    sid = (np.repeat(start, degs) + (np.arange(row_ptr[-1]) - np.repeat(row_ptr[:-1], degs))).astype(np.int32)
    return row_ptr, sid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dat")
    ap.add_argument("--synthetic", nargs=3, type=int, metavar=("N", "M", "DEG"))
    a = ap.parse_args()
    if a.dat:
        row_ptr, sid = load_dat(a.dat)
    elif a.synthetic:
        row_ptr, sid = synthetic(*a.synthetic)
    else:
        ap.error("--dat or --synthetic")
    deg = np.diff(row_ptr)
    size, counts = classes_from_rows(row_ptr, sid)
    H = int(row_ptr[-1])
    print(f"reads {len(deg)}  hits {H}  mean degree {deg.mean():.2f}  classes {len(counts)}")
    print(f"reads per class: mean over classes {counts.mean():.2f}, mean over reads {size.mean():.2f}, "
          f"median over reads {np.median(size):.0f}")
    for thr in (1, 2, 4, 8, 16, 32):
        frac_hits = deg[size >= thr].sum() / H
        print(f"  hits in classes of >= {thr:2d} reads: {100 * frac_hits:5.1f} %")
    # bytes per hit if classes of >= 4 reads are stored as dense blocks (ids once per class, 8 B per hit) and the
    # rest stays CSR (12 B per hit + 16 B per read)
    dense = size >= 4
    b_dense = 8 * deg[dense].sum() + 8 * dense.sum() + sum(4 * d for d in deg[dense] / size[dense])
    b_csr = 12 * deg[~dense].sum() + 16 * (~dense).sum()
    print(f"bytes per hit: CSR {(12 * H + 16 * len(deg)) / H:.2f} -> blocked {(b_dense + b_csr) / H:.2f}")
    print(f"reductions per hit: 1.00 -> {((deg[dense] / size[dense]).sum() + deg[~dense].sum()) / H:.3f}")


if __name__ == "__main__":
    sys.exit(main())
