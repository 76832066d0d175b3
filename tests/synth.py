"""Seeded synthetic hit matrices (numpy) shaped like the BASELINE configs, for kernel-level tests."""
import numpy as np


def random_matrix(N, M, avg_deg, seed=0, zipf=False, family=True, tiny_frac=0.02, zero_rows=0.0):
    """CSR (row_ptr, sid, conprb, ncpv).  Rows hit consecutive transcript ids (isoform families are
    contiguous in RSEM's numbering), conprb spans 1e-60..1e-3, a few entries are < 1e-300 (clamped)."""
    rng = np.random.default_rng(seed)
    if zipf:
        k = np.arange(1, 201)
        p = k ** -1.1
        p /= p.sum()
        deg = rng.choice(k, size=N, p=p)
    else:
        deg = 1 + rng.poisson(max(avg_deg - 1, 0), size=N)
    deg = np.minimum(deg, M)
    if zero_rows > 0:
        deg[rng.random(N) < zero_rows] = 0
    row_ptr = np.zeros(N + 1, np.uint64)
    row_ptr[1:] = np.cumsum(deg)
    H = int(row_ptr[-1])
    start = rng.integers(1, M + 1, size=N)
    start = np.minimum(start, M - deg + 1).clip(1)
    rows = np.repeat(np.arange(N), deg)
    within = np.arange(H) - np.repeat(row_ptr[:-1].astype(np.int64), deg)
    if family:
        sid = (start[rows] + within).astype(np.int32)
    else:
        sid = rng.integers(1, M + 1, size=H).astype(np.int32)
    sign = np.where(rng.random(H) < 0.5, 1, -1).astype(np.int32)
    sid = sid * sign
    conprb = 10.0 ** rng.uniform(-60, -3, size=H)
    tiny = rng.random(H) < tiny_frac
    conprb[tiny] = 10.0 ** rng.uniform(-320, -295, size=int(tiny.sum()))
    conprb[rng.random(H) < 0.01] = 0.0
    ncpv = 10.0 ** rng.uniform(-80, -40, size=N)
    ncpv[rng.random(N) < 0.05] = 0.0
    return row_ptr, sid, conprb, ncpv


def init_theta(M, n0, n_tot):
    """EM.cpp:342-346"""
    theta = np.empty(M + 1)
    theta[0] = max(n0 / n_tot, 1e-8)
    theta[1:] = (1.0 - theta[0]) / M
    return theta
