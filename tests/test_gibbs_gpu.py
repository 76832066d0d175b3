"""K5/K6 through the C ABI against the CPU oracle: same seeds -> the same draws (count vectors bit-identical),
including matrices where reads move in and out of the noise transcript all the time (the case the
component-parallel sampler has to roll back and redo)."""
import ctypes as C
import os

import numpy as np
import pytest

import synth
from rsem_b200.capi import GibbsOut, GibbsParams

pytestmark = pytest.mark.gpu


def _ofg_like(N, M, deg, seed, noise_scale):
    """rows of (sid, conprb) with the noise entry first (Gibbs.cpp:119-131); noise_scale moves the noise weight
    from negligible to dominant so that membership changes are rare / frequent"""
    rng = np.random.default_rng(seed)
    row_ptr, sid, conprb, ncpv = synth.random_matrix(N, M, deg, seed=seed, tiny_frac=0.0)
    conprb = np.where(conprb <= 0, 1e-30, conprb)
    rows_sid, rows_val, rp = [], [], [0]
    for i in range(N):
        a, b = int(row_ptr[i]), int(row_ptr[i + 1])
        s, v = list(np.abs(sid[a:b])), list(conprb[a:b])
        if rng.random() < 0.9:
            s.insert(0, 0)
            v.insert(0, float(np.median(conprb[a:b]) * noise_scale * rng.uniform(0.1, 10)))
        rows_sid += s
        rows_val += v
        rp.append(len(rows_sid))
    return np.array(rp, np.uint64), np.array(rows_sid, np.int32), np.array(rows_val, np.float64)


def _run_gpu(ctx, rp, sid, val, M, n0, init, alpha, totc, eel, mw, genes, burnin, gap, samples, seeds):
    ctx.gibbs_upload(rp, sid, val, M)
    nc = len(samples)
    samples = np.ascontiguousarray(samples, np.int32)
    seeds = np.ascontiguousarray(seeds, np.uint32)
    init, genes = np.ascontiguousarray(init, np.int32), np.ascontiguousarray(genes, np.int32)
    alpha, eel, mw = (np.ascontiguousarray(x, np.float64) for x in (alpha, eel, mw))
    p = GibbsParams()
    p.M, p.burnin, p.gap, p.n_chains = M, burnin, gap, nc
    p.chain_samples = samples.ctypes.data_as(C.POINTER(C.c_int32))
    p.chain_seeds = seeds.ctypes.data_as(C.POINTER(C.c_uint32))
    p.n0, p.totc = n0, totc
    p.init_counts = init.ctypes.data_as(C.POINTER(C.c_int32))
    p.pseudo_counts = alpha.ctypes.data_as(C.POINTER(C.c_double))
    p.eel, p.mw = eel.ctypes.data_as(C.POINTER(C.c_double)), mw.ctypes.data_as(C.POINTER(C.c_double))
    p.n_genes = len(genes) - 1
    p.gene_start = genes.ctypes.data_as(C.POINTER(C.c_int32))
    total = int(samples.sum())
    cv = np.zeros((total, M + 1), np.int32)
    sums = [np.zeros(M + 1) for _ in range(4)] + [np.zeros(len(genes) - 1)]
    o = GibbsOut()
    o.count_vectors = cv.ctypes.data_as(C.POINTER(C.c_int32))
    o.sum_c, o.sum_c2, o.sum_tpm, o.sum_fpkm, o.sum_gene_c2 = (s.ctypes.data_as(C.POINTER(C.c_double)) for s in sums)
    ctx.gibbs_run(p, o)
    return cv, sums


@pytest.fixture(scope="module")
def ctx(built):
    import rsem_b200
    c = rsem_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("mode", ["parallel", "serial"])
@pytest.mark.parametrize("noise_scale,seed", [(1e-30, 3), (1.0, 10), (30.0, 213)])
def test_chains_match_oracle(ctx, oracle, mode, noise_scale, seed, monkeypatch):
    N, M = 3000, 150
    rp, sid, val = _ofg_like(N, M, 4, seed=seed, noise_scale=noise_scale)
    n0 = 40.0
    init = np.zeros(M + 1, np.int32)
    init[[5, 77]] = -1                                   # omitted transcripts never appear in a row
    keep = ~np.isin(sid, [5, 77])
    # drop omitted ids from the rows
    new_rp = [0]
    for i in range(N):
        new_rp.append(new_rp[-1] + int(keep[int(rp[i]):int(rp[i + 1])].sum()))
    rp, sid, val = np.array(new_rp, np.uint64), sid[keep], val[keep]
    if np.any(np.diff(rp.astype(np.int64)) == 0):
        pytest.skip("generator produced an empty row")
    alpha = np.full(M + 1, 0.7)
    totc = (M + 1 - 2) * 0.7 + n0 + N
    rng = np.random.default_rng(1)
    eel = np.concatenate([[0.0], rng.uniform(200, 2000, M)])
    mw = np.ones(M + 1)
    genes = np.arange(1, M + 2, 3, dtype=np.int32)
    genes[-1] = M + 1
    samples, burnin, gap = [3, 2, 2], 6, 2
    seeds = oracle.chain_seeds(2024, 3)
    monkeypatch.setenv("RSEM_B200_GIBBS", mode)
    cv, sums = _run_gpu(ctx, rp, sid, val, M, n0, init, alpha, totc, eel, mw, genes, burnin, gap, samples, seeds)
    at, ref_sums = 0, None
    for t, ns in enumerate(samples):
        cv_ref, s = oracle.gibbs_chain(rp, sid, val, M, n0, init, alpha, totc, eel, mw, genes, burnin, gap, ns, int(seeds[t]))
        assert np.array_equal(cv[at:at + ns], cv_ref), f"chain {t} differs"
        at += ns
        ref_sums = s if ref_sums is None else [x + y for x, y in zip(ref_sums, s)]
    for got, ref in zip(sums, ref_sums):
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)
    # the noise count really moves in the stressed cases (otherwise the roll-back path is not exercised)
    if noise_scale >= 1.0:
        assert len(set(cv[:, 0].tolist())) > 1


def _family_matrix(N, M, seed, max_family, junk):
    """disjoint isoform families (one connected component each; <= 31 transcripts: register-resident counts, more: counts
    in L2), reads hit a run of members of one family, the noise entry first; a few rows longer than 16 entries"""
    rng = np.random.default_rng(seed)
    sizes, tot = [], 0
    while tot < M:
        k = min(int(rng.integers(1, max_family + 1)), M - tot)
        sizes.append(k)
        tot += k
    sizes = np.array(sizes)
    start = np.concatenate([[1], 1 + np.cumsum(sizes)[:-1]])
    wgt = rng.random(len(sizes)) ** 3 + 1e-3
    fam = rng.choice(len(sizes), size=N, p=wgt / wgt.sum())
    rows_sid, rows_val, rp = [], [], [0]
    for i in range(N):
        k = int(sizes[fam[i]])
        d = k if rng.random() < 0.7 else int(rng.integers(1, k + 1))
        first = int(start[fam[i]]) + int(rng.integers(0, k - d + 1))
        s = [0] + list(range(first, first + d))
        v = list(10.0 ** rng.uniform(-12, -3, d))
        nz = 10.0 ** rng.uniform(-14, -5, 1)[0] if rng.random() < junk else 10.0 ** rng.uniform(-60, -40, 1)[0]
        if rng.random() < 0.05:   # a row without a noise entry
            s, v = s[1:], v
        else:
            v = [nz] + v
        rows_sid += s
        rows_val += v
        rp.append(len(rows_sid))
    return np.array(rp, np.uint64), np.array(rows_sid, np.int32), np.array(rows_val, np.float64)


@pytest.mark.parametrize("pf", ["2", "1", "0"])
@pytest.mark.parametrize("max_family,junk", [(12, 0.02), (40, 0.3)])
def test_family_components_match_oracle(ctx, oracle, pf, max_family, junk, monkeypatch):
    """the component walk of the parallel sampler (RSEM_B200_GIBBS_PF = 2 pipelined + converged, 1 pipelined per segment, 0 row-at-a-time) on the matrix shape
    it is built for: many components of very different sizes, segments of hundreds of rows (many 16-slot batches),
    rows that fill a lane group exactly and rows longer than one"""
    N, M = 30000, 1500
    rp, sid, val = _family_matrix(N, M, seed=max_family, max_family=max_family, junk=junk)
    n0 = 1500.0
    init = np.zeros(M + 1, np.int32)
    alpha = np.ones(M + 1)
    totc = (M + 1) * 1.0 + n0 + N
    rng = np.random.default_rng(2)
    eel = np.concatenate([[0.0], rng.uniform(200, 2000, M)])
    mw = np.ones(M + 1)
    genes = np.arange(1, M + 2, 4, dtype=np.int32)
    genes[-1] = M + 1
    samples, burnin, gap = [2, 2], 3, 1
    seeds = oracle.chain_seeds(7, 2)
    monkeypatch.setenv("RSEM_B200_GIBBS", "parallel")
    monkeypatch.setenv("RSEM_B200_GIBBS_PF", pf)
    cv, sums = _run_gpu(ctx, rp, sid, val, M, n0, init, alpha, totc, eel, mw, genes, burnin, gap, samples, seeds)
    at = 0
    for t, ns in enumerate(samples):
        cv_ref, _ = oracle.gibbs_chain(rp, sid, val, M, n0, init, alpha, totc, eel, mw, genes, burnin, gap, ns, int(seeds[t]))
        assert np.array_equal(cv[at:at + ns], cv_ref), f"chain {t} differs"
        at += ns
    assert len(set(cv[:, 0].tolist())) > 1   # reads do move in and out of the noise transcript
    lens = np.diff(rp.astype(np.int64))
    assert lens.max() > 16 or max_family <= 12
