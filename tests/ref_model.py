"""ctypes Model / Reads / Refs builders for the oracle and the C ABI, plus a small Python restatement of the
HOST-side model bookkeeping (initial parameters, estimateFromReads, finish) used only to pin the oracle's
update statistics against a one-round run of the reference.  Test infrastructure."""
import ctypes as C

import numpy as np

from oracle_binding import Reads, Refs
from rsem_b200.capi import LenDist, Model, ModelStats


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Keep:
    """owns the numpy arrays a ctypes struct points into"""

    def __init__(self):
        self.arrays = []

    def f64(self, a):
        a = np.ascontiguousarray(a, np.float64)
        self.arrays.append(a)
        return _dp(a)

    def arr(self, a, dt, ct):
        a = np.ascontiguousarray(a, dt)
        self.arrays.append(a)
        return a.ctypes.data_as(C.POINTER(ct))


def make_model(m, seed_len, keep):
    """dict from rsem_files.read_model (or initial_model) -> ctypes Model"""
    mod = Model()
    mod.model_type, mod.M, mod.seed_len = m["model_type"], m["M"], seed_len
    mod.est_rspd, mod.rspd_B = int(m["est_rspd"]), m["B"]
    mod.has_mld = int(m["mld"] is not None)
    mod.pro_len = m["pro_len"]
    mod.ori[0], mod.ori[1] = m["ori0"], 1.0 - m["ori0"]
    g = m["gld"]
    mod.gld = LenDist(g["lb"], g["ub"], g["span"], keep.f64(g["pdf"]), keep.f64(g["cdf"]))
    if m["mld"] is not None:
        d = m["mld"]
        mod.mld = LenDist(d["lb"], d["ub"], d["span"], keep.f64(d["pdf"]), keep.f64(d["cdf"]))
    mod.rspd_pdf, mod.rspd_cdf = keep.f64(m["rspd_pdf"]), keep.f64(m["rspd_cdf"])
    mod.profile, mod.noise_profile, mod.mw = keep.f64(m["profile"]), keep.f64(m["noise"]), keep.f64(m["mw"])
    return mod


def make_reads(r, keep):
    rd = Reads()
    rd.n_mates, rd.has_qual = r["n_mates"], int(r["has_qual"])
    for k in range(r["n_mates"]):
        rd.off[k] = keep.arr(r["off"][k], np.uint64, C.c_uint64)
        rd.base[k] = keep.arr(r["base"][k], np.uint8, C.c_uint8)
        if r["has_qual"]:
            rd.qual[k] = keep.arr(r["qual"][k], np.uint8, C.c_uint8)
    rd.lowq = keep.arr(r["lowq"], np.uint8, C.c_uint8)
    return rd


def make_refs(s, keep):
    rf = Refs()
    rf.M = s["M"]
    rf.seq_off = keep.arr(s["seq_off"], np.uint64, C.c_uint64)
    rf.seq = keep.arr(s["seq"], np.uint8, C.c_uint8)
    rf.full_len = keep.arr(s["full_len"], np.int32, C.c_int32)
    rf.tot_len = keep.arr(s["tot_len"], np.int32, C.c_int32)
    rf.mask_off = keep.arr(s["mask_off"], np.uint64, C.c_uint64)
    rf.mask_words = keep.arr(s["mask_words"], np.uint32, C.c_uint32)
    return rf


def make_stats(m, mp, keep):
    st = ModelStats()
    hasq = m["model_type"] & 1
    arrs = dict(profile=np.zeros(2500 if hasq else m["pro_len"] * 25), noise=np.zeros(500 if hasq else 5),
                gld=np.zeros(mp["maxL"] - (mp["minL"] - 1) + 1), rspd=np.zeros(m["B"] + 2))
    st.profile, st.noise_profile = keep.f64(arrs["profile"]), keep.f64(arrs["noise"])
    st.gld_pdf, st.rspd_pdf = keep.f64(arrs["gld"]), keep.f64(arrs["rspd"])
    st.gld_lb, st.gld_span = mp["minL"] - 1, mp["maxL"] - (mp["minL"] - 1)
    # the arrays handed to ctypes are the contiguous copies kept in `keep`
    n = len(keep.arrays)
    return st, dict(profile=keep.arrays[n - 4], noise=keep.arrays[n - 3], gld=keep.arrays[n - 2], rspd=keep.arrays[n - 1])


# ---- host-side model bookkeeping, restated (only no-polyA, mean < 0 cases: mw == 1) ------------------------
def _trim(d):
    """LenDist::trim, LenDist.h:265-294"""
    pdf, cdf = d["pdf"], d["cdf"]
    nz = np.nonzero(pdf[1:] >= 1e-300)[0]
    newlb, newub = nz[0], nz[-1] + 1
    d["pdf"] = np.concatenate([[0.0], pdf[newlb + 1:newub + 1]])
    d["cdf"] = np.concatenate([[0.0], cdf[newlb + 1:newub + 1]])
    d["span"] = newub - newlb
    d["lb"] += newlb
    d["ub"] = d["lb"] + d["span"]
    return d


def lendist_from_counts(lb, ub, counts):
    pdf = np.asarray(counts, np.float64).copy()
    pdf[0] = 0
    pdf = pdf / pdf.sum()
    return _trim(dict(lb=lb, ub=ub, span=ub - lb, pdf=pdf, cdf=np.cumsum(pdf)))


def uniform_lendist(minL, maxL):
    span = maxL - (minL - 1)
    pdf = np.full(span + 1, 1.0 / span)
    pdf[0] = 0
    cdf = np.arange(span + 1) / span
    return dict(lb=minL - 1, ub=maxL, span=span, pdf=pdf, cdf=cdf)


def initial_model(read_type, mp, M, alignable, unalign):
    """Model(ModelParams&, true) + estimateFromReads for references without poly(A) and mean < 0"""
    hasq, paired = read_type & 1, read_type >= 2
    m = dict(model_type=read_type, M=M, ori0=mp["probF"], est_rspd=mp["estRSPD"], B=mp["B"] if mp["estRSPD"] else 20)
    B = m["B"]
    m["rspd_pdf"] = np.concatenate([[0.0], np.full(B, 1.0 / B), [0.0]])
    m["rspd_cdf"] = np.concatenate([[0.0], np.arange(1, B + 1) / B, [0.0]])
    # length distribution of reads / mates from all non-low-quality reads
    lo, hi = (mp["mate_minL"], mp["mate_maxL"]) if paired else (mp["minL"], mp["maxL"])
    counts = np.zeros(hi - (lo - 1) + 1)
    noise_c = np.zeros(500 if hasq else 5)
    for tag, r in ((0, unalign), (1, alignable)):
        if r is None:
            continue
        for k in range(r["n_mates"]):
            lens = np.diff(r["off"][k].astype(np.int64))
            ok = r["lowq"] == 0
            np.add.at(counts, lens[ok] - (lo - 1), 1.0)
            if tag == 0:
                keep = np.repeat(ok, lens)
                b = r["base"][k][keep].astype(np.int64)
                idx = r["qual"][k][keep].astype(np.int64) * 5 + b if hasq else b
                np.add.at(noise_c, idx, 1.0)
    ld = lendist_from_counts(lo - 1, hi, counts)
    if paired:
        m["gld"], m["mld"] = uniform_lendist(mp["minL"], mp["maxL"]), ld
    else:
        m["gld"], m["mld"] = ld, None
    if hasq:
        p = np.zeros((100, 5, 5))
        for q in range(100):
            probO = np.exp(-q / 10.0 * np.log(10.0))
            probC = (1.0 - probO) * (1 - 1e-5)
            probO = probO / 3 * (1 - 1e-5)
            p[q, :4, :4] = probO
            p[q, np.arange(4), np.arange(4)] = probC
            p[q, :4, 4] = 1e-5
            p[q, 4, :4] = (1 - 1e-5) / 4
            p[q, 4, 4] = 1e-5
        m["pro_len"] = 0
        m["noise"] = ((noise_c.reshape(100, 5) + 1.0) / (noise_c.reshape(100, 5) + 1.0).sum(1, keepdims=True)).ravel()
    else:
        L = mp["maxL"]
        p = np.zeros((L, 5, 5))
        p[:, :4, :4] = 0.01 / 3 * (1 - 1e-5)
        p[:, np.arange(4), np.arange(4)] = 0.99 * (1 - 1e-5)
        p[:, :4, 4] = 1e-5
        p[:, 4, :4] = (1 - 1e-5) / 4
        p[:, 4, 4] = 1e-5
        m["pro_len"] = L
        m["noise"] = (noise_c + 1.0) / (noise_c + 1.0).sum()
    m["profile"] = p.ravel()
    m["noise_c"] = noise_c
    m["mw"] = np.ones(M + 1)
    return m


def finish_from_stats(m, st, mp):
    """Model::init(); collect(); finish() (EM.cpp:400-404) -> dict of the re-estimated tables"""
    hasq, paired = m["model_type"] & 1, m["model_type"] >= 2
    out = {}
    prof = st["profile"].reshape(-1, 5)
    s = prof.sum(1, keepdims=True)
    out["profile"] = np.where(s >= 1e-300, prof / np.where(s >= 1e-300, s, 1.0), 0.0).ravel()
    if hasq:
        tot = (st["noise"] + m["noise_c"]).reshape(100, 5)
        s = tot.sum(1, keepdims=True)
        out["noise"] = np.where(s > 0, tot / np.where(s > 0, s, 1.0), st["noise"].reshape(100, 5)).ravel()
    else:
        tot = st["noise"] + m["noise_c"]
        out["noise"] = tot / tot.sum()
    if paired:
        out["gld"] = lendist_from_counts(mp["minL"] - 1, mp["maxL"], st["gld"])
    if m["est_rspd"]:
        pdf = st["rspd"].copy()
        pdf[1:m["B"] + 1] /= pdf[1:m["B"] + 1].sum()
        out["rspd_pdf"] = pdf
    return out
