"""Unpack a golden case (tests/golden/<name>.tar.gz, written by tools/make_golden.py) and load everything the
oracle / the C ABI need from it."""
import os
import tarfile

import numpy as np

import ref_model as rm
import rsem_files as rf

GOLDEN = os.path.join(rf.ROOT, "tests", "golden")
CASES = {"se_noq": 0, "se_q_rspd": 1, "pe_noq": 2, "pe_q_rspd": 3, "se_q_polyA": 1}


class Case:
    def __init__(self, name, tmpdir):
        self.name, self.rt = name, CASES[name]
        with tarfile.open(os.path.join(GOLDEN, name + ".tar.gz")) as tar:
            tar.extractall(tmpdir, filter="data")
        self.dir = d = os.path.join(str(tmpdir), name)
        self.paired, self.hasq = self.rt >= 2, bool(self.rt & 1)
        self.mp = rf.read_mparams(f"{d}/s.temp/s.mparams")
        self.refs = rf.read_seq(f"{d}/ref/r.seq")
        self.M = self.refs["M"]
        self.N0, self.N1 = (int(x) for x in open(f"{d}/s.stat/s.cnt").read().split()[:2])
        self.row_ptr, self.sid, self.pos, self.insertL = rf.read_dat(f"{d}/s.temp/s.dat", self.paired)
        self.reads = rf.read_reads(d, self.rt, "alignable", self.refs["has_polyA"], self.mp["seedLen"])
        self.unalign = rf.read_reads(d, self.rt, "un", self.refs["has_polyA"], self.mp["seedLen"]) if self.N0 else None
        self.grp = np.array(open(f"{d}/ref/r.grp").read().split(), np.int32)
        self.omit = np.array(open(f"{d}/s.temp/s.omit").read().split(), np.int64)

    def out(self, f):
        return os.path.join(self.dir, "out", f)

    def theta0(self):
        """EM.cpp:342-346"""
        th = np.empty(self.M + 1)
        th[0] = max(self.N0 / (self.N0 + self.N1), 1e-8)
        th[1:] = (1 - th[0]) / self.M
        return th

    def ofg_dense(self, conprb, ncpv):
        """rows of the .ofg file as (sid, value) lists from dense per-hit arrays (EM.cpp:435-457)"""
        rows = []
        for i in range(len(self.row_ptr) - 1):
            r = []
            if ncpv[i] >= 1e-300:
                r.append((0, ncpv[i]))
            for j in range(int(self.row_ptr[i]), int(self.row_ptr[i + 1])):
                if conprb[j] >= 1e-300:
                    r.append((abs(int(self.sid[j])), conprb[j]))
            if r:
                rows.append(r)
        return rows

    def eel(self, gld):
        """calcExpectedEffectiveLengths, WriteResults.h:24-53"""
        lb, ub, span, pdf, cdf = gld["lb"], gld["ub"], gld["span"], gld["pdf"], gld["cdf"]
        clen = np.concatenate([[0.0], np.cumsum(pdf[1:] * (lb + np.arange(1, span + 1)))])
        eel = np.zeros(self.M + 1)
        for i in range(1, self.M + 1):
            tot, full = int(self.refs["tot_len"][i]), int(self.refs["full_len"][i])
            p1 = max(min(tot - full + 1, ub) - lb, 0)
            p2 = max(min(tot, ub) - lb, 0)
            if p2 == 0:
                continue
            v = full * cdf[p1] + ((cdf[p2] - cdf[p1]) * (tot + 1) - (clen[p2] - clen[p1]))
            eel[i] = 0.0 if v < 1.0 else v
        return eel
