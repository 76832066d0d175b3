"""Readers for RSEM's intermediate / output files, and helpers that run the reference binaries
(oracle/_ref) and the drop-in executables (bin/) on a generated dataset.  Test infrastructure."""
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
BIN_DIR = os.path.join(ROOT, "bin")
GEN = os.path.join(ROOT, "tools", "gen_dataset")


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, b)) for b in ("rsem-run-em-rounds", "rsem-run-gibbs", "rsem-build-read-index"))


def gen_dataset(out, **kw):
    args = [GEN, "--out", out]
    for k, v in kw.items():
        args += ["--" + k.replace("_", "-"), str(v)]
    subprocess.check_call(args, stderr=subprocess.DEVNULL)
    return out


def read_files(d, read_type, tag="alignable"):
    ext = "fq" if read_type & 1 else "fa"
    if read_type >= 2:
        return [f"{d}/s.temp/s_{tag}_1.{ext}", f"{d}/s.temp/s_{tag}_2.{ext}"]
    return [f"{d}/s.temp/s_{tag}.{ext}"]


def clone(src, dst):
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    return dst


def run_em(d, read_type, which, rounds=None, min_rounds=None, threads=1, gibbs_out=True, extra=(), check=True):
    """which = 'ref' (oracle/_ref/rsem-run-em-rounds) or 'ours' (bin/rsem-run-em).  Returns stdout."""
    env = dict(os.environ)
    if rounds is not None:
        env["RSEM_MAX_ROUND"] = str(rounds)
        env["RSEM_MIN_ROUND"] = str(min_rounds if min_rounds is not None else min(20, rounds))
    if which == "ref":
        exe = os.path.join(REF_DIR, "rsem-run-em-rounds")
        subprocess.check_call([os.path.join(REF_DIR, "rsem-build-read-index"), "32", str(read_type & 1), "1", *read_files(d, read_type)])
    else:
        exe = os.path.join(BIN_DIR, "rsem-run-em")
    cmd = [exe, "ref/r", str(read_type), "s", "s.temp/s", "s.stat/s", "-p", str(threads)]
    if gibbs_out:
        cmd.append("--gibbs-out")
    cmd += list(extra)
    p = subprocess.run(cmd, cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if check and p.returncode != 0:
        raise RuntimeError(f"{exe} failed ({p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    return p


def run_gibbs(d, which, burnin, nsamples, gap, threads, seed, extra=(), check=True):
    exe = os.path.join(REF_DIR if which == "ref" else BIN_DIR, "rsem-run-gibbs")
    cmd = [exe, "ref/r", "s.temp/s", "s.stat/s", str(burnin), str(nsamples), str(gap), "-p", str(threads), "--seed", str(seed), "-q", *extra]
    p = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if check and p.returncode != 0:
        raise RuntimeError(f"{exe} failed ({p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    return p


def read_theta(path):
    with open(path) as f:
        n = int(f.readline())
        raw = np.array(f.readline().split(), dtype=np.float64)
        pol = np.array(f.readline().split(), dtype=np.float64)
    assert len(raw) == n and len(pol) == n
    return raw, pol


def read_tokens(path):
    with open(path) as f:
        return np.array(f.read().split(), dtype=np.float64)


def read_res(path):
    """row-major result file -> list of rows (lists of strings)"""
    with open(path) as f:
        return [line.rstrip("\n").split("\t") for line in f]


def read_ofg(path):
    with open(path) as f:
        M, N0 = f.readline().split()
        row_ptr, sid, conprb = [0], [], []
        for line in f:
            t = line.split()
            sid += [int(x) for x in t[0::2]]
            conprb += [float(x) for x in t[1::2]]
            row_ptr.append(len(sid))
    return int(M), int(N0), np.array(row_ptr, np.uint64), np.array(sid, np.int32), np.array(conprb, np.float64)


def read_dat(path, paired):
    with open(path) as f:
        hdr = f.readline().split()
        N = int(hdr[0])
        row_ptr, sid, pos, ins = [0], [], [], []
        for _ in range(N):
            t = [int(x) for x in f.readline().split()]
            k, rest = t[0], t[1:]
            step = 3 if paired else 2
            sid += rest[0::step]
            pos += rest[1::step]
            if paired:
                ins += rest[2::step]
            row_ptr.append(len(sid))
    return (np.array(row_ptr, np.uint64), np.array(sid, np.int32), np.array(pos, np.int32),
            np.array(ins, np.int32) if paired else None)


def close_rel(a, b, rtol, floor=1e-7, atol=1e-12):
    """north_star tolerance: relative error on entries >= floor of the reference, absolute below"""
    a, b = np.asarray(a), np.asarray(b)
    big = np.abs(b) >= floor
    ok_big = np.all(np.abs(a[big] - b[big]) <= rtol * np.abs(b[big]))
    ok_small = np.all(np.abs(a[~big] - b[~big]) <= atol + rtol * np.abs(b[~big]))
    return bool(ok_big and ok_small)


def max_rel(a, b, floor=1e-7):
    a, b = np.asarray(a), np.asarray(b)
    big = np.abs(b) >= floor
    return float(np.max(np.abs(a[big] - b[big]) / np.abs(b[big]))) if big.any() else 0.0
