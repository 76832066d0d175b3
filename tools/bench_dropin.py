#!/usr/bin/env python
"""Whole-executable timing of the drop-in against the reference on one generated dataset.

    python tools/bench_dropin.py --read-type 1 --N1 1000000 --M 50000 --avg-family 10 --read-len 100 --rounds 20

Runs oracle/_ref/rsem-run-em-rounds (-p nproc) and bin/rsem-run-em on copies of the same inputs with the same fixed
round count, takes per-round times from the timestamps of their `ROUND =` lines, checks theta parity and prints one
JSON line.  (Tooling for profiles/; not part of the product path.)"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rsem_files as rf  # noqa: E402


def timed_run(cmd, cwd, env):
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
    stamps = {}
    for line in p.stdout:
        if line.startswith("ROUND = "):
            stamps[int(line.split(",")[0].split("=")[1])] = time.perf_counter() - t0
    p.wait()
    return time.perf_counter() - t0, stamps, p.returncode


def summarize(wall, st, rounds):
    d = {"wall_s": round(wall, 3), "to_round1_s": round(st[1], 3)}
    if rounds >= 10:
        d["model_round_ms"] = round((st[10] - st[1]) / 9 * 1e3, 2)
    if rounds >= 13:  # round 11 recomputes conprb with the final model, rounds >= 12 are frozen-conprb rounds
        d["rounds_11_to_last_s"] = round(st[rounds] - st[10], 4)
    d["after_last_round_s"] = round(wall - st[rounds], 3)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--read-type", type=int, default=1)
    ap.add_argument("--N1", type=int, default=1_000_000)
    ap.add_argument("--M", type=int, default=50_000)
    ap.add_argument("--avg-family", type=float, default=10)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--est-rspd", type=int, default=0)
    ap.add_argument("--ref-threads", type=int, default=0, help="-p of the reference (0 = all host threads)")
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory(prefix="rsem_dropin_") as tmp:
        base = rf.gen_dataset(os.path.join(tmp, "base"), read_type=a.read_type, M=a.M, N1=a.N1, N0=a.N1 // 20,
                              avg_family=a.avg_family, read_len=a.read_len, est_rspd=a.est_rspd, seed=11)
        n_hits = int(open(f"{base}/s.temp/s.dat").readline().split()[1])
        env = dict(os.environ, RSEM_MAX_ROUND=str(a.rounds), RSEM_MIN_ROUND=str(a.rounds))
        subprocess.check_call([os.path.join(rf.REF_DIR, "rsem-build-read-index"), "32", str(a.read_type & 1), "1",
                               *rf.read_files(base, a.read_type)])
        ref, ours = rf.clone(base, os.path.join(tmp, "ref")), rf.clone(base, os.path.join(tmp, "ours"))
        args = ["ref/r", str(a.read_type), "s", "s.temp/s", "s.stat/s", "--gibbs-out"]
        w_r, s_r, rc_r = timed_run([os.path.join(rf.REF_DIR, "rsem-run-em-rounds"), *args, "-p", str(a.ref_threads or cores)], ref, env)
        w_o0, s_o, rc_o = timed_run([os.path.join(rf.BIN_DIR, "rsem-run-em"), *args, "-p", "8"], ours, env)   # first CUDA process on the box
        w_o, s_o, rc_o = timed_run([os.path.join(rf.BIN_DIR, "rsem-run-em"), *args, "-p", "8"], ours, env)
        assert rc_r == 0 and rc_o == 0, (rc_r, rc_o)
        tr, _ = rf.read_theta(f"{ref}/s.stat/s.theta")
        to, _ = rf.read_theta(f"{ours}/s.stat/s.theta")
        out = {"read_type": a.read_type, "reads": a.N1, "transcripts": a.M, "hits": n_hits, "rounds": a.rounds,
               "host_cores": cores, "reference": summarize(w_r, s_r, a.rounds), "b200": summarize(w_o, s_o, a.rounds),
               "theta_max_rel_err": rf.max_rel(to, tr)}
        out["b200"]["first_run_wall_s"] = round(w_o0, 3)
        out["reference"]["threads"] = a.ref_threads or cores
        out["speedup_wall"] = round(w_r / w_o, 2)
        if a.rounds >= 10 and out["b200"]["model_round_ms"] > 0:
            out["speedup_model_round"] = round(out["reference"]["model_round_ms"] / out["b200"]["model_round_ms"], 1)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
