#!/usr/bin/env python
"""bench.py - EM E+M step throughput of librsem_b200 on B200 (and the reference's CPU arm).

Metric (BASELINE.json): EM reads*hits/s (= hits streamed per second by the frozen-conprb E+M
round, rounds >= 12 of the reference's EM loop) and the equivalent EM iterations/s.

  python bench.py --gpus 1 --steps 20 --warmup 3          # this repo, workload C3
  torchrun ... bench.py --gpus N ...                      # weak scaling, one C3 shard per GPU
  python bench.py --impl reference --steps 20 --warmup 3  # oracle/_ref/rsem-run-em on the host cores

A "step" is one EM round = one pass of K2 (E-step + count accumulation) [+ NCCL allreduce of the
count vector when N > 1] + K4 (theta update, convergence test) over the resident hit matrix.
The hit matrix (12.8 GB at C3) is far larger than L2 (126 MB), so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (reads, transcripts, mean hits/read)            BASELINE.json configs[0..2]
    "C1": (100_000, 5_000, 5),
    "C2": (10_000_000, 50_000, 10),
    "C3": (50_000_000, 200_000, 20),
}
E2E_ROUNDS = 100  # frozen-conprb rounds per end-to-end job (the reference runs >= 20, typically 100s-1000s)


def alg_bytes(N, H, M):
    """12 B per hit + 16 B per read + 16 B per transcript"""
    return 12 * H + 16 * N + 16 * (M + 1)


# --------------------------------------------------------------------------------------------------
def gen_matrix_torch(torch, dev, N, M, deg, seed):
    """C3-shaped matrix generated on the device: degree 1 + Poisson(deg - 1), a row hits consecutive
    transcript ids (isoform families are contiguous), conprb ~ 10^U(-60,-3), ncpv ~ 10^U(-80,-40)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    degs = 1 + torch.poisson(torch.full((N,), float(deg - 1), device=dev), generator=g).to(torch.int64)
    degs.clamp_(max=min(M, 1000))
    row_ptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(degs, 0, out=row_ptr[1:])
    H = int(row_ptr[-1].item())
    start = torch.randint(1, M + 1, (N,), device=dev, generator=g, dtype=torch.int64)
    start = torch.minimum(start, M - degs + 1).clamp_(min=1)
    sid = torch.empty(H, dtype=torch.int32, device=dev)
    CH = 5_000_000  # rows per chunk: bounds the int64 temporaries
    for a in range(0, N, CH):
        b = min(N, a + CH)
        ha, hb = int(row_ptr[a].item()), int(row_ptr[b].item())
        rows = torch.repeat_interleave(torch.arange(a, b, device=dev), degs[a:b])
        within = torch.arange(ha, hb, device=dev) - row_ptr[rows]
        s = (start[rows] + within).to(torch.int32)
        sign = torch.randint(0, 2, (hb - ha,), device=dev, generator=g, dtype=torch.int32) * 2 - 1
        sid[ha:hb] = s * sign
        del rows, within, s, sign
    conprb = torch.empty(H, dtype=torch.float64, device=dev)
    for a in range(0, H, 100_000_000):
        b = min(H, a + 100_000_000)
        u = torch.rand(b - a, device=dev, generator=g, dtype=torch.float64)
        conprb[a:b] = torch.pow(10.0, -60.0 + 57.0 * u)
        del u
    ncpv = torch.pow(10.0, -80.0 + 40.0 * torch.rand(N, device=dev, generator=g, dtype=torch.float64))
    return row_ptr, sid, conprb, ncpv, H


def sort_rows_torch(torch, row_ptr, sid, conprb, ncpv, by):
    """Experiment: rows reordered by their first transcript id (locality of the theta gathers / count reductions)
    or by their degree (uniform row batches)."""
    N = row_ptr.numel() - 1
    degs = row_ptr[1:] - row_ptr[:-1]
    key = sid[row_ptr[:-1]].abs() if by == "start" else degs
    perm = torch.argsort(key, stable=True)
    nd = degs[perm]
    nrp = torch.zeros_like(row_ptr)
    torch.cumsum(nd, 0, out=nrp[1:])
    nsid = torch.empty_like(sid)
    ncon = torch.empty_like(conprb)
    CH = 5_000_000
    for a in range(0, N, CH):
        b = min(N, a + CH)
        ha, hb = int(nrp[a].item()), int(nrp[b].item())
        src = torch.repeat_interleave(row_ptr[perm[a:b]] - nrp[a:b], nd[a:b]) + torch.arange(ha, hb, device=sid.device)
        nsid[ha:hb] = sid[src]
        ncon[ha:hb] = conprb[src]
        del src
    return nrp, nsid, ncon, ncpv[perm].contiguous()


class ClockSampler:
    """nvidia-smi sampling while the GPU runs the timed workload (B200_PROFILING.md 'clocks' recipe).
    nvidia-smi needs ~0.2 s to start and samples every 100 ms, so it is started before the warm-up and only
    samples whose time stamp lies inside [t0, t1] (GPU under the benchmark's load) are used."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p, self.f = index, None, None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _rows(self):
        out = []
        try:
            with open(self.f.name) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 10:
                        continue
                    try:
                        ts = time.mktime(time.strptime(c[0].split(".")[0], "%Y/%m/%d %H:%M:%S")) + float("0." + c[0].split(".")[1])
                        out.append((ts, float(c[2]), float(c[3]), c[6:10]))
                    except (ValueError, IndexError):
                        continue
        except Exception:
            pass
        return out

    def wait_first(self, timeout=3.0):
        t = time.time()
        while self.p and time.time() - t < timeout and not self._rows():
            time.sleep(0.05)

    def count_since(self, t0):
        return sum(1 for r in self._rows() if r[0] >= t0)

    def stop(self, t0, t1):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r for r in self._rows() if t0 <= r[0] <= t1 + 0.05]
        os.unlink(self.f.name)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "timed region + identical rounds right after it (nvidia-smi -lms 100)"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
def cpu_baseline(oracle_mod, row_ptr, sid, conprb, ncpv, M, n0, budget_s=12.0):
    """the oracle port (OpenMP, all host cores) on a bounded sample of the same workload"""
    orc = oracle_mod.Oracle()
    cores = os.cpu_count() or 1
    N = len(row_ptr) - 1
    theta = np.empty(M + 1)
    theta[0] = max(n0 / (N + n0), 1e-8)
    theta[1:] = (1 - theta[0]) / M
    H = int(row_ptr[-1])
    t0 = time.perf_counter()
    theta, _, _ = orc.em_rounds(row_ptr, sid, conprb, ncpv, theta, n0, 12, 2, 1 << 30, 1 << 30, n_threads=cores)
    per = (time.perf_counter() - t0) / 2
    rounds = int(max(3, min(200, budget_s / max(per, 1e-4))))
    t0 = time.perf_counter()
    orc.em_rounds(row_ptr, sid, conprb, ncpv, theta, n0, 14, rounds, 1 << 30, 1 << 30, n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": H * rounds / dt, "unit": "hits/s", "cores": cores, "kind": "port",
            "sample": f"first {N} reads / {H} hits of the workload matrix, {rounds} frozen-conprb rounds, "
                      f"oracle/librsem_oracle.so with {cores} OpenMP threads"}


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import rsem_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    N, M, deg = WORKLOADS[name]
    if args.scale != 1.0:
        N = max(1000, int(N * args.scale))
    if args.deg:
        deg = args.deg

    row_ptr, sid, conprb, ncpv, H = gen_matrix_torch(torch, dev, N, M, deg, seed=1234 + rank)
    if args.sort_rows:
        row_ptr, sid, conprb, ncpv = sort_rows_torch(torch, row_ptr, sid, conprb, ncpv, args.sort_rows)
    n0 = N / 20
    ctx = rsem_b200.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [ctx.lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], world, rank)
    ctx.adopt_device_matrix(N, H, M, row_ptr.data_ptr(), sid.data_ptr(), conprb.data_ptr(), ncpv.data_ptr())
    if args.variant:
        ctx.set_estep_variant(args.variant)
    theta0 = np.empty(M + 1)
    theta0[0] = max(n0 / (N + n0), 1e-8)
    theta0[1:] = (1 - theta0[0]) / M
    ctx.set_theta(theta0)

    # host copies for the e2e leg and the CPU baseline sample (before the device tensors are dropped)
    e2e_host = None
    sample = None
    e2e_err = "skipped (--no-e2e)"
    if not args.no_e2e:
        try:
            e2e_host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (row_ptr, sid, conprb, ncpv)]
            for h, t in zip(e2e_host, (row_ptr, sid, conprb, ncpv)):
                h.copy_(t)
        except Exception as e:  # not enough host memory for 12.8 GB pinned
            e2e_host = None
            e2e_err = str(e).splitlines()[0]
    if rank == 0 and not args.no_cpu_baseline:
        ns = min(N, 2_000_000)
        hs = int(row_ptr[ns].item())
        sample = (row_ptr[: ns + 1].cpu().numpy().astype(np.uint64), sid[:hs].cpu().numpy(), conprb[:hs].cpu().numpy(),
                  ncpv[:ns].cpu().numpy())
    del row_ptr, sid, conprb, ncpv
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + timed region (device-resident inputs) ------------------------------------------
    BIG = 1 << 30
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    ctx.em_rounds(12, args.warmup, BIG, BIG, n0)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_load0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        stats, _ = ctx.em_rounds(12 + args.warmup, args.steps, BIG, BIG, n0)
        e1.record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    # keep the same load running until nvidia-smi has a few samples of it (a 20-round region lasts 70 ms)
    extra = 0
    while world == 1 and sampler.count_since(t_load0) < 5 and time.time() - t_load0 < 2.5:
        ctx.em_rounds(12 + args.warmup + args.steps + extra, 20, BIG, BIG, n0)
        extra += 20
    if world > 1:  # every rank must take part in the allreduce of each round: a fixed continuation
        ctx.em_rounds(12 + args.warmup + args.steps, 200, BIG, BIG, n0)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert len(stats) == args.steps
    ms_per_step = ms / args.steps
    total_hits = H * world  # every rank holds an equally shaped shard (weak scaling)
    value = total_hits / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (K2), per-launch CUDA events on the launching stream ------
    ctx.set_profiling(True)
    ctx.estep_timing(reset=True)
    ctx.em_rounds(12 + args.warmup + args.steps, max(3, min(10, args.steps)), BIG, BIG, n0)
    ctx.sync()
    k2_ms, k2_n = ctx.estep_timing(reset=True)
    ctx.set_profiling(False)
    peak, peak_src = measured_peak_gbs()
    ab = alg_bytes(N, H, M)
    achieved = ab / (k2_ms / k2_n * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "kernel": "estep_rows_kernel (K2)",
                "algorithmic_bytes_per_launch": ab, "k2_ms_per_launch": round(k2_ms / k2_n, 4), "peak_source": peak_src}
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            tr = json.load(f)
            if tr.get("workload") == name:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
    except Exception:
        pass

    # ---- end to end through the C ABI with HOST buffers --------------------------------------------
    e2e = None
    if e2e_host is not None:
        hr, hsid, hc, hn = e2e_host
        bytes_in = hr.numel() * 8 + hsid.numel() * 4 + hc.numel() * 8 + hn.numel() * 8 + (M + 1) * 8
        bytes_out = (M + 1) * 8

        def job():
            ctx.upload_hits_ptr(N, H, M, hr.data_ptr(), hsid.data_ptr())
            ctx.upload_conprb_ptr(hc.data_ptr(), hn.data_ptr())
            ctx.set_theta(theta0)
            ctx.em_rounds(12, E2E_ROUNDS, BIG, BIG, n0)
            return ctx.get_theta()

        job()  # warm-up
        barrier()
        n_jobs = 2
        t0 = time.perf_counter()
        for _ in range(n_jobs):
            th = job()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert abs(th.sum() - 1.0) < 1e-9
        e2e = {"value": total_hits * E2E_ROUNDS * n_jobs / dt, "unit": "hits/s", "h2d_bytes_per_step": bytes_in,
               "d2h_bytes_per_step": bytes_out, "rounds_per_job": E2E_ROUNDS, "jobs": n_jobs,
               "note": "job = upload CSR + conprb from pinned host memory, build tiles, run rounds, read theta back"}
    else:
        e2e = {"value": None, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "no end-to-end leg: " + e2e_err}

    cpu = None
    if rank == 0 and sample is not None:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding
        cpu = cpu_baseline(oracle_binding, *sample, M, len(sample[3]) / 20)

    if rank == 0:
        out = {
            "metric": "em_reads_hits_per_sec", "value": value, "unit": "hits/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "em_iterations_per_sec": 1e3 / ms_per_step,
            "config": {"workload": f"{name}: {N} reads x {M} transcripts, {H} hits per GPU (mean degree {H / N:.2f}), "
                                   "frozen-conprb E+M round (EM.cpp rounds >= 12)",
                       "reads_per_gpu": N, "transcripts": M, "hits_per_gpu": H,
                       "l2_policy": "inputs (12.8 GB) larger than L2, no flush needed" if H * 12 > 2e8 else "inputs fit L2",
                       "parallelism": f"reads sharded over {world} GPU(s), ncclAllReduce(count) per round" if world > 1 else "1 GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "class_layout": ctx.class_layout_info(),
        }
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """the reference's own CPU implementation (oracle/_ref/rsem-run-em-rounds, all host cores) on a
    bounded sample of the workload: same transcriptome size, fewer reads.  Per-round time is taken from
    the timestamps of the reference's own 'ROUND =' stdout lines for the frozen-conprb rounds (>= 12)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    N, M, deg = WORKLOADS[name]
    exe = os.path.join(ROOT, "oracle", "_ref", "rsem-run-em-rounds")
    gen = os.path.join(ROOT, "tools", "gen_dataset")
    idx = os.path.join(ROOT, "oracle", "_ref", "rsem-build-read-index")
    if not (os.path.exists(exe) and os.path.exists(gen) and os.path.exists(idx)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries or tools/gen_dataset missing"}))
        return
    cores = os.cpu_count() or 1
    n_sample = min(N, args.ref_reads)
    rt = 3 if name == "C3" else (1 if name == "C2" else 0)
    rl = 50 if name == "C1" else 100
    with tempfile.TemporaryDirectory(prefix="rsem_ref_bench_") as d:
        subprocess.check_call([gen, "--out", d, "--read-type", str(rt), "--M", str(M), "--N1", str(n_sample), "--N0",
                               str(n_sample // 20), "--avg-family", str(deg), "--read-len", str(rl), "--seed", "11"],
                              stderr=subprocess.DEVNULL)
        files = ([f"{d}/s.temp/s_alignable_1.fq", f"{d}/s.temp/s_alignable_2.fq"] if rt == 3 else
                 [f"{d}/s.temp/s_alignable.fq"] if rt == 1 else [f"{d}/s.temp/s_alignable.fa"])
        subprocess.check_call([idx, "32", str(rt & 1), "1", *files])
        with open(f"{d}/s.temp/s.dat") as f:
            hdr = f.readline().split()
        H = int(hdr[1])
        total_rounds = 11 + args.warmup + args.steps
        env = dict(os.environ, RSEM_MAX_ROUND=str(total_rounds), RSEM_MIN_ROUND=str(total_rounds))
        p = subprocess.Popen([exe, f"{d}/ref/r", str(rt), f"{d}/s", f"{d}/s.temp/s", f"{d}/s.stat/s", "-p", str(cores)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env, text=True, bufsize=1)
        stamps = {}
        for line in p.stdout:
            if line.startswith("ROUND = "):
                stamps[int(line.split(",")[0].split("=")[1])] = time.perf_counter()
        p.wait()
    first = 11 + args.warmup
    dt = stamps[first + args.steps] - stamps[first]
    ms_per_step = dt / args.steps * 1e3
    value = H / (ms_per_step * 1e-3)
    model_ms = (stamps[10] - stamps[1]) / 9 * 1e3
    sample = (f"{n_sample} reads / {H} hits (read_type {rt}, {M} transcripts) generated by tools/gen_dataset; "
              f"reference rsem-run-em -p {cores}; rounds 12.. timed from its ROUND lines; "
              f"model rounds 2-10 took {model_ms:.0f} ms each")
    out = {"impl": "reference", "metric": "em_reads_hits_per_sec", "value": value, "unit": "hits/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{name} (bounded sample: {n_sample} of {N} reads, all {M} transcripts)"},
           "cpu_baseline": {"value": value, "unit": "hits/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": value, "unit": "hits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale the number of reads (debugging only)")
    ap.add_argument("--variant", type=int, default=0, help="E-step kernel variant (0 auto, 1 TMA-staged, 2 direct)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sort-rows", default="", choices=["", "start", "deg"],
                    help="experiment: reorder the reads by first transcript id or by degree")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel experiments)")
    ap.add_argument("--deg", type=int, default=0, help="experiment: override the mean number of hits per read")
    ap.add_argument("--ref-reads", type=int, default=2_000_000, help="reads in the reference arm's bounded sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
