"""Host-side parsers / formatters of the drop-in executables (rsem_b200/host/files.cpp) without a GPU.

`bin/rsem-b200-host-selftest` runs load_dat, parse_reads, write_ofg and load_ofg with a given number of host threads and
dumps the arrays; here they are compared with an independent numpy parse of the same text files (reference formats:
HitContainer.h:62-91, SingleReadQ.h:38-55, EM.cpp:435-457, Gibbs.cpp:101-137) and across thread counts."""
import os
import subprocess

import numpy as np
import pytest

import rsem_files as rf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "rsem-b200-host-selftest")
CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}


def _run(d, read_type, threads, seed_len, out):
    subprocess.check_call([EXE, f"{d}/s.temp/s", str(read_type), str(threads), str(seed_len), out],
                          stdout=subprocess.DEVNULL)


def _load(prefix, name, dtype):
    return np.fromfile(f"{prefix}.{name}", dtype=dtype)


@pytest.mark.parametrize("read_type", [1, 2])
def test_parsers_match_an_independent_parse_for_every_thread_count(built, tmp_path, read_type):
    d = rf.gen_dataset(str(tmp_path / "d"), read_type=read_type, M=300, N1=60000, N0=500, avg_family=6, read_len=60, seed=5)
    paired, hasq = read_type >= 2, bool(read_type & 1)
    # ---- independent parse of .dat
    with open(f"{d}/s.temp/s.dat") as f:
        n1, nh, rt = (int(x) for x in f.readline().split())
        deg, sid, pos, ins = [], [], [], []
        for line in f:
            t = line.split()
            k = int(t[0])
            deg.append(k)
            body = np.array(t[1:], dtype=np.int64).reshape(k, 3 if paired else 2)
            sid.append(body[:, 0]); pos.append(body[:, 1])
            if paired:
                ins.append(body[:, 2])
    assert len(deg) == n1 and rt == read_type
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
    sid, pos = np.concatenate(sid).astype(np.int32), np.concatenate(pos).astype(np.int32)
    assert len(sid) == nh
    # ---- independent parse of the alignable reads
    mates = []
    for path in rf.read_files(d, read_type):
        lines = open(path).read().split("\n")
        step = 4 if hasq else 2
        seqs = lines[1::step][:n1]
        quals = lines[3::step][:n1] if hasq else None
        mates.append((seqs, quals))
    outs = {}
    for threads in (1, 3, 8):
        out = str(tmp_path / f"t{threads}")
        _run(d, read_type, threads, 25, out)
        outs[threads] = out
        assert np.array_equal(_load(out, "row_ptr.u64", np.uint64), row_ptr)
        assert np.array_equal(_load(out, "sid.i32", np.int32), sid)
        assert np.array_equal(_load(out, "pos.i32", np.int32), pos)
        if paired:
            assert np.array_equal(_load(out, "insertL.i32", np.int32), np.concatenate(ins).astype(np.int32))
        for m, (seqs, quals) in enumerate(mates):
            off = _load(out, f"off{m}.u64", np.uint64)
            lens = np.array([len(s) for s in seqs], dtype=np.uint64)
            assert np.array_equal(off, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64))
            base = _load(out, f"base{m}.u8", np.uint8)
            want = np.frombuffer("".join(seqs).upper().encode(), dtype=np.uint8)
            lut = np.full(256, 255, np.uint8)
            for ch, c in CODE.items():
                lut[ord(ch)] = c
            assert np.array_equal(base, lut[want])
            if hasq:
                qual = _load(out, f"qual{m}.u8", np.uint8)
                assert np.array_equal(qual, np.frombuffer("".join(quals).encode(), dtype=np.uint8) - 33)
        lowq = _load(out, "lowq.u8", np.uint8)
        assert len(lowq) == n1 and set(np.unique(lowq)) <= {0, 1}
        # reads shorter than the seed length (25) are low quality (SingleReadQ.h:63-95); none here (read_len 60)
        assert lowq.sum() == 0
        # ---- .ofg round trip: entries < 1e-300 and rows without entries are dropped (EM.cpp:435-457)
        con = _load(out, "in_con.f64", np.float64)
        ncp = _load(out, "in_ncpv.f64", np.float64)
        o_rp = _load(out, "ofg_row_ptr.u64", np.uint64)
        o_sid = _load(out, "ofg_sid.i32", np.int32)
        o_con = _load(out, "ofg_con.f64", np.float64)
        exp_sid, exp_con, exp_rp = [], [], [0]
        for i in range(n1):
            a, b = int(row_ptr[i]), int(row_ptr[i + 1])
            row_s, row_c = [], []
            if ncp[i] >= 1e-300:
                row_s.append(0); row_c.append(ncp[i])
            keep = con[a:b] >= 1e-300
            row_s += list(sid[a:b][keep]); row_c += list(con[a:b][keep])
            if row_s:
                exp_sid += row_s; exp_con += row_c; exp_rp.append(len(exp_sid))
        assert np.array_equal(o_rp, np.array(exp_rp, np.uint64))
        assert np.array_equal(np.abs(o_sid), np.abs(np.array(exp_sid, np.int32)))
        assert np.allclose(o_con, np.array(exp_con), rtol=1e-14, atol=0)  # 15 significant digits in the text
    # identical files for every thread count
    ref = open(outs[1] + ".ofg", "rb").read()
    for threads in (3, 8):
        assert open(outs[threads] + ".ofg", "rb").read() == ref


def test_ofg_sidecar_is_the_text_parse(built, tmp_path):
    """write_ofg leaves `<file>.ofg.b200` next to the text: the rows in upload layout with the doubles the TEXT
    denotes (15 significant digits), so rsem-run-gibbs draws exactly what the reference draws from the text file.  The
    selftest loads both and compares them bit for bit; RSEM_B200_SIDECAR=0 writes / reads no side-car."""
    d = rf.gen_dataset(str(tmp_path / "d"), read_type=0, M=200, N1=20000, N0=300, avg_family=5, read_len=50, seed=9)
    out = str(tmp_path / "o")
    p = subprocess.run([EXE, f"{d}/s.temp/s", "0", "3", "25", out], stdout=subprocess.PIPE, text=True, check=True)
    assert "ofg_sidecar present 1 identical 1" in p.stdout
    hdr = np.fromfile(out + ".ofg.b200", dtype=np.uint64, count=6)
    assert bytes(hdr[:1].tobytes()) == b"RSEMOFG1" and hdr[5] == os.path.getsize(out + ".ofg")
    rows, entries = int(hdr[3]), int(hdr[4])
    assert rows == sum(1 for _ in open(out + ".ofg")) - 1 and entries > rows
    # the stored doubles are what Python parses from the text, not the unrounded inputs
    vals = np.fromfile(out + ".ofg.b200", dtype=np.float64, offset=48 + 8 * (rows + 1) + 4 * entries)
    text_vals = np.array([float(t) for line in list(open(out + ".ofg"))[1:] for t in line.split()[1::2]])
    assert len(vals) == entries and np.array_equal(vals, text_vals)
    assert not np.array_equal(vals, np.fromfile(out + ".in_con.f64")[: len(vals)])
    out2 = str(tmp_path / "o2")
    p = subprocess.run([EXE, f"{d}/s.temp/s", "0", "3", "25", out2], stdout=subprocess.PIPE, text=True, check=True,
                       env=dict(os.environ, RSEM_B200_SIDECAR="0"))
    assert "ofg_sidecar present 0 identical 1" in p.stdout and not os.path.exists(out2 + ".ofg.b200")
    assert open(out2 + ".ofg", "rb").read() == open(out + ".ofg", "rb").read()
