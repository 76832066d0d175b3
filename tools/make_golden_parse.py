#!/usr/bin/env python
"""Generate tests/golden/parse_*.tar.gz: a small SAM file with the reference index files rsem-parse-alignments reads, and the
files the REFERENCE's rsem-parse-alignments (oracle/_ref, built from /root/reference by oracle/Makefile) writes for it.
Runs in the build container only; the committed fixtures pin bin/rsem-parse-alignments wherever oracle/_ref is absent.

    python tools/make_golden_parse.py
"""
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rsem_files as rf  # noqa: E402
from test_parse_alignments import _tag_some_unaligned  # noqa: E402

CASES = {
    # name: (read_type, generator options, extra argv)
    "parse_pe_q": (3, dict(M=60, N1=500, N0=40, read_len=40, var_len=20, spurious=0.05, omit=4, seed=31), ()),
    "parse_se_noq_tag": (0, dict(M=50, N1=400, N0=120, read_len=36, seed=32), ("-tag", "XM")),
}


def main():
    ref = os.path.join(rf.REF_DIR, "rsem-parse-alignments")
    if not os.path.exists(ref):
        sys.exit("oracle/_ref/rsem-parse-alignments is missing: run `make -C oracle ref` in the build container")
    for name, (rt, opts, extra) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            d = rf.gen_dataset(os.path.join(tmp, "d"), read_type=rt, sam=1, **opts)
            pack = os.path.join(tmp, name)
            os.makedirs(f"{pack}/ref"); os.makedirs(f"{pack}/out/t"); os.makedirs(f"{pack}/out/s")
            for f in ("r.ti", "r.grp"):
                shutil.copy(f"{d}/ref/{f}", f"{pack}/ref/{f}")
            if extra:
                _tag_some_unaligned(f"{d}/aln.sam", f"{pack}/aln.sam", rt >= 2)
            else:
                shutil.copy(f"{d}/aln.sam", f"{pack}/aln.sam")
            subprocess.check_call([ref, f"{pack}/ref/r", f"{pack}/out/t/s", f"{pack}/out/s/s", f"{pack}/aln.sam", str(rt), "-q", *extra])
            with open(f"{pack}/README", "w") as f:
                f.write(f"case {name}: read_type {rt}, tools/gen_dataset options {opts}, argv extra {list(extra)}\n"
                        "out/ = files written by the reference's rsem-parse-alignments (RSEM v1.3.1 @ 8bc1e21)\n")
            with tarfile.open(os.path.join(ROOT, "tests", "golden", name + ".tar.gz"), "w:gz") as tar:
                tar.add(pack, arcname=name)
            print(name, os.path.getsize(os.path.join(ROOT, "tests", "golden", name + ".tar.gz")), "bytes")


if __name__ == "__main__":
    main()
