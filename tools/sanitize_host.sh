#!/bin/bash
# ThreadSanitizer and AddressSanitizer / UBSan runs of bin/rsem-parse-alignments' sources (reader threads, SAM converter threads,
# writer thread) on a generated SAM file and the BAM made from it.  CPU only.  Usage: tools/sanitize_host.sh [workdir]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W="${1:-$(mktemp -d)}"; mkdir -p "$W"
SRC="$ROOT/rsem_b200/host/main_parse.cpp $ROOT/rsem_b200/host/bam.cpp $ROOT/rsem_b200/host/files.cpp $ROOT/rsem_b200/host/model.cpp $ROOT/rsem_b200/host/results.cpp $ROOT/rsem_b200/host/sidecar.cpp"
LINK="-I$ROOT/include -L$ROOT/rsem_b200 -lrsem_b200 -Wl,-rpath,$ROOT/rsem_b200 -lz -pthread"
g++ -O1 -g -fsanitize=thread -std=c++17 -o "$W/parse_tsan" $SRC $LINK
g++ -O1 -g -fsanitize=address,undefined -std=c++17 -o "$W/parse_asan" $SRC $LINK
"$ROOT/tools/gen_dataset" --out "$W/ds" --read-type 3 --M 500 --N1 40000 --N0 2000 --avg-family 6 --read-len 60 --seed 3 --sam 1 > /dev/null 2>&1
"$ROOT/bin/rsem-b200-host-selftest" --bam-copy "$W/ds/aln.sam" "$W/ds.bam" 4 > /dev/null
mkdir -p "$W/o/t" "$W/o/s"
rc=0
for exe in parse_tsan parse_asan; do
  for aln in "$W/ds/aln.sam" "$W/ds.bam"; do
    n=$(ASAN_OPTIONS=detect_leaks=0 RSEM_B200_IO_THREADS=6 "$W/$exe" "$W/ds/ref/r" "$W/o/t/s" "$W/o/s/s" "$aln" 3 -q 2>&1 | grep -cE "WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error" || true)
    echo "$exe $(basename "$aln"): $n reports"
    [ "$n" = "0" ] || rc=1
  done
done
exit $rc
