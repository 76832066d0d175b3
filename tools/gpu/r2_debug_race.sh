#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic"
for i in 1 2 3 4; do $B 2>&1 | grep -E "CLASSDBG|Error|metric" | cut -c1-200; done
for T in 768 768 513 513; do RSEM_B200_CLASS_THREADS=$T $B 2>&1 | grep -E "CLASSDBG|Error|metric" | cut -c1-200; done
