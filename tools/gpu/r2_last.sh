#!/bin/bash
# host binaries after the .ofg side-car: Gibbs drop-in tests + acceptance through the Perl driver
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_dropin_gpu.py tests/test_acceptance_gpu.py -x -q -k "gibbs or perl_driver" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/r2w_tests.log
