#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/r02e_tests.log
timeout 500 python bench.py --steps 3 --warmup 1 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
echo "bench exit $?"; tail -c 400 gpurun_out/r02e_bench_n1.err
