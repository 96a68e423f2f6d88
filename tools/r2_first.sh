#!/bin/bash
# round 2, first GPU call: baseline of the scatter-regime kernels before the rewrite
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --no-also > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
echo "bench exit $?"; tail -c 600 gpurun_out/r02a_bench_n1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_transport_pool -s 1 -c 1 -f -o gpurun_out/r02a_skin200_pool python tools/prof_run.py --workload skin200 --packets 2000000 --calls 2 > gpurun_out/ncu_skin.log 2>&1
echo "ncu skin exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_transport_pool -s 1 -c 1 -f -o gpurun_out/r02a_phantom400_pool python tools/prof_run.py --workload phantom400 --packets 200000 --calls 2 > gpurun_out/ncu_phantom.log 2>&1
echo "ncu phantom exit $?"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02a_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r02a_tests.log
