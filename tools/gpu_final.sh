#!/bin/bash
# Final single-GPU evidence run of a round: GPU suite, bench line + reference arm, launch list, one full ncu capture.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r01f}
mkdir -p gpurun_out
: > gpurun_out/steps.log
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/test_gpu_all.log 2>&1
echo "gpu-suite exit $?" | tee -a gpurun_out/steps.log
timeout 420 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
echo "bench exit $?" | tee -a gpurun_out/steps.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
echo "reference exit $?" | tee -a gpurun_out/steps.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch-list exit $?" | tee -a gpurun_out/steps.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_transport_column_parked -s 2 -c 1 -f -o gpurun_out/${TAG}_column_parked_homog200 python tools/prof_run.py --workload homog200 --packets 100000000 --calls 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu-full exit $?" | tee -a gpurun_out/steps.log
tail -3 gpurun_out/test_gpu_all.log
