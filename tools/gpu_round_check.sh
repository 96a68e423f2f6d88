#!/bin/bash
# One gpurun call: the whole GPU suite, the bench line + reference arm, sanitizer on the new paths, launch list.
# Every step writes under gpurun_out/ as it goes, so a clamped call still leaves what it finished.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-r01e}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/steps.log
timeout 900 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/test_gpu_all.log 2>&1
echo "gpu-suite exit $?" | tee -a gpurun_out/steps.log
timeout 420 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
echo "bench exit $?" | tee -a gpurun_out/steps.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
echo "reference exit $?" | tee -a gpurun_out/steps.log
{ echo "== Gaussian beam / periodic boundaries / depth-limited upload, memcheck"; timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_next.py 1500 2>&1 | tail -60
  echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_next.py 600 2>&1 | tail -60; } > gpurun_out/sanitizer_${TAG}.txt 2>&1
echo "sanitizer exit $?" | tee -a gpurun_out/steps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch-list exit $?" | tee -a gpurun_out/steps.log
tail -3 gpurun_out/test_gpu_all.log; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitizer_${TAG}.txt; grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitizer_${TAG}.txt
