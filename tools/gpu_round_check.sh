#!/bin/bash
# One gpurun call: new-option parity tests, the bench line + reference arm, then the whole GPU suite.
# Every step writes under gpurun_out/ as it goes, so a clamped call still leaves what it finished.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_next.py -q --durations=12 > gpurun_out/test_next.log 2>&1
echo "next-tests exit $?" | tee -a gpurun_out/steps.log
timeout 420 python bench.py > gpurun_out/bench_r01e_n1.json 2> gpurun_out/bench_r01e_n1.err
echo "bench exit $?" | tee -a gpurun_out/steps.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01e_reference.json 2> gpurun_out/bench_r01e_reference.err
echo "reference exit $?" | tee -a gpurun_out/steps.log
timeout 900 python -m pytest tests -m gpu -q --durations=15 --deselect tests/test_gpu_next.py > gpurun_out/test_gpu_all.log 2>&1
echo "gpu-suite exit $?" | tee -a gpurun_out/steps.log
tail -5 gpurun_out/test_next.log; tail -8 gpurun_out/test_gpu_all.log; head -c 600 gpurun_out/bench_r01e_n1.json
