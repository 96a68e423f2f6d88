#!/bin/bash
# what the driver runs at round end: the GPU suite, smoke(), the bench line and the reference arm
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests.log 2>&1
echo "gpu suite exit $?"; tail -3 gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
{ echo "== memcheck"; timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_r02.py 2000 2>&1 | tail -40; echo "== racecheck"; timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_r02.py 500 2>&1 | tail -40; } > gpurun_out/sanitizer_r02_final.txt 2>&1
grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitizer_r02_final.txt
S=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo "bench exit $? in $(( $(date +%s) - S )) s"
grep "bench " gpurun_out/r02_final_bench_n1.err | cut -c1-160
S=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_reference.json 2> gpurun_out/r02_final_reference.err; echo "reference exit $? in $(( $(date +%s) - S )) s"
