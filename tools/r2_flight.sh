#!/bin/bash
# flight kernel: parity tests, then A/B timing against the work-queue kernel on skin200 / phantom400
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_production.py -q -x > gpurun_out/r02b_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/r02b_tests.log
for opt in "flight=0" "flight=1" "flight=1 --option flight_inter=1" "flight=1 --option flight_regs=2" "flight=1 --option flight_regs=4" "flight=1 --option walk_min=4" "flight=1 --option walk_min=12"; do
  echo "== skin200 $opt"; timeout 120 python tools/prof_run.py --workload skin200 --packets 20000000 --calls 2 --option $opt
done
for opt in "flight=0" "flight=1" "flight=1 --option flight_inter=0" "flight=1 --option flight_regs=3" "flight=1 --option flight_regs=4"; do
  echo "== phantom400 $opt"; timeout 200 python tools/prof_run.py --workload phantom400 --packets 1000000 --calls 2 --option $opt
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_transport_flight -s 1 -c 1 -f -o gpurun_out/r02d_skin200_flight python tools/prof_run.py --workload skin200 --packets 8000000 --calls 2 --option flight=1 > gpurun_out/ncu_skin_flight.log 2>&1
