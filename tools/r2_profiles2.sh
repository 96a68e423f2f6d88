#!/bin/bash
# final-state evidence: ncu --set full of the flight kernel AT THE BENCH'S LAUNCH SIZE (1e9 skin200 packets), phantom400 at 2e7,
# and the launch list of the bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_transport_flight -s 0 -c 1 -f -o gpurun_out/r02g_skin200_flight_1e9 python tools/prof_run.py --workload skin200 --packets 1000000000 --calls 1 > gpurun_out/ncu1.log 2>&1; echo "ncu skin 1e9 $?"
timeout 600 ncu --set full --clock-control none -k regex:k_transport_flight -s 0 -c 1 -f -o gpurun_out/r02g_phantom400_flight_2e7 python tools/prof_run.py --workload phantom400 --packets 20000000 --calls 1 > gpurun_out/ncu3.log 2>&1; echo "ncu phantom 2e7 $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_r02g.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list $?"
