#!/bin/bash
# round-2 evidence: aggregated-RED A/B, ncu --set full of the flight kernel (skin200, phantom400), launch list of the bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for opt in "flight_agg=0" "flight_agg=1"; do
  echo "== skin200 $opt"; timeout 120 python tools/prof_run.py --workload skin200 --packets 20000000 --calls 3 --option $opt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_transport_flight -s 1 -c 1 -f -o gpurun_out/r02_skin200_flight python tools/prof_run.py --workload skin200 --packets 8000000 --calls 2 > gpurun_out/ncu1.log 2>&1; echo "ncu skin $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_transport_flight -s 1 -c 1 -f -o gpurun_out/r02_skin200_flight_agg python tools/prof_run.py --workload skin200 --packets 8000000 --calls 2 --option flight_agg=1 > gpurun_out/ncu2.log 2>&1; echo "ncu skin agg $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_transport_flight -s 1 -c 1 -f -o gpurun_out/r02_phantom400_flight python tools/prof_run.py --workload phantom400 --packets 1000000 --calls 2 > gpurun_out/ncu3.log 2>&1; echo "ncu phantom $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --packets 100000000 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list $?"
