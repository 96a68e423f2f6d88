#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_transport_flight -s 1 -c 1 -f -o gpurun_out/r02b_skin200_flight python tools/prof_run.py --workload skin200 --packets 2000000 --calls 2 --option flight=1 > gpurun_out/ncu_skin_flight.log 2>&1
echo "ncu skin exit $?"; tail -2 gpurun_out/ncu_skin_flight.log
