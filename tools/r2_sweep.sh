#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for opt in "flight_launch_min=1" "flight_launch_min=2" "flight_launch_min=3" "flight_launch_min=4" "flight_launch_min=6"; do
  echo "== skin200 $opt"; timeout 120 python tools/prof_run.py --workload skin200 --packets 40000000 --calls 2 --option $opt | tail -1
done
for opt in "flight_launch_min=1" "flight_launch_min=3"; do
  echo "== phantom400 $opt"; timeout 200 python tools/prof_run.py --workload phantom400 --packets 2000000 --calls 2 --option $opt | tail -1
done
