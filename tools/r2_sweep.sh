#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for opt in "walk_min=8" "walk_min=6" "walk_min=10" "chunk=32" "chunk=128" "chunk=256"; do
  echo "== skin200 $opt"; timeout 120 python tools/prof_run.py --workload skin200 --packets 40000000 --calls 2 --option $opt | tail -1
done
for opt in "walk_min=8" "walk_min=4" "walk_min=12" "walk_min=16" "walk_min=20" "flight_regs=3 --option walk_min=12" "chunk=32"; do
  echo "== phantom400 $opt"; timeout 200 python tools/prof_run.py --workload phantom400 --packets 2000000 --calls 2 --option $opt | tail -1
done
