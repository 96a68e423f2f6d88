#!/bin/bash
# Is a GPU slower when its neighbour is busy?  N=1 bench on GPU 0 alone, on GPU 1 alone, then both at once (no NCCL).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { CUDA_VISIBLE_DEVICES=$1 python bench.py --no-cpu-baseline --no-also --steps 30 2>/dev/null | python -c "
import json,sys
b=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$2 gpu $1: value %.4g ms %.4f kernel %.4f clocks %s' % (b['value'], b['ms_per_step'], b['breakdown_ms_per_step']['kernel'], b['clocks']))"; }
run 0 alone
run 1 alone
run 0 together & run 1 together & wait
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active --format=csv
