#!/bin/bash
# 2-GPU check: the multi-rank tests, then the bench line at N=2
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_column.py -q -x > gpurun_out/multi_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/multi_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/bench_r01f_n2.json 2> gpurun_out/bench_r01f_n2.err
echo "bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_r01f_n2.json'):
    if l.startswith('{'):
        b=json.loads(l); print('value %.4g ms %.4f breakdown %s e2e %.4g ms %.4f' % (b['value'], b['ms_per_step'], b['breakdown_ms_per_step'], b['e2e']['value'], b['e2e']['ms_per_step']))
PY
