#!/bin/bash
# quick A/B: boundary + column tests, then the bench line without the CPU baseline
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_column.py tests/test_gpu_driver.py tests/test_gpu_next.py -q -x > gpurun_out/quick_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/quick_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-also > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "bench exit $?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/quick_bench.json'))
print('value %.4g ms %.4f e2e %.4g ms %.4f parts %s' % (b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['ms_per_step'], b['e2e']['parts_ms']))
PY
