#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
for opt in "column_park=0" "column_park=-1"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline --no-also --option $opt > gpurun_out/abm.json 2>/dev/null
python - <<PY
import json
for l in open('gpurun_out/abm.json'):
    if l.startswith('{'):
        b=json.loads(l)
        print('$opt', 'value %.4g ms %.4f kernel(max) %.4f allreduce(max) %.4f rank0 kernel %.4f planes %s' % (b['value'], b['ms_per_step'], b['breakdown_ms_per_step']['kernel'], b['breakdown_ms_per_step']['allreduce'], b['roofline']['kernel_ms'], b['breakdown_ms_per_step'].get('allreduce_planes_of_box')))
PY
done
