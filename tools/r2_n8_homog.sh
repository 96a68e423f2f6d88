#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for e in "X=1" "NCCL_P2P_DISABLE=1"; do
env $e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 2 --warmup 1 --packets 80000000 --no-e2e --also homog200 > gpurun_out/r02_n8h.json 2> gpurun_out/r02_n8h.err; echo "$e exit $?"
python - <<PY
import json
b=json.loads(open("gpurun_out/r02_n8h.json").read().splitlines()[-1])
h=b["also"]["homog200"]
print({k:h[k] for k in ("kernel","kernel_ms","ms_per_step","allreduce_ms","e2e_ms_per_step","e2e_root_io_ms_per_step")})
print("by rank", [round(x,3) for x in h["kernel_ms_by_rank"]])
print("skin by rank", [round(x,2) for x in b["breakdown_ms_per_step"]["kernel_by_rank"]])
PY
done
