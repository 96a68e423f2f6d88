#!/bin/bash
# 8 GPUs: also.homog200 (per-rank kernel times of k_transport_column_parked under a communicator) under NCCL settings
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02_n8_homog_variants.txt
for e in "X=1" "NCCL_MAX_NCHANNELS=2" "NCCL_BUFFSIZE=262144" "NCCL_ALGO=Tree" "NCCL_PROTO=LL"; do
env $e timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 2 --warmup 1 --packets 8000000 --no-e2e --no-cpu-baseline --also homog200 > gpurun_out/r02_n8h.json 2> gpurun_out/r02_n8h.err; echo "$e exit $?" | tee -a gpurun_out/r02_n8_homog_variants.txt
python - <<PY | tee -a gpurun_out/r02_n8_homog_variants.txt
import json
b=json.loads(open("gpurun_out/r02_n8h.json").read().splitlines()[-1])
h=b["also"]["homog200"]
print({k:round(h[k],3) if isinstance(h[k],float) else h[k] for k in ("kernel_ms","ms_per_step","allreduce_ms","e2e_ms_per_step","e2e_root_io_ms_per_step")})
print("by rank", [round(x,3) for x in h["kernel_ms_by_rank"]])
PY
done
